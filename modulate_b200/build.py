"""Build the sm_100a shared library (kernels + C ABI + C++ facade) in-tree with nvcc.

    python -m modulate_b200.build [--force]

The result, ``modulate_b200/libmodulate_b200.so``, is git-ignored but travels to the GPU box
with the gpurun snapshot.  There is no JIT and no fallback: if the library cannot be built or
loaded, importing the compute API raises.
"""
from __future__ import annotations

import fcntl
import glob
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libmodulate_b200.so")
CLI = os.path.join(PKG, "bin", "modulate")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def _deps() -> list[str]:
    return _sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h")) + glob.glob(os.path.join(CSRC, "cli", "*.cpp"))


def _fingerprint() -> str:
    """Content hash of every source the binaries depend on (+ the flags): robust against snapshot
    copies that do not preserve modification times."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in sorted(_deps()):
        h.update(os.path.relpath(d, ROOT).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stamp(target: str) -> str:
    return target + ".stamp"


def stale(target: str) -> bool:
    if not os.path.exists(target) or not os.path.exists(_stamp(target)):
        return True
    with open(_stamp(target)) as f:
        return f.read().strip() != _fingerprint()


def build(force: bool = False, verbose: bool = False) -> str:
    """Build (if stale) under an exclusive file lock: several ranks importing at once must not
    compile into the same output concurrently."""
    lock_path = os.path.join(PKG, ".build.lock")
    with open(lock_path, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            fp = _fingerprint()
            if force or stale(LIB):
                tmp = LIB + f".tmp{os.getpid()}"
                cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                       "-o", tmp, *_sources()]
                if verbose:
                    print(" ".join(cmd), flush=True)
                subprocess.check_call(cmd)
                os.replace(tmp, LIB)  # atomic: a process that already mapped the old file keeps it
                with open(_stamp(LIB), "w") as f:
                    f.write(fp)
            cli_src = sorted(glob.glob(os.path.join(CSRC, "cli", "*.cpp")))
            if cli_src and (force or stale(CLI)):
                os.makedirs(os.path.dirname(CLI), exist_ok=True)
                tmp = CLI + f".tmp{os.getpid()}"
                cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                       "-o", tmp, *cli_src, "-L", PKG, "-lmodulate_b200", "-Wl,-rpath,$ORIGIN/..", "-ldl", "-pthread"]
                if verbose:
                    print(" ".join(cmd), flush=True)
                subprocess.check_call(cmd)
                os.replace(tmp, CLI)
                with open(_stamp(CLI), "w") as f:
                    f.write(fp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
