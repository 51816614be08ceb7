"""tools/unpack_timing.py -- wall-clock of the CLI on a synthetic 1 GiB / 10 000-entry archive in tmpfs
(BASELINE config 5 shape: -unpack, -dtaset, -pack_add), to see the host pipeline around the kernel."""
import os, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import synth, arkfixture

CLI = os.path.join(ROOT, "modulate_b200", "bin", "modulate")
base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
root = tempfile.mkdtemp(prefix="modark_", dir=base)
try:
    total = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 30)
    sizes = [int(x) for x in synth.entry_sizes_loguniform(10_000, total, lo=1 << 10, hi=1 << 20, seed=7)]
    key = 0x0BADF00D
    t0 = time.perf_counter()
    arkfixture.write_archive(root, n_files=10_000, n_parts=2, seed=5, body_key=key, sizes=sizes)
    print(f"fixture written in {time.perf_counter() - t0:.1f} s under {root}")
    for label, args in (("unpack (bodies ciphered)", ["-bodykey", str(key), "-unpack", "out"]),
                        ("unpack again (warm page cache)", ["-bodykey", str(key), "-unpack", "out2"]),
                        ("pack_add -packall", ["-bodykey", str(key), "-packall", "-pack_add", "out", "re"])):
        t0 = time.perf_counter()
        r = subprocess.run(["bash", "-c", "time " + " ".join([CLI, *args])], cwd=root, capture_output=True, text=True, env=dict(os.environ, MOD_TRACE="1"))
        print("   ", " | ".join(l for l in r.stderr.splitlines() if l.startswith("[mod] Extract") or l.startswith("[mod] Unpack") or l.startswith("real")))
        dt = time.perf_counter() - t0
        print(f"{label:34s} rc={r.returncode}  {dt:6.2f} s  {total / dt / 1e9:6.2f} GB/s")
        if r.returncode:
            print(r.stdout[-500:], r.stderr[-500:])
finally:
    shutil.rmtree(root, ignore_errors=True)
