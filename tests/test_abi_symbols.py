"""CPU: the C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        names += re.findall(r"\b(mod_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_something():
    fns = declared_functions()
    assert "mod_cycle" in fns and "mod_plan_run" in fns and len(fns) >= 20


def test_library_exports_every_declared_symbol():
    from modulate_b200 import _abi
    lib = _abi.load()
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/ but not exported by the .so"
        assert name in _abi.SIGNATURES, f"{name} has no ctypes prototype"
    assert sorted(_abi.SIGNATURES) == declared_functions()


def test_abi_version_and_desc_layout():
    from modulate_b200 import _abi, DESC_DTYPE
    assert _abi.load().mod_abi_version() == 2
    assert ctypes.sizeof(_abi.ModDesc) == 24 == DESC_DTYPE.itemsize


def test_no_cpu_fallback_when_no_gpu():
    """Compute entry points must fail loudly (never silently compute on the CPU) without a GPU."""
    import modulate_b200 as mb
    try:
        n = mb.device_count()
    except mb.ModError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present")
    buf = bytearray(64)
    with pytest.raises(mb.ModError):
        mb.cycle(buf, 64, 1)
    assert bytes(buf) == bytes(64)  # untouched


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under modulate_b200/ may reference it."""
    pkg = os.path.join(ROOT, "modulate_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "libcycle_ref" not in text, f
