"""Tiny device-memory helper for the GPU tests, built on the C ABI only (no torch needed)."""
from __future__ import annotations

import ctypes

import numpy as np

from modulate_b200 import _abi


class DeviceBuffer:
    def __init__(self, nbytes: int):
        self.lib = _abi.load()
        self.nbytes = int(nbytes)
        self.ptr = self.lib.mod_device_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError(self.lib.mod_last_error().decode())

    @classmethod
    def from_numpy(cls, a: np.ndarray) -> "DeviceBuffer":
        a = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        b = cls(a.nbytes)
        b.upload(a)
        return b

    def upload(self, a: np.ndarray, offset: int = 0) -> None:
        a = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        assert offset + a.nbytes <= self.nbytes
        if a.nbytes:
            _abi.check(self.lib.mod_memcpy_h2d(self.ptr + offset, a.ctypes.data, a.nbytes, None))
            _abi.check(self.lib.mod_stream_sync(None))

    def download(self, offset: int = 0, nbytes: int | None = None) -> np.ndarray:
        n = self.nbytes - offset if nbytes is None else nbytes
        out = np.empty(n, dtype=np.uint8)
        if n:
            _abi.check(self.lib.mod_memcpy_d2h(out.ctypes.data, self.ptr + offset, n, None))
            _abi.check(self.lib.mod_stream_sync(None))
        return out

    def fill_payload(self, seed_offset: int = 0, chunk: int = 64 << 20) -> None:
        """Fill with the synthetic stream synth.payload(seed_offset + i)."""
        import synth
        for o in range(0, self.nbytes, chunk):
            n = min(chunk, self.nbytes - o)
            self.upload(synth.payload(seed_offset + o, n), o)

    def free(self) -> None:
        if self.ptr:
            self.lib.mod_device_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def sync() -> None:
    _abi.check(_abi.load().mod_stream_sync(None))
