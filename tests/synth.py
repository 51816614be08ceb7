"""Deterministic synthetic inputs shared by tests, bench.py and the golden-vector script.

Payload bytes come from a counter-based generator (splitmix64 of seed ^ word index), so any
window of any buffer can be regenerated independently -- which is what lets a shard, a sampled
range of a 16 GiB set, or the GPU box reproduce exactly the bytes used here.
"""
from __future__ import annotations

import numpy as np

SEED = 0x4D6F64756C617465  # "Modulate"
PS3_KEY = 0xC64EED30       # reference Settings.h:19
PS4_KEY = 0x90CFC0AB       # reference Settings.h:20
EDGE_KEYS = [0, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 1, 0x7FFFFFFE, PS3_KEY, PS4_KEY]

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def payload(offset: int, n: int, seed: int = SEED) -> np.ndarray:
    """Bytes [offset, offset+n) of the infinite synthetic stream for `seed`."""
    if n == 0:
        return np.zeros(0, dtype=np.uint8)
    w0 = offset // 8
    w1 = (offset + n + 7) // 8
    out = np.empty((w1 - w0) * 8, dtype=np.uint8)
    step = 1 << 22
    for s in range(w0, w1, step):
        e = min(w1, s + step)
        idx = np.arange(s, e, dtype=np.uint64) ^ np.uint64(seed)
        out[(s - w0) * 8:(e - w0) * 8] = splitmix64(idx).view(np.uint8)
    lo = offset - w0 * 8
    return out[lo:lo + n]


def i32(key: int) -> int:
    key &= 0xFFFFFFFF
    return key - (1 << 32) if key & 0x80000000 else key


def entry_keys(n: int, seed: int = SEED) -> np.ndarray:
    """Per-entry keys: (int32) splitmix64(seed ^ i), entries 0..4 forced to the edge keys
    0, 0x7fffffff, 0x80000000, -1, 1."""
    k = (splitmix64(np.arange(n, dtype=np.uint64) ^ np.uint64(seed ^ 0xA5A5A5A5)) & np.uint64(0xFFFFFFFF))
    k = k.astype(np.uint32).view(np.int32).copy()
    edge = np.array([0, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 1], dtype=np.uint32).view(np.int32)
    k[:min(n, 5)] = edge[:min(n, 5)]
    return k


def entry_sizes_loguniform(n: int, total: int, lo: int = 1 << 10, hi: int = 1 << 20, seed: int = 7) -> np.ndarray:
    """n sizes, log-uniform in [lo, hi], rescaled so they sum to exactly `total` (config 2)."""
    rng = np.random.default_rng(seed)
    s = np.exp(rng.uniform(np.log(lo), np.log(hi), size=n))
    s = np.maximum(1, np.floor(s * (total / s.sum()))).astype(np.int64)
    s[-1] += total - int(s.sum())
    assert s.min() >= 1 and int(s.sum()) == total
    return s


def packed_offsets(sizes: np.ndarray) -> np.ndarray:
    """Byte-packed running offsets in BuildArk order (reference CArk.cpp:807-811: no padding)."""
    off = np.zeros(len(sizes), dtype=np.int64)
    if len(sizes) > 1:
        off[1:] = np.cumsum(sizes[:-1])
    return off
