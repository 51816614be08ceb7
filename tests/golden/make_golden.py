"""Generate tests/golden/cycle_kat.json by RUNNING THE UNMODIFIED REFERENCE cipher.

Run in the dev container (needs /root/reference; builds oracle/_ref/libcycle_ref.so from
/root/reference/Modulate/CEncryptionCycler.cpp via oracle/Makefile):

    python tests/golden/make_golden.py

Every expected byte in the JSON is an output of the reference's own CEncryptionCycler::Cycle
(CEncryptionCycler.cpp:4-14); nothing is computed by this repo's oracle or kernels.  The GPU box
has no /root/reference, so the committed JSON is what travels.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
import synth   # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main() -> None:
    assert oracle.have_ref(), "reference sources not available"
    out = {"source": "unmodified /root/reference/Modulate/CEncryptionCycler.cpp via oracle/_ref/libcycle_ref.so",
           "keystream_first64": {}, "deep_bytes": {}, "roundtrip": [], "unaligned": [], "batch": {}}

    keys = synth.EDGE_KEYS + [2, 16807, 0x12345678, 0xDEADBEEF, 0x7FFFFFFD, 0x80000001, 127773, 0xFFFF0000]
    for k in keys:
        out["keystream_first64"][f"0x{k:08x}"] = oracle.keystream(k, 64, use_ref=True).tobytes().hex()

    # bytes deep in the stream (the reference reaches them by stepping; we only record them)
    n = (1 << 30) + 1
    for k in (synth.PS4_KEY, synth.PS3_KEY, 1):
        ks = oracle.keystream(k, n, use_ref=True)
        pos = [9999, 65535, 65536, (1 << 20), (1 << 24) + 3, (1 << 28) - 1, (1 << 30) - 16, 1 << 30]
        out["deep_bytes"][f"0x{k:08x}"] = {str(p): int(ks[p]) for p in pos}
        out["deep_bytes"][f"0x{k:08x}"]["window_at_2^29"] = ks[(1 << 29):(1 << 29) + 32].tobytes().hex()
        del ks

    # encrypt / decrypt round trips on synthetic payloads (BASELINE config 1 = 64 KiB, PS4 key)
    for size, key in [(65536, synth.PS4_KEY), (65536, synth.PS3_KEY), (1, 1), (15, 0xDEADBEEF), (17, 0x80000000),
                      (4097, 0x12345678), (1000003, synth.PS4_KEY), (1 << 22, 0xFFFFFFFF)]:
        plain = synth.payload(0, size)
        enc = oracle.cycle(plain, key, use_ref=True)
        dec = oracle.cycle(enc, key, use_ref=True)
        assert (dec == plain).all()
        out["roundtrip"].append({"size": size, "key": f"0x{key:08x}", "plain_sha256": sha(plain),
                                 "cycled_sha256": sha(enc), "cycled_head": enc[:16].tobytes().hex(),
                                 "cycled_tail": enc[-16:].tobytes().hex()})

    # one Cycle call on a window that starts at an arbitrary byte of a larger buffer
    big = synth.payload(0, 1 << 16)
    for start, size, key in [(1, 100, synth.PS4_KEY), (3, 4096, 0x12345678), (15, 33, 1), (16, 31, 0xFFFFFFFF),
                             (255, 40001, synth.PS3_KEY), (7, 7, 0), (13, 5000, 0x7FFFFFFE)]:
        buf = big.copy()
        w = oracle.cycle(buf[start:start + size], key, use_ref=True)
        buf[start:start + size] = w
        out["unaligned"].append({"start": start, "size": size, "key": f"0x{key:08x}", "buffer_sha256": sha(buf)})

    # a small descriptor batch: gather entries out of a packed image with per-entry keys, each
    # entry ciphered by one reference Cycle() call (the CArk.cpp:494 gather + a Cycle per entry)
    rng = np.random.default_rng(1234)
    sizes = np.concatenate([[0, 1, 15, 16, 17, 31, 32, 33, 511, 512, 513, 4095, 4096, 4097],
                            rng.integers(1, 20000, size=50)]).astype(np.int64)
    src_off = synth.packed_offsets(sizes)
    order = rng.permutation(len(sizes))
    dst_off = np.zeros(len(sizes), dtype=np.int64)
    run = 5  # destination packed in a different order, starting at an odd offset
    for i in order:
        dst_off[i] = run
        run += int(sizes[i])
    keys_e = synth.entry_keys(len(sizes))
    src = synth.payload(0, int(sizes.sum()))
    dst = np.full(run + 11, 0xEE, dtype=np.uint8)
    for i in range(len(sizes)):
        s, d, l = int(src_off[i]), int(dst_off[i]), int(sizes[i])
        if l:
            dst[d:d + l] = oracle.cycle(src[s:s + l], int(keys_e[i]), use_ref=True)
    out["batch"] = {"src_off": src_off.tolist(), "dst_off": dst_off.tolist(), "len": sizes.tolist(),
                    "key": [int(k) for k in keys_e], "src_bytes": int(src.size), "dst_bytes": int(dst.size),
                    "dst_fill": 0xEE, "dst_sha256": sha(dst)}

    with open(os.path.join(HERE, "cycle_kat.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "cycle_kat.json"))


if __name__ == "__main__":
    main()
