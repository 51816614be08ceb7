"""Synthetic HDR + ARK sets on disk for the facade / CLI tests (built with the oracle)."""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np

import oracle
import synth
from oracle import ark_oracle as ao

DIRS = ["ps4/config", "ps4/songs/credits", "ps4/songs/tut0", "ps4/songs/custom1", "ps4/ui", "ps4/ui/textures", "ps4"]
EXTS = ["dta_dta_ps4", "mogg", "png_ps4", "bin", "moggsong"]


def make_names(n: int, seed: int = 1) -> List[str]:
    rng = np.random.default_rng(seed)
    names = set()
    while len(names) < n:
        d = DIRS[int(rng.integers(0, len(DIRS)))]
        stem = "".join(chr(int(c)) for c in rng.integers(97, 123, size=int(rng.integers(3, 10))))
        if rng.random() < 0.2:
            stem = stem.capitalize()
        names.add(f"{d}/{stem}.{EXTS[int(rng.integers(0, len(EXTS)))]}")
    return sorted(names)


def write_archive(root: str, *, ps4: bool = True, n_files: int = 60, n_parts: int = 3, seed: int = 1,
                  body_key: int = 0, sizes: Optional[Sequence[int]] = None):
    """Write main_<plat>.hdr + part files under `root`.  Returns (Header, plain payload list)."""
    os.makedirs(root, exist_ok=True)
    plat = "ps4" if ps4 else "ps3"
    rng = np.random.default_rng(seed)
    names = make_names(n_files, seed)
    if sizes is None:
        sizes = [int(x) for x in rng.integers(0, 40000, size=n_files)]
        for i in range(0, n_files, 9):
            sizes[i] = 0  # zero-size entries exist in real headers (offset 0, marker 0)
    offsets = synth.packed_offsets(np.array(sizes, dtype=np.int64))
    total = int(sum(sizes))
    payloads = [synth.payload(int(o) + 1000 * seed, int(s)).tobytes() for o, s in zip(offsets, sizes)]
    image = bytearray(total)
    for o, s, p in zip(offsets, sizes, payloads):
        body = np.frombuffer(p, dtype=np.uint8)
        if body_key and s:
            body = oracle.cycle(body, body_key)
        image[int(o):int(o) + s] = body.tobytes()
    # parts: equal shares, the last takes the remainder
    share = total // n_parts
    part_sizes = [share] * (n_parts - 1) + [total - share * (n_parts - 1)]
    parts = [(f"main_{plat}_{i}.ark", s) for i, s in enumerate(part_sizes)]
    entries = [ao.Entry(name=n, offset=(int(o) if s else 0), size=int(s)) for n, o, s in zip(names, offsets, sizes)]
    hdr = ao.Header(ps4=ps4, parts=parts, entries=entries)
    plain = ao.serialise_header(hdr)
    cipher = bytearray(plain)
    cipher[4:] = oracle.cycle(np.frombuffer(plain[4:], dtype=np.uint8), ao.KEY_PS4 if ps4 else ao.KEY_PS3).tobytes()
    with open(os.path.join(root, f"main_{plat}.hdr"), "wb") as f:
        f.write(cipher)
    pos = 0
    for pth, s in parts:
        with open(os.path.join(root, pth), "wb") as f:
            f.write(image[pos:pos + s])
        pos += s
    return hdr, payloads, plain


def read_header(path: str) -> ao.Header:
    raw = open(path, "rb").read()
    magic = int.from_bytes(raw[:4], "little")
    plain = raw[:4] + oracle.cycle(np.frombuffer(raw[4:], dtype=np.uint8), ao.platform_key(magic)).tobytes()
    return ao.parse_header(plain), plain
