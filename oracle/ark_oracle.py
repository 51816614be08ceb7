"""oracle/ark_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (plain Python / numpy) of the parts of the reference's ``CArk`` that sit either
side of the hot path: the ``.hdr`` on-disk format, ``LoadArkData``'s part concatenation,
``ExtractFiles``' gather and ``BuildArk``'s offset / part-size assignment.  Only ``tests/`` imports it.

Parity status: PINNED by executing the reference.  ``make -C oracle ark_ref`` compiles the
reference's own ``CArk.cpp`` (Win32 calls resolved by the stand-ins in ``oracle/ref_ark/``) and
``tests/golden/make_ark_golden.py`` commits the headers it wrote, the tables it parsed and the part
sizes it assigned (``tests/golden/ark/``); ``tests/test_ref_fixtures.py`` checks this restatement's
reader, PS3 and PS4 writers and ``build_ark`` against them (header bytes 12..27, uninitialised stack
in the reference, masked).  The cipher under the header is ``oracle.cycle``, itself checked against
the unmodified reference cipher.

Citations are relative to /root/reference/Modulate/.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

MAGIC_PS3 = 0xC64EED30  # Settings.h:16
MAGIC_PS4 = 0x6F303F55  # Settings.h:17
KEY_PS3 = 0xC64EED30    # Settings.h:19
KEY_PS4 = 0x90CFC0AB    # Settings.h:20
UNENCRYPTED_VERSION = 9  # CArk.cpp:309, :903
HASH_PS3 = 0x7D401F60   # CArk.cpp:719
HASH_PS4 = 0xDDB682F0   # CArk.cpp:719
MAX_ARK_SIZE = 512 << 20  # kuMaxArkSize, CArk.cpp:19


class ArkError(Exception):
    """Carries the reference's eError name (Error.h:5-20)."""


@dataclass
class Entry:
    name: str
    offset: int = 0
    size: int = 0
    flags1: int = -1
    flags2: int = -1
    hash: int = 0


@dataclass
class Header:
    ps4: bool
    parts: List[Tuple[str, int]] = field(default_factory=list)  # (path, size)
    entries: List[Entry] = field(default_factory=list)


def platform_key(magic: int) -> int:
    """CArk.cpp:336 -- PS3 magic selects the PS3 key, anything else accepted selects PS4."""
    return KEY_PS3 if magic == MAGIC_PS3 else KEY_PS4


# ---- reading (CArk::Load after the Cycle call, CArk.cpp:341-416) ---------------------------------

class _Reader:
    def __init__(self, data: bytes, pos: int):
        self.d, self.p = data, pos

    def i32(self) -> int:
        v = struct.unpack_from("<i", self.d, self.p)[0]
        self.p += 4
        return v

    def u32(self) -> int:
        v = struct.unpack_from("<I", self.d, self.p)[0]
        self.p += 4
        return v

    def i64(self) -> int:
        v = struct.unpack_from("<q", self.d, self.p)[0]
        self.p += 8
        return v

    def string(self) -> str:
        # sFileDefinition::InitialiseFromData ReadString / sStringList::GetValue (CArk.cpp:533-551,
        # :614-632): i32 length, bytes; the VALUE is clipped to 255 chars but the cursor advances by
        # the full length.
        n = self.i32()
        raw = self.d[self.p:self.p + n]
        self.p += n
        raw = raw[:255]
        raw = raw.split(b"\0", 1)[0]  # std::string(char*) stops at the first NUL
        return raw.decode("latin-1")


def parse_header(plain: bytes) -> Header:
    """`plain` is the whole .hdr file AFTER Cycle() has been applied from byte 4 on."""
    magic = struct.unpack_from("<I", plain, 0)[0]
    if magic not in (MAGIC_PS3, MAGIC_PS4):
        raise ArkError("eError_UnknownVersionNumber")  # CArk.cpp:328-334
    r = _Reader(plain, 4)
    r.u32()              # muVersion      (sHeaderBase, CArk.h:27-32) -- not checked by Load
    r.u32()              # miNumChecksums
    r.p += 16            # mChecksumData
    n_arks = r.i32()
    if n_arks < 0 or n_arks > 100:
        raise ArkError("eError_ValueOutOfBounds")  # CArk.cpp:345-349
    # ark sizes: sIntList {miNum, values}; GetValue(ii) fails if ii >= miNum (CArk.cpp:511-521)
    n_sizes = r.i32()
    if n_sizes < n_arks:
        raise ArkError("eError_ValueOutOfBounds")
    sizes = [r.i32() & 0xFFFFFFFF for _ in range(n_sizes)]
    # ark paths: sStringList
    n_paths = r.i32()
    if n_paths < n_arks:
        raise ArkError("eError_ValueOutOfBounds")
    paths = [r.string() for _ in range(n_paths)]
    # checksums list is skipped whole, then miNum more ints ("hashes"), then a word that must be 0
    n_ck = r.i32()
    r.p += 4 * n_ck      # GetDataSize() = 4 + 4*miNum   (CArk.cpp:381)
    r.p += 4 * n_ck      # "Skip hashes"                 (CArk.cpp:382)
    if r.i32() != 0:
        raise ArkError("eError_InvalidData")  # CArk.cpp:384-388
    n_files = r.i32()
    if n_files < 0 or n_files > 25000:
        raise ArkError("eError_ValueOutOfBounds")  # CArk.cpp:395-399
    entries = []
    for _ in range(n_files):
        # sFileDefinition::InitialiseFromData, CArk.cpp:634-647
        off = r.i64()
        name = r.string()
        flags1 = r.i32()
        size = r.i32()       # read as unsigned, stored in an int (CArk.cpp:637)
        h = r.u32()
        if len(name) == 0 or off < 0 or size < 0:
            raise ArkError("eError_InvalidData")
        entries.append(Entry(name=name, offset=off, size=size, flags1=flags1, hash=h))
    n_f2 = r.i32()
    f2 = [r.i32() for _ in range(max(n_f2, 0))]
    for i, e in enumerate(entries):
        if i >= n_f2:
            raise ArkError("eError_ValueOutOfBounds")  # sIntList::GetValue, CArk.cpp:410-416
        e.flags2 = f2[i]
    return Header(ps4=(magic == MAGIC_PS4), parts=list(zip(paths[:n_arks], sizes[:n_arks])), entries=entries)


# ---- writing (CArk::SaveArk::lSaveHeader, CArk.cpp:901-1136) -------------------------------------

def file_hash(name: str, n_files: int) -> int:
    """lCalculateFileHash, CArk.cpp:832-843: int arithmetic with C truncating remainder; the
    do/while consumes the first char even of an empty string (the terminating NUL)."""
    def trunc_rem(a: int, b: int) -> int:
        q = abs(a) // abs(b)
        if (a < 0) != (b < 0):
            q = -q
        return a - q * b

    def wrap32(v: int) -> int:
        v &= 0xFFFFFFFF
        return v - (1 << 32) if v & 0x80000000 else v

    raw = name.encode("latin-1") + b"\0"
    h = 0
    i = 0
    while True:
        c = raw[i]
        c = c - 256 if c >= 128 else c  # plain char is signed on the reference's targets
        h = wrap32(wrap32(h * 0x7F) + c)
        h = trunc_rem(h, n_files)
        i += 1
        if raw[i] == 0:
            break
    return h


def serialise_entry(e: Entry, ps4: bool) -> bytes:
    """sFileDefinition::Serialise, CArk.cpp:685-721."""
    nm = e.name.encode("latin-1")
    marker = (HASH_PS4 if ps4 else HASH_PS3) if e.size else 0
    return struct.pack("<qi", e.offset, len(nm)) + nm + struct.pack("<iII", e.flags1, e.size & 0xFFFFFFFF, marker)


def bucket_order_ps3(entries: Sequence[Entry]) -> List[int]:
    """PS3 branch of the sort, CArk.cpp:1047-1060: by (name-hash bucket, position in mpFiles)."""
    n = len(entries)
    return sorted(range(n), key=lambda i: (file_hash(entries[i].name, n), i))


def path_order_ps4(entries: Sequence[Entry]) -> List[int]:
    """PS4 branch of the sort, CArk.cpp:969-1045: at every depth a leaf (file) sorts before a
    sub-directory, names compare with _stricmp, full ties by flags1 then flags2.  The reference's
    comparator returns true for equal elements (not a strict weak ordering), so entries that tie
    completely -- the same name twice -- have no defined order; for distinct names this is a total
    order and the reference-written PS4 fixtures confirm it."""
    import functools

    def stricmp(a: str, b: str) -> int:
        la, lb = a.lower(), b.lower()
        return (la > lb) - (la < lb)

    paths = [e.name.split("/") for e in entries]

    def cmp(i: int, j: int) -> int:
        a, b = paths[i], paths[j]
        d = 0
        while True:
            a_leaf, b_leaf = d + 1 == len(a), d + 1 == len(b)
            if a_leaf != b_leaf:
                return -1 if a_leaf else 1
            c = stricmp(a[d], b[d])
            if c:
                return c
            if a_leaf:
                ka = (entries[i].flags1, entries[i].flags2, i)
                kb = (entries[j].flags1, entries[j].flags2, j)
                return (ka > kb) - (ka < kb)
            d += 1

    return sorted(range(len(entries)), key=functools.cmp_to_key(cmp))


def serialise_header(hdr: Header, order: Optional[Sequence[int]] = None, checksum: bytes = b"\0" * 16) -> bytes:
    """Plaintext header bytes in the order lSaveHeader emits them.  `order` is the entry order
    (default: the platform's -- PS3 by name-hash bucket, PS4 by path).  The 16 checksum bytes are
    uninitialised stack in the reference (CArk.cpp:911-921) and zero here."""
    n = len(hdr.entries)
    out = bytearray()
    out += struct.pack("<I", MAGIC_PS4 if hdr.ps4 else MAGIC_PS3)
    out += struct.pack("<II", UNENCRYPTED_VERSION, 1) + checksum + struct.pack("<i", len(hdr.parts))
    out += struct.pack("<i", len(hdr.parts)) + b"".join(struct.pack("<I", s & 0xFFFFFFFF) for _, s in hdr.parts)
    out += struct.pack("<i", len(hdr.parts))
    for p, _ in hdr.parts:
        pb = p.encode("latin-1")
        out += struct.pack("<i", len(pb)) + pb
    out += struct.pack("<i", len(hdr.parts)) + b"\0" * (4 * len(hdr.parts))  # checksums, CArk.cpp:947-953
    out += struct.pack("<i", len(hdr.parts)) + b"\0" * (4 * len(hdr.parts))  # string counts, :955-961
    out += struct.pack("<i", n)
    if order is None:
        order = path_order_ps4(hdr.entries) if hdr.ps4 else bucket_order_ps3(hdr.entries)
    hashes = [file_hash(hdr.entries[i].name, n) for i in order]
    # bucket chain threading, CArk.cpp:1064-1110 (restated literally, including the way a bucket
    # that re-appears later is looked up through the FIRST pair pushed for it)
    hash_offsets: List[List[int]] = []
    prev = -1
    flags1_out = {}
    for pos, idx in enumerate(order):
        h = hashes[pos]
        flags = -1
        for pair in hash_offsets:
            if pair[0] == h:
                flags = pair[1]
                pair[1] = pos
                break
        if prev != -1:
            flags = prev
        flags1_out[idx] = flags
        e = hdr.entries[idx]
        out += serialise_entry(Entry(e.name, e.offset, e.size, flags, e.flags2, e.hash), hdr.ps4)
        prev = pos
        last = pos == len(order) - 1
        if last or hashes[pos + 1] != h:
            hash_offsets.append([h, pos])
            prev = -1
    out += struct.pack("<i", n)
    for bucket in range(n):
        head = -1
        for pair in hash_offsets:
            if pair[0] == bucket:
                head = pair[1]
                break
        out += struct.pack("<i", head)
    return bytes(out)


# ---- data movement -----------------------------------------------------------------------------

def load_ark_data(part_blobs: Sequence[bytes]) -> np.ndarray:
    """CArk::LoadArkData, CArk.cpp:723-758: parts back-to-back in one flat buffer."""
    return np.frombuffer(b"".join(part_blobs), dtype=np.uint8).copy()


def extract(entries: Sequence[Entry], ark: np.ndarray) -> List[bytes]:
    """CArk::ExtractFiles' payload write, CArk.cpp:494: bytes [offset, offset+size) of the flat image."""
    return [ark[e.offset:e.offset + e.size].tobytes() for e in entries]


def plan_part_sizes(total: int, n_parts: int) -> List[int]:
    """ConstructFromDirectory's allowance plan, CArk.cpp:211-217: remaining / parts left."""
    out, remaining = [], total
    for i in range(n_parts):
        s = remaining // (n_parts - i)
        out.append(s)
        remaining -= s
    return out


def build_ark(sizes: Sequence[int], allowances: Sequence[int]) -> Tuple[List[int], List[int]]:
    """CArk::BuildArk, CArk.cpp:784-824: running byte-packed offsets (zero-size entries get 0), and a
    part is closed when its running size EXCEEDS the allowance, the overshoot being carried into the
    next part's allowance.  Returns (offsets, part sizes)."""
    parts = list(allowances)
    offsets = []
    ptr = 0
    part_start = 0
    idx = 0
    allowed = parts[0]
    for s in sizes:
        if s == 0:
            offsets.append(0)
            continue
        offsets.append(ptr)
        ptr += s
        if ptr - part_start > allowed:
            size = ptr - part_start
            parts[idx] = size
            idx += 1
            allowed += parts[idx] - size  # reads the next allowance: IndexError = the reference's overrun
            part_start = ptr
    parts[idx] = ptr - part_start
    return offsets, parts
