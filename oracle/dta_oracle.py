"""oracle/dta_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-Python restatement of the reference's binary DTA ("DTB") reader and writer:
``CDtaFile::Load`` / ``AddTreeNode`` (CDtaFile.cpp:57-100, :393-509) and ``CDtaFile::Save`` /
``SaveToStream`` (CDtaFile.cpp:362-391, :1302-1326; CDtaFile.h:262-284).  ``CDtaFile.cpp`` cannot be
compiled unmodified here (MSVC-only explicit specialisations without ``template<>``,
CDtaFile.h:262-284), so ``make -C oracle ark_ref`` stages it, fixes those four declarations
(``oracle/ref_ark/patch.sed``) and builds it; parity is PINNED by the DTB fixtures the reference's
own Load + Save produced (``tests/golden/dtb/``, ``tests/test_ref_fixtures.py``).  Only ``tests/``
imports this module.

A tree is ``("tree", type, node_id, [children])``; leaves are ``("int", type, value)``,
``("float", 1, value)``, ``("str", type, bytes)``.
"""
from __future__ import annotations

import struct
from typing import List, Tuple

INT_TYPES = (0, 6, 8, 9)        # ENodeType_Integer0/6/8/9, CDtaFile.h:10-23
FLOAT_TYPE = 1
STR_TYPES = (5, 18, 33, 35)     # String, Id, IncludeFile, Define
TREE_TYPES = (16, 17)


class DtaError(Exception):
    pass


def _read_tree(data: bytes, pos: int, ttype: int):
    """AddTreeNode, CDtaFile.cpp:393-509."""
    n, node_id = struct.unpack_from("<hh", data, pos)
    pos += 4
    if n <= 0:
        raise DtaError("eError_InvalidData")          # :398-401
    children = []
    for _ in range(n):
        (ctype,) = struct.unpack_from("<i", data, pos)
        pos += 4
        if ctype in STR_TYPES:
            (ln,) = struct.unpack_from("<i", data, pos)
            pos += 4
            if ln < 0:
                raise DtaError("eError_InvalidData")  # :431-434
            raw = data[pos:pos + ln]
            if len(raw) != ln:
                raise DtaError("truncated")
            children.append(("str", ctype, raw.split(b"\0", 1)[0]))  # std::string(char*) stops at NUL (:445-449)
            pos += ln
        elif ctype in TREE_TYPES:
            pos += 4                                   # :464
            sub, pos = _read_tree(data, pos, ctype)
            children.append(sub)
        elif ctype in INT_TYPES:
            (v,) = struct.unpack_from("<i", data, pos)
            pos += 4
            children.append(("int", ctype, v))
        elif ctype == FLOAT_TYPE:
            v = data[pos:pos + 4]                      # keep the bit pattern (NaN-safe comparison)
            if len(v) != 4:
                raise DtaError("truncated")
            pos += 4
            children.append(("float", 1, v))
        else:
            raise DtaError("eError_InvalidData")      # :503-504
    return ("tree", ttype, node_id, children), pos


def parse(data: bytes) -> List[Tuple]:
    """CDtaFile::Load, CDtaFile.cpp:75-97: skip 5 bytes; the first tree is Tree1; every further
    top-level tree is preceded by (type, word)."""
    pos = 5
    ttype = 16
    trees = []
    try:
        while pos < len(data):
            if ttype not in TREE_TYPES:
                raise DtaError("eError_InvalidData")
            tree, pos = _read_tree(data, pos, ttype)
            trees.append(tree)
            if pos >= len(data):
                break
            (ttype,) = struct.unpack_from("<i", data, pos)
            pos += 8
    except struct.error as exc:
        raise DtaError("truncated") from exc
    return trees


def _write_tree(tree, out: bytearray) -> None:
    """CDtaNodeBase::SaveToStream, CDtaFile.cpp:1310-1326 + the typed leaves CDtaFile.h:262-284."""
    _, _, node_id, children = tree
    out += struct.pack("<hh", len(children), node_id)
    for c in children:
        if c[0] == "tree":
            out += struct.pack("<ii", c[1], 1)
            _write_tree(c, out)
        elif c[0] == "str":
            out += struct.pack("<ii", c[1], len(c[2])) + c[2]
        elif c[0] == "int":
            out += struct.pack("<ii", c[1], c[2])
        else:
            out += struct.pack("<i", c[1]) + c[2]


def serialise(trees: List[Tuple]) -> bytes:
    """CDtaFile::Save, CDtaFile.cpp:364-374, for the single-tree files the reference round-trips;
    further top-level trees get the (type, 1) words Load expects between them."""
    out = bytearray(b"\x01" + struct.pack("<i", 1))
    for i, t in enumerate(trees):
        if i:
            out += struct.pack("<ii", t[1], 1)
        _write_tree(t, out)
    return bytes(out)


def save_like_reference(trees: List[Tuple]) -> bytes:
    """CDtaFile::Save exactly as written, CDtaFile.cpp:364-374: the root's children are streamed back
    to back, WITHOUT the (type, 1) words Load consumes between top-level trees -- so only single-tree
    files round-trip through the reference (pinned by tests/golden/dtb/two_trees)."""
    out = bytearray(b"\x01" + struct.pack("<i", 1))
    for t in trees:
        _write_tree(t, out)
    return bytes(out)


def find_node(tree, name: bytes):
    """CDtaNodeBase::FindNode, CDtaFile.cpp:32-55: depth-first, first string leaf equal to name.
    Returns (parent_children_list, index) or None."""
    for i, c in enumerate(tree[3]):
        if c[0] == "tree":
            hit = find_node(c, name)
            if hit:
                return hit
        elif c[0] == "str" and c[2] == name:
            return tree[3], i
    return None
