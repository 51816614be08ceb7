"""CPU: property tests (hypothesis) of the host-side shard planner and key jump-ahead."""
import numpy as np
from hypothesis import given, settings, strategies as st

import modulate_b200 as mb
import oracle


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 5000), st.integers(-2**31, 2**31 - 1)), min_size=1, max_size=40),
       st.integers(1, 9), st.integers(0, 15))
def test_shards_partition_the_batch(entries, world, dst_shift):
    sizes = np.array([e[0] for e in entries], dtype=np.int64)
    keys = np.array([e[1] for e in entries], dtype=np.int64)
    off = np.zeros(len(sizes), np.int64)
    off[1:] = np.cumsum(sizes[:-1])
    descs = mb.make_descs(off, off + dst_shift, sizes, keys)
    total = int(sizes.sum())
    src = np.arange(total, dtype=np.uint32).astype(np.uint8)
    want = oracle.cycle_batch(descs, src, np.zeros(total + dst_shift, np.uint8))
    got = np.zeros(total + dst_shift, np.uint8)
    covered = 0
    for r in range(world):
        shard = mb.shard_descs(descs, r, world)
        covered += int(shard["len"].sum())
        oracle.cycle_batch(shard, src, got)
    assert covered == total
    assert (got == want).all()


@settings(max_examples=200, deadline=None)
@given(st.integers(-2**31, 2**31 - 1), st.integers(0, 2**40), st.integers(0, 2**20))
def test_key_jump_composes(key, a, b):
    assert mb.key_jump(mb.key_jump(key, a), b) == mb.key_jump(key, a + b)
    assert mb.key_jump(key, 0) == oracle.key_residue(key)


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 2**40), st.integers(1, 16))
def test_shard_range_partitions(total, world):
    prev = 0
    for r in range(world):
        b, e = mb.shard_range(total, r, world)
        assert b == prev and b <= e
        prev = e
    assert prev == total


@settings(max_examples=120, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 9000), st.integers(-2**31, 2**31 - 1), st.integers(0, 40)), min_size=1, max_size=60),
       st.sampled_from([1024, 2048, 4096, 16384]), st.integers(0, 127), st.sampled_from([16, 128]))
def test_group_cut_partitions_the_batch(entries, group_bytes, dst_phase, modulus):
    """mod_group_descs (how the host-pointer batch path forms its pipeline groups): the pieces, ciphered one by one,
    give the batch's bytes; pieces stay in order and tile their entries; a group boundary that falls inside an entry
    sits on an aligned destination address; no group holds more than group_bytes + modulus bytes of payload."""
    sizes = np.array([e[0] for e in entries], dtype=np.int64)
    keys = np.array([e[1] for e in entries], dtype=np.int64)
    gaps = np.array([e[2] for e in entries], dtype=np.int64)
    src_off = np.zeros(len(sizes), np.int64)
    src_off[1:] = np.cumsum(sizes[:-1])
    dst_off = np.cumsum(gaps) + src_off  # monotone destination with holes
    descs = mb.make_descs(src_off, dst_off, sizes, keys)
    total = int(sizes.sum())
    src = (np.arange(total, dtype=np.uint32) * 7 + 3).astype(np.uint8)
    dst_len = int((dst_off + sizes).max()) + 1
    want = oracle.cycle_batch(descs, src, np.zeros(dst_len, np.uint8))
    pieces, closes = mb.group_descs(descs, group_bytes, dst_phase, modulus)
    assert len(pieces) == len(closes) and closes[-1] == 1
    got = np.zeros(dst_len, np.uint8)
    oracle.cycle_batch(pieces, src, got)
    assert (got == want).all()
    assert int(pieces["len"].sum()) == total
    # order preserved, pieces tile the source stream, every piece keeps its entry's src -> dst displacement
    run = 0
    ends_src = src_off + sizes
    for p in pieces:
        n = int(p["len"])
        if n == 0:
            continue
        assert int(p["src_off"]) == run
        i = int(np.searchsorted(ends_src, run, side="right"))  # the entry that holds source byte `run`
        assert run + n <= int(ends_src[i]) and int(p["dst_off"]) - run == int(dst_off[i]) - int(src_off[i])
        run += n
    assert run == total
    # boundaries and group sizes
    acc = 0
    ends = {int(o + n) for o, n in zip(dst_off, sizes)}
    for p, c in zip(pieces, closes):
        acc += int(p["len"])
        if c:
            end = int(p["dst_off"]) + int(p["len"])
            if end not in ends:  # the boundary is inside an entry: it must be aligned
                assert (dst_phase + end) % modulus == 0
            assert acc <= group_bytes + modulus
            acc = 0
        else:
            assert acc < group_bytes


def test_group_cut_known_cases():
    """Hand-checked cuts: an entry that crosses the fill point is split at the aligned address just below it (just
    above it when that would leave nothing), an entry that ends exactly on the fill point closes the group whole,
    empty entries ride along, and the tail piece carries the jumped key."""
    key = 0x13579BDF
    # one 5000-byte entry at dst 100, groups of 2048, 128-byte alignment of (dst_phase = 28) + offset
    descs = mb.make_descs([0], [100], [5000], [key])
    pieces, closes = mb.group_descs(descs, 2048, 28, 128)
    # fill point 100 + 2048 = 2148 -> (28 + 2148) % 128 = 0: already aligned; next 2148 + 2048 = 4196, also aligned
    assert [(int(p["dst_off"]), int(p["len"])) for p in pieces] == [(100, 2048), (2148, 2048), (4196, 904)]
    assert list(closes) == [1, 1, 1]
    assert [int(p["key"]) for p in pieces] == [key, mb.key_jump(key, 2048), mb.key_jump(key, 4096)]
    # same entry, phase 0: cuts move down to multiples of 128 of the destination offset
    pieces, closes = mb.group_descs(descs, 2048, 0, 128)
    assert [(int(p["dst_off"]), int(p["len"])) for p in pieces] == [(100, 1948), (2048, 2048), (4096, 1004)]
    assert all((int(p["dst_off"]) + int(p["len"])) % 128 == 0 for p in pieces[:-1])
    # small entries: 3 x 700 bytes packed from 0 with an empty one in between; the third entry crosses 2048
    descs = mb.make_descs([0, 700, 700, 1400, 2100], [0, 700, 700, 1400, 2100], [700, 0, 700, 700, 50], [1, 2, 3, 4, 5])
    pieces, closes = mb.group_descs(descs, 2048, 0, 128)
    assert [(int(p["dst_off"]), int(p["len"])) for p in pieces] == [(0, 700), (700, 0), (700, 700), (1400, 648), (2048, 52), (2100, 50)]
    assert list(closes) == [0, 0, 0, 1, 0, 1]
    assert int(pieces[4]["key"]) == mb.key_jump(4, 648) and int(pieces[3]["key"]) == 4
    # an entry that ends exactly on the fill point is not split
    descs = mb.make_descs([0, 2048], [0, 2048], [2048, 10], [7, 8])
    pieces, closes = mb.group_descs(descs, 2048, 0, 128)
    assert [(int(p["dst_off"]), int(p["len"])) for p in pieces] == [(0, 2048), (2048, 10)] and list(closes) == [1, 1]
    # a cut that would be empty moves UP instead: entry of 300 bytes at dst 2000, fill point 2048 -> aligned 2048 - 2000 = 48
    descs = mb.make_descs([0, 2000], [0, 2000], [2000, 300], [7, 8])
    pieces, closes = mb.group_descs(descs, 2048, 0, 128)
    assert [(int(p["dst_off"]), int(p["len"])) for p in pieces] == [(0, 2000), (2000, 48), (2048, 252)]
    descs = mb.make_descs([0], [2047], [300], [9])  # fill point 2047 + 2048 = 4095: beyond the entry -> whole, closes at the end
    pieces, closes = mb.group_descs(descs, 2048, 0, 128)
    assert [(int(p["dst_off"]), int(p["len"])) for p in pieces] == [(2047, 300)] and list(closes) == [1]
