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
