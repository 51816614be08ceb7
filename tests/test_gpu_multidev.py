"""GPU tests of the round-2 C ABI additions: per-device contexts, the in-process multi-GPU entry
points (mod_cycle_sharded / mod_cycle_batch_sharded; they use every visible device, so on a
one-GPU box they exercise the same code with world == 1), windowed plan runs (the slot-ring
primitive of CArk::ExtractFiles / BuildArk) and the drain-on-error paths.  Bar: bit-exact."""
import ctypes

import numpy as np
import pytest

import oracle
import synth
from gpuutil import DeviceBuffer, sync
from modulate_b200 import _abi

pytestmark = pytest.mark.gpu


def pinned(n):
    L = _abi.load()
    p = L.mod_host_alloc(n)
    assert p
    arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(n,))
    return p, arr


def test_cycle_sharded_vs_oracle(mb):
    n = (40 << 20) + 13
    data = synth.payload(11, n)
    want = oracle.cycle(data, synth.PS3_KEY)
    for mask in (0, 1):  # all visible devices; device 0 only
        buf = data.copy()
        mb.cycle_sharded(buf, n, synth.PS3_KEY, mask)
        assert (buf == want).all(), mask
    # a mask that selects nothing is an error, not a silent no-op
    with pytest.raises(mb.ModError):
        mb.cycle_sharded(data.copy(), n, 1, 1 << 40)
    # device buffers belong to one GPU: refused
    dev = DeviceBuffer(64)
    with pytest.raises(mb.ModError):
        mb.cycle_sharded(dev.ptr, 64, 1, 0)
    dev.free()


def test_cycle_sharded_short_and_odd_lengths(mb):
    for n in (1, 15, 16, 17, 100, 4097):
        data = synth.payload(n, n)
        buf = data.copy()
        mb.cycle_sharded(buf, n, 0xDEADBEEF, 0)
        assert (buf == oracle.cycle(data, 0xDEADBEEF)).all(), n


@pytest.mark.parametrize("layout", ["packed", "inplace", "holes"])
def test_cycle_batch_sharded_vs_oracle(mb, layout, monkeypatch):
    monkeypatch.setenv("MOD_GROUP_BYTES", str(1 << 20))
    n = 700
    sizes = synth.entry_sizes_loguniform(n, 24 << 20, lo=16, hi=1 << 18, seed=9)
    sizes[17] = 0
    src_off = synth.packed_offsets(sizes) + 5
    if layout == "packed":
        dst_off = synth.packed_offsets((sizes + 15) & ~15)
    elif layout == "inplace":
        dst_off = src_off
    else:
        dst_off = src_off * 2 + 3
    keys = synth.entry_keys(n, seed=77)
    descs = mb.make_descs(src_off, dst_off, sizes, keys)
    src = synth.payload(1, int((src_off + sizes).max()) + 3)
    if layout == "inplace":
        want = oracle.cycle_batch(descs, src, src.copy())
        buf = src.copy()
        mb.cycle_batch_sharded(descs, buf, buf)
        assert (buf == want).all()
        return
    dst = np.full(int((dst_off + sizes).max()) + 9, 0x3C, np.uint8)
    want = oracle.cycle_batch(descs, src, dst.copy())
    mb.cycle_batch_sharded(descs, src, dst)
    assert (dst == want).all()


def test_batch_host_large_entries_are_cut(mb, monkeypatch):
    """Entries larger than a group are cut into jumped-key pieces so that their upload, kernel and
    download overlap; the result must not change."""
    monkeypatch.setenv("MOD_GROUP_BYTES", str(1 << 20))
    sizes = np.array([5 << 20, 100, (3 << 20) + 7, 1, (1 << 20) + (1 << 19) + 1], np.int64)
    src_off = synth.packed_offsets(sizes) + 1
    dst_off = synth.packed_offsets(sizes) + 9
    descs = mb.make_descs(src_off, dst_off, sizes, [synth.PS4_KEY, 5, 0x80000000, 7, 0x1234567])
    src = synth.payload(2, int((src_off + sizes).max()))
    dst = np.full(int((dst_off + sizes).max()) + 5, 0x11, np.uint8)
    want = oracle.cycle_batch(descs, src, dst.copy())
    mb.cycle_batch(descs, src, dst)
    assert (dst == want).all()


def test_config3_16gib_through_sharded_api(mb):
    """BASELINE config 3 at FULL size through the in-process multi-GPU API: 32 parts x 512 MiB
    (kuMaxArkSize, CArk.cpp:19), one key per part, in place in pinned HOST memory, split over every
    visible device by mod_cycle_batch_sharded.  Zero plaintext -> the buffer is the keystream:
    sampled windows (both sides of every part boundary and of every shard cut) against the
    closed-form oracle; a second pass restores all-zero (full-buffer check)."""
    part, n_parts = 512 << 20, 32
    total = part * n_parts
    keys = synth.entry_keys(n_parts, seed=303)
    off = np.arange(n_parts, dtype=np.int64) * part
    descs = mb.make_descs(off, off, np.full(n_parts, part, np.int64), keys)
    p, host = pinned(total)
    try:
        host[:] = 0
        mb.cycle_batch_sharded(descs, p, p, total, total)
        world = mb.device_count()
        probes = {0, total - 4096}
        for k in range(1, n_parts):
            probes.update((k * part - 2048, k * part))
        for r in range(1, world):
            cut = total * r // world
            probes.update((cut - 2048, cut))
        rng = np.random.default_rng(5)
        probes.update(int(x) for x in rng.integers(0, total - 4096, size=64))
        for o in sorted(probes):
            k = o // part
            n = min(4096, (k + 1) * part - o)
            want = oracle.cycle_at(np.zeros(n, np.uint8), int(keys[k]), o - k * part)
            assert (host[o:o + n] == want).all(), o
        mb.cycle_batch_sharded(descs, p, p, total, total)
        step = 1 << 30
        for o in range(0, total, step):
            assert not host[o:o + step].any(), o
    finally:
        del host
        _abi.load().mod_host_free(p)


def test_plan_run_window_slot_ring(mb):
    """One plan for the archive, run group by group with only a window of the source and of the
    destination resident (what CArk::ExtractFiles does with its slot ring) == one full run."""
    n = 400
    sizes = synth.entry_sizes_loguniform(n, 16 << 20, lo=1, hi=1 << 18, seed=3)
    sizes[5] = 0
    src_off = synth.packed_offsets(sizes) + 7
    dst_off = synth.packed_offsets(sizes)  # byte-packed staging, like the extract slots
    keys = synth.entry_keys(n, seed=31)
    descs = mb.make_descs(src_off, dst_off, sizes, keys)
    src_np = synth.payload(4, int((src_off + sizes).max()) + 1)
    dst_bytes = int(sizes.sum())
    want = oracle.cycle_batch(descs, src_np, np.zeros(dst_bytes, np.uint8))
    plan = mb.Plan(descs, src_np.size, dst_bytes)
    got = np.zeros(dst_bytes, np.uint8)
    group = 50
    for e0 in range(0, n, group):
        e1 = min(n, e0 + group)
        t0, t1 = plan.tile_range(e0, e1)
        s_lo, s_hi = int(src_off[e0]), int(src_off[e1 - 1] + sizes[e1 - 1])
        d_lo, d_hi = int(dst_off[e0]), int(dst_off[e1 - 1] + sizes[e1 - 1])
        d_lo16 = d_lo & ~15
        s_win = DeviceBuffer.from_numpy(src_np[s_lo:s_hi])
        d_win = DeviceBuffer.from_numpy(np.zeros(d_hi - d_lo16, np.uint8))
        plan.run_window(t0, t1, s_win.ptr, s_lo, s_hi - s_lo, d_win.ptr, d_lo16, d_hi - d_lo16)
        sync()
        out = d_win.download()
        got[d_lo:d_hi] = out[d_lo - d_lo16:]
        assert not out[:d_lo - d_lo16].any()  # bytes of the previous group's last chunk are not touched
        s_win.free()
        d_win.free()
    assert (got == want).all()
    # windows that do not cover the tile range are refused on the host, before any launch
    t0, t1 = plan.tile_range(0, 100)
    s_win = DeviceBuffer(1 << 20)
    with pytest.raises(mb.ModError):
        plan.run_window(t0, t1, s_win.ptr, 0, 10, s_win.ptr, 0, 1 << 20)
    with pytest.raises(mb.ModError):
        plan.run_window(t0, plan.num_tiles + 1, s_win.ptr, 0, 1 << 20, s_win.ptr, 0, 1 << 20)
    s_win.free()
    plan.close()


def test_failure_mid_batch_drains_and_recovers(mb, monkeypatch):
    """A failure at group 3 of a pipelined host batch: the call reports it, nothing is still in
    flight when it returns (the caller's buffers can be reused at once), destination bytes of the
    groups that never ran keep their value, and the next call works."""
    monkeypatch.setenv("MOD_GROUP_BYTES", str(1 << 20))
    n = 300
    sizes = np.full(n, 40_000, np.int64)
    off = synth.packed_offsets(sizes)
    descs = mb.make_descs(off, off, sizes, synth.entry_keys(n, seed=2))
    src = synth.payload(6, int(sizes.sum()))
    dst = np.full(src.size, 0xEE, np.uint8)
    want = oracle.cycle_batch(descs, src, dst.copy())
    monkeypatch.setenv("MOD_TEST_FAIL_GROUP", "3")
    with pytest.raises(mb.ModError):
        mb.cycle_batch(descs, src, dst)
    done = dst == want
    untouched = dst == 0xEE
    assert (done | untouched).all()
    per_group = (1 << 20) // 40_000 + 1
    assert untouched[3 * per_group * 40_000:].all()  # groups >= 3 never ran
    monkeypatch.delenv("MOD_TEST_FAIL_GROUP")
    mb.cycle_batch(descs, src, dst)
    assert (dst == want).all()
    # a descriptor that leaves its buffer, in the middle of the batch: refused before anything runs
    bad = descs.copy()
    bad["src_off"][150] = src.size
    dst2 = np.full(src.size, 0xEE, np.uint8)
    with pytest.raises(mb.ModError):
        mb.cycle_batch(bad, src, dst2)
    assert (dst2 == 0xEE).all()
    mb.cycle_batch(descs, src, dst2)
    assert (dst2 == want).all()


def test_contexts_survive_device_switch(mb):
    """Per-device contexts: using a second GPU does not tear down the first one's streams,
    workspaces or plans (round 1 freed everything on a device change)."""
    if mb.device_count() < 2:
        pytest.skip("needs two GPUs")
    data = synth.payload(0, 3 << 20)
    want = oracle.cycle(data, synth.PS4_KEY)
    descs = mb.make_descs([0], [0], [data.size], [synth.PS4_KEY])
    mb.init(0)
    plan0 = mb.Plan(descs, data.size, data.size)
    d0 = DeviceBuffer.from_numpy(data)
    mb.init(1)
    d1 = DeviceBuffer.from_numpy(data)
    mb.cycle(d1.ptr, data.size, synth.PS4_KEY)
    assert (d1.download() == want).all()
    host = data.copy()
    mb.cycle(host, data.size, synth.PS4_KEY)  # host path on device 1
    assert (host == want).all()
    with pytest.raises(mb.ModError):  # a plan is bound to its device
        plan0.run(d0.ptr, d0.ptr)
    mb.cycle(d0.ptr, data.size, synth.PS4_KEY)  # a resident buffer is cycled where it lives
    mb.init(0)
    assert (d0.download() == want).all()
    plan0.run(d0.ptr, d0.ptr)
    sync()
    assert (d0.download() == data).all()
    for b in (d0, d1):
        b.free()
    plan0.close()


def test_general_kernel_on_coaligned_tiles(mb, monkeypatch):
    """The launch picks the co-aligned kernel when every entry agrees with the base pointers mod 16 and
    the general (shared-memory staged, funnel-shifting) kernel otherwise; both must give the same bytes.
    MOD_FORCE_GENERAL sends co-aligned work through the general kernel."""
    n = 300
    sizes = synth.entry_sizes_loguniform(n, 6 << 20, lo=1, hi=1 << 17, seed=13)
    off = synth.packed_offsets(sizes) + 3
    descs = mb.make_descs(off, off, sizes, synth.entry_keys(n, seed=5))
    src_np = synth.payload(8, int((off + sizes).max()) + 5)
    want = oracle.cycle_batch(descs, src_np, src_np.copy())
    for force in (False, True):
        if force:
            monkeypatch.setenv("MOD_FORCE_GENERAL", "1")
        buf = DeviceBuffer.from_numpy(src_np)
        plan = mb.Plan(descs, src_np.size, src_np.size)
        plan.run(buf.ptr, buf.ptr)
        sync()
        assert (buf.download() == want).all(), force
        mb.cycle_device(buf.ptr, buf.ptr, src_np.size, 77)  # contiguous kernel, same two flavours
        sync()
        assert (buf.download() == oracle.cycle(want, 77)).all(), force
        plan.close()
        buf.free()
