"""CPU: host-side logic of the multi-GPU path -- key jump-ahead, offset-range and descriptor
sharding -- checked by replaying each rank's shard through the oracle; plus a world_size-2
gloo run of the same partitioning with one process per rank."""
import os
import subprocess
import sys

import numpy as np

import modulate_b200 as mb
import oracle
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_key_jump_matches_oracle():
    rng = np.random.default_rng(3)
    for key in synth.EDGE_KEYS + [int(x) for x in rng.integers(0, 1 << 32, size=32)]:
        for pos in (0, 1, 16, 4095, 4096, 1 << 20, (1 << 31) - 3, (1 << 31) - 2, (1 << 31) + 5, (1 << 40) + 123):
            assert mb.key_jump(key, pos) == oracle.key_jump(key, pos), (hex(key), pos)


def test_shard_range_covers_and_aligns():
    for total in (0, 1, 15, 16, 1000, (1 << 20) + 7, 16 << 30):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                b, e = mb.shard_range(total, r, world)
                assert b == prev and e >= b
                if r < world - 1 and e < total:
                    assert e % 16 == 0
                prev = e
            assert prev == total


def _random_descs(rng, n, max_len):
    sizes = rng.integers(0, max_len, size=n)
    sizes[rng.integers(0, n, size=max(1, n // 10))] = 0
    src_off = synth.packed_offsets(sizes)
    dst_off = src_off + 3
    return mb.make_descs(src_off, dst_off, sizes, synth.entry_keys(n)), int(sizes.sum())


def test_shard_descs_union_is_byte_identical():
    rng = np.random.default_rng(17)
    for n, max_len in ((1, 100000), (7, 50000), (200, 3000), (50, 200000)):
        descs, total = _random_descs(rng, n, max_len)
        src = synth.payload(0, total)
        want = np.zeros(total + 3, np.uint8)
        oracle.cycle_batch(descs, src, want)
        for world in (1, 2, 3, 8):
            got = np.zeros(total + 3, np.uint8)
            payloads = []
            for r in range(world):
                shard = mb.shard_descs(descs, r, world)
                payloads.append(int(shard["len"].sum()))
                oracle.cycle_batch(shard, src, got)
                # interior cuts land on 16-byte destination boundaries
                for d in shard:
                    full = descs[(descs["src_off"] <= d["src_off"]) & (descs["src_off"] + descs["len"] > d["src_off"])]
                    if len(full) and d["src_off"] != full[0]["src_off"]:
                        assert int(d["dst_off"]) % 16 == 0
            assert sum(payloads) == total
            assert (got == want).all(), (n, world)
            if total > 64 * world:
                assert max(payloads) - min(payloads) <= max(int(descs["len"].max()) if n > world * 4 else 32, 32)


def test_gloo_world2_offset_range_sharding(tmp_path):
    """Two processes (gloo, CPU): each cycles its own offset range of one stream with the jumped
    key through the ORACLE (no GPU here); rank 0 gathers and compares with the unsharded stream.
    Exercises exactly the host logic bench.py uses at N > 1."""
    script = tmp_path / "w2.py"
    script.write_text(f"""
import os, sys
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
import numpy as np, torch, torch.distributed as dist
import modulate_b200 as mb, oracle, synth
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
total, key = 1_000_003, synth.PS4_KEY
b, e = mb.shard_range(total, rank, world)
mine = oracle.cycle(synth.payload(b, e - b), mb.key_jump(key, b))
# descriptor sharding too
sizes = np.array([5, 70000, 0, 300000, 123, 629875]); assert sizes.sum() == total
descs = mb.make_descs(synth.packed_offsets(sizes), synth.packed_offsets(sizes), sizes, synth.entry_keys(len(sizes)))
shard = mb.shard_descs(descs, rank, world)
src = synth.payload(0, total); part = np.zeros(total, np.uint8); oracle.cycle_batch(shard, src, part)
t = torch.from_numpy(part.astype(np.int32)); dist.reduce(t, 0)   # disjoint shards: sum == union
sizes_t = torch.tensor([e - b]); alls = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
dist.all_gather(alls, sizes_t)
if rank == 0:
    chunks = [torch.from_numpy(mine)] + [torch.zeros(int(alls[r]), dtype=torch.uint8) for r in range(1, world)]
    for r in range(1, world): dist.recv(chunks[r], r)
    got = torch.cat(chunks).numpy()
    want = oracle.cycle(synth.payload(0, total), key)
    assert (got == want).all()
    full = np.zeros(total, np.uint8); oracle.cycle_batch(descs, src, full)
    assert (t.numpy().astype(np.uint8) == full).all()
    print('WORLD2 OK')
else:
    dist.send(torch.from_numpy(mine), 0)
dist.barrier()
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    import socket
    with socket.socket() as sock:  # a free rendezvous port (a fixed one may be taken on a shared box)
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "WORLD2 OK" in out.stdout
