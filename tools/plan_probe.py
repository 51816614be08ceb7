"""tools/plan_probe.py -- cost of mod_plan_create for large descriptor tables (config 4: 1M entries)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import modulate_b200 as mb
import synth

mb.init(0)
for n in (10_000, 100_000, 1_000_000):
    rng = np.random.default_rng(1)
    sizes = rng.integers(1 << 10, (64 << 10) + 1, size=n).astype(np.int64)
    off = synth.packed_offsets(sizes)
    total = int(sizes.sum())
    descs = mb.make_descs(off, off, sizes, synth.entry_keys(n))
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        p = mb.Plan(descs, total, total)
        ts.append(time.perf_counter() - t0)
        tiles = p.num_tiles
        p.close()
    print(f"n={n:>9}  payload {total / 2**30:6.2f} GiB  tiles {tiles:>8}  plan_create best {min(ts) * 1e3:7.2f} ms  first {ts[0] * 1e3:7.2f} ms")
