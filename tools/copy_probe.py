"""tools/copy_probe.py -- torch device copy bandwidth, burst vs sustained, with clocks/power (context for roofline.peak)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import ClockSampler

n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
for reps in (10, 100, 500, 1500):
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    cs = ClockSampler(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cs.start()
    e0.record()
    for _ in range(reps):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    info = cs.stop()
    ms = e0.elapsed_time(e1) / reps
    print(f"reps {reps:5d}: {2 * n / ms / 1e6:8.1f} GB/s (read+write)  clocks {info}")
