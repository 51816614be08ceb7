"""tools/latency_probe.py -- device-resident Cycle() time vs buffer size (CUDA events, in place)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import modulate_b200 as mb

mb.init(0)
st = torch.cuda.current_stream()
for size in (4096, 65536, 384 * 1024, 1 << 20, 4 << 20, 16 << 20, 64 << 20, 256 << 20, 1 << 30):
    buf = torch.zeros(size, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        mb.cycle_device(buf.data_ptr(), buf.data_ptr(), size, 0x90CFC0AB, st.cuda_stream)
    torch.cuda.synchronize()
    n = 50
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(n):
        mb.cycle_device(buf.data_ptr(), buf.data_ptr(), size, 0x90CFC0AB, st.cuda_stream)
    b.record(st)
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / n * 1e3
    print(f"{size:>12} B  {us:9.2f} us  {size / us / 1e3:9.1f} GB/s")
