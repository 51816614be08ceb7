"""tools/copy_align_probe.py -- does the byte alignment of a pinned<->HBM copy matter?  Duplex cudaMemcpyAsync of
16 MiB pieces whose host and device addresses share the phase (address & 63) given on the command line.

    python tools/copy_align_probe.py [phase ...]      (default: 0 16 4 7)
"""
import sys
import time

import torch

phases = [int(x) for x in sys.argv[1:]] or [0, 16, 4, 7]
PIECE, N = 16 << 20, 64
h_in = torch.empty(PIECE * N + 4096, dtype=torch.uint8).pin_memory()
h_out = torch.empty(PIECE * N + 4096, dtype=torch.uint8).pin_memory()
d_in = torch.empty(PIECE * 4 + 4096, dtype=torch.uint8, device="cuda")
d_out = torch.empty(PIECE * 4 + 4096, dtype=torch.uint8, device="cuda")
up, down = torch.cuda.Stream(), torch.cuda.Stream()
for ph in phases:
    def once():
        for i in range(N):
            o, s = i * PIECE + ph, (i % 4) * PIECE + ph
            n = PIECE - 64
            with torch.cuda.stream(up):
                d_in[s:s + n].copy_(h_in[o:o + n], non_blocking=True)
            with torch.cuda.stream(down):
                h_out[o:o + n].copy_(d_out[s:s + n], non_blocking=True)
    once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"phase {ph:2d}: duplex {PIECE * N / dt / 1e9:6.1f} GB/s each way", flush=True)
