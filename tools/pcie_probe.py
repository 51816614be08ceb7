"""tools/pcie_probe.py -- raw pinned<->HBM copy ceilings of this box, for 1..N GPUs at once.

    python tools/pcie_probe.py [--gpus 8] [--json profiles/pcie.json]

The denominator of bench.py's e2e figure: the e2e path moves every payload byte host->device and
device->host, both directions in flight at once on every GPU.  For n = 1, 2, 4, ... GPUs this probe
runs nothing but cudaMemcpyAsync (one H2D stream and one D2H stream per GPU, 64 MiB copies from / to
per-GPU pinned buffers, one process) and reports
    h2d_gbs / d2h_gbs          one direction alone, summed over the n GPUs
    duplex_payload_gbs         both directions at once: bytes moved ONE way per second, summed over GPUs
                               (= the payload GB/s an ideal e2e pipeline could reach on this host)
"""
import argparse
import json
import time

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
ap.add_argument("--json", default="")
ap.add_argument("--mib", type=int, default=1024, help="bytes per GPU per direction per repetition (MiB)")
args = ap.parse_args()

N = min(args.gpus, torch.cuda.device_count())
PER = args.mib << 20
CHUNK = 64 << 20

bufs = []
for g in range(N):
    with torch.cuda.device(g):
        bufs.append({
            "h_in": torch.empty(PER, dtype=torch.uint8).pin_memory(),
            "h_out": torch.empty(PER, dtype=torch.uint8).pin_memory(),
            "d_in": torch.empty(PER, dtype=torch.uint8, device=f"cuda:{g}"),
            "d_out": torch.empty(PER, dtype=torch.uint8, device=f"cuda:{g}"),
            "s_up": torch.cuda.Stream(device=g), "s_down": torch.cuda.Stream(device=g),
        })
        bufs[-1]["h_in"].fill_(g + 1)


def run(gpus, up, down, reps=4):
    def once():
        for o in range(0, PER, CHUNK):
            for g in gpus:
                b = bufs[g]
                if up:
                    with torch.cuda.stream(b["s_up"]):
                        b["d_in"][o:o + CHUNK].copy_(b["h_in"][o:o + CHUNK], non_blocking=True)
                if down:
                    with torch.cuda.stream(b["s_down"]):
                        b["h_out"][o:o + CHUNK].copy_(b["d_out"][o:o + CHUNK], non_blocking=True)

    def sync():
        for g in gpus:
            torch.cuda.synchronize(g)
    once()
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    sync()
    dt = (time.perf_counter() - t0) / reps
    return PER * len(gpus) / dt / 1e9  # bytes one way, all GPUs, per second


out = {"h2d_gbs": {}, "d2h_gbs": {}, "duplex_payload_gbs": {}, "per_gpu_mib": args.mib, "chunk_mib": CHUNK >> 20,
       "gpu": torch.cuda.get_device_name(0), "how": "tools/pcie_probe.py: cudaMemcpyAsync only, pinned host buffers, one process"}
n = 1
while n <= N:
    gpus = list(range(n))
    out["h2d_gbs"][str(n)] = round(run(gpus, True, False), 2)
    out["d2h_gbs"][str(n)] = round(run(gpus, False, True), 2)
    out["duplex_payload_gbs"][str(n)] = round(run(gpus, True, True), 2)
    print(f"{n} GPU(s): H2D {out['h2d_gbs'][str(n)]:7.1f}  D2H {out['d2h_gbs'][str(n)]:7.1f}  "
          f"duplex {out['duplex_payload_gbs'][str(n)]:7.1f} GB/s each way (sum over GPUs)")
    n *= 2
if args.json:
    with open(args.json, "w") as f:
        json.dump(out, f, indent=1)
