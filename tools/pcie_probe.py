"""tools/pcie_probe.py -- raw pinned<->HBM copy bandwidth of this box (context for the e2e number)."""
import time
import torch

n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


def both():
    h2d(); d2h()


def chunked(chunk):
    def f():
        for o in range(0, n, chunk):
            with torch.cuda.stream(s1):
                d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
            with torch.cuda.stream(s2):
                h2[o:o + chunk].copy_(d2[o:o + chunk], non_blocking=True)
    return f


print(f"H2D 1GiB: {n / t(h2d) / 1e9:.1f} GB/s")
print(f"D2H 1GiB: {n / t(d2h) / 1e9:.1f} GB/s")
print(f"both directions concurrently: {n / t(both) / 1e9:.1f} GB/s each")
for c in (4 << 20, 16 << 20, 64 << 20):
    print(f"both, {c >> 20} MiB chunks: {n / t(chunked(c)) / 1e9:.1f} GB/s each")
