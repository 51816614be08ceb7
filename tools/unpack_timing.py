"""tools/unpack_timing.py -- the facade's file pipelines on a synthetic 1 GiB / 10 000-entry archive in a RAM-backed
file system (BASELINE config 5 shape), in process through include/modulate_ark.h, for several thread / slot settings.

    python tools/unpack_timing.py [total_bytes]

Prints GB/s of payload for mod_ark_unpack and mod_ark_pack (after one warm-up call each) per setting of the
MOD_IO_READERS / MOD_IO_WRITERS / MOD_IO_SLOTS / MOD_IO_GROUP_MIB tuning variables."""
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import arkfixture  # noqa: E402
import modulate_b200 as mb  # noqa: E402
import synth  # noqa: E402

base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
root = tempfile.mkdtemp(prefix="modark_", dir=base)
try:
    total = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 30)
    sizes = [int(x) for x in synth.entry_sizes_loguniform(10_000, total, lo=1 << 10, hi=1 << 20, seed=7)]
    key = 0x0BADF00D
    arkfixture.write_archive(root, n_files=10_000, n_parts=2, seed=5, body_key=key, sizes=sizes)
    hdr = os.path.join(root, "main_ps4.hdr")
    mb.init(0)
    settings = [{}] + [{"MOD_IO_READERS": r, "MOD_IO_WRITERS": w} for r, w in ((2, 4), (4, 8), (4, 16), (8, 16), (8, 24), (12, 32))] + \
        [{"MOD_IO_SLOTS": s} for s in (3, 6, 12)] + [{"MOD_IO_GROUP_MIB": g} for g in (8, 16, 64)]
    print(f"host threads: {os.cpu_count()}, file system: {base}")
    for cfg in settings:
        for k in ("MOD_IO_READERS", "MOD_IO_WRITERS", "MOD_IO_SLOTS", "MOD_IO_GROUP_MIB"):
            os.environ.pop(k, None)
        for k, v in cfg.items():
            os.environ[k] = str(v)
        out, re_dir = os.path.join(root, "out"), os.path.join(root, "re")
        best_u, best_p = 0.0, 0.0
        for rep in range(3):  # rep 0 warms up (page-locked slots, page cache)
            shutil.rmtree(out, ignore_errors=True)
            t0 = time.perf_counter()
            mb.ark_unpack(hdr, root, out, key)
            t1 = time.perf_counter()
            shutil.rmtree(re_dir, ignore_errors=True)
            os.makedirs(re_dir)
            t2 = time.perf_counter()
            mb.ark_pack(hdr, out, re_dir, "main_ps4.hdr", pack_all=True, ignore_new_files=False, body_key=key)
            t3 = time.perf_counter()
            if rep:
                best_u, best_p = max(best_u, total / (t1 - t0) / 1e9), max(best_p, total / (t3 - t2) / 1e9)
        print(f"{str(cfg or 'defaults'):60s} unpack {best_u:6.2f} GB/s   pack {best_p:6.2f} GB/s", flush=True)
finally:
    shutil.rmtree(root, ignore_errors=True)
