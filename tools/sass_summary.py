#!/usr/bin/env python
"""tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt -- instruction mix, registers, spills and shared
memory of every kernel in the library (cuobjdump -sass / -res-usage; runs without a GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "modulate_b200", "libmodulate_b200.so")
KEYS = ["IMAD.WIDE.U32", "LEA.HI", "PRMT", "LOP3.LUT", "SHF.R.W.U32", "IMAD.HI.U32", "LDG.E.128", "STG.E.128",
        "LDG.E.128.CONSTANT", "LDG.E.CONSTANT", "LDS.128", "UBLKCP.S.G", "SYNCS.EXCH.64", "SYNCS.ARRIVE.TRANS64",
        "SYNCS.PHASECHK.TRANS64.TRYWAIT", "CCTL.E.PF2", "STG.E", "STG.E.U8", "LDG.E.U8", "BAR.SYNC.DEFER_BLOCKING",
        "CALL.REL.NOINC", "HMMA", "UTCMMA", "UTMALDG"]


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
print(f"# SASS summary of {os.path.relpath(so, ROOT)} (cuobjdump -sass / -res-usage)")
print("# nvcc 12.9 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo; regenerate with tools/sass_summary.py")
print("arch:", ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", sass)))), "\n")
print("## resources (registers / stack = spill / static shared memory)")
for m in re.finditer(r"Function (\S+):\n\s*(.*)", res):
    if "modk" in m.group(1):
        print(f"- {demangle(m.group(1))}: {m.group(2).strip()}")
print()
for f in re.split(r"\n\s*Function : ", sass)[1:]:
    name = f.split("\n", 1)[0].strip()
    if "modk" not in name:
        continue
    ops = collections.Counter()
    for line in f.split("\n"):
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1)] += 1
    print(f"## {demangle(name)}: {sum(ops.values())} instructions")
    shown = []
    for k in KEYS:
        n = sum(v for o, v in ops.items() if o == k or o.startswith(k + "."))
        if n:
            shown.append(f"{k} {n}")
    print("  " + ", ".join(shown))
    print("  local-memory (spill) instructions:", sum(v for o, v in ops.items() if o.startswith(("STL", "LDL"))), "\n")

print("""Reading: the hot loop is IMAD.WIDE.U32 (x 2a) + LEA.HI (Mersenne fold) per byte, three PRMT per four bytes (the
first-level PRMT 0x7340 also gathers the states' top bytes), one three-input LOP3 for the hazard OR and one for the XOR;
128-bit global loads / stores.  Every batched / inline kernel holds TWO copies of that loop: the predicate-free one for
interior tiles of big entries (4 x LDG.E.128 / LDS.128, 4 x STG.E.128, no per-chunk branches; ~280 instructions per
thread for 64 bytes = 4.4 per byte, ncu: 2.38e9 warp instructions for 16 GiB) and the predicated one for an entry's
first / last tile.  The general kernels stage the tile's source span with one bulk-async copy per CTA (UBLKCP) signalled
on an mbarrier (SYNCS.*) and read it back with LDS.128; SHF.R.W funnel shifts re-align it (one code variant per word
shift, hence the instruction count).  CCTL.E.PF2 is the L2 prefetch of the tile records.
The co-aligned kernels carry three more copies of the predicated loop, with 1, 2 and 3 rounds of 128 chunks, for partly
filled tiles.
Local-memory instructions: none in the interior-tile paths.  The co-aligned kernels spill one keystream word (STL + LDL)
in the four-round predicated edge-tile copy only; the general kernels (48 registers since they run at 10 CTAs/SM) spill one to
three words in the edge-tile copies of the loop, and 18 of their local-memory instructions belong to the out-of-line
store_partial helper (caller-saved registers around the byte stores of an entry's partial first / last chunk: at most
two calls per entry).
No tcgen05 / HMMA: there is no contraction on this path (byte-wise integer cipher).""")
