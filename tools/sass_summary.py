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
