#!/bin/bash
# tools/cli_timing.sh -- cold-start wall time of the CLI (one process per command, like the reference tool) on a synthetic
# 1 GiB / 10 000-entry archive with ciphered bodies in /dev/shm:   gpurun -- bash tools/cli_timing.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
D=/dev/shm/mod_cli_timing; rm -rf $D; mkdir -p $D
python - <<PY
import sys; sys.path.insert(0, "$ROOT"); sys.path.insert(0, "$ROOT/tests")
import arkfixture, synth
sizes = [int(x) for x in synth.entry_sizes_loguniform(10_000, 1 << 30, lo=1 << 10, hi=1 << 20, seed=7)]
arkfixture.write_archive("$D", n_files=10_000, n_parts=2, seed=5, body_key=0x0BADF00D, sizes=sizes)
PY
python -c "import modulate_b200.build as b; b.build()"
CLI=$ROOT/modulate_b200/bin/modulate
cd $D
for i in 1 2 3; do
  rm -rf out re; mkdir re
  s=$(date +%s.%N); MOD_TRACE=1 $CLI -bodykey 195948557 -unpack out 2> trace_u.txt > /dev/null; e=$(date +%s.%N)
  echo "unpack wall $(python -c "print('%.3f' % ($e - $s))") s   $(grep 'ExtractFiles:' trace_u.txt | sed 's/.*slots; //')"
  python - $s $e trace_u.txt <<'PY'
import re, sys
s, e = float(sys.argv[1]), float(sys.argv[2]); t = open(sys.argv[3]).read()
st = {m.group(1): float(m.group(2)) for m in re.finditer(r"\[mod\] cli (.+?) at ([0-9.]+)", t)}
print("    exec -> main %.3f s, main -> return %.3f s, return -> process gone %.3f s;" % (st["main entered"] - s, st["main returns"] - st["main entered"], e - st["main returns"]),
      "; ".join(l[6:] for l in t.splitlines() if "context +" in l))
PY
  s=$(date +%s.%N); MOD_TRACE=1 $CLI -bodykey 195948557 -packall -pack_add out re 2> trace_p.txt > /dev/null; e=$(date +%s.%N)
  echo "pack   wall $(python -c "print('%.3f' % ($e - $s))") s   $(grep 'SaveArk:' trace_p.txt | sed 's/.*): //')"
done
cd /; rm -rf $D
