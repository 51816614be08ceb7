tools/gpu_ab.sh "gf0 gf1 gf0 gf1" "cfg2"
for v in gf0 gf1; do
echo -n "$v cfg2 sustained: "
MODULATE_B200_LIB=$PWD/variants/libmod_$v.so timeout 300 python - <<'PY'
import sys, os, json, argparse
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import bench
args = argparse.Namespace(gpus=1, steps=20, warmup=5, impl="ours", workload="cfg2", kernel_only=False, no_extras=True)
c = bench.Ctx()
r = bench.measure_workload(c, args, "cfg2", full=True)
s = r["roofline"]["sustained"]
print("burst %.3f sustained %.3f at %.0f MHz %.0f W e2e %.1f" % (r["roofline"]["frac"], s["frac"], s["clocks"]["sm_mhz"], s["clocks"]["power_w_max"], r["e2e"]["value"]))
PY
done
