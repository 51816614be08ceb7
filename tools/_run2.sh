tools/gpu_ab.sh "p0 p0nh p1k p2k p2knh p4k p8k p2kc10 u2c12 u2c16 t64c16" "cfg2 cfg3 cfg4"
