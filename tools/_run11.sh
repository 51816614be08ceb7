tools/gpu_ab.sh "coal10 coal12 coal14 coal10 coal12" "cfg3 cfg4"
tools/gpu_ab.sh "coal10 gen10 coal10 gen10" "cfg2"
