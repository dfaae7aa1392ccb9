# quick cycle + ncu capture in one call
bash scripts/gpu_quick.sh
bash scripts/gpu_prof.sh ${1:-cur}
