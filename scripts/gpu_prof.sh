#!/bin/bash
# one full ncu capture of the C3 solve kernel (20k problems) + of the setup kernels, summarised on the box
mkdir -p gpurun_out
NAME=${1:-r02}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ldp_solve -s 1 -c 1 -f -o gpurun_out/prof_solve_$NAME python bench.py --problems 20000 --steps 1 --warmup 1 --no-e2e --no-cpu --no-configs > gpurun_out/ncu_full_$NAME.log 2>&1
tail -2 gpurun_out/ncu_full_$NAME.log
python scripts/ncu_summary.py gpurun_out/prof_solve_$NAME.ncu-rep 40 > gpurun_out/ncu_solve_${NAME}_summary.txt 2>&1
python scripts/ncu_byfunc.py gpurun_out/prof_solve_$NAME.ncu-rep >> gpurun_out/ncu_solve_${NAME}_summary.txt 2>&1
ITERS=$(python -c "import json; print(json.loads(open('gpurun_out/ncu_full_$NAME.log').read().strip().splitlines()[-1])['roofline']['mean_iterations'])" 2>/dev/null || echo 128)
python scripts/ncu_to_json.py gpurun_out/prof_solve_$NAME.ncu-rep 20000 $ITERS gpurun_out/ncu_solve_$NAME.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"qp_(factor|product)" -s 2 -c 2 -f -o gpurun_out/prof_setup_$NAME python bench.py --problems 20000 --steps 1 --warmup 1 --no-e2e --no-cpu --no-configs > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/prof_setup_$NAME.ncu-rep 25 > gpurun_out/ncu_setup_${NAME}_summary.txt 2>&1
head -12 gpurun_out/ncu_solve_${NAME}_summary.txt; head -14 gpurun_out/ncu_setup_${NAME}_summary.txt
