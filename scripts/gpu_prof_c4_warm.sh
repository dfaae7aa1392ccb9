#!/bin/bash
# one full ncu capture of the team solve kernel on a WARM-started C4 batch (3 waves of problems), summarised on the box
mkdir -p gpurun_out
NAME=${1:-c4_warm}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ldp_solve -s 3 -c 1 -f -o gpurun_out/prof_$NAME python scripts/bench_c4.py --n 1332 --reps 1 > gpurun_out/ncu_$NAME.log 2>&1
tail -2 gpurun_out/ncu_$NAME.log
python scripts/ncu_summary.py gpurun_out/prof_$NAME.ncu-rep 45 > gpurun_out/ncu_${NAME}_summary.txt 2>&1
python scripts/ncu_byfunc.py gpurun_out/prof_$NAME.ncu-rep >> gpurun_out/ncu_${NAME}_summary.txt 2>&1
cat gpurun_out/ncu_${NAME}_summary.txt
