#!/bin/bash
# one full ncu capture of the factor kernel alone: C3 (a warp per problem) and C4 (a team per problem)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qp_factor" -s 1 -c 1 -f -o gpurun_out/prof_factor_c3 python bench.py --problems 20000 --steps 1 --warmup 1 --no-e2e --no-cpu --no-configs --no-workspace --no-weak > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/prof_factor_c3.ncu-rep 30 > gpurun_out/ncu_factor_c3_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qp_factor" -s 1 -c 1 -f -o gpurun_out/prof_factor_c4 python scripts/bench_c4.py --n 2664 --reps 1 > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/prof_factor_c4.ncu-rep 30 > gpurun_out/ncu_factor_c4_summary.txt 2>&1
cat gpurun_out/ncu_factor_c3_summary.txt gpurun_out/ncu_factor_c4_summary.txt
