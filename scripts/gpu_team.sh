#!/bin/bash
# first GPU pass of the team kernel: parity tests, then C4 timing with and without the team mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "team or c4 or shapes_vs_oracle or odd_batch" 2>&1 | tail -40 > gpurun_out/team_tests.log
timeout 600 python scripts/bench_c4.py --n 50000 > gpurun_out/c4_team.json 2> gpurun_out/c4_team.err
DAQP_B200_TEAM=0 timeout 600 python scripts/bench_c4.py --n 10000 --reps 1 > gpurun_out/c4_warp.json 2> gpurun_out/c4_warp.err
tail -5 gpurun_out/team_tests.log; cat gpurun_out/c4_team.json gpurun_out/c4_warp.json; tail -3 gpurun_out/c4_team.err
