#!/bin/bash
# full GPU suite + quick C3 bench + C4 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3_quick.json 2> gpurun_out/bench_c3_quick.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_c3_quick.json'))
    r = d['roofline']
    print("C3:", round(d['value']), "QP/s; solve ms", round(r['kernel_ms_per_launch'], 2), "setup ms", round(r['setup_kernel_ms_per_launch'], 2))
except Exception as e:
    print("bench failed", e); print(open('gpurun_out/bench_c3_quick.err').read()[-2000:])
PY
timeout 600 python scripts/bench_c4.py --n 50000 > gpurun_out/c4_team.json 2> gpurun_out/c4_team.err; cat gpurun_out/c4_team.json; tail -3 gpurun_out/c4_team.err
