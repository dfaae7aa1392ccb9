#!/bin/bash
# round-end evidence in one gpurun call: ncu of the shipped kernels, the full GPU suite, the default bench line, its launch list
mkdir -p gpurun_out
bash scripts/gpu_prof.sh r02 > gpurun_out/prof_r02.log 2>&1
cp gpurun_out/ncu_solve_r02.json profiles/ncu_solve_r02.json 2>/dev/null
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu_r02_final.log
tail -5 gpurun_out/pytest_gpu_r02_final.log
timeout 900 python bench.py > gpurun_out/bench_r02_1gpu_final.json 2> gpurun_out/bench_r02_1gpu_final.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_r02_1gpu_final.json'))
    print("C3 value", round(d['value']), "e2e", round(d['e2e']['value']), "frac", round(d['roofline']['frac'], 3), "c4", round(d['c4']['cold']['value']), round(d['c4']['warm']['value']))
except Exception as e:
    print("bench failed", e); print(open('gpurun_out/bench_r02_1gpu_final.err').read()[-1500:])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"qp_|ldp_|minrep|init_active|first_viol|bnb_|max_soft" -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 > gpurun_out/b_ncu.log 2>&1
grep -c gpu__time gpurun_out/launches_r02.csv
