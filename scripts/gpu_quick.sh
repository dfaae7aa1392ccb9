# quick GPU cycle: parity tests + short bench (no cpu/e2e legs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > gpurun_out/bq.json 2> gpurun_out/bq.err; tail -3 gpurun_out/bq.err; python -c "
import json; d=json.loads(open('gpurun_out/bq.json').read()); r=d['roofline']; print('value',d['value'],'ms/step',d['ms_per_step'],'solve ms',r['kernel_ms_per_launch'],'setup ms',r['setup_kernel_ms_per_launch'],'frac',r['frac'],'warps',r['resident_problems_per_sm'], d.get('parity'))"
