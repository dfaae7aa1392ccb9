# GPU cycle for the minrep path: its tests first, then the whole GPU suite, a quick bench (regression check), minrep throughput
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_minrep.py -q -m gpu -x > gpurun_out/pytest_minrep.log 2>&1; tail -25 gpurun_out/pytest_minrep.log
timeout 300 python -m pytest tests/test_minrep.py -q -m gpu -k throughput -s > gpurun_out/minrep.log 2>&1; tail -4 gpurun_out/minrep.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bq.json 2> gpurun_out/bq.err; tail -3 gpurun_out/bq.err; python -c "
import json; d=json.loads(open('gpurun_out/bq.json').read()); r=d['roofline']; print('value',d['value'],'ms/step',d['ms_per_step'],'solve ms',r['kernel_ms_per_launch'],'setup ms',r['setup_kernel_ms_per_launch'],'frac',r['frac'],'warps',r['resident_problems_per_sm'], d.get('parity'), d.get('workspace',{}).get('ms_per_step'))"
