# GPU cycle for the shared workspace: its tests, the whole GPU suite, then the full bench line (with the shared leg)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "shared_workspace or workspace" > gpurun_out/pytest_shared.log 2>&1; tail -25 gpurun_out/pytest_shared.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu > gpurun_out/bench_shared.json 2> gpurun_out/bench_shared.err; tail -3 gpurun_out/bench_shared.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_shared.json').read()); r=d['roofline']; print('value',d['value'],'e2e',d['e2e']['value'],'solve ms',r['kernel_ms_per_launch'],'frac',r['frac']); print('workspace',d.get('workspace')); print('shared',d.get('shared_workspace'))"
