mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_warmstart.py -q -m gpu -x > gpurun_out/pytest_warm.log 2>&1; tail -25 gpurun_out/pytest_warm.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
