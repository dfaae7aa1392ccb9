# two-GPU evidence: the NCCL sharding test and the weak-scaling bench at N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "sharded" > gpurun_out/pytest_gpu2.log 2>&1; tail -3 gpurun_out/pytest_gpu2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-workspace > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; tail -2 gpurun_out/bench_g2.err; cat gpurun_out/bench_g2.json
