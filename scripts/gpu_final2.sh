# round-1 final evidence (second pass): tests, smoke, bench line, reference arm, ncu launch list, side-kernel numbers
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 300 python -m pytest tests/test_minrep.py -q -m gpu -k throughput -s > gpurun_out/minrep.log 2>&1; tail -4 gpurun_out/minrep.log
timeout 300 python scripts/bench_warmstart.py --out gpurun_out/warmstart.json 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'qp_setup|ldp_solve|ldp_update|max_soft|minrep|init_active' -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/b_ncu.log 2>&1; tail -4 gpurun_out/launches.csv
