# final evidence: bench line, ncu launch list of the same command (our kernels only), full captures of both kernels
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'qp_setup|ldp_solve|ldp_update|max_soft' -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/b_ncu.log 2>&1; tail -4 gpurun_out/launches.csv
bash scripts/gpu_prof.sh ${1:-final}
bash scripts/gpu_prof_setup.sh ${1:-final}
