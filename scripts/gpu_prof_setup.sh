mkdir -p gpurun_out
NAME=${1:-cur}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qp_setup -s 1 -c 1 -f -o gpurun_out/prof_setup_$NAME python bench.py --problems 20000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_setup_$NAME.log 2>&1; tail -2 gpurun_out/ncu_setup_$NAME.log
