# bench every build_variants/*.so (solve + setup kernel time per 100k C3 batch)
mkdir -p gpurun_out
for so in build_variants/*.so; do
  n=$(basename $so .so)
  DAQP_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err || tail -3 gpurun_out/var_$n.err
  python -c "
import json; d=json.loads(open('gpurun_out/var_$n.json').read()); r=d['roofline']; print('$n', 'solve ms %.2f' % r['kernel_ms_per_launch'], 'setup ms %.2f' % r['setup_kernel_ms_per_launch'], 'frac %.4f' % r['frac'])" || true
done
