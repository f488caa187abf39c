#!/bin/bash
# bench every library build under build/lib_*.so (kernel tuning experiments)
mkdir -p gpurun_out
for lib in build/lib_*.so; do
  tag=$(basename $lib .so)
  SPGG_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/var_$tag.log 2> gpurun_out/var_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/var_$tag.log').read().strip().splitlines()[-1]); r=d['roofline']
    print('$tag value %.3e k_step %.1f us frac %.3f gmax %.1f us whole %.3f'%(d['value'],r['kernel_us'],r['frac'],r['gmax_kernel_us'],r['whole_step_frac']))
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/var_$tag.err').read()[-500:])
PY
done
