#!/bin/bash
# Bench every build/libspgg_<tag>.so given on the command line (plus the in-tree library as "base"):
# us per k_step launch, us per k_gmax launch, us per iteration.  Experiment builds only
# (scripts/build_variant.sh); what-if variants (-DSPGG_X_*) compute wrong results on purpose.
# usage: bash scripts/gpu_variants.sh <out-tag> <tag> [<tag> ...]
OUT=gpurun_out/variants_$1.txt; shift
mkdir -p gpurun_out; : > $OUT
run() {
  timeout 100 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-12s k_step %7.2f us  exact pair %s  iteration %7.2f us  frac %.3f whole %.3f spec %s' % ('$1', r['kernel_us'], r.get('exact_pair_us'), 10*d['ms_per_step'], r['frac'], r['whole_step_frac'], r.get('speculation')))" >> $OUT 2>&1 || echo "$1 FAILED" >> $OUT
}
[ -z "$NOBASE" ] && run base
for t in "$@"; do SPGG_B200_LIB=$PWD/build/libspgg_$t.so run $t; done
cat $OUT
