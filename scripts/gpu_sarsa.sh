#!/bin/bash
# the other TD rules (general kernel k_step): us per iteration, in-tree library vs another build.  usage: bash scripts/gpu_sarsa.sh [other.so]
OTHER=${1:-}
for spec in 'fp32 {"algorithm":"sarsa"}' 'fp64 {"algorithm":"sarsa"}' 'fp32 {"algorithm":"double_qlearning"}'; do
  set -- $spec
  echo "== $1 $2"
  python scripts/prof_fp64.py 4096 12 $1 "$2" 2>&1 | tail -1 | cut -c1-110
  if [ -n "$OTHER" ]; then SPGG_B200_LIB=$OTHER python scripts/prof_fp64.py 4096 12 $1 "$2" 2>&1 | tail -1 | cut -c1-110; fi
done
