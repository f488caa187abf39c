#!/bin/bash
# batch policy experiments of the resident kernel: cluster size x threads per CTA
mkdir -p gpurun_out
for cfg in "8 512" "8 256" "16 512" "16 256" "16 384" "16 128"; do
  set -- $cfg
  if [ "$1" = "16" ]; then export SPGG_RES_CS16=1; else unset SPGG_RES_CS16; fi
  export SPGG_RES_THREADS=$2
  echo "== CS=$1 threads=$2"; timeout 300 python scripts/res_bench.py c1_L200x60 c1_L100x10 c1_L200x18 2>&1 | grep "resident"
done
