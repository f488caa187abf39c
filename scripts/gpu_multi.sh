#!/bin/bash
# multi-GPU visit: NCCL strip parity tests + strip bench.  usage: bash scripts/gpu_multi.sh <ngpu> [L]
N=${1:-2}; LL=${2:-32768}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_${N}.txt 2>&1
echo "== pytest strips"; timeout 900 python -m pytest tests/test_gpu_strips.py -m gpu -q -x --timeout 800 > gpurun_out/pytest_strips_${N}.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_strips_${N}.log
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --inner 20 --L $LL > gpurun_out/bench_${N}.log 2> gpurun_out/bench_${N}.err; echo "rc=$?"; tail -2 gpurun_out/bench_${N}.log; tail -5 gpurun_out/bench_${N}.err
