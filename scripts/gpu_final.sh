#!/bin/bash
# Round-end style visit: smoke, all GPU tests, headline bench, small-lattice benches, ncu evidence.
TAG=${1:-r01s}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench c4"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -5 gpurun_out/bench.err
for w in c1 c2 sweep; do
  echo "== bench $w"; timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --inner 1000 > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "rc=$?"; tail -2 gpurun_out/bench_$w.err
  SPGG_NO_RESIDENT=1 timeout 600 python bench.py --workload $w --steps 3 --warmup 2 --inner 200 > gpurun_out/bench_${w}_periter.log 2> gpurun_out/bench_${w}_periter.err
done
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.log 2>&1; echo "rc=$?"
echo "== ncu k_step"; bash scripts/gpu_profile.sh $TAG > gpurun_out/profile_$TAG.log 2>&1; tail -3 gpurun_out/profile_$TAG.log
echo "== ncu resident"
RCMD="python bench.py --workload sweep --steps 1 --warmup 1 --inner 300"
$RCMD > gpurun_out/plain_res_${TAG}.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_res_${TAG}.csv $RCMD > gpurun_out/ncu_list_res_${TAG}.log 2>&1; echo "list rc=$?"
$RCMD > gpurun_out/plain2_res_${TAG}.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_resident -s 2 -c 1 -o gpurun_out/prof_res_${TAG} -f $RCMD > gpurun_out/ncu_full_res_${TAG}.log 2>&1; echo "full rc=$?"
ls gpurun_out | wc -l
