#!/bin/bash
TAG=$1
python -c "
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value %.3e e2e %.3e ms/step %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']))
r=d['roofline']; print('k_step %.1f us frac %.3f  gmax %.1f us  whole %.3f'%(r['kernel_us'],r['frac'],r['gmax_kernel_us'],r['whole_step_frac']))
"
python scripts/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep 2>/dev/null | head -22 | grep -E "duration|inst_executed.sum|issue_active|warps_active|registers"
python scripts/ncu_sass.py gpurun_out/prof_$TAG.ncu-rep "neighbor-aware-reinforcement-learning-fosters-cooperation-in-spatial-public-goods-games-_b200/libspgg_b200.so" k_step_fastILi1ELb0ELb1ELb1 > /tmp/sass_$TAG.txt
awk '{print $2}' /tmp/sass_$TAG.txt | sort -n | uniq -c | sort -k2 -n | tail -10
awk '{s[$4]+=$3; t+=$3} END {for (k in s) if (s[k]>t*0.02) printf "%6.1f%% %s\n", 100*s[k]/t, k}' /tmp/sass_$TAG.txt | sort -rn | head -8
