#!/bin/bash
# usage: tools/ab_env.sh VAR [workload] -- the bench step with VAR=1 and VAR=0, alternating, 2 rounds x 5 trials
VAR=$1; W=${2:-train_4096x9}
for i in 1 2; do for v in 1 0; do
  env $VAR=$v python bench.py --workload $W --trials 5 --no-ref-eager --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$VAR=$v', '$W', round(d['ms_per_step'],3), [round(t,3) for t in d['trials_ms_per_step']], 'gemm_ms', round(r['gemm_ms_per_step'],3), 'launches', r['launches_per_step'], 'exec_frac', round(r['executed_frac'],3), 'gpu_launches', d['gpu_launches'])"
done; done
