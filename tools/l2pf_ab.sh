#!/bin/bash
# A/B of the producer's L2 prefetch: wait-cycle traces of the dominant shapes, then the training step both ways.
for pf in 1 0; do
  echo "=== RPG_GEMM_L2PF=$pf"
  for args in "155648 512 512 plain" "155648 512 512 dual" "155648 192 512 plain" "36864 512 512 plain"; do
    RPG_GEMM_L2PF=$pf python tools/gemm_trace.py $args | grep -v "total\|staging"
  done
  RPG_GEMM_L2PF=$pf RPG_GEMM_WS=0 python tools/gemm_trace.py 155648 512 512 plain | grep -v "total\|staging"
done
for i in 1 2; do for pf in 1 0; do
  RPG_GEMM_L2PF=$pf python bench.py --trials 5 --no-ref-eager --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('l2pf=$pf', round(d['ms_per_step'],3), [round(t,3) for t in d['trials_ms_per_step']], 'gemm_ms', round(d['roofline']['gemm_ms_per_step'],3), 'exec_frac', round(d['roofline']['executed_frac'],3))"
done; done
