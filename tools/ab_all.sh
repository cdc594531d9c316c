#!/bin/bash
# cumulative A/B: this round's GPU-side changes all off vs all on (default), alternating, 100-step trials
for i in 1 2 3; do
  for mode in on off; do
    if [ $mode = off ]; then E="RPG_TN_GROUP=0 RPG_GEMM_L2PF=0 RPG_MERGED_UPDATE_DGRAD=0"; else E=""; fi
    env $E python bench.py --steps 100 --trials 3 --no-ref-eager --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$mode', round(d['ms_per_step'],3), [round(t,3) for t in d['trials_ms_per_step']], 'e2e', round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'])"
  done
done
