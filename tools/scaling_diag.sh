#!/bin/bash
# usage: tools/scaling_diag.sh <N> <tag> -- where does the N-GPU step lose time against one GPU of the same box?
N=$1; tag=${2:-diag}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload train_4096x9 --trials 3 --no-ref-eager "$@"; }
python bench.py --workload train_4096x9 --trials 3 --no-ref-eager --no-cpu-baseline > gpurun_out/${tag}_n1_samebox.json 2> gpurun_out/${tag}_n1_samebox.err
run > gpurun_out/${tag}_n${N}_overlap.json 2> gpurun_out/${tag}_n${N}_overlap.err
run --no-overlap-allreduce > gpurun_out/${tag}_n${N}_nooverlap.json 2> gpurun_out/${tag}_n${N}_nooverlap.err
RPG_BENCH_SKIP_ALLREDUCE=1 run > gpurun_out/${tag}_n${N}_noreduce.json 2> gpurun_out/${tag}_n${N}_noreduce.err
NCCL_DEBUG=INFO run --steps 5 --trials 1 > /dev/null 2> gpurun_out/${tag}_n${N}_nccl_info.log
for f in gpurun_out/${tag}_n*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], d["ms_per_step"], d["trials_ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
grep -i "nvls\|channels\|Connected" gpurun_out/${tag}_n${N}_nccl_info.log | head -12
