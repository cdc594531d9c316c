#!/bin/bash
# usage: tools/collect_scaling.sh <N> <tag>   -- runs the multi-GPU workloads on N GPUs of one box
N=$1; tag=${2:-r2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
for w in train_4096x9 train_2048x17 train_2048x17_strong infer_8192x9 infer_65536x9_strong; do
  run --workload $w --trials 3 --no-ref-eager > gpurun_out/${tag}_n${N}_${w}.json 2> gpurun_out/${tag}_n${N}_${w}.err || echo "FAILED $w"
done
if [ "${NOOVERLAP:-0}" = "1" ]; then
for w in train_4096x9 train_2048x17_strong; do
  run --workload $w --trials 3 --no-ref-eager --no-overlap-allreduce > gpurun_out/${tag}_n${N}_${w}_nooverlap.json 2> gpurun_out/${tag}_n${N}_${w}_nooverlap.err || echo "FAILED no-overlap $w"
done
fi
