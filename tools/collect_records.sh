#!/bin/bash
# Runs every BASELINE workload on ONE GPU and stores the JSON lines under gpurun_out/ (copied into profiles/ afterwards).
# usage: tools/collect_records.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
for w in train_4096x9 train_2048x17 infer_4096x9 infer_fp32_4096x9 infer_8192x9 infer_65536x9_strong train_knn4_4096x9 train_fp32_4096x9; do
  python bench.py --workload $w --trials 5 --no-ref-eager --no-cpu-baseline > gpurun_out/${tag}_n1_${w}.json 2> gpurun_out/${tag}_n1_${w}.err || echo "FAILED $w"
done
