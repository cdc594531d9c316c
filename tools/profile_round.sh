#!/bin/bash
# ncu captures of one round (run under gpurun, one GPU).  usage: tools/profile_round.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
# (1) every launch of a training step: duration + DRAM bytes
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${tag}_launches_train_4096x9.csv python bench.py --steps 2 --warmup 1 --trials 1 --no-cpu-baseline --no-ref-eager > /dev/null 2> gpurun_out/${tag}_launches.err
# (2) which tensor-pipe counters exist on this part
ncu --query-metrics 2>/dev/null | grep -i -E "tensor|tmem|utc" > gpurun_out/${tag}_tensor_metric_names.txt
# (3) the plain edge-level GEMM: every tensor-related counter + the full set
ncu --metrics regex:sm__.*tensor.*,regex:smsp__.*tensor.*,sm__cycles_elapsed.avg,sm__cycles_active.avg,gpu__time_duration.sum --clock-control none \
    -k regex:gemm_tc -c 3 --csv --log-file gpurun_out/${tag}_gemm_plain_tensor_counters.csv python tools/gemm_one.py 155648 512 512 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 2 -o gpurun_out/${tag}_gemm_plain python tools/gemm_one.py 155648 512 512 > /dev/null 2>&1
# (4) the bandwidth kernels of one step (full set)
ncu --set full --clock-control none -k regex:"segment_sum|attention_series|edge_init|adam|pack_weights" -c 14 -o gpurun_out/${tag}_aux \
    python bench.py --steps 1 --warmup 1 --trials 1 --no-cpu-baseline --no-ref-eager > /dev/null 2>&1
# (5) the grouped weight-gradient launch of a layer backward (full set)
ncu --set full --clock-control none --import-source on -k regex:tn_pair -c 2 -o gpurun_out/${tag}_tn_group \
    python bench.py --steps 1 --warmup 1 --trials 1 --no-cpu-baseline --no-ref-eager > /dev/null 2>&1
