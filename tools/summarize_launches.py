"""Summarises an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list: per-kernel share of one
training step (between consecutive pose_loss_kernel launches).  usage: summarize_launches.py in.csv [out.txt]"""
import collections
import csv
import json
import re
import sys

src = sys.argv[1]
with open(src) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
recs = collections.OrderedDict()
for row in csv.DictReader(lines):
    r = recs.setdefault(int(row["ID"]), {"name": row["Kernel Name"]})
    r[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    r["unit:" + row["Metric Name"]] = row["Metric Unit"]
rows = list(recs.values())
marks = [i for i, r in enumerate(rows) if "pose_loss_kernel" in r["name"]]
a, b = (marks[-2], marks[-1]) if len(marks) >= 2 else (0, len(rows))
seg = rows[a:b]


def short(n):
    return re.sub(r"\(.*", "", n).split("::")[-1][:48]


def to_ns(r):
    v, u = r.get("gpu__time_duration.sum", 0.0), r.get("unit:gpu__time_duration.sum", "ns")
    return v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)


def to_bytes(r, k):
    v, u = r.get(k, 0.0), r.get("unit:" + k, "byte")
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


tot = sum(to_ns(r) for r in seg)
agg = collections.OrderedDict()
for r in seg:
    k = short(r["name"])
    c = agg.setdefault(k, [0, 0.0, 0.0, 0.0, 0.0])
    c[0] += 1
    c[1] += to_ns(r)
    c[2] += to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")
    # tensor pipe: time-weighted active percentage and the executed bf16 FLOPs the hardware counted
    c[3] += to_ns(r) * r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
    c[4] += r.get("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum", 0.0)
out = [f"# one training step (train_4096x9, bf16): {len(seg)} launches, {tot / 1e6:.3f} ms summed "
       f"(ncu: serialised, cold cache -- compare SHARES)", f"{'ms':>8} {'share':>6} {'n':>4} {'DRAM MB':>9}  kernel"]
have_tc = any(v[4] for v in agg.values())
if have_tc:
    out[-1] = f"{'ms':>8} {'share':>6} {'n':>4} {'DRAM MB':>9} {'tensor-pipe %':>13} {'TFLOP/s':>8}  kernel"
for k, (c, ns, by, tc, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    extra = f" {tc / ns:13.1f} {fl / ns / 1e3:8.0f}" if have_tc else ""
    out.append(f"{ns / 1e6:8.3f} {100 * ns / tot:5.1f}% {c:4d} {by / 1e6:9.1f}{extra}  {k}")
gem = [v for k, v in agg.items() if k.startswith("gemm_tc_kernel") or k.startswith("tn_pair_group_kernel")]
if gem:
    n = sum(v[0] for v in gem)
    out.append(f"# tcgen05 GEMM kernels (gemm_tc_kernel<...> + tn_pair_group_kernel): {n} launches, share {100 * sum(v[1] for v in gem) / tot:.1f}%, "
               f"DRAM traffic per launch {sum(v[2] for v in gem) / n / 1e6:.1f} MB")
    summary = {"kernel": "gemm_tc_kernel + tn_pair_group_kernel", "launches_per_step": n, "share_of_step": sum(v[1] for v in gem) / tot,
               "dram_bytes_per_launch": sum(v[2] for v in gem) / n}
    if have_tc:
        gns = sum(v[1] for v in gem)
        summary["tensor_pipe_active_pct_gemm_time"] = sum(v[3] for v in gem) / gns      # hardware counter, time-weighted
        summary["tensor_pipe_active_pct_step_time"] = sum(v[3] for v in agg.values()) / tot
        summary["hw_counted_tflops_gemm_time"] = sum(v[4] for v in gem) / gns / 1e3
        out.append(f"# tensor pipe (sm__pipe_tensor_cycles_active, % of peak at the clock of the run): "
                   f"{summary['tensor_pipe_active_pct_gemm_time']:.1f} % over the GEMM time, "
                   f"{summary['tensor_pipe_active_pct_step_time']:.1f} % over the step; hardware-counted "
                   f"{summary['hw_counted_tflops_gemm_time']:.0f} TFLOP/s over the GEMM time")
    out.append("# json: " + json.dumps(summary))
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
