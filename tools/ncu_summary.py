"""Key metrics + top stall locations of one kernel from an .ncu-rep. usage: ncu_summary.py rep.ncu-rep 'title' [out.txt]"""
import csv
import io
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = r[0], r[1], r[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum", "sm__inst_executed_pipe_tensor.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "launch__registers_per_thread", "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__cluster_size", "smsp__inst_executed.sum"]
out = [f"# {title}", f"# source: {rep} (ncu --set full --clock-control none)"]
for h, u, v in zip(hdr, units, vals):
    if any(h.endswith(w) for w in want) and v:
        out.append(f"{h:92s} {v:>18s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, x in enumerate(rows) if "# Samples" in x)
h2 = rows[hi]
ix = {h: i for i, h in enumerate(h2)}
data = [x for x in rows[hi + 1:] if len(x) == len(h2) and x[ix["# Samples"]].strip().isdigit()]
tot = sum(int(x[ix["# Samples"]]) for x in data) or 1
out.append(f"# top stall-sample locations ({tot} samples)")
for x in sorted(data, key=lambda x: -int(x[ix["# Samples"]]))[:14]:
    n = int(x[ix["# Samples"]])
    out.append(f"{n:7d} {100 * n / tot:5.1f}%  {x[ix['Source']].strip()[:100]}")
text = "\n".join(out)
print(text)
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(text + "\n")
