"""One line per profiled launch of an .ncu-rep (ncu --set full): duration, DRAM bytes, achieved DRAM GB/s, registers,
occupancy and issue activity.  usage: ncu_table.py report.ncu-rep [title]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h, units = rows[hdr], rows[hdr + 1]
col = {n: i for i, n in enumerate(h)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("smsp__issue_active.avg.pct", "issue %"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %")]


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)


print(f"# {title}\n# source: {rep} (ncu --set full --clock-control none; per launch, cold cache)")
print(f"{'kernel':44s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s} {'regs':>5s} {'warps%':>7s} {'issue%':>7s} {'tensor%':>8s} {'L2hit%':>7s}")
for r in rows[hdr + 2:]:
    if len(r) < len(h):
        continue
    name = r[col["Kernel Name"]].replace("void ", "").replace("rpg::", "")[:44]
    us = to_us(r[col[want[0][0]]], units[col[want[0][0]]])
    rd = to_bytes(r[col[want[1][0]]], units[col[want[1][0]]])
    wr = to_bytes(r[col[want[2][0]]], units[col[want[2][0]]])
    rest = []
    for m, _ in want[3:]:
        rest.append(float(r[col[m]].replace(",", "")) if m in col and r[col[m]] not in ("", "n/a") else float("nan"))
    print(f"{name:44s} {us:8.1f} {rd / 1e6:8.1f} {wr / 1e6:8.1f} {(rd + wr) / us / 1e3:7.0f} {rest[0]:5.0f} {rest[1]:7.1f} {rest[2]:7.1f} {rest[3]:8.1f} {rest[4]:7.1f}")
