"""Who waits for whom inside gemm_tc_kernel: runs one GEMM shape on the -DRPG_GEMM_TRACE build of the library
(tools/build_trace_lib.sh) and prints, averaged over the CTAs, the share of its lifetime each role spends waiting.
usage: gemm_trace.py M N K [plain|dual|fused|resid]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import relpose_gnn_b200._lib as L   # noqa: E402

L.LIB_PATH = os.path.join(ROOT, "tools", "_trace", "librpg_b200_trace.so")
from relpose_gnn_b200 import ops   # noqa: E402
from relpose_gnn_b200.graph import GraphBatch   # noqa: E402

M, N, K = [int(v) for v in sys.argv[1:4]]
variant = sys.argv[4] if len(sys.argv) > 4 else "plain"
dev = torch.device("cuda:0")
lib = L.load()
lib.rpg_debug_gemm_trace.restype = C.c_int
lib.rpg_debug_gemm_trace.argtypes = [C.c_void_p, C.c_int, C.c_int]
A = torch.randn(M, K, device=dev).bfloat16()
B = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
kw = {}
if variant == "fused":
    g = GraphBatch.fully_connected(M // 72, 9, dev)
    P = torch.randn(g.n_node_rows, 2 * N, device=dev).bfloat16()
    kw = dict(bias=torch.randn(N, device=dev), gadd=[(P[:, :N], "src"), (P[:, N:], "dst")], graph=g, relu=True)
elif variant == "resid":
    kw = dict(resid=torch.randn(M, N, device=dev).bfloat16())
elif variant == "dual":
    kw = dict(bias=torch.randn(N, device=dev), out_relu=torch.empty(M, N, dtype=torch.bfloat16, device=dev))


def run():
    ops.gemm_nt(A, B, out=out, **kw)


for _ in range(3):
    run()
torch.cuda.synchronize()
lib.rpg_debug_gemm_trace(None, 0, 1)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
reps = 5
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / reps
buf = np.zeros((148, 8), dtype=np.uint64)
assert lib.rpg_debug_gemm_trace(buf.ctypes.data, 148, 0) == 0
t = buf.astype(np.float64) / reps
names = ["producer: wait free stage", "producer: total", "mma: wait operands", "mma: wait drained accumulator", "mma: total",
         "epilogue w0: wait accumulator", "epilogue w0: wait staging tile (TMA store)", "epilogue w0: total"]
print(f"{variant} {M}x{N}x{K}: {us:.1f} us/launch ({2.0 * M * N * K / us / 1e6:.0f} TFLOP/s incl. launch gaps)")
lead = t[t[:, 4] > 0]          # CTAs that issue MMAs (pair kernels: the leaders)
for i, n in enumerate(names):
    rows = lead if i in (2, 3, 4) else t[t[:, 7] > 0]
    tot = rows[:, 4 if i in (2, 3, 4) else (1 if i < 2 else 7)]
    print(f"  {n:46s} {rows[:, i].mean():10.0f} clk  {100.0 * rows[:, i].mean() / max(tot.mean(), 1):5.1f} % of the role's lifetime")
