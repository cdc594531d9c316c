"""Runs one GEMM shape a few times (for ncu captures). usage: gemm_one.py M N K [fused|resid|dual]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relpose_gnn_b200 import ops
from relpose_gnn_b200.graph import GraphBatch
M, N, K = [int(v) for v in sys.argv[1:4]]
variant = sys.argv[4] if len(sys.argv) > 4 else ""
fused = variant == "fused"
dev = torch.device("cuda:0")
A = torch.randn(M, K, device=dev).bfloat16(); B = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
kw = {}
if fused:
    g = GraphBatch.fully_connected(M // 72, 9, dev)
    P = torch.randn(g.n_node_rows, 2 * N, device=dev).bfloat16()
    kw = dict(bias=torch.randn(N, device=dev), gadd=[(P[:, :N], "src"), (P[:, N:], "dst")], graph=g, relu=True)
if variant == "resid":
    kw = dict(resid=torch.randn(M, N, device=dev).bfloat16())
if variant == "dual":
    kw = dict(bias=torch.randn(N, device=dev), out_relu=torch.empty(M, N, dtype=torch.bfloat16, device=dev))
for _ in range(5):
    ops.gemm_nt(A, B, out=out, **kw)
torch.cuda.synchronize()
print("done")
