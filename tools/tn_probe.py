"""One weight-gradient product C[M,N] += A[:, :M]^T B[:, :N] over R rows, a few times (for ncu / timing).
usage: tn_probe.py M N R"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relpose_gnn_b200 import ops
M, N, R = [int(v) for v in sys.argv[1:4]]
dev = torch.device("cuda:0")
A = torch.randn(R, M, device=dev).bfloat16(); B = torch.randn(R, N, device=dev).bfloat16()
out = torch.zeros(M, N, device=dev); ws = ops.wgrad_ws(max(M, N, 512), dev)
for _ in range(3):
    ops.wgrad(A, B, out, ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10):
    ops.wgrad(A, B, out, ws)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
print(f"{M}x{N} R={R}: {us:.1f} us per call (kernel + fold), {2.0 * M * N * R / us / 1e6:.0f} TFLOP/s, operands {R * (M + N) * 2 / 1e6:.0f} MB -> {R * (M + N) * 2 / us / 1e6:.2f} TB/s")
