"""GPU probe for the tcgen05 GEMM: runs NT / TN cases from simplest to hardest against torch fp32 matmul and
prints one line per case (flushes as it goes, so a trap or hang shows where it broke).
usage: python tools/gemm_probe.py [nt|tn|all]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relpose_gnn_b200 import ops  # noqa: E402
from relpose_gnn_b200.graph import GraphBatch  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16


def rnd(*shape, scale=1.0, seed=[0]):
    seed[0] += 1
    g = torch.Generator(device="cpu").manual_seed(seed[0])
    return (torch.randn(*shape, generator=g) * scale).to(dev).to(BF)


def report(name, got, ref, tol):
    torch.cuda.synchronize()
    err = (got.float() - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-9
    ok = err / scale < tol and bool(torch.isfinite(got.float()).all())
    print(f"{'PASS' if ok else 'FAIL'} {name}: max_abs_err={err:.4g} ref_max={scale:.4g} rel={err / scale:.3g}", flush=True)
    return ok


def nt_cases():
    ok = True
    for (M, N, K, bn) in [(128, 256, 64, 0), (128, 256, 256, 0), (256, 512, 512, 0), (384, 64, 128, 0),
                          (300, 192, 128, 0), (128, 48, 64, 0), (200, 16, 128, 0), (1000, 1536, 512, 0),
                          (128, 128, 64, 128), (36864, 512, 512, 0)]:
        A, B = rnd(M, K), rnd(N, K, scale=K ** -0.5)
        out = torch.empty(M, N, dtype=BF, device=dev)
        print(f"... NT M={M} N={N} K={K} block_n={bn}", flush=True)
        ops.gemm_nt(A, B, out=out, block_n=bn)
        ok &= report(f"NT plain M={M} N={N} K={K}", out, A.float() @ B.float().t(), 1e-2)
    # segments + full epilogue
    M, N = 72 * 5, 256
    g = GraphBatch.fully_connected(5, 9, dev)
    A0, A1, A2 = rnd(M, 128), rnd(M, 64), rnd(M, 192)
    B = rnd(N, 384, scale=0.05)
    bias = torch.randn(N, device=dev)
    P = rnd(45, 2 * N)
    resid, mask = rnd(M, N), rnd(M, N)
    rs = torch.rand(9, device=dev) + 0.5
    ref = torch.cat([A0, A1, A2], 1).float() @ B.float().t() + bias
    ei = g.edge_index()
    ref = ref + P[:, :N].float()[ei[0]] + P[:, N:].float()[ei[1]] + resid.float()
    ref = ref * rs[torch.arange(M, device=dev) % 9].unsqueeze(1)
    ref = ref * (mask.float() > 0)
    out, outr = torch.empty(M, N, dtype=BF, device=dev), torch.empty(M, N, dtype=BF, device=dev)
    o32 = torch.empty(M, N, dtype=torch.float32, device=dev)
    print("... NT fused epilogue", flush=True)
    ops.gemm_nt(None, B, segs=[A0, A1, A2], bias=bias, gadd=[(P[:, :N], "src"), (P[:, N:], "dst")], graph=g,
                resid=resid, row_scale=rs, mask=mask, out=out, out_relu=outr, out_f32=o32)
    ok &= report("NT fused out", out, ref, 1e-2)
    ok &= report("NT fused out_relu", outr, ref.clamp(min=0), 1e-2)
    ok &= report("NT fused out_f32", o32, ref, 1e-4)
    out2 = torch.empty(M, N, dtype=BF, device=dev)
    ops.gemm_nt(None, B, segs=[A0, A1, A2], bias=bias, relu=True, out=out2)
    ok &= report("NT relu", out2, (torch.cat([A0, A1, A2], 1).float() @ B.float().t() + bias).clamp(min=0), 1e-2)
    # residual only (pair kernels fetch it by TMA): ragged M, N below one block, and a multi-block shape
    for (Mr, Nr, Kr) in [(300, 192, 128), (1000, 512, 256), (36864, 512, 512)]:
        Ar, Br, Rr = rnd(Mr, Kr), rnd(Nr, Kr, scale=Kr ** -0.5), rnd(Mr, Nr)
        o_r = torch.empty(Mr, Nr, dtype=BF, device=dev)
        ops.gemm_nt(Ar, Br, resid=Rr, out=o_r)
        ok &= report(f"NT resid M={Mr} N={Nr} K={Kr}", o_r, Ar.float() @ Br.float().t() + Rr.float(), 1e-2)
    # gathered adds as one-hot K panels == the epilogue gather path
    for (Gn, Nn, keepmask) in [(5, 9, None), (40, 9, np.array([1, 0, 1, 1, 0, 0, 1, 0, 1, 0] * 3 + [1] * 6, bool)), (3, 17, None)]:
        gg = GraphBatch.fully_connected(Gn, Nn, dev, keepmask)
        Mg = gg.n_edge_rows
        Ag = rnd(Mg, 128)
        Bg = rnd(256, 128, scale=0.1)
        Pg = rnd(gg.n_node_rows, 512)
        refg = Ag.float() @ Bg.float().t() + bias
        eig = gg.edge_index()
        refg = (refg + Pg[:, :256].float()[eig[0]] + Pg[:, 256:].float()[eig[1]]).clamp(min=0)
        o1 = torch.empty(Mg, 256, dtype=BF, device=dev)
        ops.gemm_nt(Ag, Bg, bias=bias, gpanel=[(Pg[:, :256], "src"), (Pg[:, 256:], "dst")], graph=gg, relu=True, out=o1)
        ok &= report(f"NT one-hot gather panels G={Gn} N={Nn} Ep={gg.Ep}", o1, refg, 1e-2)
    # bit patterns: out_bits == packbits(out > 0); mask_bits has the same effect as the bf16 mask
    M2, N2 = 72 * 5, 256
    bits = torch.zeros(M2, N2 // 8, dtype=torch.uint8, device=dev)
    out3 = torch.empty(M2, N2, dtype=BF, device=dev)
    ops.gemm_nt(None, B, segs=[A0, A1, A2], bias=bias, relu=True, out=out3, out_bits=bits)
    want = (out3.float() > 0).view(M2, N2 // 8, 8).to(torch.uint8)
    want = (want << torch.arange(8, device=dev, dtype=torch.uint8)).sum(-1).to(torch.uint8)
    okb = bool(torch.equal(bits, want))
    print(("PASS" if okb else "FAIL") + " NT out_bits", flush=True)
    ok &= okb
    o_m = torch.empty(M2, N2, dtype=BF, device=dev)
    o_b = torch.empty(M2, N2, dtype=BF, device=dev)
    ops.gemm_nt(None, B, segs=[A0, A1, A2], resid=resid, mask=out3, out=o_m)
    ops.gemm_nt(None, B, segs=[A0, A1, A2], resid=resid, mask_bits=bits, out=o_b)
    okb = bool(torch.equal(o_m, o_b))
    print(("PASS" if okb else "FAIL") + " NT mask_bits == bf16 mask", flush=True)
    ok &= okb
    # timing of the big shape
    M, N, K = 294912, 512, 512
    A, B = rnd(M, K), rnd(N, K, scale=K ** -0.5)
    out = torch.empty(M, N, dtype=BF, device=dev)
    for _ in range(3):
        ops.gemm_nt(A, B, out=out)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10):
        ops.gemm_nt(A, B, out=out)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 10
    print(f"INFO NT {M}x{N}x{K}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s  "
          f"{(M * K + M * N) * 2 / ms / 1e6:.0f} GB/s", flush=True)
    ref = A[:4096].float() @ B.float().t()
    ok &= report("NT big (first 4096 rows)", out[:4096], ref, 1e-2)
    t0 = time.time()
    for _ in range(10):
        torch.matmul(A, B.t())
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(10):
        torch.matmul(A, B.t())
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 10
    print(f"INFO cuBLAS same shape: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    return ok


def tn_cases():
    ok = True
    for (R, M, N, splits) in [(64, 128, 64, 1), (64, 128, 256, 1), (256, 128, 256, 1), (1000, 192, 256, 3),
                              (5000, 512, 512, 7), (777, 48, 128, 2), (4096, 1536, 512, 4), (300, 512, 16, 2),
                              (640, 128, 64, 7), (36864, 512, 64, 74)]:   # last two: splits with no k-blocks at all
        A, B = rnd(R, M), rnd(R, N)
        ws = torch.empty(splits, M, N, dtype=torch.float32, device=dev)
        print(f"... TN R={R} M={M} N={N} splits={splits}", flush=True)
        ops.gemm_tn_partials(A, B, ws, M, N, splits)
        ok &= report(f"TN R={R} M={M} N={N} splits={splits}", ws.sum(0), A.float().t() @ B.float(), 2e-3)
    # wgrad wrapper with accumulation into a strided output
    R, M, N = 294912, 512, 512
    A, B = rnd(R, M, scale=0.1), rnd(R, N, scale=0.1)
    out = torch.ones(M, 2 * N, dtype=torch.float32, device=dev)
    ws = ops.wgrad_ws(512, dev)
    ops.wgrad(A, B, out[:, N:], ws)
    ref = A.float().t() @ B.float() + 1
    ok &= report("wgrad accumulate strided", out[:, N:], ref, 2e-3)
    for (R2, M2) in [(5000, 512), (777, 192), (36864, 48)]:
        A2, B2 = rnd(R2, M2), rnd(R2, 256)
        o2 = torch.zeros(M2, 256, dtype=torch.float32, device=dev)
        bias2 = torch.ones(M2, dtype=torch.float32, device=dev)
        ops.wgrad(A2, B2, o2, ws, bias=bias2)
        ok &= report(f"wgrad fused column sums R={R2} M={M2}", bias2, A2.float().sum(0) + 1, 1e-3)
        ok &= report(f"wgrad with column sums: dW R={R2} M={M2}", o2, A2.float().t() @ B2.float(), 2e-3)
    ok &= report("wgrad untouched half", out[:, :N], torch.ones(M, N, device=dev), 1e-6)
    for _ in range(2):
        ops.wgrad(A, B, out[:, N:], ws)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10):
        ops.wgrad(A, B, out[:, N:], ws)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 10
    print(f"INFO wgrad {R}x{M}x{N}: {ms:.3f} ms  {2 * R * M * N / ms / 1e9:.1f} TFLOP/s", flush=True)
    return ok


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), flush=True)
    ok = True
    if which in ("nt", "all"):
        ok &= nt_cases()
    if which in ("tn", "all"):
        ok &= tn_cases()
    print("ALL PASS" if ok else "SOME FAILED", flush=True)
    sys.exit(0 if ok else 1)
