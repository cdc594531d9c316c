"""Channel attention kernels (`-m gpu`): the series form (csrc/rpg_attention.cu) against an fp64 softmax for rows in every
series class (K = 6 / 10 / 14 / 20) and beyond the series bound (exact exp2 path), mixed inside the same warps, forward and
backward, for every width the path supports; and against the former exp2 kernels (aux given)."""
import numpy as np
import pytest
import torch

import relpose_gnn_b200 as rpg
from relpose_gnn_b200 import graph as G, ops

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def reference(gtp, dy, c):
    g, th, ph = (gtp[:, i * c:(i + 1) * c].double().clone().requires_grad_(True) for i in range(3))
    s = torch.softmax(ph.unsqueeze(2) * th.unsqueeze(1), dim=-1)          # att.py:25-26
    y = (s * g.unsqueeze(1)).sum(-1)                                       # att.py:30
    (y * dy.double()).sum().backward()
    return y.detach(), torch.cat([g.grad, th.grad, ph.grad], 1)


def make_rows(Et, c, seed):
    """Rows whose range R = max|phi| * (max theta - min theta) / 2 cycles through all classes, row by row."""
    gen = torch.Generator().manual_seed(seed)
    gtp = torch.randn(Et, 3 * c, generator=gen)
    targets = torch.tensor([0.01, 0.12, 0.5, 1.2, 2.8, 5.0, 12.0])[torch.arange(Et) % 7]
    th, ph = gtp[:, c:2 * c], gtp[:, 2 * c:]
    R = ph.abs().max(1).values * (th.max(1).values - th.min(1).values) / 2
    s = (targets / R).sqrt().unsqueeze(1)
    gtp[:, c:2 * c] = th * s + torch.randn(Et, 1, generator=gen)         # an offset on theta: the centring must absorb it
    gtp[:, 2 * c:] = ph * s
    return gtp


def rel_rows(a, b):
    return ((a.double() - b).norm(dim=1) / b.norm(dim=1).clamp(min=1e-30))


@pytest.mark.parametrize("c", [16, 64, 128, 256])
def test_series_attention_matches_fp64_softmax_in_every_class(c):
    N, Gn = 5, 31
    src, dst = G.fc_template(N)
    graph = G.GraphBatch(src, dst, Gn, N, dev())
    Et = graph.n_edge_rows                                                  # 620 rows: the last warp is partial
    gtp = make_rows(Et, c, 7 + c)
    dyn = torch.randn(Gn * N, c, generator=torch.Generator().manual_seed(c))
    ei = graph.edge_index().cpu()
    dy = dyn[ei[1]]
    y_ref, d_ref = reference(gtp, dy, c)
    cp = ops.pad64(c)
    y = torch.zeros(Et, cp, dtype=torch.bfloat16, device=dev())
    y_lo = torch.zeros_like(y)
    ops.attention_fwd(gtp.to(dev()), c, y, y_lo=y_lo)                       # (hi, lo) planes: fp32-level comparison
    got = (y.float() + y_lo.float())[:, :c].cpu()
    err = rel_rows(got, y_ref)
    assert err.max().item() < 3e-5, err.max().item()                        # hi + lo carries ~16 bits
    assert rel_rows(y.float()[:, :c].cpu(), y_ref).max().item() < 6e-3     # bf16 output
    c3p = ops.pad64(3 * c)
    dgtp = torch.zeros(Et, c3p, dtype=torch.bfloat16, device=dev())
    ops.attention_bwd(gtp.to(dev()), dyn.to(dev()), graph, c, dgtp)
    gerr = rel_rows(dgtp.float()[:, :3 * c].cpu(), d_ref)
    assert gerr.max().item() < 8e-3, gerr.max().item()                      # bf16 output of an fp32 computation
    # bf16 projections (the layer's bf16 mode): same kernels on rows rounded to bf16 -- compare with the fp64 softmax OF
    # THOSE rounded rows, so that only the kernel arithmetic is measured
    g16 = gtp.bfloat16()
    y_ref16, d_ref16 = reference(g16.float(), dy, c)
    yb = torch.zeros_like(y)
    ops.attention_fwd_bf16(g16.to(dev()), c, yb)
    assert rel_rows(yb.float()[:, :c].cpu(), y_ref16).max().item() < 6e-3
    db = torch.zeros_like(dgtp)
    ops.attention_bwd_bf16(g16.to(dev()), dyn.to(dev()), graph, c, db)
    assert rel_rows(db.float()[:, :3 * c].cpu(), d_ref16).max().item() < 8e-3
    # the former exp2 kernels (selected by passing the statistics buffer) agree with the series form
    aux = torch.empty(Et, 4 * c, dtype=torch.float32, device=dev())
    if c <= 256:
        y2 = torch.zeros_like(y)
        ops.attention_fwd(gtp.to(dev()), c, y2, aux=aux)
        assert rel_rows(y2.float()[:, :c].cpu(), y.float()[:, :c].cpu().double()).max().item() < 8e-3
        dg2 = torch.zeros_like(dgtp)
        ops.attention_bwd(gtp.to(dev()), dyn.to(dev()), graph, c, dg2, aux=aux)
        assert rel_rows(dg2.float()[:, :3 * c].cpu(), dgtp.float()[:, :3 * c].cpu().double()).max().item() < 1.6e-2
