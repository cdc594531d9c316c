"""Diagnostic (GPU): relative errors of every output / gradient of the layer against the fp64 oracle, plus
standalone checks of the attention kernels against torch autograd.  Prints, never asserts."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import relpose_gnn_b200 as rpg  # noqa: E402
from oracle import restatement as R  # noqa: E402
from relpose_gnn_b200 import ops  # noqa: E402
from relpose_gnn_b200.graph import GraphBatch  # noqa: E402
from relpose_gnn_b200.layers import PARAM_ORDER  # noqa: E402

dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp(min=1e-30)).item()


def attention_check(c=64, Et=500, N=9):
    g = GraphBatch.fully_connected((Et + 71) // 72, N, dev)
    Et = g.n_edge_rows
    gen = torch.Generator(device="cuda").manual_seed(1)
    gtp = torch.randn(Et, 3 * c, device=dev, generator=gen, requires_grad=True)
    dyn = torch.randn(g.n_node_rows, c, device=dev, generator=gen)
    y = torch.zeros(Et, max(c, 64), dtype=torch.bfloat16, device=dev)
    aux = torch.empty(Et, 4 * c, device=dev)
    ops.attention_fwd(gtp.detach(), c, y, aux=aux)
    gg, th, ph = gtp[:, :c], gtp[:, c:2 * c], gtp[:, 2 * c:]
    s = torch.softmax(ph.unsqueeze(2) * th.unsqueeze(1), -1)
    y_ref = (s * gg.unsqueeze(1)).sum(-1)
    print(f"attention fwd c={c}: rel {rel(y[:, :c].float(), y_ref):.3e}")
    ei = g.edge_index()
    dy = dyn[ei[1]]
    (y_ref * dy).sum().backward()
    dgtp = torch.zeros(Et, ops.pad64(3 * c), dtype=torch.bfloat16, device=dev)
    ops.attention_bwd(gtp.detach(), dyn, g, c, dgtp)
    for nm, sl in (("dg", slice(0, c)), ("dtheta", slice(c, 2 * c)), ("dphi", slice(2 * c, 3 * c))):
        print(f"attention bwd c={c} {nm}: rel {rel(dgtp[:, sl].float(), gtp.grad[:, sl]):.3e}")
    dgtp2 = torch.zeros_like(dgtp)
    ops.attention_bwd(gtp.detach(), dyn, g, c, dgtp2, aux=aux)
    for nm, sl in (("dg", slice(0, c)), ("dtheta", slice(c, 2 * c)), ("dphi", slice(2 * c, 3 * c))):
        print(f"attention bwd (saved statistics) c={c} {nm}: rel {rel(dgtp2[:, sl].float(), gtp.grad[:, sl]):.3e}")


def layer_check(D, N, Gn, seed, ct_out_scale=1.0, ct_e_scale=1.0, quantize_inputs=False):
    case = R.synth_layer_case(D, N, Gn, seed)
    params, x, e, ei = case["params"], case["x"], case["e"], case["edge_index"]
    if quantize_inputs:   # feed the oracle bf16-representable inputs/weights: isolates arithmetic from input rounding
        params = {k: v.bfloat16().double() if k.endswith("weight") else v for k, v in params.items()}
        x, e = x.bfloat16().double(), e.bfloat16().double()
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xo, eo = x.clone().requires_grad_(True), e.clone().requires_grad_(True)
    out_o, en_o, inter = R.layer_forward(p, xo, ei, eo, return_intermediates=True)
    ((out_o * case["ct_out"] * ct_out_scale).sum() + (en_o * case["ct_e"] * ct_e_scale).sum()).backward()
    m = rpg.simpleConvEdge_upt(D, D, D)
    m.load_state_dict({k: v.float() for k, v in params.items()})
    m = m.to(dev)
    xg = x.float().to(dev).requires_grad_(True)
    eg = e.float().to(dev).requires_grad_(True)
    out, en = m(xg, ei.to(dev), eg)
    ((out * (case["ct_out"] * ct_out_scale).float().to(dev)).sum() + (en * (case["ct_e"] * ct_e_scale).float().to(dev)).sum()).backward()
    print(f"--- layer D={D} N={N} G={Gn} ct_out={ct_out_scale} ct_e={ct_e_scale} quantized_inputs={quantize_inputs}")
    print(f"  out {rel(out, out_o):.3e}  e_new {rel(en, en_o):.3e}  dx {rel(xg.grad, xo.grad):.3e}  de {rel(eg.grad, eo.grad):.3e}")
    for k in PARAM_ORDER:
        print(f"  grad {k:32s} {rel(m.get_parameter(k).grad, p[k].grad):.3e}   |ref| {p[k].grad.norm().item():.3e}")


if __name__ == "__main__":
    attention_check(64)
    attention_check(16, Et=200, N=4)
    layer_check(128, 9, 2, 100)
    layer_check(128, 9, 2, 100, ct_out_scale=0.0)
    layer_check(128, 9, 2, 100, ct_e_scale=0.0)
    layer_check(128, 9, 2, 100, quantize_inputs=True)
    layer_check(512, 9, 4, 300)
    layer_check(512, 9, 4, 300, quantize_inputs=True)
