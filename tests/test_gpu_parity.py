"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI, against
(1) the golden fixtures produced by the real reference and (2) the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): bf16 mode within 2e-2 relative.  relative = ||a-b||_F / ||b||_F per tensor.
"""
import os

import numpy as np
import pytest
import torch

import relpose_gnn_b200 as rpg
from oracle import restatement as R
from relpose_gnn_b200 import graph as G
from relpose_gnn_b200.layers import PARAM_ORDER

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_BF16 = 2e-2
# Gradients against the plain fp64 reference: a bf16 forward flips the sign of ~0.3 % of the near-zero ReLU
# pre-activations, and each flip changes d(relu) by O(1) => relative Frobenius error ~ sqrt(flip fraction) = 3-6 %
# (measured, tests/diag/layer_diag.py; drops to the 3e-3 arithmetic floor before the first ReLU).  This is a property
# of bf16 arithmetic, not of the kernels, so the fixture comparison is loose ...
TOL_GRAD_FLIP = 1e-1
TOL_GRAD_SMALL = 2.5e-1        # per-tensor bound for tiny gradients (attention biases) dominated by flip noise
# ... and the rigorous gradient check imposes the kernel's own activation pattern on the oracle (R._relu):
TOL_GRAD = 1.5e-2


def rel(a, b):
    a = torch.as_tensor(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b.detach().cpu() if torch.is_tensor(b) else b), dtype=torch.float64)
    return ((a - b).norm() / b.norm().clamp(min=1e-30)).item()


def dev():
    return torch.device("cuda:0")


def make_layer(D, params):
    m = rpg.simpleConvEdge_upt(D, D, D)
    m.load_state_dict({k: v.float() for k, v in params.items()})
    return m.to(dev())


@pytest.mark.parametrize("tag", ["D128_N9_G2", "D128_N4_G3", "D512_N8_G2", "D512_N17_G1"])
def test_layer_against_reference_fixture(tag):
    fx = np.load(os.path.join(GOLD, f"layer_{tag}.npz"))
    D, N, Gn, seed = [int(v) for v in fx["meta"]]
    case = R.synth_layer_case(D, N, Gn, seed)
    m = make_layer(D, case["params"])
    x = case["x"].float().to(dev()).requires_grad_(True)
    e = case["e"].float().to(dev()).requires_grad_(True)
    ei = case["edge_index"].to(dev())
    out, e_new = m(x, ei, e)
    assert out.dtype == torch.float32 and out.shape == (Gn * N, D) and e_new.shape == (ei.size(1), D)
    assert rel(out, fx["out_f64"]) < TOL_BF16
    assert rel(e_new, fx["e_new_f64"]) < TOL_BF16
    ((out * case["ct_out"].float().to(dev())).sum() + (e_new * case["ct_e"].float().to(dev())).sum()).backward()
    assert rel(x.grad, fx["dx"]) < TOL_GRAD_FLIP
    assert rel(e.grad, fx["de"]) < TOL_GRAD_FLIP
    num = den = 0.0
    errs = {}
    for k in PARAM_ORDER:
        g = m.get_parameter(k).grad
        if "grad." + k in fx.files:
            ref = torch.from_numpy(fx["grad." + k])
            errs[k] = rel(g, ref)
            num += (g.double().cpu() - ref).norm().item() ** 2
            den += ref.norm().item() ** 2
        else:
            u, w = R.grad_probe_vectors(g.shape)
            g64 = g.double().cpu().numpy()
            errs[k + "@v"] = np.linalg.norm(g64 @ w - fx["gradrows." + k]) / np.linalg.norm(fx["gradrows." + k])
            errs[k + "u@"] = np.linalg.norm(u @ g64 - fx["gradcols." + k]) / np.linalg.norm(fx["gradcols." + k])
    # the attention projections receive tiny gradients (|g| ~ 1e-2 of the MLP ones) that the flip noise dominates
    bad = {k: v for k, v in errs.items() if v > (2 * TOL_GRAD_SMALL if k.startswith("att.") else TOL_GRAD_SMALL)}
    assert not bad, bad
    if den:
        assert (num / den) ** 0.5 < TOL_GRAD_FLIP      # all parameter gradients as one vector


@pytest.mark.parametrize("D,N,Gn,drop_edges", [(512, 9, 5, False), (512, 17, 3, False), (512, 8, 16, True),
                                               (256, 9, 33, True), (128, 3, 50, False), (1024, 8, 2, False),
                                               (2048, 8, 2, True), (384, 9, 4, True), (640, 5, 6, False)])
def test_layer_against_oracle(D, N, Gn, drop_edges):
    """Forward against the plain oracle; backward against the oracle with the kernel's activation pattern imposed.
    Inputs, weights and cotangents are bf16-representable so that only the kernel arithmetic is measured."""
    from relpose_gnn_b200 import ops
    from relpose_gnn_b200.layers import layer_backward_raw, layer_forward_raw
    seed = 1000 + D + N + Gn
    q = lambda t: t.bfloat16().double()                                         # noqa: E731
    params = {k: (q(v) if k.endswith("weight") else v.float().double())
              for k, v in R.synth_params(R.LAYER_SHAPES(D), seed, torch.float64).items()}
    x = q(R.synth_inputs(Gn, N, D, seed + 1, torch.float64)[0])
    tmpl = R.fc_edge_index(N)
    if drop_edges:
        keep = R.edge_dropout_keep(N * (N - 1) // 2, np.random.RandomState(seed).random_sample(N * (N - 1) // 2))
        tmpl = R.apply_edge_dropout(tmpl, keep)
    ei = R.batched_edge_index(tmpl, Gn, N)
    gen = torch.Generator().manual_seed(seed + 2)
    e = q(torch.relu(torch.randn(ei.size(1), D, generator=gen, dtype=torch.float64)))
    ct_o = q(torch.randn(Gn * N, D, generator=gen, dtype=torch.float64))
    ct_e = q(torch.randn(ei.size(1), D, generator=gen, dtype=torch.float64))
    # CUDA path through the raw C-ABI wrappers (bf16 in / bf16 out)
    m = make_layer(D, params)
    graph = G.from_edge_index(ei.to(dev()), Gn * N)
    assert graph.G == Gn and graph.N == N and graph.Ep == tmpl.size(1)
    lw = m._packed(dev()).refresh(m)
    acts = layer_forward_raw(lw, graph, x.to(dev()).bfloat16(), e.to(dev()).bfloat16())
    grads = {k: torch.zeros_like(m.get_parameter(k)) for k in PARAM_ORDER}
    dx, de = layer_backward_raw(lw, graph, acts, ct_o.to(dev()).bfloat16(), ct_e.to(dev()).bfloat16(), grads)
    # oracle (fp64 CPU)
    out_o, en_o = R.layer_forward(params, x, ei, e)
    assert rel(acts["out"].float(), out_o) < TOL_BF16 and rel(acts["e_new"].float(), en_o) < TOL_BF16
    masks = {k: (acts[k] > 0).cpu() for k in ("h1", "h2", "h3")}
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xo, eo = x.clone().requires_grad_(True), e.clone().requires_grad_(True)
    out_m, en_m = R.layer_forward(p, xo, ei, eo, relu_masks=masks)
    ((out_m * ct_o).sum() + (en_m * ct_e).sum()).backward()
    assert rel(dx.float(), xo.grad) < TOL_GRAD
    assert rel(de.float(), eo.grad) < TOL_GRAD
    errs = {k: rel(grads[k], p[k].grad) for k in PARAM_ORDER}
    # attention bias gradients are sums of strongly cancelling terms (each dg_j ~ mean of dy): allow 4x
    bad = {k: v for k, v in errs.items() if v > (4 * TOL_GRAD if k.startswith("att.") and k.endswith("bias") else TOL_GRAD)}
    assert not bad, bad


def test_layer_is_deterministic_and_inputs_untouched():
    D, N, Gn = 512, 9, 64
    case = R.synth_layer_case(D, N, Gn, 77)
    m = make_layer(D, case["params"])
    x = case["x"].float().to(dev())
    e = case["e"].float().to(dev())
    ei = case["edge_index"].to(dev())
    x0, e0 = x.clone(), e.clone()
    runs = []
    for _ in range(2):
        xr, er = x.clone().requires_grad_(True), e.clone().requires_grad_(True)
        out, en = m(xr, ei, er)
        (out.sum() + en.sum()).backward()
        runs.append((out, en, xr.grad, er.grad, m.get_parameter("mlp.0.weight").grad.clone()))
        m.zero_grad()
    for a, b in zip(*runs):
        assert torch.equal(a, b)                      # no atomics anywhere: bitwise reproducible
    assert torch.equal(x, x0) and torch.equal(e, e0)


def test_graph_permutation_equivariance_at_full_size():
    """Config-C sized batch (4096 graphs x 9 nodes, D=512): graphs are independent, so permuting the graphs of the
    batch permutes the outputs bit-for-bit (size-independent property; the oracle cannot run this size quickly)."""
    D, N, Gn = 512, 9, 4096
    params = R.synth_params(R.LAYER_SHAPES(D), 5, torch.float32)
    m = make_layer(D, params)
    g = G.GraphBatch.fully_connected(Gn, N, dev())
    ei = G.attach(g.edge_index(), g)
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(Gn * N, D, device=dev(), generator=gen).bfloat16()
    e = torch.randn(Gn * g.Ep, D, device=dev(), generator=gen).relu().bfloat16()
    with torch.no_grad():
        out, en = m(x, ei, e)
        perm = torch.randperm(Gn, device=dev(), generator=gen)
        xp = x.view(Gn, N, D)[perm].reshape(-1, D).contiguous()
        ep = e.view(Gn, g.Ep, D)[perm].reshape(-1, D).contiguous()
        outp, enp = m(xp, ei, ep)
    assert torch.equal(outp.view(Gn, N, D), out.view(Gn, N, D)[perm])
    assert torch.equal(enp.view(Gn, g.Ep, D), en.view(Gn, g.Ep, D)[perm])
    assert torch.isfinite(out.float()).all()
    # spot-check 3 graphs of the big batch against the oracle
    idx = [0, 1777, 4095]
    xs = x.view(Gn, N, D)[idx].reshape(-1, D).double().cpu()
    es = e.view(Gn, g.Ep, D)[idx].reshape(-1, D).double().cpu()
    o_ref, e_ref = R.layer_forward({k: v.double() for k, v in params.items()}, xs,
                                   R.batched_edge_index(R.fc_edge_index(N), 3, N), es)
    assert rel(out.view(Gn, N, D)[idx].reshape(-1, D).float(), o_ref) < TOL_BF16
    assert rel(en.view(Gn, g.Ep, D)[idx].reshape(-1, D).float(), e_ref) < TOL_BF16


def test_edge_index_validation_and_generality():
    D, N, Gn = 128, 4, 3
    params = R.synth_params(R.LAYER_SHAPES(D), 1, torch.float64)
    m = make_layer(D, params)
    x = torch.randn(Gn * N, D, device=dev())
    ei = R.batched_edge_index(R.fc_edge_index(N), Gn, N).to(dev())
    e = torch.randn(ei.size(1), D, device=dev()).relu()
    with pytest.raises(TypeError):
        m(x, ei.int(), e)
    with pytest.raises(ValueError):
        m(x, ei[:, :0], e[:0])
    oob = ei.clone()
    oob[1, 5] = Gn * N                                         # node index out of range
    with pytest.raises(ValueError, match="template"):
        m(x, oob, e)
    out, en = m(x, ei, e)                                      # the good one passes and is cached on the tensor
    assert ei.rpg_graph.G == Gn and ei.rpg_graph.N == N and ei.rpg_graph.Ep == N * (N - 1)
    # a batch whose graphs differ is not a uniform template of 3 graphs; it is still ONE valid graph of 12 nodes,
    # so it runs as G=1 with its own template and must agree with the oracle's generic gather/scatter
    odd = ei.clone()
    odd[0, 17] = (odd[0, 17] + 1) % (Gn * N)
    out2, en2 = m(x, odd, e)
    assert odd.rpg_graph.G == 1 and odd.rpg_graph.N == Gn * N and odd.rpg_graph.Ep == odd.size(1)
    o_ref, e_ref = R.layer_forward(params, x.double().cpu(), odd.cpu(), e.double().cpu())
    assert rel(out2, o_ref) < TOL_BF16 and rel(en2, e_ref) < TOL_BF16
    o_ref, e_ref = R.layer_forward(params, x.double().cpu(), ei.cpu(), e.double().cpu())
    assert rel(out, o_ref) < TOL_BF16 and rel(en, e_ref) < TOL_BF16


@pytest.mark.parametrize("tag,droprate,edrop", [("D128_N9_G2", 0.0, False), ("D128_N8_G3_drop", 0.5, True)])
def test_stack_against_reference_fixture(tag, droprate, edrop):
    fx = np.load(os.path.join(GOLD, f"stack_{tag}.npz"))
    D, N, Gn, seed = [int(v) for v in fx["meta"]]
    case = R.synth_stack_case(D, N, Gn, seed, droprate, edrop)
    model = rpg.RelPoseGNN(D, D, D, droprate=droprate, gnn_recursion=2)
    model.load_state_dict({k: v.float() for k, v in case["params"].items()}, strict=False)
    model = model.to(dev())
    crit = rpg.PoseNetCriterion(sax=0.0, saq=-2.0).to(dev())
    x = case["x"].float().to(dev()).requires_grad_(True)
    ei = case["edge_index"].to(dev())
    kx = case["keep_x"].to(dev()) if droprate > 0 else None
    ke = case["keep_e"].to(dev()) if droprate > 0 else None
    pn, pe, ei_out = model(x, ei, keep_x=kx, keep_e=ke)
    assert ei_out is ei
    assert rel(pn, fx["pose_nodes"]) < TOL_BF16
    assert rel(pe, fx["pose_edges"]) < TOL_BF16
    loss, t_loss, q_loss = crit(pe, case["poses"].float().to(dev()), ei)
    assert np.allclose([loss.item(), t_loss.item(), q_loss.item()], fx["loss"], rtol=TOL_BF16, atol=2e-3)
    loss.backward()
    assert rel(x.grad, fx["dx"]) < 1.5e-1        # ReLU / L1 sign flips (see TOL_GRAD_FLIP), two rounds deep
    assert abs(crit.sax.grad.item() - fx["dsax"][0]) < 2e-2 and abs(crit.saq.grad.item() - fx["dsaq"][0]) < 2e-2
    num = den = 0.0
    for k in case["params"]:
        ref = torch.from_numpy(fx["grad." + k])
        g = model.get_parameter(k).grad
        if ref.abs().max() == 0:
            assert g is None or g.abs().max().item() == 0, k     # node heads / last node update get no gradient
        else:
            assert rel(g, ref) < 3e-1, k
            num += (g.double().cpu() - ref).norm().item() ** 2
            den += ref.norm().item() ** 2
    assert (num / den) ** 0.5 < 1.5e-1


def test_stack_backward_against_mask_matched_oracle():
    """Whole-stack gradients with the kernel's activation patterns imposed on the oracle (bf16-representable inputs);
    node AND edge pose heads receive cotangents so every parameter of the path gets a gradient."""
    D, N, Gn = 256, 9, 7
    q = lambda t: t.bfloat16().double()                                         # noqa: E731
    case = R.synth_stack_case(D, N, Gn, 999, droprate=0.5, edge_dropout=True)
    params = {k: (q(v) if k.endswith("weight") and not k.startswith("fc_") else v.float().double())
              for k, v in case["params"].items()}
    x = q(case["x"])
    ei = case["edge_index"]
    model = rpg.RelPoseGNN(D, D, D, droprate=0.5).to(dev())
    model.load_state_dict({k: v.float() for k, v in params.items()}, strict=False)
    model.keep_debug_activations = True
    xg = x.to(dev()).bfloat16().requires_grad_(True)
    pn, pe, _ = model(xg, ei.to(dev()), keep_x=case["keep_x"].to(dev()), keep_e=case["keep_e"].to(dev()))
    gen = torch.Generator().manual_seed(5)
    ct_n = torch.randn(pn.shape, generator=gen).double()
    ct_e = torch.randn(pe.shape, generator=gen).double()
    ((pn * ct_n.float().to(dev())).sum() + (pe * ct_e.float().to(dev())).sum()).backward()
    dbg = model.debug_activations
    masks = {"e0": (dbg["e0"] > 0).cpu(),
             "rounds": [{"h1": (a["h1"] > 0).cpu(), "h2": (a["h2"] > 0).cpu(), "h3": (a["h3"] > 0).cpu(),
                         "x": (a["out_relu"] > 0).cpu(), "e": (a["e_new_relu"] > 0).cpu()} for a in dbg["rounds"]]}
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xo = x.clone().requires_grad_(True)
    pn_o, pe_o, _, _ = R.stack_forward(p, xo, ei, 2, 0.5, case["keep_x"], case["keep_e"], relu_masks=masks)
    assert rel(pn, pn_o) < TOL_BF16 and rel(pe, pe_o) < TOL_BF16
    ((pn_o * ct_n).sum() + (pe_o * ct_e).sum()).backward()
    assert rel(xg.grad.float(), xo.grad) < 2e-2
    for k in params:
        assert rel(model.get_parameter(k).grad, p[k].grad) < 2e-2, k


@pytest.mark.parametrize("D,N,Gn", [(512, 9, 6), (1024, 8, 3), (2048, 8, 2)])
def test_stack_against_oracle_full_width_with_dropout(D, N, Gn):
    """D = 1024 / 2048 are the widths of the released R2 / R3 checkpoints (the tested set)."""
    case = R.synth_stack_case(D, N, Gn, 4242, droprate=0.5, edge_dropout=True)
    model = rpg.RelPoseGNN(D, D, D, droprate=0.5).to(dev())
    model.load_state_dict({k: v.float() for k, v in case["params"].items()}, strict=False)
    x = case["x"].float().to(dev())
    ei = case["edge_index"].to(dev())
    with torch.no_grad():
        pn, pe, _ = model(x, ei, keep_x=case["keep_x"].to(dev()), keep_e=case["keep_e"].to(dev()))
    pn_o, pe_o, _, _ = R.stack_forward(case["params"], case["x"], case["edge_index"], 2, 0.5, case["keep_x"], case["keep_e"])
    assert rel(pn, pn_o) < TOL_BF16 and rel(pe, pe_o) < TOL_BF16
    # in-kernel counter-based dropout == the mask rpg_dropout_mask materialises for the oracle
    from relpose_gnn_b200 import ops
    model.dropout_seed = 100
    with torch.no_grad():
        pn2, pe2, _ = model(x, ei)
    seed = model.last_seed                     # the (counter, rank)-hashed seed the kernels received
    kx = ops.dropout_mask(seed, 0.5, Gn * N, D, dev()).cpu().bool()
    ke = ops.dropout_mask(seed + 1, 0.5, ei.size(1), D, dev()).cpu().bool()
    assert 0.45 < kx.float().mean().item() < 0.55 and 0.45 < ke.float().mean().item() < 0.55
    pn_o2, pe_o2, _, _ = R.stack_forward(case["params"], case["x"], case["edge_index"], 2, 0.5, kx, ke)
    assert rel(pn2, pn_o2) < TOL_BF16 and rel(pe2, pe_o2) < TOL_BF16


def test_gemm_probes():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gemm_probe", os.path.join(os.path.dirname(GOLD), "..", "tools", "gemm_probe.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.nt_cases()
    assert mod.tn_cases()


def test_fused_gradient_accumulation_matches_autograd_path():
    """attach_grad_bucket: kernels accumulate straight into the flat bucket; must equal the grads autograd returns."""
    from relpose_gnn_b200 import parallel
    D, N, Gn = 256, 9, 11
    case = R.synth_stack_case(D, N, Gn, 77, droprate=0.5, edge_dropout=True)
    sd = {k: v.float() for k, v in case["params"].items()}
    x = case["x"].float().to(dev())
    ei = case["edge_index"].to(dev())
    poses = case["poses"].float().to(dev())
    kx, ke = case["keep_x"].to(dev()), case["keep_e"].to(dev())

    def run(fused):
        model = rpg.RelPoseGNN(D, D, D, droprate=0.5).to(dev())
        model.load_state_dict(sd, strict=False)
        crit = rpg.PoseNetCriterion(0.0, -2.0).to(dev())
        params = list(model.parameters()) + list(crit.parameters())
        if fused:
            bucket = parallel.FlatGradBucket(params)
            model.attach_grad_bucket(bucket)
        for _ in range(2):                      # two steps: accumulation semantics (+=) must match too
            pn, pe, _ = model(x, ei, keep_x=kx, keep_e=ke)
            loss, _, _ = crit(pe, poses, ei)
            loss.backward()
        return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).flatten() for p in params])

    a, b = run(False), run(True)
    # same kernels, different association of the two steps' sums ((g1 + a) + b vs g1 + (a + b)): not bitwise
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)
    assert (a - b).norm() / b.norm() < 1e-6


# ------------------------------------------------------------------------------------------------ fp32 mode
TOL_FP32 = 1e-4      # BASELINE.json north_star: fp32/TF32 mode within 1e-4 relative


@pytest.mark.parametrize("D,N,Gn,drop_edges", [(512, 9, 5, False), (512, 17, 2, True), (128, 8, 9, False)])
def test_fp32_mode_layer_against_oracle(D, N, Gn, drop_edges):
    """fp32 mode (split-bf16 operands, fp32 accumulation) against the fp64 oracle AND against the reference fixture's
    own fp32 run where one exists; bf16 mode on the same inputs for contrast."""
    seed = 5000 + D + N
    params = R.synth_params(R.LAYER_SHAPES(D), seed, torch.float64)
    x, _ = R.synth_inputs(Gn, N, D, seed + 1, torch.float64)
    tmpl = R.fc_edge_index(N)
    if drop_edges:
        keep = R.edge_dropout_keep(N * (N - 1) // 2, np.random.RandomState(seed).random_sample(N * (N - 1) // 2))
        tmpl = R.apply_edge_dropout(tmpl, keep)
    ei = R.batched_edge_index(tmpl, Gn, N)
    gen = torch.Generator().manual_seed(seed + 2)
    e = torch.relu(torch.randn(ei.size(1), D, generator=gen, dtype=torch.float64))
    out_o, en_o = R.layer_forward(params, x, ei, e)
    m = make_layer(D, params)
    m.precision = "fp32"
    with torch.no_grad():
        out, en = m(x.float().to(dev()), ei.to(dev()), e.float().to(dev()))
    assert out.dtype == torch.float32
    assert rel(out, out_o) < TOL_FP32 and rel(en, en_o) < TOL_FP32, (rel(out, out_o), rel(en, en_o))
    m.precision = "bf16"
    with torch.no_grad():
        out_b, _ = m(x.float().to(dev()), ei.to(dev()), e.float().to(dev()))
    assert rel(out_b, out_o) > 10 * rel(out, out_o)          # the fp32 mode really is a different, tighter arithmetic


def test_fp32_mode_layer_against_reference_fixture():
    fx = np.load(os.path.join(GOLD, "layer_D512_N8_G2.npz"))
    D, N, Gn, seed = [int(v) for v in fx["meta"]]
    case = R.synth_layer_case(D, N, Gn, seed)
    m = make_layer(D, case["params"])
    m.precision = "fp32"
    with torch.no_grad():
        out, en = m(case["x"].float().to(dev()), case["edge_index"].to(dev()), case["e"].float().to(dev()))
    assert rel(out, fx["out_f64"]) < TOL_FP32 and rel(en, fx["e_new_f64"]) < TOL_FP32
    assert rel(out, fx["out_f32"]) < TOL_FP32                 # the reference's own fp32 CPU run


@pytest.mark.parametrize("droprate", [0.0, 0.5])
def test_fp32_mode_stack_against_oracle(droprate):
    """BASELINE config-B shape in miniature: the whole GNN stack in fp32 mode, poses within 1e-4."""
    D, N, Gn = 512, 9, 6
    case = R.synth_stack_case(D, N, Gn, 6100, droprate=droprate, edge_dropout=False)
    model = rpg.RelPoseGNN(D, D, D, droprate=droprate).to(dev())
    model.load_state_dict({k: v.float() for k, v in case["params"].items()}, strict=False)
    model.precision = "fp32"
    kx = case["keep_x"].to(dev()) if droprate > 0 else None
    ke = case["keep_e"].to(dev()) if droprate > 0 else None
    with torch.no_grad():
        pn, pe, _ = model(case["x"].float().to(dev()), case["edge_index"].to(dev()), keep_x=kx, keep_e=ke)
    pn_o, pe_o, _, _ = R.stack_forward(case["params"], case["x"], case["edge_index"], 2, droprate, case["keep_x"], case["keep_e"])
    assert rel(pn, pn_o) < TOL_FP32 and rel(pe, pe_o) < TOL_FP32, (rel(pn, pn_o), rel(pe, pe_o))


@pytest.mark.parametrize("D,N,Gn,drop_edges", [(512, 9, 5, False), (256, 9, 7, True), (128, 4, 11, False), (128, 9, 6, "tiny")])
def test_fp32_mode_layer_backward_against_oracle(D, N, Gn, drop_edges):
    """fp32-mode gradients (split-bf16 dgrad / wgrad on the same kernels) against the fp64 oracle.
    (1) with the kernel's own ReLU patterns imposed: every gradient within 1e-4 -- the arithmetic is fp32-accurate;
    (2) against the plain oracle, NO masks: 2e-3.  What remains there is not arithmetic: a pre-activation closer to zero
        than the fp32-mode error (~5e-6 relative) flips its ReLU, each flip changes d relu by O(1), and the Frobenius error
        goes like sqrt(flip fraction) ~ 1e-3 (the reference's own fp32 run differs from fp64 the same way)."""
    from relpose_gnn_b200.layers import layer_backward_split_raw, layer_forward_split_raw
    from relpose_gnn_b200 import ops
    seed = 7000 + D + N
    params = R.synth_params(R.LAYER_SHAPES(D), seed, torch.float64)
    params = {k: v.float().double() for k, v in params.items()}                 # fp32-representable
    x = R.synth_inputs(Gn, N, D, seed + 1, torch.float64)[0].float().double()
    tmpl = R.fc_edge_index(N)
    if drop_edges == "tiny":          # three undirected edges survive: no one-hot panels (epilogue gathers), isolated nodes
        keep = np.zeros(N * (N - 1) // 2, bool)
        keep[[0, 5, 11]] = True
        tmpl = R.apply_edge_dropout(tmpl, keep)
    elif drop_edges:
        keep = R.edge_dropout_keep(N * (N - 1) // 2, np.random.RandomState(seed).random_sample(N * (N - 1) // 2))
        tmpl = R.apply_edge_dropout(tmpl, keep)
    ei = R.batched_edge_index(tmpl, Gn, N)
    gen = torch.Generator().manual_seed(seed + 2)
    e = torch.relu(torch.randn(ei.size(1), D, generator=gen, dtype=torch.float64)).float().double()
    ct_o = torch.randn(Gn * N, D, generator=gen, dtype=torch.float64).float().double()
    ct_e = torch.randn(ei.size(1), D, generator=gen, dtype=torch.float64).float().double()
    m = make_layer(D, params)
    m.precision = "fp32"
    graph = G.from_edge_index(ei.to(dev()), Gn * N)
    lw = m._packed_split(dev()).refresh(m, training=True)
    acts = layer_forward_split_raw(lw, graph, ops.to_split(x.float().to(dev())), ops.to_split(e.float().to(dev())), for_backward=True)
    grads = {k: torch.zeros_like(m.get_parameter(k)) for k in PARAM_ORDER}
    dx, de = layer_backward_split_raw(lw, graph, acts, ops.to_split(ct_o.float().to(dev())), ops.to_split(ct_e.float().to(dev())), grads)
    dx, de = ops.from_split(*dx), ops.from_split(*de)
    out = ops.from_split(*acts["out"])
    en = ops.from_split(*acts["e_new"])

    def oracle(masks):
        p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        xo, eo = x.clone().requires_grad_(True), e.clone().requires_grad_(True)
        o, n = R.layer_forward(p, xo, ei, eo, relu_masks=masks)
        ((o * ct_o).sum() + (n * ct_e).sum()).backward()
        return o, n, xo.grad, eo.grad, {k: p[k].grad for k in PARAM_ORDER}

    masks = {k: (ops.from_split(*acts[k]) > 0).cpu() for k in ("h1", "h2", "h3")}
    o, n, gx, ge, gp = oracle(masks)
    assert rel(out, o) < TOL_FP32 and rel(en, n) < TOL_FP32
    assert rel(dx, gx) < TOL_FP32 and rel(de, ge) < TOL_FP32, (rel(dx, gx), rel(de, ge))
    errs = {k: rel(grads[k], gp[k]) for k in PARAM_ORDER}
    bad = {k: v for k, v in errs.items() if v > (4 * TOL_FP32 if k.startswith("att.") and k.endswith("bias") else TOL_FP32)}
    assert not bad, bad
    o, n, gx, ge, gp = oracle(None)                                           # plain oracle, no masks
    assert rel(dx, gx) < 2e-3 and rel(de, ge) < 2e-3, (rel(dx, gx), rel(de, ge))
    num = sum((grads[k].double().cpu() - gp[k]).norm().item() ** 2 for k in PARAM_ORDER)
    den = sum(gp[k].norm().item() ** 2 for k in PARAM_ORDER)
    assert (num / den) ** 0.5 < 2e-3, (num / den) ** 0.5
    # and through the module API (autograd): same numbers
    m.zero_grad()
    xg = x.float().to(dev()).requires_grad_(True)
    eg = e.float().to(dev()).requires_grad_(True)
    o2, n2 = m(xg, ei.to(dev()), eg)
    ((o2 * ct_o.float().to(dev())).sum() + (n2 * ct_e.float().to(dev())).sum()).backward()
    assert rel(xg.grad, dx.double().cpu()) < 1e-6 and rel(m.get_parameter("mlp.0.weight").grad, grads["mlp.0.weight"].double().cpu()) < 1e-6


@pytest.mark.parametrize("node_ct", [True, False])
def test_fp32_mode_training_step_against_mask_matched_oracle(node_ct):
    """The whole stack in fp32 mode with feature dropout + edge dropout: loss and every parameter gradient against the
    oracle with the kernel's ReLU patterns imposed (1e-4 on the gradient vector, 4e-4 per tensor).  node_ct = False is the
    reference's training objective (edge poses only, train.py:256-264): the last round has no gradient on `out`."""
    from relpose_gnn_b200 import ops
    D, N, Gn = 256, 9, 6
    case = R.synth_stack_case(D, N, Gn, 8100, droprate=0.5, edge_dropout=True)
    params = {k: v.float().double() for k, v in case["params"].items()}
    x = case["x"].float().double()
    ei = case["edge_index"]
    model = rpg.RelPoseGNN(D, D, D, droprate=0.5).to(dev())
    model.load_state_dict({k: v.float() for k, v in params.items()}, strict=False)
    model.precision = "fp32"
    model.keep_debug_activations = True
    xg = x.float().to(dev()).requires_grad_(True)
    pn, pe, _ = model(xg, ei.to(dev()), keep_x=case["keep_x"].to(dev()), keep_e=case["keep_e"].to(dev()))
    gen = torch.Generator().manual_seed(5)
    ct_n = torch.randn(pn.shape, generator=gen).double()
    ct_e = torch.randn(pe.shape, generator=gen).double()
    if not node_ct:
        ct_n = torch.zeros_like(ct_n)
        (pe * ct_e.float().to(dev())).sum().backward()
    else:
        ((pn * ct_n.float().to(dev())).sum() + (pe * ct_e.float().to(dev())).sum()).backward()
    dbg = model.debug_activations
    fs = lambda pr: (ops.from_split(*pr) > 0).cpu()                            # noqa: E731
    masks = {"e0": fs(dbg["e0"]),
             "rounds": [{"h1": fs(a["h1"]), "h2": fs(a["h2"]), "h3": fs(a["h3"]), "x": fs(a["out_relu"]), "e": fs(a["e_new_relu"])}
                        for a in dbg["rounds"]]}
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xo = x.clone().requires_grad_(True)
    pn_o, pe_o, _, _ = R.stack_forward(p, xo, ei, 2, 0.5, case["keep_x"], case["keep_e"], relu_masks=masks)
    assert rel(pn, pn_o) < TOL_FP32 and rel(pe, pe_o) < TOL_FP32
    ((pn_o * ct_n).sum() + (pe_o * ct_e).sum()).backward()
    assert rel(xg.grad, xo.grad) < TOL_FP32, rel(xg.grad, xo.grad)
    num = den = 0.0
    for k in params:
        g = model.get_parameter(k).grad
        if p[k].grad is None or p[k].grad.abs().max() == 0:
            assert g is None or g.abs().max().item() == 0, k     # node heads / last node update without a node cotangent
            continue
        assert rel(g, p[k].grad) < 4e-4, (k, rel(g, p[k].grad))
        num += (g.double().cpu() - p[k].grad).norm().item() ** 2
        den += p[k].grad.norm().item() ** 2
    assert (num / den) ** 0.5 < TOL_FP32, (num / den) ** 0.5


def test_qexp_and_eval_composition_against_reference_fixtures():
    """pose_utils.qexp and the evaluation composition of test.py:227-243."""
    fx = np.load(os.path.join(GOLD, "qexp.npz"))
    q = rpg.qexp(torch.from_numpy(fx["v"]).float().to(dev()))
    assert np.allclose(q.cpu().numpy(), fx["q"], rtol=0, atol=1e-6)
    ec = np.load(os.path.join(GOLD, "eval_compose.npz"))
    fc = np.load(os.path.join(GOLD, "fc_enumeration.npz"))
    for case in range(5):
        n, ref_node, _ = [int(v) for v in ec[f"case{case}_meta"]]
        ei = torch.from_numpy(fc[f"fc_N{n}"]).to(dev())
        pred, targ = rpg.compose_query_pose(torch.from_numpy(ec[f"case{case}_output_R"]).to(dev()),
                                            torch.from_numpy(ec[f"case{case}_target"]).to(dev()), ei, ref_node,
                                            ec[f"case{case}_pose_m"], ec[f"case{case}_pose_s"])
        assert np.allclose(pred.cpu().numpy()[0], ec[f"case{case}_pred7"], rtol=0, atol=2e-6)
        assert np.allclose(targ.cpu().numpy()[0], ec[f"case{case}_targ7"][0], rtol=0, atol=2e-6)
    # a batch of graphs with an edge-dropout template against the oracle
    Gn, N = 37, 9
    keep = R.edge_dropout_keep(N * (N - 1) // 2, np.random.RandomState(3).random_sample(N * (N - 1) // 2))
    tmpl = R.apply_edge_dropout(R.fc_edge_index(N), keep)
    ei = R.batched_edge_index(tmpl, Gn, N)
    gen = torch.Generator().manual_seed(5)
    pe = torch.randn(ei.size(1), 6, generator=gen) * 0.3
    pa = torch.randn(Gn * N, 6, generator=gen) * 0.5
    into0 = int((tmpl[1] == 0).sum())
    for ref_node in range(min(into0, 3)):
        pred, targ = rpg.compose_query_pose(pe.to(dev()), pa.to(dev()), ei.to(dev()), ref_node, [0.1, -0.2, 0.3], [2.0, 1.5, 0.5])
        po, to = R.compose_eval_batch(pe.numpy(), pa.numpy(), tmpl.numpy(), Gn, N, ref_node, [0.1, -0.2, 0.3], [2.0, 1.5, 0.5])
        assert np.allclose(pred.cpu().numpy(), po, rtol=0, atol=2e-6) and np.allclose(targ.cpu().numpy(), to, rtol=0, atol=2e-6)
    with pytest.raises(ValueError):
        rpg.compose_query_pose(pe.to(dev()), pa.to(dev()), ei.to(dev()), into0)
    # error metrics of the evaluation loop (test.py:202-203, 262-265) against the reference functions' own numbers
    t_err, q_err = rpg.pose_errors(torch.from_numpy(ec["err_pred"]).to(dev()), torch.from_numpy(ec["err_targ"]).to(dev()))
    assert np.allclose(t_err.cpu().numpy(), ec["err_t"], rtol=1e-6, atol=1e-6)
    assert np.allclose(q_err.cpu().numpy(), ec["err_q"], rtol=1e-5, atol=0.1)       # acos near 1: fp32 inputs => ~0.05 deg floor


# ------------------------------------------------------------------ sibling layers
def _sibling_modules(kind, D, params):
    m = rpg.simpleConvEdge(D, D, D) if kind == "convedge" else rpg.simpleConv(D, D)
    m.load_state_dict({k: v.float() for k, v in params.items()})
    return m.to(dev())


@pytest.mark.parametrize("kind,tag", [("convedge", "D128_N9_G2"), ("convedge", "D128_N5_G3"),
                                      ("conv", "D128_N9_G2"), ("conv", "D128_N4_G3")])
def test_sibling_layers_against_reference_fixture(kind, tag):
    """simpleConvEdge / simpleConv (my_gnn_layer.py:242-274, 394-412) drop-ins against the real reference's output."""
    fx = np.load(os.path.join(GOLD, f"{kind}_{tag}.npz"))
    D, N, Gn, seed = [int(v) for v in fx["meta"]]
    case = R.synth_sibling_case(kind, D, N, Gn, seed)
    m = _sibling_modules(kind, D, case["params"])
    assert set(m.state_dict().keys()) == set(case["params"].keys())
    x = case["x"].float().to(dev()).requires_grad_(True)
    ei = case["edge_index"].to(dev())
    if kind == "convedge":
        e = case["e"].float().to(dev()).requires_grad_(True)
        out, e_new = m(x, ei, e)
        assert rel(e_new, fx["e_new_f64"]) < TOL_BF16
        ((out * case["ct_out"].float().to(dev())).sum() + (e_new * case["ct_e"].float().to(dev())).sum()).backward()
        assert rel(e.grad, fx["de"]) < TOL_GRAD_FLIP
    else:
        out = m(x, ei)
        (out * case["ct_out"].float().to(dev())).sum().backward()
    assert out.dtype == torch.float32 and out.shape == (Gn * N, D)
    assert rel(out, fx["out_f64"]) < TOL_BF16
    assert rel(x.grad, fx["dx"]) < TOL_GRAD_FLIP
    num = den = 0.0
    for k, p in m.named_parameters():
        ref = torch.from_numpy(fx["grad." + k])
        num += (p.grad.double().cpu() - ref).norm().item() ** 2
        den += ref.norm().item() ** 2
    assert (num / den) ** 0.5 < TOL_GRAD_FLIP


@pytest.mark.parametrize("kind,D,N,Gn,drop_edges", [("convedge", 512, 9, 5, False), ("convedge", 256, 8, 21, True),
                                                    ("conv", 512, 9, 5, False), ("conv", 128, 17, 3, True)])
def test_sibling_layers_against_mask_matched_oracle(kind, D, N, Gn, drop_edges):
    """Tight gradient check: bf16-representable inputs, the kernel's own ReLU patterns imposed on the fp64 oracle."""
    seed = 3000 + D + N + Gn
    q = lambda t: t.bfloat16().double()                                         # noqa: E731
    shapes = R.CONV_EDGE_SHAPES(D) if kind == "convedge" else R.CONV_SHAPES(D)
    params = {k: (q(v) if k.endswith("weight") else v.float().double())
              for k, v in R.synth_params(shapes, seed, torch.float64).items()}
    x = q(R.synth_inputs(Gn, N, D, seed + 1, torch.float64)[0])
    tmpl = R.fc_edge_index(N)
    if drop_edges:
        keep = R.edge_dropout_keep(N * (N - 1) // 2, np.random.RandomState(seed).random_sample(N * (N - 1) // 2))
        tmpl = R.apply_edge_dropout(tmpl, keep)
    ei = R.batched_edge_index(tmpl, Gn, N)
    gen = torch.Generator().manual_seed(seed + 2)
    e = q(torch.relu(torch.randn(ei.size(1), D, generator=gen, dtype=torch.float64)))
    ct_o = q(torch.randn(Gn * N, D, generator=gen, dtype=torch.float64))
    ct_e = q(torch.randn(ei.size(1), D, generator=gen, dtype=torch.float64))
    m = _sibling_modules(kind, D, params)
    xg = x.float().to(dev()).requires_grad_(True)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xo = x.clone().requires_grad_(True)
    if kind == "convedge":
        from relpose_gnn_b200.layers import PARAM_ORDER_EDGE, layer_backward_raw, layer_forward_raw
        graph = G.from_edge_index(ei.to(dev()), Gn * N)
        lw = m._packed(dev()).refresh(m)
        acts = layer_forward_raw(lw, graph, x.to(dev()).bfloat16(), e.to(dev()).bfloat16())
        grads = {k: torch.zeros_like(m.get_parameter(k)) for k in PARAM_ORDER_EDGE}
        dx, de = layer_backward_raw(lw, graph, acts, ct_o.to(dev()).bfloat16(), ct_e.to(dev()).bfloat16(), grads)
        out_o, en_o = R.conv_edge_forward(params, x, ei, e)
        assert rel(acts["out"].float(), out_o) < TOL_BF16 and rel(acts["e_new"].float(), en_o) < TOL_BF16
        masks = {k: (acts[k] > 0).cpu() for k in ("h1", "h2")}
        eo = e.clone().requires_grad_(True)
        out_m, en_m = R.conv_edge_forward(p, xo, ei, eo, relu_masks=masks)
        ((out_m * ct_o).sum() + (en_m * ct_e).sum()).backward()
        assert rel(de.float(), eo.grad) < TOL_GRAD
        errs = {k: rel(grads[k], p[k].grad) for k in PARAM_ORDER_EDGE}
    else:
        out = m(xg, ei.to(dev()))
        (out * ct_o.float().to(dev())).sum().backward()
        dx = xg.grad
        assert rel(out, R.conv_forward(params, x, ei)) < TOL_BF16
        # the kernel's activation pattern: recompute h in fp64 from the bf16 node projections' sign is not observable
        # from outside, so impose the oracle's own pattern where |pre-activation| is not tiny and compare loosely below
        out_m = R.conv_forward(p, xo, ei)
        (out_m * ct_o).sum().backward()
        errs = {k: rel(v.grad, p[k].grad) for k, v in m.named_parameters()}
    tol_dx = TOL_GRAD if kind == "convedge" else TOL_GRAD_FLIP
    assert rel(dx.float(), xo.grad) < tol_dx
    lim = TOL_GRAD if kind == "convedge" else TOL_GRAD_FLIP
    bad = {k: v for k, v in errs.items() if v > (4 * lim if k.startswith("att.") and k.endswith("bias") else lim)}
    assert not bad, bad


# ------------------------------------------------------------------ dynamic kNN rewiring
@pytest.mark.parametrize("N,k,D,Gn", [(8, 4, 512, 40), (9, 4, 1024, 7), (17, 6, 128, 5), (4, 3, 64, 3)])
def test_knn_graph_matches_restated_torch_cluster(N, k, D, Gn):
    gen = torch.Generator().manual_seed(N * 100 + k)
    x = torch.randn(Gn * N, D, generator=gen)
    ei = rpg.knn_graph(x.to(dev()), k, num_nodes_per_graph=N)
    ref = R.knn_graph(x, k, Gn, N)
    assert ei.dtype == torch.int64 and tuple(ei.shape) == (2, Gn * N * k)
    assert torch.equal(ei.cpu(), ref)
    batch = torch.arange(Gn).repeat_interleave(N).to(dev())
    assert torch.equal(rpg.knn_graph(x.to(dev()), k, batch=batch), ei)


def test_knn_graph_ties_and_duplicate_points_go_to_the_lower_index():
    """VERDICT r1 missing 5: equal distances.  Integer coordinates make every squared distance exact in fp32 and fp64
    alike, so ties are real ties on both sides: duplicated points (distance 0), whole graphs of identical points, and
    lattice points at equal distance.  Rule (restated torch_cluster CUDA kernel: candidates scanned in index order,
    strict <): the lower node index first."""
    N, k, D, Gn = 9, 4, 64, 50
    gen = torch.Generator().manual_seed(99)
    x = torch.randint(-2, 3, (Gn * N, D), generator=gen).float()
    x[:, 8:] = 0                                            # 8 informative coordinates in {-2..2}: many equal distances
    x[N:2 * N] = x[N]                                       # graph 1: nine identical points
    x[2 * N + 3] = x[2 * N + 1]                             # graph 2: a duplicated point
    x[3 * N:4 * N, :] = 0
    x[3 * N:4 * N, 0] = torch.arange(N).float()             # graph 3: a line -- left and right neighbour tie
    ei = rpg.knn_graph(x.to(dev()), k, num_nodes_per_graph=N).cpu()
    ref = R.knn_graph(x, k, Gn, N)
    d = ((x[ref[0]] - x[ref[1]]) ** 2).sum(1).view(-1, k)
    assert (d[:, 1:] == d[:, :-1]).any()                    # the case under test: ties inside the k nearest
    assert torch.equal(ei, ref)
    assert ei[0, N * k:N * k + k].tolist() == [N + 1, N + 2, N + 3, N + 4]           # identical points: lowest indices, self skipped
    assert ei[0, (3 * N + 4) * k:(3 * N + 4) * k + k].tolist() == [3 * N + 3, 3 * N + 5, 3 * N + 2, 3 * N + 6]


@pytest.mark.parametrize("sizes,k", [([9, 3, 1, 17, 2, 5], 4), ([2, 2, 7], 1), ([64, 1, 30], 8), ([5, 5, 5, 6], 5)])
def test_knn_graph_on_batches_of_unequal_graph_sizes(sizes, k):
    """A general PyG `batch` vector: graphs of different sizes, graphs with fewer than k + 1 nodes (they contribute all
    their other nodes), single-node graphs (no edges)."""
    gen = torch.Generator().manual_seed(sum(sizes) + k)
    n = sum(sizes)
    x = torch.randn(n, 128, generator=gen)
    x[: sizes[0]] = torch.randint(-1, 2, (sizes[0], 128), generator=gen).float()        # ties in the first graph
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    ei = rpg.knn_graph(x.to(dev()), k, batch=batch.to(dev())).cpu()
    ref = R.knn_graph_batch(x, k, batch)
    assert ei.shape == ref.shape == (2, sum(s * min(k, s - 1) for s in sizes))
    assert torch.equal(ei, ref)
    assert bool((batch[ei[0]] == batch[ei[1]]).all()) and bool((ei[0] != ei[1]).all())
    with pytest.raises(ValueError):
        rpg.knn_graph(x.to(dev()), k, batch=batch.flip(0).to(dev()))                     # unsorted batch vector


def test_stack_with_knn_rewiring_against_oracle():
    """PoseNetX_R2.forward with knn=4 (the CLI default, train.py:377): rewired graph + the same stack, forward and
    backward (gradients against the oracle with the kernel's own activation patterns imposed)."""
    D, N, Gn, k = 256, 8, 6, 4
    q = lambda t: t.bfloat16().double()                                         # noqa: E731
    case = R.synth_stack_case(D, N, Gn, 777, droprate=0.0)
    params = {kk: (q(v) if kk.endswith("weight") and not kk.startswith("fc_") else v.float().double())
              for kk, v in case["params"].items()}
    x = q(case["x"])
    model = rpg.RelPoseGNN(D, D, D, droprate=0.0, knn=k).to(dev())
    model.load_state_dict({kk: v.float() for kk, v in params.items()}, strict=False)
    model.keep_debug_activations = True
    xg = x.float().to(dev()).requires_grad_(True)
    pn, pe, ei_out = model(xg, case["edge_index"].to(dev()))
    ei_ref = R.knn_graph(x.float(), k, Gn, N)
    assert torch.equal(ei_out.cpu(), ei_ref)
    assert tuple(pe.shape) == (Gn * N * k, 6)
    pn_o, pe_o, _, _ = R.stack_forward(params, x, ei_ref, 2, 0.0)
    assert rel(pn, pn_o) < TOL_BF16 and rel(pe, pe_o) < TOL_BF16
    gen = torch.Generator().manual_seed(9)
    ct_n = torch.randn(pn.shape, generator=gen).double()
    ct_e = torch.randn(pe.shape, generator=gen).double()
    ((pn * ct_n.float().to(dev())).sum() + (pe * ct_e.float().to(dev())).sum()).backward()
    dbg = model.debug_activations
    masks = {"e0": (dbg["e0"] > 0).cpu(),
             "rounds": [{"h1": (a["h1"] > 0).cpu(), "h2": (a["h2"] > 0).cpu(), "h3": (a["h3"] > 0).cpu(),
                         "x": (a["out_relu"] > 0).cpu(), "e": (a["e_new_relu"] > 0).cpu()} for a in dbg["rounds"]]}
    p = {kk: v.clone().requires_grad_(True) for kk, v in params.items()}
    xo = x.clone().requires_grad_(True)
    pn_m, pe_m, _, _ = R.stack_forward(p, xo, ei_ref, 2, 0.0, relu_masks=masks)
    ((pn_m * ct_n).sum() + (pe_m * ct_e).sum()).backward()
    assert rel(xg.grad.float(), xo.grad) < 2e-2
    for kk in params:
        assert rel(model.get_parameter(kk).grad, p[kk].grad) < 2e-2, kk


@pytest.mark.parametrize("kept", [[0], [0, 9], [3, 17, 30]])
def test_layer_with_isolated_nodes_and_tiny_templates(kept):
    """Edge dropout can leave nodes without incoming edges (their mean is 0, not the attention bias) and templates so
    small that a 128-row block spans more than 64 nodes (gathers fall back from one-hot panels to the epilogue)."""
    from relpose_gnn_b200.layers import layer_backward_raw, layer_forward_raw
    D, N, Gn = 128, 9, 41
    seed = 5000 + len(kept)
    q = lambda t: t.bfloat16().double()                                         # noqa: E731
    params = {k: (q(v) if k.endswith("weight") else v.float().double())
              for k, v in R.synth_params(R.LAYER_SHAPES(D), seed, torch.float64).items()}
    x = q(R.synth_inputs(Gn, N, D, seed + 1, torch.float64)[0])
    keep = np.zeros(N * (N - 1) // 2, bool)
    keep[kept] = True
    tmpl = R.apply_edge_dropout(R.fc_edge_index(N), keep)
    ei = R.batched_edge_index(tmpl, Gn, N)
    indeg = torch.bincount(tmpl[1], minlength=N)
    assert (indeg == 0).any()                                                    # the case under test
    gen = torch.Generator().manual_seed(seed + 2)
    e = q(torch.relu(torch.randn(ei.size(1), D, generator=gen, dtype=torch.float64)))
    ct_o = q(torch.randn(Gn * N, D, generator=gen, dtype=torch.float64))
    ct_e = q(torch.randn(ei.size(1), D, generator=gen, dtype=torch.float64))
    m = make_layer(D, params)
    graph = G.from_edge_index(ei.to(dev()), Gn * N)
    assert graph.Ep == 2 * len(kept)
    lw = m._packed(dev()).refresh(m)
    acts = layer_forward_raw(lw, graph, x.to(dev()).bfloat16(), e.to(dev()).bfloat16())
    grads = {k: torch.zeros_like(m.get_parameter(k)) for k in PARAM_ORDER}
    dx, de = layer_backward_raw(lw, graph, acts, ct_o.to(dev()).bfloat16(), ct_e.to(dev()).bfloat16(), grads)
    out_o, en_o, mid = R.layer_forward(params, x, ei, e, return_intermediates=True)
    iso = (indeg == 0).repeat(Gn)
    assert float(acts["a"].float().cpu()[iso].abs().max()) == 0.0               # mean over nothing
    assert rel(acts["a"].float(), mid["a"]) < TOL_BF16
    assert rel(acts["out"].float(), out_o) < TOL_BF16 and rel(acts["e_new"].float(), en_o) < TOL_BF16
    masks = {k: (acts[k] > 0).cpu() for k in ("h1", "h2", "h3")}
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xo, eo = x.clone().requires_grad_(True), e.clone().requires_grad_(True)
    out_m, en_m = R.layer_forward(p, xo, ei, eo, relu_masks=masks)
    ((out_m * ct_o).sum() + (en_m * ct_e).sum()).backward()
    assert rel(dx.float(), xo.grad) < TOL_GRAD and rel(de.float(), eo.grad) < TOL_GRAD
    errs = {k: rel(grads[k], p[k].grad) for k in PARAM_ORDER}
    bad = {k: v for k, v in errs.items() if v > (4 * TOL_GRAD if k.startswith("att.") and k.endswith("bias") else TOL_GRAD)}
    assert not bad, bad


def test_seeded_dropout_equals_explicit_masks_forward_and_backward():
    """The in-kernel counter-based feature dropout (production path) against the same masks passed explicitly (the
    path every oracle comparison uses): identical poses and gradients."""
    from relpose_gnn_b200 import ops
    D, N, Gn = 256, 9, 11
    case = R.synth_stack_case(D, N, Gn, 2024, droprate=0.5, edge_dropout=True)
    ei = case["edge_index"].to(dev())

    def run(explicit_seed=None):
        model = rpg.RelPoseGNN(D, D, D, droprate=0.5).to(dev())
        model.load_state_dict({k: v.float() for k, v in case["params"].items()}, strict=False)
        model.dropout_seed = 4711
        x = case["x"].float().to(dev()).requires_grad_(True)
        if explicit_seed is None:
            pn, pe, _ = model(x, ei)
        else:
            kx = ops.dropout_mask(explicit_seed, 0.5, Gn * N, D, dev())
            ke = ops.dropout_mask(explicit_seed + 1, 0.5, ei.size(1), D, dev())
            pn, pe, _ = model(x, ei, keep_x=kx, keep_e=ke)
        gen = torch.Generator().manual_seed(3)
        ct_n = torch.randn(pn.shape, generator=gen).to(dev())
        ct_e = torch.randn(pe.shape, generator=gen).to(dev())
        ((pn * ct_n).sum() + (pe * ct_e).sum()).backward()
        return getattr(model, "last_seed", 0), pn.detach(), pe.detach(), x.grad.detach(), {k: v.grad.detach().clone()
                                                                              for k, v in model.named_parameters()}

    seed_used, pn0, pe0, gx0, g0 = run()
    _, pn1, pe1, gx1, g1 = run(explicit_seed=seed_used)
    # The seeded path drops inside the producing GEMMs and runs the heads as bf16-weight tensor-core GEMMs, the explicit
    # path uses the stand-alone fp32-weight head kernels: same masks, bf16-level differences in the head arithmetic.
    assert rel(pn0, pn1.double().cpu()) < 5e-3 and rel(pe0, pe1.double().cpu()) < 5e-3
    assert rel(gx0, gx1.double().cpu()) < 1e-2
    for k in g0:
        assert rel(g0[k], g1[k].double().cpu()) < 1e-2, k
    # and with the tensor-core heads switched off the two paths agree to fp32 rounding
    def run_plain(explicit_seed=None):
        import relpose_gnn_b200.model as M
        orig = M.RelPoseGNN.__init__
        def patched(self, *a, **kw):
            orig(self, *a, **kw)
            self.tensor_core_heads = False
        M.RelPoseGNN.__init__ = patched
        try:
            return run(explicit_seed)
        finally:
            M.RelPoseGNN.__init__ = orig
    seed_used, pn2, pe2, gx2, g2 = run_plain()
    _, pn3, pe3, gx3, g3 = run_plain(explicit_seed=seed_used)
    assert rel(pn2, pn3.double().cpu()) < 1e-6 and rel(pe2, pe3.double().cpu()) < 1e-6
    assert rel(gx2, gx3.double().cpu()) < 1e-6
    for k in g2:
        assert rel(g2[k], g3[k].double().cpu()) < 1e-5, k


def _full_step(model, crit, x, ei, poses):
    for p in list(model.parameters()) + list(crit.parameters()):
        p.grad = None
    pn, pe, _ = model(x, ei)
    loss, _, _ = crit(pe, poses, ei)
    loss.backward()
    return loss.detach().clone(), torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])


def test_training_step_is_bitwise_deterministic_at_full_size():
    """No atomics anywhere (fixed-order segment sums, split-R folds, head reductions): two runs of the BASELINE-size
    training step (4096 graphs x 9 nodes, D = 512, edge dropout, seeded feature dropout) give identical bits."""
    D, N, Gn = 512, 9, 4096
    torch.manual_seed(0)
    model = rpg.RelPoseGNN(D, D, D, droprate=0.5).to(dev())
    crit = rpg.PoseNetCriterion(0.0, -2.0).to(dev())
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(Gn * N, D, generator=gen).to(dev()).bfloat16()
    poses = (0.1 * torch.randn(Gn * N, 6, generator=gen)).to(dev())
    keep = G.edge_dropout_keep(N * (N - 1) // 2, np.random.RandomState(3))
    ei = G.GraphBatch.fully_connected(Gn, N, dev(), keep).edge_index()
    model.dropout_seed = 99
    l0, g0 = _full_step(model, crit, x, ei, poses)
    model.dropout_seed = 99
    l1, g1 = _full_step(model, crit, x, ei, poses)
    assert torch.equal(l0, l1) and torch.equal(g0, g1)
    assert torch.isfinite(g0).all() and g0.abs().sum() > 0


def test_full_size_batch_equals_mean_of_its_halves():
    """Graphs are independent and the loss is a mean over edges: loss and every gradient of the 4096-graph batch equal
    the average over its two 2048-graph halves (size-independent property at the BASELINE size; droprate 0 because the
    seeded dropout pattern is indexed by the global row)."""
    D, N, Gn = 512, 9, 4096
    torch.manual_seed(0)
    model = rpg.RelPoseGNN(D, D, D, droprate=0.0).to(dev())
    crit = rpg.PoseNetCriterion(0.0, -2.0).to(dev())
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(Gn * N, D, generator=gen).to(dev()).bfloat16()
    poses = (0.1 * torch.randn(Gn * N, 6, generator=gen)).to(dev())
    keep = G.edge_dropout_keep(N * (N - 1) // 2, np.random.RandomState(5))
    ei = G.GraphBatch.fully_connected(Gn, N, dev(), keep).edge_index()
    ei_h = G.GraphBatch.fully_connected(Gn // 2, N, dev(), keep).edge_index()
    l, g = _full_step(model, crit, x, ei, poses)
    h = Gn // 2 * N
    la, ga = _full_step(model, crit, x[:h].contiguous(), ei_h, poses[:h].contiguous())
    lb, gb = _full_step(model, crit, x[h:].contiguous(), ei_h, poses[h:].contiguous())
    assert abs(l.item() - 0.5 * (la.item() + lb.item())) < 1e-5 * max(1.0, abs(l.item()))
    assert rel(g, (0.5 * (ga + gb)).double().cpu()) < 2e-3          # fp32 accumulation order differs between the splits


def test_edge_mask_apply_matches_reference_indexing_and_config_A():
    """(1) train.py:242-247: per-edge tensors indexed by the tiled keep mask; (2) BASELINE config A (one graph, 9 nodes,
    72 edges, D = 512) through the whole stack against the oracle."""
    N, Gn = 9, 13
    keep = R.edge_dropout_keep(N * (N - 1) // 2, np.random.RandomState(11).random_sample(N * (N - 1) // 2))
    tiled = torch.from_numpy(np.tile(np.concatenate([keep, keep]), Gn))
    gen = torch.Generator().manual_seed(4)
    for t in (torch.randn(Gn * N * (N - 1), 6, generator=gen), torch.randn(Gn * N * (N - 1), 64, generator=gen).bfloat16(),
              torch.randint(0, 1000, (Gn * N * (N - 1), 2), generator=gen, dtype=torch.int64)):
        got = rpg.apply_edge_mask(t.to(dev()), keep, Gn)
        assert torch.equal(got.cpu(), t[tiled])
    # the compacted edge_index is what GraphBatch builds from the same mask
    ei_full = R.batched_edge_index(R.fc_edge_index(N), Gn, N)
    ei_kept = rpg.apply_edge_mask(ei_full.t().contiguous().to(dev()), keep, Gn).t().contiguous()
    assert torch.equal(ei_kept.cpu(), R.batched_edge_index(R.apply_edge_dropout(R.fc_edge_index(N), keep), Gn, N))
    # ... and what the device-side builder (rpg_build_edge_index) emits
    assert torch.equal(G.GraphBatch.fully_connected(Gn, N, dev(), keep).edge_index(), ei_kept)
    # config A
    D = 512
    case = R.synth_stack_case(D, 9, 1, 4711, droprate=0.0)
    model = rpg.RelPoseGNN(D, D, D, droprate=0.0).to(dev())
    model.load_state_dict({k: v.float() for k, v in case["params"].items()}, strict=False)
    with torch.no_grad():
        pn, pe, _ = model(case["x"].float().to(dev()), case["edge_index"].to(dev()))
    pn_o, pe_o, _, _ = R.stack_forward(case["params"], case["x"], case["edge_index"], 2, 0.0)
    assert tuple(pe.shape) == (72, 6) and rel(pn, pn_o) < TOL_BF16 and rel(pe, pe_o) < TOL_BF16


@pytest.mark.parametrize("N,k,Gn", [(8, 4, 33), (9, 4, 5), (17, 6, 3)])
def test_per_graph_tables_built_on_device_equal_host_tables(N, k, Gn):
    """GraphBatch.per_graph (kNN fast path: tables by rpg_per_graph_tables, no host round trip) against the general
    host-built one-graph tables, table by table; and the validation counter."""
    gen = torch.Generator().manual_seed(N + k)
    x = torch.randn(Gn * N, 64, generator=gen).to(dev())
    ei = rpg.knn_graph(x, k, num_nodes_per_graph=N)
    fast = G.GraphBatch.per_graph(ei, Gn, N, check=True)
    host = G.GraphBatch(ei[0].cpu().numpy(), ei[1].cpu().numpy(), 1, Gn * N, dev())
    assert (fast.G, fast.N, fast.Ep) == (host.G, host.N, host.Ep)
    for name in ("src", "dst", "in_ptr", "in_idx", "out_ptr", "out_idx", "min_ptr", "min_idx", "max_ptr", "max_idx",
                 "inv_deg", "deg", "has_in"):
        assert torch.equal(fast._tables[name].cpu(), host._tables[name].cpu()), name
    assert torch.equal(fast.edge_index(), ei)
    bad = ei.clone()
    bad[0, 3] = (bad[0, 3] + N) % (Gn * N)           # an edge into another graph
    with pytest.raises(ValueError):
        G.GraphBatch.per_graph(bad, Gn, N, check=True)
