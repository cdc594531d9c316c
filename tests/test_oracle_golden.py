"""Pins oracle/restatement.py against fixtures produced by the REAL reference
(oracle/make_golden.py, run in the build container through oracle/pyg_shim.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import restatement as R

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def check_paramsums(fx, params, prefix="paramsum."):
    for k, v in params.items():
        s = fx[prefix + k]
        assert abs(v.sum().item() - s[0]) <= 1e-9 * max(1, abs(s[0])), k
        assert abs(v.abs().sum().item() - s[1]) <= 1e-9 * max(1, abs(s[1])), k


@pytest.mark.parametrize("n", [2, 3, 4, 8, 9, 17])
def test_fc_enumeration_matches_reference(n):
    fx = np.load(os.path.join(G, "fc_enumeration.npz"))
    ei = R.fc_edge_index(n)
    assert np.array_equal(ei.numpy(), fx[f"fc_N{n}"])
    # closed form of SURVEY Appendix D
    for k in range(ei.size(1)):
        assert R.fc_edge_slot(n, int(ei[0, k]), int(ei[1, k])) == k
    # the reference's commented invariant (dataset_7Scenes_multi.py:386-392): forward half == combinations
    import itertools
    half = ei[:, : n * (n - 1) // 2].t().tolist()
    assert sorted(map(tuple, half)) == sorted(itertools.combinations(range(n), 2))


def test_edge_dropout_mask_matches_reference():
    fx = np.load(os.path.join(G, "edge_dropout.npz"))
    for case in range(4):
        n, batch, _ = fx[f"case{case}_meta"]
        keep = R.edge_dropout_keep(n * (n - 1) // 2, fx[f"case{case}_draws"])
        tiled = np.tile(keep, 2 * batch).astype(np.int64)
        assert np.array_equal(tiled, fx[f"case{case}_tiled"])
    keep = R.edge_dropout_keep(6, np.ones(6), keep_factor=0.0)
    assert keep.all() and fx["none_survive_tiled"].all()


@pytest.mark.parametrize("kind,tag", [("convedge", "D128_N9_G2"), ("convedge", "D128_N5_G3"),
                                      ("conv", "D128_N9_G2"), ("conv", "D128_N4_G3")])
def test_sibling_layers_match_reference(kind, tag):
    """simpleConvEdge / simpleConv (my_gnn_layer.py:242-274, 394-412): oracle restatement vs the real reference."""
    fx = np.load(os.path.join(G, f"{kind}_{tag}.npz"))
    D, N, Gn, seed = [int(v) for v in fx["meta"]]
    case = R.synth_sibling_case(kind, D, N, Gn, seed)
    p = {k: v.clone().requires_grad_(True) for k, v in case["params"].items()}
    x = case["x"].clone().requires_grad_(True)
    if kind == "convedge":
        e = case["e"].clone().requires_grad_(True)
        out, e_new = R.conv_edge_forward(p, x, case["edge_index"], e)
        ((out * case["ct_out"]).sum() + (e_new * case["ct_e"]).sum()).backward()
        assert np.allclose(e_new.detach().numpy(), fx["e_new_f64"], rtol=1e-10, atol=1e-12)
        assert np.allclose(e.grad.numpy(), fx["de"], rtol=1e-9, atol=1e-11)
    else:
        out = R.conv_forward(p, x, case["edge_index"])
        (out * case["ct_out"]).sum().backward()
    assert np.allclose(out.detach().numpy(), fx["out_f64"], rtol=1e-10, atol=1e-12)
    assert np.allclose(x.grad.numpy(), fx["dx"], rtol=1e-9, atol=1e-11)
    for k, v in p.items():
        assert np.allclose(v.grad.numpy(), fx["grad." + k], rtol=1e-8, atol=1e-10), k


def test_eval_composition_matches_reference():
    """test.py:227-243 executed on seeded single graphs (oracle/make_golden.py: golden_eval_compose)."""
    fx = np.load(os.path.join(G, "eval_compose.npz"))
    fc = np.load(os.path.join(G, "fc_enumeration.npz"))
    for case in range(5):
        n, ref_node, _ = [int(v) for v in fx[f"case{case}_meta"]]
        pred, targ = R.compose_eval_batch(fx[f"case{case}_output_R"], fx[f"case{case}_target"], fc[f"fc_N{n}"], 1, n,
                                          ref_node, fx[f"case{case}_pose_m"], fx[f"case{case}_pose_s"])
        assert np.allclose(pred[0], fx[f"case{case}_pred7"], rtol=0, atol=2e-6)     # the reference computes in fp32
        assert np.allclose(targ[0], fx[f"case{case}_targ7"][0], rtol=0, atol=2e-6)


def test_pose_errors_match_reference():
    fx = np.load(os.path.join(G, "eval_compose.npz"))
    t_err, q_err = R.pose_errors(fx["err_pred"], fx["err_targ"])
    assert np.allclose(t_err, fx["err_t"], atol=1e-12) and np.allclose(q_err, fx["err_q"], atol=1e-9)
    assert q_err[0] < 1e-5 and q_err[1] < 1e-5


def test_qexp_matches_reference():
    fx = np.load(os.path.join(G, "qexp.npz"))
    assert np.allclose(R.qexp(fx["v"]), fx["q"], atol=1e-14)
    assert np.allclose(np.linalg.norm(R.qexp(fx["v"]), axis=-1), 1.0)


@pytest.mark.parametrize("tag", ["D128_N9_G2", "D128_N4_G3", "D512_N8_G2", "D512_N17_G1"])
def test_layer_matches_reference(tag):
    fx = np.load(os.path.join(G, f"layer_{tag}.npz"))
    D, N, Gn, seed = [int(v) for v in fx["meta"]]
    case = R.synth_layer_case(D, N, Gn, seed)
    check_paramsums(fx, case["params"])
    assert np.array_equal(case["edge_index"].numpy(), fx["edge_index"])
    p = {k: v.clone().requires_grad_(True) for k, v in case["params"].items()}
    x = case["x"].clone().requires_grad_(True)
    e = case["e"].clone().requires_grad_(True)
    out, e_new = R.layer_forward(p, x, case["edge_index"], e)
    assert rel(out.detach(), fx["out_f64"]) < 1e-12
    assert rel(e_new.detach(), fx["e_new_f64"]) < (1e-12 if fx["e_new_f64"].dtype == np.float64 else 1e-6)
    assert rel(out.detach(), fx["out_f32"]) < 1e-5        # the reference run in its native fp32
    ((out * case["ct_out"]).sum() + (e_new * case["ct_e"]).sum()).backward()
    assert rel(x.grad, fx["dx"]) < 1e-11
    assert rel(e.grad, fx["de"]) < (1e-11 if fx["de"].dtype == np.float64 else 1e-6)
    for k, v in p.items():
        if "grad." + k in fx.files:
            assert rel(v.grad, fx["grad." + k]) < 1e-11, k
        else:
            u, w = R.grad_probe_vectors(v.shape)
            assert rel(v.grad.numpy() @ w, fx["gradrows." + k]) < 1e-10, k
            assert rel(u @ v.grad.numpy(), fx["gradcols." + k]) < 1e-10, k


@pytest.mark.parametrize("tag,droprate,edrop", [("D128_N9_G2", 0.0, False), ("D128_N8_G3_drop", 0.5, True)])
def test_stack_matches_reference(tag, droprate, edrop):
    fx = np.load(os.path.join(G, f"stack_{tag}.npz"))
    D, N, Gn, seed = [int(v) for v in fx["meta"]]
    case = R.synth_stack_case(D, N, Gn, seed, droprate, edrop)
    check_paramsums(fx, case["params"])
    assert np.array_equal(case["edge_index"].numpy(), fx["edge_index"])
    p = {k: v.clone().requires_grad_(True) for k, v in case["params"].items()}
    x = case["x"].clone().requires_grad_(True)
    sax = torch.zeros(1, dtype=torch.float64, requires_grad=True)
    saq = torch.full((1,), -2.0, dtype=torch.float64, requires_grad=True)
    pn, pe, _, _ = R.stack_forward(p, x, case["edge_index"], 2, droprate, case["keep_x"], case["keep_e"])
    assert rel(pn.detach(), fx["pose_nodes"]) < 1e-12
    assert rel(pe.detach(), fx["pose_edges"]) < 1e-12
    target = R.compute_RP(case["poses"], case["edge_index"])
    # the reference builds RP in torch's default fp32 (posenet.py:1023), hence 1e-7 and not 1e-14
    assert rel(target, fx["target_R"]) < 1e-6
    loss, t_loss, q_loss = R.posenet_criterion(pe, target, sax, saq)
    assert np.allclose([loss.item(), t_loss.item(), q_loss.item()], fx["loss"], rtol=1e-7)
    loss.backward()
    assert rel(x.grad, fx["dx"]) < 1e-10
    assert np.allclose(sax.grad.numpy(), fx["dsax"], rtol=1e-10)
    assert np.allclose(saq.grad.numpy(), fx["dsaq"], rtol=1e-10)
    for k, v in p.items():
        g = v.grad if v.grad is not None else torch.zeros_like(v)
        ref = fx["grad." + k]
        if np.abs(ref).max() == 0:
            assert g.abs().max().item() == 0, k      # node heads / last node update get no gradient
        else:
            assert rel(g, ref) < 1e-10, k
