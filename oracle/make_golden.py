"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the REAL reference.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Every fixture is produced by executing the reference's own, unmodified first-party code:
  * modules (`my_gnn_layer.py`, `att.py`, `posenet.py`, `criterion.py`) are imported through
    `oracle/pyg_shim.py`;
  * code that only exists inline inside dataset/trainer methods (FC edge enumeration,
    dataset_7Scenes_multi.py:377-385,418-422; edge-dropout mask, train.py:238-242) is read
    from the reference file AT RUN TIME, dedented and exec'd against a stub ``self`` -- the
    source text is never copied into this repository.
The fixtures pin `oracle/restatement.py` (tests/test_oracle_golden.py) and are the
travelling ground truth for the GPU parity tests.
"""
import os
import sys
import textwrap
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyg_shim  # noqa: E402
from oracle import restatement as R  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF = "/root/reference/python/niantic"


def _ref_lines(path, first, last):
    with open(os.path.join(REF, path)) as f:
        lines = f.readlines()[first - 1:last]
    return textwrap.dedent("".join(lines))


def golden_fc_enumeration():
    """exec dataset_7Scenes_multi.py:377-385 + 418-422 for several N."""
    body = _ref_lines("datasets/dataset_7Scenes_multi.py", 378, 385)
    flip = _ref_lines("datasets/dataset_7Scenes_multi.py", 421, 422)
    out = {}
    for n in (2, 3, 4, 8, 9, 17):
        env = {"torch": torch, "self": types.SimpleNamespace(seq_len=n)}
        exec(body, env)
        exec(flip, env)
        out[f"fc_N{n}"] = env["edge_index"].numpy().astype(np.int64)
    np.savez(os.path.join(OUT, "fc_enumeration.npz"), **out)
    print("fc_enumeration:", {k: v.shape for k, v in out.items()})


def golden_edge_dropout():
    """exec train.py:238-242 with seeded numpy RNG; store draws + resulting tiled mask."""
    body = _ref_lines("training/train.py", 238, 242)
    out = {}
    for case, (n, batch, seed) in enumerate([(8, 8, 0), (9, 4, 1), (17, 2, 2), (3, 2, 11)]):
        n_edges = n * (n - 1) * batch
        rs = np.random.RandomState(seed)
        draws = rs.random_sample(n * (n - 1) // 2)
        np.random.seed(seed)  # the reference calls the global np.random.random
        env = {"np": np, "int": int,
               "self": types.SimpleNamespace(batch_size=batch, edge_keep_factor=0.5),
               "data": types.SimpleNamespace(edge_index=np.zeros((2, n_edges)))}
        exec(body, env)
        out[f"case{case}_meta"] = np.array([n, batch, seed])
        out[f"case{case}_draws"] = draws
        out[f"case{case}_tiled"] = np.asarray(env["surviving_edges"]).astype(np.int64)
    # the "nothing survives" branch (:240-241), forced with keep factor 0
    env = {"np": np, "int": int,
           "self": types.SimpleNamespace(batch_size=2, edge_keep_factor=0.0),
           "data": types.SimpleNamespace(edge_index=np.zeros((2, 12 * 2)))}
    np.random.seed(5)
    exec(body, env)
    out["none_survive_tiled"] = np.asarray(env["surviving_edges"]).astype(np.int64)
    np.savez(os.path.join(OUT, "edge_dropout.npz"), **out)
    print("edge_dropout: cases", len(out))


def _load_into(module, params):
    sd = {k: v.clone() for k, v in params.items()}
    missing = module.load_state_dict(sd, strict=True)
    return missing


def golden_layer(tag, D, N, G, seed, store_params):
    """simpleConvEdge_upt forward + backward on the FC batch; float64 reference run + float32 run."""
    gnn, _, _ = pyg_shim.import_reference()
    case = R.synth_layer_case(D, N, G, seed)
    params64, x, e, ei = case["params"], case["x"], case["e"], case["edge_index"]
    ct_out, ct_e = case["ct_out"], case["ct_e"]

    out = {}
    for dt, name in ((torch.float64, "f64"), (torch.float32, "f32")):
        layer = gnn.simpleConvEdge_upt(D, D, D).to(dt)
        _load_into(layer, {k: v.to(dt) for k, v in params64.items()})
        xd = x.to(dt).clone().requires_grad_(True)
        ed = e.to(dt).clone().requires_grad_(True)
        o, en = layer(xd, ei, ed)
        ((o * ct_out.to(dt)).sum() + (en * ct_e.to(dt)).sum()).backward()
        out[f"out_{name}"] = o.detach().numpy()
        out[f"e_new_{name}"] = en.detach().numpy()
        if name == "f64":
            out["dx"] = xd.grad.numpy()
            out["de"] = ed.grad.numpy()
            for k, p in layer.named_parameters():
                out["grad." + k] = p.grad.numpy()
    out["meta"] = np.array([D, N, G, seed])
    out["edge_index"] = ei.numpy()
    for k, v in params64.items():       # checksum only; regenerate with synth_params(seed)
        out["paramsum." + k] = np.array([v.sum().item(), v.abs().sum().item()])
    # x, e, cotangents and params are regenerated from the seed by R.synth_layer_case
    if not store_params:
        for k in [k for k in out if k.startswith("grad.") and k.endswith("weight")]:
            g = out.pop(k)              # keep fixture small: two seeded random projections of big grads
            u, v = R.grad_probe_vectors(g.shape)
            out["gradrows." + k[5:]] = g @ v
            out["gradcols." + k[5:]] = u @ g
        for k in ("out_f32", "e_new_f32", "de", "e_new_f64"):
            out[k] = out[k].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, f"layer_{tag}.npz"), **out)
    print(f"layer_{tag}: out {out['out_f64'].shape} e_new {out['e_new_f64'].shape}")


def golden_sibling(kind, tag, D, N, G, seed):
    """simpleConvEdge / simpleConv (my_gnn_layer.py:242-274, 394-412) forward + backward, float64 reference run."""
    gnn, _, _ = pyg_shim.import_reference()
    case = R.synth_sibling_case(kind, D, N, G, seed)
    params, x, e, ei = case["params"], case["x"], case["e"], case["edge_index"]
    layer = (gnn.simpleConvEdge(D, D, D) if kind == "convedge" else gnn.simpleConv(D, D)).double()
    _load_into(layer, params)
    xd = x.clone().requires_grad_(True)
    out = {}
    if kind == "convedge":
        ed = e.clone().requires_grad_(True)
        o, en = layer(xd, ei, ed)
        ((o * case["ct_out"]).sum() + (en * case["ct_e"]).sum()).backward()
        out["e_new_f64"] = en.detach().numpy()
        out["de"] = ed.grad.numpy()
    else:
        o = layer(xd, ei)
        (o * case["ct_out"]).sum().backward()
    out["out_f64"] = o.detach().numpy()
    out["dx"] = xd.grad.numpy()
    for k, p in layer.named_parameters():
        out["grad." + k] = p.grad.numpy()
    out["meta"] = np.array([D, N, G, seed])
    np.savez_compressed(os.path.join(OUT, f"{kind}_{tag}.npz"), **out)
    print(f"{kind}_{tag}: out {out['out_f64'].shape}")


class _StubFE(torch.nn.Module):
    """Stands in for torchvision resnet34 (train.py:173): identity features, NOT the target path."""

    def __init__(self, d):
        super().__init__()
        self.fc = torch.nn.Linear(d, d)
        self.avgpool = torch.nn.Identity()
        self._feats = None

    def forward(self, _img):
        return self._feats


def golden_stack(tag, D, N, G, seed, droprate, edge_dropout):
    """GNN part of PoseNetX_R2.forward + reference compute_RP + PoseNetCriterion, fwd and bwd."""
    _, _, posenet = pyg_shim.import_reference()
    sys.modules.setdefault("transforms3d", types.ModuleType("transforms3d"))
    for sub in ("euler", "quaternions"):
        m = types.ModuleType("transforms3d." + sub)
        sys.modules.setdefault("transforms3d." + sub, m)
        setattr(sys.modules["transforms3d"], sub, sys.modules["transforms3d." + sub])
    from niantic.modules import criterion as ref_criterion

    dt = torch.float64
    case = R.synth_stack_case(D, N, G, seed, droprate, edge_dropout, dt)
    params, x, poses, ei = case["params"], case["x"], case["poses"], case["edge_index"]
    kx, ke = case["keep_x"], case["keep_e"]
    out = {}
    if edge_dropout:
        out["edge_keep"] = case["edge_keep"]

    fe = _StubFE(D)
    model = posenet.PoseNetX_R2(fe, droprate=droprate, pretrained=False, feat_dim=D,
                                edge_feat_dim=D, node_dim=D, use_gnn=True, knn=-1,
                                gnn_recursion=2, device="cpu").to(dt)
    sd = model.state_dict()
    for k, v in params.items():
        assert sd[k].shape == v.shape, k
        sd[k] = v.clone()
    model.load_state_dict(sd)
    model.train()
    xin = x.clone().requires_grad_(True)
    fe._feats = xin

    # F.dropout's RNG stream cannot travel; swap in explicit masks with F.dropout's scaling.
    queue = [kx.to(dt), ke.to(dt)]
    real_dropout = posenet.F.dropout
    posenet.F.dropout = lambda t, p=0.5, **kw: t * queue.pop(0) / (1.0 - p)
    try:
        data = types.SimpleNamespace(x=torch.zeros(G * N, 3 * model.input_img_height * 4, dtype=dt),
                                     edge_index=ei, edge_attr=None, batch=None)
        pose_n, pose_e, ei_out = model(data)
    finally:
        posenet.F.dropout = real_dropout
    assert torch.equal(ei_out, ei)
    with torch.no_grad():   # on CPU `.to(device)` hands back the leaf itself (posenet.py:1023), so the
        target_R = model.compute_RP(poses, ei).to(dt)   # reference's in-place loop needs no_grad here
    crit = ref_criterion.PoseNetCriterion(sax=0.0, saq=-2.0, learn_beta=True).to(dt)   # train.py:68-69,198-199
    loss, t_loss, q_loss = crit(pose_e.view(1, -1, 6), target_R.view(1, -1, 6))
    loss.backward()

    out.update({"meta": np.array([D, N, G, seed]), "droprate": np.array(droprate),
                "edge_index": ei.numpy(),
                "pose_nodes": pose_n.detach().numpy(), "pose_edges": pose_e.detach().numpy(),
                "target_R": target_R.detach().numpy(),
                "loss": np.array([loss.item(), t_loss.item(), q_loss.item()]),
                "dx": xin.grad.numpy(),
                "dsax": crit.sax.grad.numpy(), "dsaq": crit.saq.grad.numpy()})
    named = dict(model.named_parameters())
    for k, v in params.items():
        out["paramsum." + k] = np.array([v.sum().item(), v.abs().sum().item()])
        g = named[k].grad
        out["grad." + k] = (g if g is not None else torch.zeros_like(v)).numpy()
    np.savez_compressed(os.path.join(OUT, f"stack_{tag}.npz"), **out)
    print(f"stack_{tag}: loss {loss.item():.6f} pose_edges {tuple(pose_e.shape)}")


def golden_qexp():
    sys.modules.setdefault("transforms3d", types.ModuleType("transforms3d"))
    for sub in ("euler", "quaternions"):
        sys.modules.setdefault("transforms3d." + sub, types.ModuleType("transforms3d." + sub))
        setattr(sys.modules["transforms3d"], sub, sys.modules["transforms3d." + sub])
    pyg_shim.install()
    from niantic.utils import pose_utils
    rs = np.random.RandomState(0)
    v = np.concatenate([rs.randn(15, 3) * 0.7, np.zeros((1, 3))])
    q = np.stack([pose_utils.qexp(t) for t in v])
    np.savez(os.path.join(OUT, "qexp.npz"), v=v, q=q)
    print("qexp:", q.shape)


def golden_eval_compose():
    """exec test.py:227-243 (reference absolute pose from the first edge into node 0, qexp, un-normalisation) on seeded
    single-graph cases (the reference evaluates one graph per batch)."""
    sys.modules.setdefault("transforms3d", types.ModuleType("transforms3d"))
    for sub in ("euler", "quaternions"):
        sys.modules.setdefault("transforms3d." + sub, types.ModuleType("transforms3d." + sub))
        setattr(sys.modules["transforms3d"], sub, sys.modules["transforms3d." + sub])
    pyg_shim.install()
    from niantic.utils import pose_utils
    body = _ref_lines("testing/test.py", 227, 243)
    fc = np.load(os.path.join(OUT, "fc_enumeration.npz"))
    out = {}
    for case, (n, ref_node, seed) in enumerate([(8, 0, 0), (9, 0, 1), (9, 3, 2), (17, 5, 3), (4, 2, 4)]):
        rs = np.random.RandomState(100 + seed)
        edges = fc[f"fc_N{n}"]
        output_R = (rs.randn(edges.shape[1], 6) * 0.3).astype(np.float32)
        target = (rs.randn(n, 6) * 0.5).astype(np.float32)
        pose_m = rs.randn(3).astype(np.float32)
        pose_s = (rs.rand(3) + 0.5).astype(np.float32)
        env = {"np": np, "qexp": pose_utils.qexp, "edges": edges, "output_R": output_R.copy(), "target": target.copy(),
               "ref_node": ref_node, "self": types.SimpleNamespace(pose_m=pose_m, pose_s=pose_s)}
        exec(body, env)
        out[f"case{case}_meta"] = np.array([n, ref_node, seed])
        out[f"case{case}_output_R"] = output_R
        out[f"case{case}_target"] = target
        out[f"case{case}_pose_m"] = pose_m
        out[f"case{case}_pose_s"] = pose_s
        out[f"case{case}_pred7"] = np.asarray(env["output"][0], dtype=np.float64)
        out[f"case{case}_targ7"] = np.asarray(env["target"], dtype=np.float64)      # all nodes; the reference keeps row 0
    # error metrics of test.py:202-203, 262-265: ||t_pred - t_gt|| and pose_utils.quaternion_angular_error
    rs = np.random.RandomState(77)
    pred = rs.randn(24, 7)
    targ = rs.randn(24, 7)
    pred[:, 3:] /= np.linalg.norm(pred[:, 3:], axis=1, keepdims=True)
    targ[:, 3:] /= np.linalg.norm(targ[:, 3:], axis=1, keepdims=True)
    targ[0, 3:] = pred[0, 3:]                       # identical rotation -> 0 degrees (|d| clamps to 1)
    targ[1, 3:] = -pred[1, 3:]                      # antipodal quaternion = same rotation
    out["err_pred"] = pred
    out["err_targ"] = targ
    out["err_t"] = np.asarray([np.linalg.norm(p - t) for p, t in zip(pred[:, :3], targ[:, :3])])
    out["err_q"] = np.asarray([pose_utils.quaternion_angular_error(p, t) for p, t in zip(pred[:, 3:], targ[:, 3:])])
    np.savez(os.path.join(OUT, "eval_compose.npz"), **out)
    print("eval_compose: cases", sum(k.endswith("_meta") for k in out))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    golden_fc_enumeration()
    golden_edge_dropout()
    golden_qexp()
    golden_eval_compose()
    golden_sibling("convedge", "D128_N9_G2", 128, 9, 2, 200)
    golden_sibling("convedge", "D128_N5_G3", 128, 5, 3, 201)
    golden_sibling("conv", "D128_N9_G2", 128, 9, 2, 210)
    golden_sibling("conv", "D128_N4_G3", 128, 4, 3, 211)
    golden_layer("D128_N9_G2", 128, 9, 2, 100, store_params=True)
    golden_layer("D128_N4_G3", 128, 4, 3, 101, store_params=True)
    golden_layer("D512_N8_G2", 512, 8, 2, 102, store_params=False)
    golden_layer("D512_N17_G1", 512, 17, 1, 103, store_params=False)
    golden_stack("D128_N9_G2", 128, 9, 2, 200, droprate=0.0, edge_dropout=False)
    golden_stack("D128_N8_G3_drop", 128, 8, 3, 201, droprate=0.5, edge_dropout=True)


if __name__ == "__main__":
    main()
