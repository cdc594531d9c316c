"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the RelPose-GNN message-passing hot path.

A plain-PyTorch (CPU, any float dtype; float64 for the yardstick) restatement of the
reference algorithm, written from the reference sources cited on each function.  It is
used by `tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s cpu_baseline /
`--impl reference` legs ONLY.  Nothing in `relpose_gnn_b200/` imports it.

Parity status: PINNED against outputs of the reference's own modules executed in the
build container through `oracle/pyg_shim.py` (fixtures under `tests/golden/`, produced by
`oracle/make_golden.py`).  The reference itself ships no tests or golden vectors for this
path, so those fixtures are the pin.

All functions take parameters as a dict keyed by the reference's ``state_dict`` names
(SURVEY.md section 8b), so a reference checkpoint can be fed in unchanged.
"""
import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# Graph structure
# --------------------------------------------------------------------------------------


def fc_edge_index(n_nodes):
    """Fully connected edge enumeration of one graph, as the reference datasets build it.

    Follows python/niantic/datasets/dataset_7Scenes_multi.py:377-385 (forward half: for
    offset d = 1..N-1, edges s -> s+d for s = 0..N-d-1) and :418-422 (the same list with
    source/destination swapped appended).  dataset_Cambridge_multi.py:240-248,273-278 is
    the identical construction.  Returns int64 [2, N(N-1)].
    """
    src, dst = [], []
    for d in range(1, n_nodes):
        for s in range(0, n_nodes - d):
            src.append(s)
            dst.append(s + d)
    fwd = torch.tensor([src, dst], dtype=torch.long)
    return torch.cat([fwd, fwd.flip(0)], dim=1)


def fc_edge_slot(n_nodes, i, j):
    """Closed form of the slot k of directed edge i -> j."""
    lo, hi = (i, j) if i < j else (j, i)
    d = hi - lo
    off = (d - 1) * (2 * n_nodes - d) // 2
    half = n_nodes * (n_nodes - 1) // 2
    return off + lo + (0 if i < j else half)


def batched_edge_index(template, n_graphs, n_nodes):
    """PyG ``Batch`` concatenation [3p]: graph g gets node offset g*N (train.py:24,132)."""
    offs = (torch.arange(n_graphs, dtype=torch.long) * n_nodes).view(-1, 1, 1)
    return (template.unsqueeze(0) + offs).permute(1, 0, 2).reshape(2, -1)


def edge_dropout_keep(n_undirected, rand_u01, keep_factor=0.5):
    """Surviving-edge mask of the train-time edge dropout, python/niantic/training/train.py:238-242.

    ``rand_u01`` are the ``np.random.random(num_edges)`` draws.  An undirected edge survives
    iff its draw is < keep_factor; if none survive all are kept (:240-241).  The mask is
    shared by both directions of an edge and by every graph of the batch (np.tile, :242).
    Returns bool [n_undirected]; directed rows u and u + n_undirected follow bit u.
    """
    surv = np.asarray(rand_u01[:n_undirected]) < keep_factor
    if surv.sum() == 0:
        surv = np.ones_like(surv)
    return surv.astype(bool)


def apply_edge_dropout(template, keep_undirected):
    """edge_index[:, mask] for one graph (documented intent, train.py:244-245, README.md:26-27)."""
    keep = np.concatenate([keep_undirected, keep_undirected])
    return template[:, torch.from_numpy(keep)]


# --------------------------------------------------------------------------------------
# Layer
# --------------------------------------------------------------------------------------


def _linear(x, w, b):
    return x @ w.t() + b


def _relu(t, mask=None):
    """ReLU; with `mask` (bool, same shape) the activation pattern is imposed instead of recomputed.  Used by the
    gradient parity tests: a reduced-precision forward flips the sign of a few near-zero pre-activations, which
    changes d(relu) on those entries by O(1); imposing the kernel's own pattern isolates the backward arithmetic."""
    return torch.relu(t) if mask is None else t * mask.to(t.dtype)


def edge_model_forward(p, x_src, x_dst, e, prefix="edge_model.edge_mlp.", mask=None):
    """simpleEdgeModel.forward, my_gnn_layer.py:236-239 (ctor :229-234)."""
    h = torch.cat([x_src, x_dst, e], dim=1)
    h = _relu(_linear(h, p[prefix + "0.weight"], p[prefix + "0.bias"]), mask)
    return _linear(h, p[prefix + "2.weight"], p[prefix + "2.bias"])


def attention_block(p, m, prefix="att."):
    """AttentionBlock.forward, att.py:16-34: rank-1 logits phi_i*theta_j, softmax over j."""
    g = _linear(m, p[prefix + "g.weight"], p[prefix + "g.bias"])            # [E, c]
    theta = _linear(m, p[prefix + "theta.weight"], p[prefix + "theta.bias"])
    phi = _linear(m, p[prefix + "phi.weight"], p[prefix + "phi.bias"])
    f = phi.unsqueeze(2) * theta.unsqueeze(1)                              # [E, c, c]  att.py:25
    s = torch.softmax(f, dim=-1)                                           # att.py:26
    y = (s * g.unsqueeze(1)).sum(-1)                                       # att.py:30
    return _linear(y, p[prefix + "W.weight"], p[prefix + "W.bias"]) + m    # att.py:32-33


def scatter_mean(msg, dst, n_rows):
    """PyG aggregate(aggr='mean') -> torch_scatter.scatter(reduce='mean') [3p]."""
    out = msg.new_zeros(n_rows, msg.size(1)).index_add(0, dst, msg)
    cnt = msg.new_zeros(n_rows).index_add(0, dst, msg.new_ones(dst.numel()))
    return out / cnt.clamp(min=1).unsqueeze(1)


def layer_forward(p, x, edge_index, e, return_intermediates=False, relu_masks=None):
    """simpleConvEdge_upt.forward, my_gnn_layer.py:293-311.  Returns (out, e_new), both pre-ReLU.
    relu_masks: optional {'h1','h2','h3'} activation patterns to impose (see _relu)."""
    rm = relu_masks or {}
    row, col = edge_index[0], edge_index[1]
    e_new = edge_model_forward(p, x[row], x[col], e, mask=rm.get("h1"))     # :295-297
    h = torch.cat([x[row], e_new], dim=1)                                   # message, :304-305 (x_j = source)
    h = _relu(_linear(h, p["mlp.0.weight"], p["mlp.0.bias"]), rm.get("h2"))
    m = _linear(h, p["mlp.2.weight"], p["mlp.2.bias"])
    z = attention_block(p, m)                                               # :306
    a = scatter_mean(z, col, x.size(0))                                     # propagate/aggregate, :301
    u = torch.cat([x, a], dim=1)                                            # update, :309-311
    u = _relu(_linear(u, p["mlp_updating.0.weight"], p["mlp_updating.0.bias"]), rm.get("h3"))
    out = _linear(u, p["mlp_updating.2.weight"], p["mlp_updating.2.bias"])
    if return_intermediates:
        return out, e_new, {"m": m, "z": z, "a": a}
    return out, e_new


def knn_graph(x, k, n_graphs, n_nodes):
    """torch_cluster.knn_graph(x, k, batch, loop=False) [3p: torch-cluster 1.5.9, requirements-cu111.txt:6; absent from
    /root/reference -- PARITY UNPINNED for this function: restated from its documented behaviour] as called at
    posenet.py:1043-1050: per graph, edges (neighbour -> centre) to the k nearest other nodes (Euclidean), grouped by
    centre, nearest first (stable: ties to the lower index)."""
    cols = []
    for g in range(n_graphs):
        xg = x[g * n_nodes:(g + 1) * n_nodes].double()
        d = ((xg.unsqueeze(1) - xg.unsqueeze(0)) ** 2).sum(-1)
        d = d + torch.diag(torch.full((n_nodes,), float("inf"), dtype=torch.float64))
        nbr = torch.sort(d, dim=1, stable=True).indices[:, :k]
        centre = torch.arange(n_nodes).view(-1, 1).expand_as(nbr)
        cols.append(torch.stack([nbr.reshape(-1), centre.reshape(-1)], 0) + g * n_nodes)
    return torch.cat(cols, dim=1)


def knn_graph_batch(x, k, batch):
    """knn_graph for a general sorted PyG `batch` vector (graphs of different sizes): a node with fewer than k other
    nodes in its graph gets all of them [3p torch-cluster 1.5.9: missing neighbours are dropped from the result].
    Same ordering and tie rule as `knn_graph`; PARITY UNPINNED like it."""
    cols = []
    batch = batch.long()
    for g in torch.unique(batch).tolist():
        ids = torch.nonzero(batch == g).flatten()
        n = ids.numel()
        kk = min(k, n - 1)
        if kk <= 0:
            continue
        xg = x[ids].double()
        d = ((xg.unsqueeze(1) - xg.unsqueeze(0)) ** 2).sum(-1)
        d = d + torch.diag(torch.full((n,), float("inf"), dtype=torch.float64))
        nbr = torch.sort(d, dim=1, stable=True).indices[:, :kk]
        centre = torch.arange(n).view(-1, 1).expand_as(nbr)
        cols.append(torch.stack([ids[nbr.reshape(-1)], ids[centre.reshape(-1)]], 0))
    return torch.cat(cols, dim=1) if cols else torch.zeros(2, 0, dtype=torch.long)


def conv_edge_forward(p, x, edge_index, e, relu_masks=None):
    """simpleConvEdge.forward, my_gnn_layer.py:253-274: edge model, message = att(mlp(cat[x_i, x_j, e'])) with
    x_i = x[edge_index[1]] (destination), x_j = x[edge_index[0]] (source) [3p PyG], mean over destinations; no update."""
    rm = relu_masks or {}
    row, col = edge_index[0], edge_index[1]
    e_new = edge_model_forward(p, x[row], x[col], e, mask=rm.get("h1"))     # :255-257
    h = torch.cat([x[col], x[row], e_new], dim=1)                           # message, :269
    h = _relu(_linear(h, p["mlp.0.weight"], p["mlp.0.bias"]), rm.get("h2"))
    m = _linear(h, p["mlp.2.weight"], p["mlp.2.bias"])
    z = attention_block(p, m)                                               # :271
    return scatter_mean(z, col, x.size(0)), e_new                           # :262-264


def conv_forward(p, x, edge_index, relu_mask=None):
    """simpleConv.forward, my_gnn_layer.py:394-412: mean over destinations of mlp(cat[x_i, x_j])."""
    row, col = edge_index[0], edge_index[1]
    h = _relu(_linear(torch.cat([x[col], x[row]], dim=1), p["mlp.0.weight"], p["mlp.0.bias"]), relu_mask)
    return scatter_mean(_linear(h, p["mlp.2.weight"], p["mlp.2.bias"]), col, x.size(0))


def CONV_EDGE_SHAPES(D):
    s = {k: v for k, v in LAYER_SHAPES(D).items() if not k.startswith("mlp_updating.")}
    s["mlp.0.weight"] = (D, 3 * D)
    return s


def CONV_SHAPES(D):
    return {"mlp.0.weight": (D, 2 * D), "mlp.0.bias": (D,), "mlp.2.weight": (D, D), "mlp.2.bias": (D,)}


def synth_sibling_case(kind, D, N, G, seed, dtype=torch.float64):
    """Inputs of the `convedge_*` / `conv_*` golden fixtures (oracle/make_golden.py:golden_sibling)."""
    case = synth_layer_case(D, N, G, seed, dtype)
    case["params"] = synth_params(CONV_EDGE_SHAPES(D) if kind == "convedge" else CONV_SHAPES(D), seed, dtype)
    return case


# --------------------------------------------------------------------------------------
# Caller side: the GNN part of PoseNetX_R2.forward
# --------------------------------------------------------------------------------------


def compute_edge_features(x, edge_index):
    """PoseNetX_R2.compute_edge_features, posenet.py:1014-1017: cat[x[min(s,t)], x[max(s,t)]]."""
    lo = torch.minimum(edge_index[0], edge_index[1])
    hi = torch.maximum(edge_index[0], edge_index[1])
    return torch.cat([x[lo], x[hi]], dim=1)


def stack_forward(p, x, edge_index, gnn_recursion=2, droprate=0.0, keep_x=None, keep_e=None, relu_masks=None):
    """The GNN portion of PoseNetX_R2.forward, posenet.py:1053-1091.

    ``p`` holds 'proj_edge.*', 'gnn1.*', 'fc_xyz.*', 'fc_wpqr.*', 'fc_xyz_R.*', 'fc_wpqr_R.*'.
    Feature dropout (posenet.py:1073-1075) cannot be reproduced bit-for-bit from ATen's
    Philox stream, so the Bernoulli keep-masks are explicit inputs: ``keep_x`` [Nn, D] and
    ``keep_e`` [Et, D] with entries in {0,1}; the kept entries are scaled by 1/(1-droprate)
    exactly as F.dropout does.  Returns (pose_nodes [Nn,6], pose_edges [Et,6], x_last, e_last).
    """
    rm = relu_masks or {}
    g = {k[len("gnn1."):]: v for k, v in p.items() if k.startswith("gnn1.")}
    e = _relu(_linear(compute_edge_features(x, edge_index),
                      p["proj_edge.weight"], p["proj_edge.bias"]), rm.get("e0"))   # :1053-1055
    for r in range(gnn_recursion):                                          # :1060-1069 (same gnn1 weights)
        rr = rm.get("rounds", [{}] * gnn_recursion)[r]
        x, e = layer_forward(g, x, edge_index, e, relu_masks=rr)
        x, e = _relu(x, rr.get("x")), _relu(e, rr.get("e"))
    if droprate > 0:                                                        # :1073-1075
        scale = 1.0 / (1.0 - droprate)
        x = x * keep_x.to(x.dtype) * scale
        e = e * keep_e.to(e.dtype) * scale
    pose_nodes = torch.cat([_linear(x, p["fc_xyz.weight"], p["fc_xyz.bias"]),
                            _linear(x, p["fc_wpqr.weight"], p["fc_wpqr.bias"])], 1)       # :1077-1079
    pose_edges = torch.cat([_linear(e, p["fc_xyz_R.weight"], p["fc_xyz_R.bias"]),
                            _linear(e, p["fc_wpqr_R.weight"], p["fc_wpqr_R.bias"])], 1)   # :1085-1086
    return pose_nodes, pose_edges, x, e


def compute_RP(poses, edge_index):
    """PoseNetX_R2.compute_RP, posenet.py:1021-1031: RP[e] = p[src] - p[dst] (vectorised)."""
    return poses[edge_index[0]] - poses[edge_index[1]]


def posenet_criterion(pred, targ, sax, saq):
    """PoseNetCriterion.forward, criterion.py:42-60 with L1 losses (mean reduction)."""
    t_loss = (pred[..., :3] - targ[..., :3]).abs().mean()
    q_loss = (pred[..., 3:] - targ[..., 3:]).abs().mean()
    loss = torch.exp(-sax) * t_loss + sax + torch.exp(-saq) * q_loss + saq
    return loss, t_loss, q_loss


def training_loss(p, x, edge_index, poses, sax, saq, gnn_recursion=2, droprate=0.0,
                  keep_x=None, keep_e=None):
    """What one reference training step differentiates, train.py:256-264 (edge loss only)."""
    _, pose_edges, _, _ = stack_forward(p, x, edge_index, gnn_recursion, droprate, keep_x, keep_e)
    target = compute_RP(poses, edge_index)
    loss, _, _ = posenet_criterion(pose_edges, target, sax, saq)
    return loss


def qexp(v):
    """pose_utils.qexp, pose_utils.py:340-348: unit quaternion [cos|v|, sinc(|v|/pi) * v]."""
    v = np.asarray(v, dtype=np.float64)
    n = np.linalg.norm(v, axis=-1, keepdims=True)
    return np.concatenate([np.cos(n), np.sinc(n / np.pi) * v], axis=-1)


def compose_query_pose(pose_edges, poses_abs, edge_index, ref_node=0):
    """test.py:227-239: first edge into node 0 -> absolute pose of the query + qexp."""
    dst0 = (edge_index[1] == 0).nonzero()[ref_node, 0]
    out = poses_abs[edge_index[0, dst0]] - pose_edges[dst0]
    return np.concatenate([out[:3].numpy(), qexp(out[3:].numpy())])


def pose_errors(pred7, targ7):
    """test.py:202-203, 262-265: translation error ||t_pred - t_gt|| and pose_utils.quaternion_angular_error
    (pose_utils.py:420-431: 2 acos(min(1, |<q1, q2>|)) in degrees) per row of [*, 7] (t, q) poses."""
    pred7, targ7 = np.asarray(pred7, dtype=np.float64), np.asarray(targ7, dtype=np.float64)
    t_err = np.linalg.norm(pred7[:, :3] - targ7[:, :3], axis=1)
    d = np.minimum(1.0, np.abs((pred7[:, 3:] * targ7[:, 3:]).sum(1)))
    return t_err, 2.0 * np.arccos(d) * 180.0 / np.pi


def compose_eval_batch(pose_edges, poses_abs, template, n_graphs, n_nodes, ref_node=0, pose_m=None, pose_s=None):
    """test.py:227-243 for a batch of graphs sharing one edge template (the reference evaluates one graph at a time):
    per graph, the `ref_node`-th template edge into node 0 gives  abs(query) = target[src] - RP_pred;  translations are
    un-normalised with (pose_s, pose_m), log-quaternions go through qexp.  Returns (pred [G, 7], target [G, 7]) fp64."""
    pose_edges = np.asarray(pose_edges, dtype=np.float64)
    poses_abs = np.asarray(poses_abs, dtype=np.float64)
    template = np.asarray(template)
    Ep = template.shape[1]
    k = np.argwhere(template[1] == 0)[ref_node, 0]
    m = np.zeros(3) if pose_m is None else np.asarray(pose_m, dtype=np.float64)
    sc = np.ones(3) if pose_s is None else np.asarray(pose_s, dtype=np.float64)
    pred, targ = np.zeros((n_graphs, 7)), np.zeros((n_graphs, 7))
    for g in range(n_graphs):
        out = poses_abs[g * n_nodes + template[0, k]] - pose_edges[g * Ep + k]
        pred[g] = np.concatenate([out[:3] * sc + m, qexp(out[3:])])
        t = poses_abs[g * n_nodes]
        targ[g] = np.concatenate([t[:3] * sc + m, qexp(t[3:])])
    return pred, targ


# --------------------------------------------------------------------------------------
# Deterministic synthetic parameters / inputs shared by fixtures, tests and the bench
# --------------------------------------------------------------------------------------

LAYER_SHAPES = lambda D: {  # noqa: E731  (state_dict layout)
    "mlp.0.weight": (D, 2 * D), "mlp.0.bias": (D,),
    "mlp.2.weight": (D, D), "mlp.2.bias": (D,),
    "mlp_updating.0.weight": (D, 2 * D), "mlp_updating.0.bias": (D,),
    "mlp_updating.2.weight": (D, D), "mlp_updating.2.bias": (D,),
    "edge_model.edge_mlp.0.weight": (D, 3 * D), "edge_model.edge_mlp.0.bias": (D,),
    "edge_model.edge_mlp.2.weight": (D, D), "edge_model.edge_mlp.2.bias": (D,),
    "att.g.weight": (D // 8, D), "att.g.bias": (D // 8,),
    "att.theta.weight": (D // 8, D), "att.theta.bias": (D // 8,),
    "att.phi.weight": (D // 8, D), "att.phi.bias": (D // 8,),
    "att.W.weight": (D, D // 8), "att.W.bias": (D,),
}


def stack_shapes(D):
    s = {"gnn1." + k: v for k, v in LAYER_SHAPES(D).items()}
    s.update({"proj_edge.weight": (D, 2 * D), "proj_edge.bias": (D,)})
    for h in ("fc_xyz", "fc_wpqr", "fc_xyz_R", "fc_wpqr_R"):
        s[h + ".weight"] = (3, D)
        s[h + ".bias"] = (3,)
    return s


def synth_params(shapes, seed, dtype=torch.float32):
    """Seeded parameters independent of any constructor's RNG consumption order.

    Weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (the scale of torch's default Linear init,
    which is what gnn1's children keep: SURVEY.md section 8a row 1); biases likewise.  Names are
    visited in sorted order with one generator so the result depends only on (shapes, seed).
    """
    gen = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith("weight"):
            fan_in = shp[1]
        else:
            fan_in = shapes[name[:-4] + "weight"][1]
        bound = 1.0 / math.sqrt(fan_in)
        out[name] = ((torch.rand(shp, generator=gen, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return out


def synth_inputs(n_graphs, n_nodes, D, seed, dtype=torch.float32):
    """x ~ N(0,1) [G*N, D]; poses ~ N(0, 0.1) [G*N, 6]."""
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n_graphs * n_nodes, D, generator=gen, dtype=torch.float64).to(dtype)
    poses = (0.1 * torch.randn(n_graphs * n_nodes, 6, generator=gen, dtype=torch.float64)).to(dtype)
    return x, poses


def synth_keep_masks(n_rows_x, n_rows_e, D, seed, droprate=0.5):
    gen = torch.Generator().manual_seed(seed)
    kx = (torch.rand(n_rows_x, D, generator=gen) >= droprate)
    ke = (torch.rand(n_rows_e, D, generator=gen) >= droprate)
    return kx, ke


def synth_layer_case(D, N, G, seed, dtype=torch.float64):
    """Inputs of the `layer_*` golden fixtures (oracle/make_golden.py:golden_layer)."""
    params = synth_params(LAYER_SHAPES(D), seed, dtype)
    x, _ = synth_inputs(G, N, D, seed + 1, dtype)
    gen = torch.Generator().manual_seed(seed + 2)
    ei = batched_edge_index(fc_edge_index(N), G, N)
    e = torch.relu(torch.randn(ei.size(1), D, generator=gen, dtype=torch.float64)).to(dtype)
    ct_out = torch.randn(G * N, D, generator=gen, dtype=torch.float64).to(dtype)
    ct_e = torch.randn(ei.size(1), D, generator=gen, dtype=torch.float64).to(dtype)
    return {"params": params, "x": x, "e": e, "edge_index": ei, "ct_out": ct_out, "ct_e": ct_e}


def synth_stack_case(D, N, G, seed, droprate=0.0, edge_dropout=False, dtype=torch.float64):
    """Inputs of the `stack_*` golden fixtures (oracle/make_golden.py:golden_stack)."""
    params = synth_params(stack_shapes(D), seed, dtype)
    x, poses = synth_inputs(G, N, D, seed + 1, dtype)
    tmpl = fc_edge_index(N)
    keep = None
    if edge_dropout:
        draws = np.random.RandomState(seed + 3).random_sample(N * (N - 1) // 2)
        keep = edge_dropout_keep(N * (N - 1) // 2, draws)
        tmpl = apply_edge_dropout(tmpl, keep)
    ei = batched_edge_index(tmpl, G, N)
    kx, ke = synth_keep_masks(G * N, ei.size(1), D, seed + 4, droprate if droprate > 0 else 0.5)
    return {"params": params, "x": x, "poses": poses, "edge_index": ei, "template": tmpl,
            "edge_keep": keep, "keep_x": kx, "keep_e": ke}


def grad_probe_vectors(shape, seed=7):
    """Seeded probe vectors (u [rows], v [cols]) used to store big weight gradients as u@g and g@v."""
    rs = np.random.RandomState(seed)
    return rs.randn(shape[0]), rs.randn(shape[1])
