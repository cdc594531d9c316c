"""Full-size parity (`-m gpu`): BASELINE configs C (4096 x 9 with an edge-dropout mask), D (2048 x 17) and E (8192 x 9 shard
and the whole 65536 x 9 batch) at their real batch sizes, through the REAL drop-in boundary -- a fresh, un-annotated int64
`edge_index` as the PyG loader hands it over (train.py:229-256), seeded in-kernel feature dropout, tensor-core heads --
against the CPU oracle on a sample of graphs.

The gathered node terms of the edge GEMMs are one-hot K panels whose selection tile depends on
`((row0 mod Ep) / gcd(128, Ep))` of the 128-row block (rpg_gemm_t.gsel): the sample is chosen so that the row blocks of
the sampled graphs hit EVERY tile index (asserted), plus the first / last graph and a few random ones.  Graphs are
independent, so the oracle runs the sampled graphs as a small batch with the same template:
  forward : node + edge poses within 2e-2 (bf16 mode, BASELINE.json north_star);
  backward: d(loss)/dx of the sampled graphs for random cotangents on both pose outputs, against the oracle with the
            kernel's own ReLU patterns imposed (see tests/test_gpu_parity.py: TOL_GRAD), within 2e-2.
"""
import math

import numpy as np
import pytest
import torch

import relpose_gnn_b200 as rpg
from oracle import restatement as R
from relpose_gnn_b200 import graph as G, ops

pytestmark = pytest.mark.gpu
TOL = 2e-2


def dev():
    return torch.device("cuda:0")


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp(min=1e-30)).item()


def sample_graphs(Gn, Ep, seed, extra=6):
    """Graph ids whose 128-row blocks cover every selection-tile index, + first / last + `extra` random ones."""
    g = math.gcd(128, Ep)
    npat = Ep // g
    period = (128 // g)                                  # graphs per lcm(128, Ep) rows
    rs = np.random.RandomState(seed)
    start = int(rs.randint(1, max(2, Gn - period - 1)))
    ids = set(range(start, min(Gn, start + period + 1))) | {0, Gn - 1} | set(int(v) for v in rs.randint(0, Gn, extra))
    ids = sorted(ids)
    hit = set()
    for gi in ids:
        for blk in range((gi * Ep) // 128, ((gi + 1) * Ep - 1) // 128 + 1):
            hit.add(((blk * 128) % Ep) // g)
    return ids, hit, npat


def q16(t):
    return t.bfloat16().double()


def run_config(Gn, N, drop_edges, train, seed, D=512):
    torch.manual_seed(seed)
    params = {k: (q16(v) if k.endswith("weight") else v.float().double())
              for k, v in R.synth_params(R.stack_shapes(D), seed, torch.float64).items()}
    model = rpg.RelPoseGNN(D, D, D, droprate=0.5).to(dev())
    model.load_state_dict({k: v.float() for k, v in params.items()}, strict=False)
    model.keep_debug_activations = train
    tmpl = R.fc_edge_index(N)
    H = N * (N - 1) // 2
    if drop_edges:
        keep = R.edge_dropout_keep(H, np.random.RandomState(seed + 1).random_sample(H))
        tmpl = R.apply_edge_dropout(tmpl, keep)
    Ep = tmpl.size(1)
    # the fresh boundary: a plain int64 tensor, no GraphBatch attached
    ei = R.batched_edge_index(tmpl, Gn, N).to(dev())
    assert getattr(ei, "rpg_graph", None) is None
    gen = torch.Generator(device="cuda").manual_seed(seed + 2)
    x = torch.randn(Gn * N, D, device=dev(), generator=gen).bfloat16()
    ids, hit, npat = sample_graphs(Gn, Ep, seed + 3)
    assert hit == set(range(npat)), (sorted(hit), npat)          # every one-hot selection tile is exercised
    idt = torch.tensor(ids, device=dev())
    node_rows = (idt.view(-1, 1) * N + torch.arange(N, device=dev())).reshape(-1)
    edge_rows = (idt.view(-1, 1) * Ep + torch.arange(Ep, device=dev())).reshape(-1)

    model.dropout_seed = 1000 + seed
    if train:
        xg = x.clone().requires_grad_(True)
        pn, pe, ei_out = model(xg, ei)
    else:
        with torch.no_grad():
            pn, pe, ei_out = model(x, ei)
    assert ei_out is ei
    gr = ei.rpg_graph
    assert (gr.G, gr.N, gr.Ep) == (Gn, N, Ep)
    assert gr.struct.sel_src, "the benchmark configs must run on the one-hot panel path"
    assert torch.isfinite(pn).all() and torch.isfinite(pe).all()

    # the seeded keep decisions of the sampled rows, materialised for the oracle
    seed_used = model.last_seed
    kx = ops.dropout_mask(seed_used, 0.5, Gn * N, D, dev())[node_rows].cpu().bool()
    ke = ops.dropout_mask(seed_used + 1, 0.5, Gn * Ep, D, dev())[edge_rows].cpu().bool()
    xs = x[node_rows].double().cpu()
    ei_s = R.batched_edge_index(tmpl, len(ids), N)
    pn_o, pe_o, _, _ = R.stack_forward(params, xs, ei_s, 2, 0.5, kx, ke)
    assert rel(pn[node_rows], pn_o) < TOL, rel(pn[node_rows], pn_o)
    assert rel(pe[edge_rows], pe_o) < TOL, rel(pe[edge_rows], pe_o)
    if not train:
        return

    ct_n = torch.randn(pn.shape, device=dev(), generator=gen)
    ct_e = torch.randn(pe.shape, device=dev(), generator=gen)
    ((pn * ct_n).sum() + (pe * ct_e).sum()).backward()
    assert torch.isfinite(xg.grad.float()).all()
    dbg = model.debug_activations
    masks = {"e0": (dbg["e0"][edge_rows] > 0).cpu(),
             "rounds": [{"h1": (a["h1"][edge_rows] > 0).cpu(), "h2": (a["h2"][edge_rows] > 0).cpu(),
                         "h3": (a["h3"][node_rows] > 0).cpu(), "x": (a["out_relu"][node_rows] > 0).cpu(),
                         "e": (a["e_new_relu"][edge_rows] > 0).cpu()} for a in dbg["rounds"]]}
    xo = xs.clone().requires_grad_(True)
    pn_m, pe_m, _, _ = R.stack_forward(params, xo, ei_s, 2, 0.5, kx, ke, relu_masks=masks)
    ((pn_m * ct_n[node_rows].double().cpu()).sum() + (pe_m * ct_e[edge_rows].double().cpu()).sum()).backward()
    err = rel(xg.grad[node_rows].float(), xo.grad)
    assert err < TOL, err
    model.debug_activations = None


@pytest.mark.parametrize("name,Gn,N,drop_edges,train", [
    ("C_train_4096x9_edge_dropout", 4096, 9, True, True),
    ("C_train_4096x9_full_template", 4096, 9, False, True),
    ("D_train_2048x17", 2048, 17, False, True),
    ("D_train_2048x17_edge_dropout", 2048, 17, True, True),
    ("E_infer_8192x9_shard", 8192, 9, False, False),
    ("E_infer_65536x9_whole", 65536, 9, False, False),
])
def test_full_size_config_against_oracle_through_the_fresh_boundary(name, Gn, N, drop_edges, train):
    run_config(Gn, N, drop_edges, train, seed=sum(map(ord, name)) % 1000)


def test_fresh_boundary_async_validation_reports_a_bad_batch_one_call_later():
    """rpg.set_validation('async'): no read-back in the call itself; a batch that violates the template property raises
    at the next call / check_pending()."""
    D, N, Gn = 128, 5, 40
    params = R.synth_params(R.LAYER_SHAPES(D), 1, torch.float32)
    m = rpg.simpleConvEdge_upt(D, D, D)
    m.load_state_dict(params)
    m = m.to(dev())
    tmpl = R.fc_edge_index(N)
    ei = R.batched_edge_index(tmpl, Gn, N).to(dev())
    x = torch.randn(Gn * N, D, device=dev())
    e = torch.randn(ei.size(1), D, device=dev()).relu()
    out0, _ = m(x, ei, e)                              # cold call: infers (G, N) and caches the batch shape
    prev = rpg.set_validation("async")
    try:
        ei2 = ei.clone()                               # a fresh tensor of the same batch: device path, nothing read back
        out1, _ = m(x, ei2, e)
        rpg.check_pending(block=True)
        assert torch.equal(out0, out1)
        assert ei2.rpg_graph.src_np is None            # tables were built on the device
        bad = ei.clone()
        bad[1, 3 * tmpl.size(1) + 2] += 1              # graph 3 deviates from the template (graph 0)
        m(x, bad, e)                                   # runs (garbage for that graph) ...
        with pytest.raises(ValueError, match="asynchronously"):
            rpg.check_pending(block=True)              # ... and is reported here
    finally:
        rpg.set_validation(prev)
    # sync mode: the same bad batch is rejected in the call (as one graph of Gn * N nodes it is still a valid graph, so
    # it runs on the general path with its own template, like a cold call would)
    out2, _ = m(x, bad.clone(), e)
    assert torch.isfinite(out2).all()


def test_device_built_tables_equal_host_tables():
    N = 9
    H = N * (N - 1) // 2
    keep = R.edge_dropout_keep(H, np.random.RandomState(11).random_sample(H))
    tmpl = R.apply_edge_dropout(R.fc_edge_index(N), keep)
    Gn = 64
    ei = R.batched_edge_index(tmpl, Gn, N).to(dev())
    host = G.GraphBatch(tmpl[0].numpy(), tmpl[1].numpy(), Gn, N, dev())
    G._shape_cache[(Gn * N, str(dev()))] = (Gn, N)
    gdev = G.from_edge_index(ei, Gn * N)
    assert gdev.src_np is None
    for k in ("src", "dst", "in_ptr", "in_idx", "out_ptr", "out_idx", "min_ptr", "min_idx", "max_ptr", "max_idx",
              "inv_deg", "deg", "has_in", "sel_src", "sel_dst"):
        assert torch.equal(gdev._tables[k], host._tables[k]), k
    assert gdev.struct.sel_patterns == host.struct.sel_patterns and gdev.struct.sel_div == host.struct.sel_div


def test_mask_edge_index_matches_reference_indexing():
    N, Gn = 9, 33
    H = N * (N - 1) // 2
    keep = R.edge_dropout_keep(H, np.random.RandomState(3).random_sample(H))
    full = R.batched_edge_index(R.fc_edge_index(N), Gn, N)
    tiled = np.tile(keep, 2 * Gn)                       # train.py:242
    want = full[:, torch.from_numpy(tiled)]
    got = rpg.mask_edge_index(full.to(dev()), keep, Gn)
    assert torch.equal(got.cpu(), want)


def test_fused_adam_matches_torch_adam_and_refreshes_packed_weights():
    from relpose_gnn_b200 import parallel
    D, N, Gn = 128, 5, 12
    case = R.synth_stack_case(D, N, Gn, 5, droprate=0.5, edge_dropout=True)
    sd = {k: v.float() for k, v in case["params"].items()}

    def make():
        m = rpg.RelPoseGNN(D, D, D, droprate=0.5).to(dev())
        m.load_state_dict(sd, strict=False)
        return m, rpg.PoseNetCriterion(0.0, -2.0).to(dev())

    x = case["x"].float().to(dev())
    poses = case["poses"].float().to(dev())
    ei = case["edge_index"].to(dev())
    kx, ke = case["keep_x"].to(dev()), case["keep_e"].to(dev())
    m1, c1 = make()
    p1 = list(m1.parameters()) + list(c1.parameters())
    o1 = torch.optim.Adam(p1, lr=1e-3, weight_decay=1e-2)
    m2, c2 = make()
    p2 = list(m2.parameters()) + list(c2.parameters())
    bucket = parallel.FlatGradBucket(p2)
    m2.attach_grad_bucket(bucket)
    o2 = rpg.FusedAdam(p2, lr=1e-3, weight_decay=1e-2, grad_bucket=bucket, modules=[m2])
    for step in range(3):
        o1.zero_grad()
        l1 = c1(m1(x, ei, keep_x=kx, keep_e=ke)[1], poses, ei)[0]
        l1.backward()
        o1.step()
        o2.zero_grad()
        l2 = c2(m2(x, ei, keep_x=kx, keep_e=ke)[1], poses, ei)[0]
        l2.backward()
        o2.step()
        # same loss sequence: the second and third forward see the UPDATED weights (operands re-packed after the step)
        assert abs(l1.item() - l2.item()) < 2e-3 * max(1.0, abs(l1.item())), (step, l1.item(), l2.item())
    for (n1, a), b in zip(list(m1.named_parameters()) + list(c1.named_parameters()), p2):
        assert rel(b, a) < 2e-3, n1


def test_zero_grad_set_to_none_does_not_detach_the_bucket():
    """train.py:252 calls optimizer.zero_grad() (set_to_none=True by default in torch >= 2): the bucket must re-attach."""
    from relpose_gnn_b200 import parallel
    D, N, Gn = 128, 5, 12
    case = R.synth_stack_case(D, N, Gn, 6, droprate=0.0, edge_dropout=False)
    m = rpg.RelPoseGNN(D, D, D, droprate=0.0).to(dev())
    m.load_state_dict({k: v.float() for k, v in case["params"].items()}, strict=False)
    crit = rpg.PoseNetCriterion(0.0, -2.0).to(dev())
    params = list(m.parameters()) + list(crit.parameters())
    bucket = parallel.FlatGradBucket(params)
    m.attach_grad_bucket(bucket)
    x, poses, ei = case["x"].float().to(dev()), case["poses"].float().to(dev()), case["edge_index"].to(dev())

    def grads():
        crit(m(x, ei)[1], poses, ei)[0].backward()
        return bucket.allreduce().clone()

    bucket.zero()
    g0 = grads()
    torch.optim.SGD(params, lr=0.0).zero_grad()         # set_to_none=True: every p.grad is None now
    assert all(p.grad is None for p in params)
    g1 = grads()                                        # autograd allocates fresh tensors; allreduce() copies them back
    assert torch.equal(g0, g1)
    assert all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))


def test_replayed_repack_equals_a_full_rebuild_after_optimizer_steps():
    """The bf16 operand copies are re-converted after every optimizer step from RECORDED descriptors (same parameter
    storage, new values).  After FusedAdam steps and after an in-place torch update the forward must equal, bit for bit,
    the forward of the same model with every packed operand rebuilt from scratch (`invalidate_packed`)."""
    torch.manual_seed(3)
    D, N, Gn = 128, 9, 33
    model = rpg.RelPoseGNN(D, D, D, droprate=0.0).to(dev())
    crit = rpg.PoseNetCriterion(0.0, -2.0).to(dev())
    params = list(model.parameters()) + list(crit.parameters())
    opt = rpg.FusedAdam(params, lr=1e-2, modules=[model])
    src, dst = rpg.fc_template(N)
    ei = rpg.batched_edge_index(src, dst, Gn, N).to(dev())
    x = torch.randn(Gn * N, D, device=dev()).bfloat16()
    poses = 0.1 * torch.randn(Gn * N, 6, device=dev())
    for it in range(3):
        opt.zero_grad()
        pn, pe, eu = model(x, ei)
        loss, _, _ = crit(pe, poses, eu)
        loss.backward()
        opt.step()
        if it == 1:
            with torch.no_grad():
                model.gnn1.mlp[0].weight.mul_(1.01)           # version counter path (what torch optimizers do)
        with torch.no_grad():
            a_n, a_e, _ = model(x, ei)                          # replayed descriptors
            model.invalidate_packed()
            b_n, b_e, _ = model(x, ei)                          # rebuilt from scratch
        assert torch.equal(a_n, b_n) and torch.equal(a_e, b_e), it
    assert not torch.equal(a_e, torch.zeros_like(a_e))
