"""The GNN part of `PoseNetX_R2` (posenet.py:920-1091) as one fused stack on the B200 kernels.

`RelPoseGNN` mirrors the reference model's parameter names for this path -- `proj_edge`, `gnn1..gnnL`
(only `gnn1` is ever called, posenet.py:1060-1069), `fc_xyz`, `fc_wpqr`, `fc_xyz_R`, `fc_wpqr_R` -- so
`load_state_dict(reference_checkpoint['model_state_dict'], strict=False)` fills it.  The ResNet34 feature
extractor is NOT part of this path; `forward` takes the node embeddings it would produce.

    forward(x [G*N, D], edge_index [2, G*Ep]) -> (pose_nodes [G*N, 6], pose_edges [G*Ep, 6], edge_index)

as posenet.py:1091.  `training_step` adds compute_RP (posenet.py:1021-1031), PoseNetCriterion
(criterion.py:42-60) and the whole backward, i.e. what train.py:256-274 differentiates for the GNN.
"""
import torch
from torch import nn

from . import graph as graph_mod, ops
from .layers import (PARAM_ORDER, _invalidate_hook, fast_params, layer_backward_raw, layer_backward_split_raw, layer_forward_raw,
                     layer_forward_split_raw, simpleConvEdge_upt)
from .ops import BF16

_HEADS = ("fc_xyz", "fc_wpqr", "fc_xyz_R", "fc_wpqr_R")


def _lib_ws_floats(D):
    from . import _lib
    return _lib.load().rpg_layer_bwd_ws_floats(D, 0, 0)


def _arena_edge_rows(graph):
    """Edge rows the per-step arenas are sized for: the full template of the batch shape when the batch is an
    edge-dropped version of it (at most 4x the present edges), else the present edge count (kNN graphs: constant)."""
    Et = graph.n_edge_rows
    full = graph.G * graph.N * (graph.N - 1)
    return full if Et <= full <= 4 * Et else Et


class _StackFn(torch.autograd.Function):
    """proj_edge init -> R x (gnn1, ReLU, ReLU) -> feature dropout -> 4 pose heads, with a hand-written backward.
    ReLUs are folded into GEMM epilogues (forward: second store; backward: mask on the input that was a ReLU)."""

    @staticmethod
    @ops.scoped
    def forward(ctx, x, model, graph, drop, *params):
        D, R = model.node_dim, model.gnn_recursion
        dev = x.device
        Nt, Et = graph.n_node_rows, graph.n_edge_rows
        xb = ops.to_bf16(x)
        lw = model.gnn1._packed(dev).refresh(model.gnn1)
        sw = model._packed_stack(dev)
        # one allocation for every activation of the forward.  Sized for the FULL template of this batch shape: the edge
        # count changes with every edge-dropout mask, and a request of a new size each step sends the caching allocator
        # down its slow path (split / merge, ~30 us); the same size every step is a free-list hit.
        Ec = _arena_edge_rows(graph)
        arena = ops.Arena(dev, R * ops.layer_fwd_bytes(D, Nt, Ec) + Nt * 2 * D * 2 + Ec * (D * 2 + D // 8) + 4096)
        # edge-feature initialiser (posenet.py:1014-1017,1053-1055), factorised per node
        pmm = arena.take(Nt, 2 * D)
        ops.gemm_nt(xb, sw["Wmm"], out=pmm)
        e = arena.take(Et, D)
        e_bits = arena.take(Et, D // 8, torch.uint8) if any(ctx.needs_input_grad) else None   # patterns: backward only
        ops.edge_init_fwd(pmm, model.proj_edge.bias.data, graph, D, e, e_bits)
        acts = []
        xin, x_bits = xb, None
        p_drop, keep_x, keep_e, seed = drop
        # Seeded feature dropout (the production path) is applied by the GEMMs that produce the last round's outputs,
        # and the heads then are plain tensor-core GEMMs.  Explicit keep masks (parity tests) and droprate 0 keep the
        # stand-alone head kernels.
        fused = p_drop > 0 and keep_x is None and model.tensor_core_heads
        for r in range(R):                                       # same gnn1 weights each round (posenet.py:1060-1069)
            a = layer_forward_raw(lw, graph, xin, e, want_relu_copies=True, x_bits=x_bits, e_bits=e_bits, arena=arena,
                                  for_backward=any(ctx.needs_input_grad),
                                  drop=(p_drop, seed, seed + 1) if (fused and r == R - 1) else None)
            acts.append(a)
            xin, e, x_bits, e_bits = a["out_relu"], a["e_new_relu"], a.get("out_bits"), a.get("e_new_bits")
        if fused:
            pose8_n = torch.empty(Nt, 8, dtype=torch.float32, device=dev)
            pose8_e = torch.empty(Et, 8, dtype=torch.float32, device=dev)
            ops.gemm_nt(xin, sw["w6n_p"], bias=sw["b6n_p"], out_f32=pose8_n)
            ops.gemm_nt(e, sw["w6e_p"], bias=sw["b6e_p"], out_f32=pose8_e)
            pose_n, pose_e = pose8_n[:, :6], pose8_e[:, :6]      # row pitch 8: rpg_pose_criterion takes the pitch
        else:
            pose_n = ops.head_fwd(xin, sw["w6n"], sw["b6n"], keep=keep_x, seed=seed, p_drop=p_drop)
            pose_e = ops.head_fwd(e, sw["w6e"], sw["b6e"], keep=keep_e, seed=seed + 1, p_drop=p_drop)
        ctx.model, ctx.graph, ctx.drop = model, graph, drop
        ctx.fused_heads = fused
        ctx.head_bits = (x_bits, e_bits)
        ctx.saved = (xb, pmm, acts, xin, e, lw, sw)
        ctx.x_dtype = x.dtype
        if model.keep_debug_activations:
            model.debug_activations = {"e0": acts[0]["e"], "rounds": acts[:]}
        ctx.set_materialize_grads(False)
        return pose_n, pose_e

    @staticmethod
    @ops.scoped
    def backward(ctx, d_pose_n, d_pose_e):
        model, graph = ctx.model, ctx.graph
        xb, pmm, acts, x_last, e_last, lw, sw = ctx.saved
        p_drop, keep_x, keep_e, seed = ctx.drop
        D, R = model.node_dim, model.gnn_recursion
        dev = xb.device
        names = model._param_names()
        params = model._ordered_params()
        direct = model.fused_grad_accumulation and all(p.grad is not None for p in params)
        if direct:      # accumulate straight into param.grad (views of one flat bucket): no per-parameter torch kernels
            grads = {n: p.grad for n, p in zip(names, params)}
        else:
            flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
            grads, off = {}, 0
            for n, p in zip(names, params):
                grads[n] = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
        lgrads = {n: grads["gnn1." + n] for n in PARAM_ORDER}
        Nt, Et = graph.n_node_rows, graph.n_edge_rows
        Ec = _arena_edge_rows(graph)                               # full-template size: see the forward
        arena = ops.Arena(dev, R * ops.layer_bwd_bytes(D, Nt, Ec) + 4 * _lib_ws_floats(D) + Nt * 3 * D * 2 + 4096)
        ws = arena.take(1, _lib_ws_floats(D), torch.float32)

        # heads (posenet.py:1077-1086): gradient w.r.t. the pre-ReLU layer outputs (mask_relu)
        d_e = d_x = None
        if ctx.fused_heads:
            thresh = min(int(p_drop * 256.0 + 0.5), 255)
            scale = 256.0 / (256.0 - thresh)       # the rate the seeded dropout actually applies (rpg_gemm_t.drop_p)
            xb_bits, eb_bits = ctx.head_bits
            if d_pose_e is not None:
                d_e = ops.head_bwd_tc(d_pose_e.contiguous().float(), e_last, eb_bits, sw["w6eT"], scale,
                                      grads["fc_xyz_R.weight"], grads["fc_wpqr_R.weight"], grads["fc_xyz_R.bias"],
                                      grads["fc_wpqr_R.bias"])
            if d_pose_n is not None:
                d_x = ops.head_bwd_tc(d_pose_n.contiguous().float(), x_last, xb_bits, sw["w6nT"], scale,
                                      grads["fc_xyz.weight"], grads["fc_wpqr.weight"], grads["fc_xyz.bias"],
                                      grads["fc_wpqr.bias"])
            d_pose_e = d_pose_n = None
        if d_pose_e is not None:
            d_e = ops.head_bwd(d_pose_e.contiguous().float(), e_last, sw["w6e"], grads["fc_xyz_R.weight"],
                               grads["fc_wpqr_R.weight"], grads["fc_xyz_R.bias"], grads["fc_wpqr_R.bias"],
                               keep=keep_e, seed=seed + 1, p_drop=p_drop, mask_relu=True)
        if d_pose_n is not None:
            d_x = ops.head_bwd(d_pose_n.contiguous().float(), x_last, sw["w6n"], grads["fc_xyz.weight"],
                               grads["fc_wpqr.weight"], grads["fc_xyz.bias"], grads["fc_wpqr.bias"],
                               keep=keep_x, seed=seed, p_drop=p_drop, mask_relu=True)
        if d_e is None and d_x is None:
            return (None,) * (4 + len(names))

        for r in range(R - 1, -1, -1):
            d_x, d_e = layer_backward_raw(lw, graph, acts[r], d_x, d_e, lgrads, mask_dx=(r > 0), mask_de=True, arena=arena)
            acts[r] = None                                      # free this round's activations
        # Every gradient except proj_edge's is final here (the criterion's and the heads' came first): with an attached
        # bucket in overlap mode their all-reduce starts now and runs under the edge-initialiser backward below.
        if direct and model._early_reduce is not None:
            bucket, ranges = model._early_reduce
            bucket.begin_allreduce(ranges)

        # edge-feature initialiser backward: d_e is already masked by (e0 > 0)
        dpmm = arena.take(Nt, 2 * D)
        ops.segment_sum2(d_e, graph, "min", dpmm[:, :D], "max", dpmm[:, D:])
        dx = arena.take(Nt, D)
        ops.gemm_nt(dpmm, sw["WmmT"], resid=d_x, out=dx)
        gw = grads["proj_edge.weight"]
        # both column halves of proj_edge.weight in one product; the bias sums come from the first half (every edge has
        # exactly one lower endpoint)
        ops.wgrad_blocks(dpmm, xb, [gw[:, :D], gw[:, D:]], ws, bias=grads["proj_edge.bias"])
        dx = ops.to_f32(dx) if ctx.x_dtype == torch.float32 else dx
        if direct:
            return (dx, None, None, None) + (None,) * len(names)
        return (dx, None, None, None) + tuple(grads[n] for n in names)


class _StackFnSplit(torch.autograd.Function):
    """The stack in fp32 mode (split-bf16 arithmetic on the same tcgen05 kernels; the reference trains in fp32,
    train.py:266-274): every value a (hi, lo) bf16 pair, Linear = [A_hi | A_lo | A_hi] [W_hi | W_hi | W_lo]^T, weight
    gradients hi^T hi + lo^T hi + hi^T lo.  Feature dropout + heads run in the stand-alone head kernels (fp32 weights)."""

    @staticmethod
    @ops.scoped
    def forward(ctx, x, model, graph, drop, *params):
        D, R = model.node_dim, model.gnn_recursion
        dev = x.device
        Nt, Et = graph.n_node_rows, graph.n_edge_rows
        need_bwd = any(ctx.needs_input_grad)
        lw = model.gnn1._packed_split(dev).refresh(model.gnn1, training=need_bwd)
        sw = model._packed_stack_split(dev, training=need_bwd)
        xs = ops.to_split(x.float())
        pmm = torch.empty(Nt, 2 * D, dtype=torch.float32, device=dev)
        ops.gemm_nt(None, sw["Wmm3"], segs=[xs[0], xs[1], xs[0]], out_f32=pmm)
        e = (torch.empty(Et, D, dtype=BF16, device=dev), torch.empty(Et, D, dtype=BF16, device=dev))
        e_bits = torch.empty(Et, D // 8, dtype=torch.uint8, device=dev) if need_bwd else None
        ops.edge_init_fwd_split(pmm, model.proj_edge.bias.data, graph, D, e[0], e[1], e_bits)
        acts = []
        xin, x_bits = xs, None
        for r in range(R):
            a = layer_forward_split_raw(lw, graph, xin, e, want_relu_copies=True, for_backward=need_bwd, x_bits=x_bits,
                                        e_bits=e_bits)
            acts.append(a)
            xin, e, x_bits, e_bits = a["out_relu"], a["e_new_relu"], a.get("out_bits"), a.get("e_new_bits")
        p_drop, keep_x, keep_e, seed = drop
        pose_n = ops.head_fwd(xin[0], sw["w6n"], sw["b6n"], keep=keep_x, seed=seed, p_drop=p_drop, feat_lo=xin[1])
        pose_e = ops.head_fwd(e[0], sw["w6e"], sw["b6e"], keep=keep_e, seed=seed + 1, p_drop=p_drop, feat_lo=e[1])
        ctx.model, ctx.graph, ctx.drop = model, graph, drop
        ctx.saved = (xs, acts, xin, e, lw, sw)
        ctx.x_dtype = x.dtype
        if model.keep_debug_activations:
            model.debug_activations = {"e0": acts[0]["e"], "rounds": acts[:]}
        ctx.set_materialize_grads(False)
        return pose_n, pose_e

    @staticmethod
    @ops.scoped
    def backward(ctx, d_pose_n, d_pose_e):
        model, graph = ctx.model, ctx.graph
        xs, acts, x_last, e_last, lw, sw = ctx.saved
        p_drop, keep_x, keep_e, seed = ctx.drop
        D, R = model.node_dim, model.gnn_recursion
        dev = xs[0].device
        names = model._param_names()
        params = model._ordered_params()
        direct = model.fused_grad_accumulation and all(p.grad is not None for p in params)
        if direct:
            grads = {n: p.grad for n, p in zip(names, params)}
        else:
            flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
            grads, off = {}, 0
            for n, p in zip(names, params):
                grads[n] = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
        lgrads = {n: grads["gnn1." + n] for n in PARAM_ORDER}
        Nt, Et = graph.n_node_rows, graph.n_edge_rows
        d_e = d_x = None
        if d_pose_e is not None:
            d_e = ops.head_bwd_split(d_pose_e.contiguous().float(), e_last, sw["w6e"], grads["fc_xyz_R.weight"],
                                     grads["fc_wpqr_R.weight"], grads["fc_xyz_R.bias"], grads["fc_wpqr_R.bias"],
                                     keep=keep_e, seed=seed + 1, p_drop=p_drop, mask_relu=True)
        if d_pose_n is not None:
            d_x = ops.head_bwd_split(d_pose_n.contiguous().float(), x_last, sw["w6n"], grads["fc_xyz.weight"],
                                     grads["fc_wpqr.weight"], grads["fc_xyz.bias"], grads["fc_wpqr.bias"],
                                     keep=keep_x, seed=seed, p_drop=p_drop, mask_relu=True)
        if d_e is None and d_x is None:
            return (None,) * (4 + len(names))
        for r in range(R - 1, -1, -1):
            d_x, d_e = layer_backward_split_raw(lw, graph, acts[r], d_x, d_e, lgrads, mask_dx=(r > 0), mask_de=True)
            acts[r] = None
        if direct and model._early_reduce is not None:
            bucket, ranges = model._early_reduce
            bucket.begin_allreduce(ranges)
        # edge-feature initialiser backward: d_e is already masked by (e0 > 0)
        def pair(rows, cols):
            return torch.empty(rows, cols, dtype=BF16, device=dev), torch.empty(rows, cols, dtype=BF16, device=dev)
        dpmm = pair(Nt, 2 * D)
        ops.segment_sum_split(d_e, graph, "min", (dpmm[0][:, :D], dpmm[1][:, :D]))
        ops.segment_sum_split(d_e, graph, "max", (dpmm[0][:, D:], dpmm[1][:, D:]))
        dx = pair(Nt, D)
        ops.gemm_nt(None, sw["WmmT3"], segs=[dpmm[0], dpmm[1], dpmm[0]], resid=d_x[0], resid_lo=d_x[1], out=dx[0], out_lo=dx[1])
        ws = torch.empty(3 * _lib_ws_floats(D), dtype=torch.float32, device=dev)
        gw = grads["proj_edge.weight"]
        ops.wgrad_split((dpmm[0][:, :D], dpmm[1][:, :D]), xs, gw[:, :D], ws, bias=grads["proj_edge.bias"])
        ops.wgrad_split((dpmm[0][:, D:], dpmm[1][:, D:]), xs, gw[:, D:], ws)
        dx = ops.from_split(*dx)
        if direct:
            return (dx, None, None, None) + (None,) * len(names)
        return (dx, None, None, None) + tuple(grads[n] for n in names)


class RelPoseGNN(nn.Module):
    """GNN stack of PoseNetX_R2 (constructor arguments as posenet.py:923-930 where they concern this path)."""

    def __init__(self, feat_dim=1024, edge_feat_dim=1024, node_dim=1024, droprate=0.5, gnn_recursion=2, L=1, knn=-1):
        super().__init__()
        self.knn = knn                           # > 0: rewire every graph to its k-NN graph in embedding space (posenet.py:1046-1048)
        if not (feat_dim == edge_feat_dim == node_dim):
            raise ValueError("the reference instantiates feat_dim == edge_feat_dim == node_dim (train.py:174-189)")
        self.node_dim, self.droprate, self.gnn_recursion, self.n_layers = node_dim, droprate, gnn_recursion, L
        self.proj_edge = nn.Linear(feat_dim * 2, edge_feat_dim)
        for layer in range(L):
            setattr(self, f"gnn{layer + 1}", simpleConvEdge_upt(node_dim, edge_feat_dim, node_dim))
        for h in _HEADS:
            setattr(self, h, nn.Linear(node_dim, 3))
        # initialisation as posenet.py:981-997: kaiming_normal_ + zero bias on the bare Linear modules listed there
        for m in [self.proj_edge] + [getattr(self, h) for h in _HEADS]:
            nn.init.kaiming_normal_(m.weight.data)
            nn.init.constant_(m.bias.data, 0)
        self._stack_cache = {}
        self.register_load_state_dict_post_hook(_invalidate_hook)
        # Seed of the in-kernel feature dropout: a counter hashed together with the data-parallel rank, so that replicas
        # draw different masks and consecutive steps draw unrelated ones (set `dropout_seed` for reproducible runs).
        self.dropout_seed = 0x5EED
        self.dropout_rank = None                 # None: torch.distributed rank if initialised, else 0
        self.keep_debug_activations = False      # tests: keep the saved activations of the last forward
        self.fused_grad_accumulation = False     # see attach_grad_bucket
        self.tensor_core_heads = True            # seeded dropout fused into the last GEMMs + heads as GEMMs (see _StackFn)
        self.precision = "bf16"                  # or "fp32": split-bf16 arithmetic, inference only for now

    _early_reduce = None

    def attach_grad_bucket(self, bucket, overlap=False):
        """Makes backward accumulate every gradient of this module directly into `param.grad` (which a
        parallel.FlatGradBucket has pointed at one flat buffer) instead of returning fresh tensors to autograd:
        the weight-gradient kernels already accumulate (+=), so a step needs no per-parameter add kernels."""
        mine = {id(p) for p in self.parameters()}
        have = {id(p) for p in bucket.params}
        if not mine <= have:
            raise ValueError("the bucket must cover every parameter of this module")
        self.fused_grad_accumulation = True
        # overlap: the all-reduce of everything but proj_edge's gradients is launched from inside the backward (see
        # _StackFn.backward); finish with bucket.finish_allreduce() instead of bucket.allreduce()
        self._early_reduce = (bucket, bucket.ranges_excluding(list(self.proj_edge.parameters()))) if overlap else None
        return self

    def _param_names(self):
        return (["proj_edge.weight", "proj_edge.bias"] + ["gnn1." + n for n in PARAM_ORDER] +
                [f"{h}.{s}" for h in _HEADS for s in ("weight", "bias")])

    def _ordered_params(self):
        names = self.__dict__.get("_rpg_names")
        if names is None:
            names = self.__dict__["_rpg_names"] = tuple(self._param_names())
        return list(fast_params(self, names).values())

    _pack_epoch = 0

    def invalidate_packed(self):
        """Forces every packed operand of the stack (and of gnn1..gnnL) to be rebuilt at the next forward; see
        simpleConvEdge_upt.invalidate_packed."""
        self._pack_epoch = self._pack_epoch + 1
        self.__dict__.pop("_rpg_leaf_cache", None)
        for layer in range(self.n_layers):
            getattr(self, f"gnn{layer + 1}").invalidate_packed()

    _value_epoch = 0

    def mark_values_changed(self):
        """See simpleConvEdge_upt.mark_values_changed (FusedAdam calls it after every step)."""
        self._value_epoch = self._value_epoch + 1
        for layer in range(self.n_layers):
            getattr(self, f"gnn{layer + 1}").mark_values_changed()

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_stack_cache"] = {}                   # packed operands are a cache: not pickled / deep-copied
        state.pop("_rpg_leaf_cache", None)
        return state

    def _packed_stack(self, device):
        """bf16 proj_edge operands and the [6, D] fp32 head matrices; refreshed (ONE launch) when parameters change."""
        D = self.node_dim
        ps = [self.proj_edge.weight] + [getattr(self, h).weight for h in _HEADS] + [getattr(self, h).bias for h in _HEADS]
        versions = (self._pack_epoch, tuple(p.data_ptr() for p in ps), self._value_epoch) + tuple(p._version for p in ps)
        ent = self._stack_cache.get(str(device))
        if ent is not None and ent["versions"] == versions:
            return ent
        if ent is not None and ent.get("replay") is not None and ent["versions"][:2] == versions[:2]:
            ops.replay_packs(ent["replay"], ent["Wmm"])      # same addresses, new values: the recorded windows again
            ent["versions"] = versions
            return ent
        if ent is None:
            f32 = torch.float32
            ent = {"Wmm": torch.empty(2 * D, D, dtype=BF16, device=device),
                   "WmmT": torch.empty(D, 2 * D, dtype=BF16, device=device)}
            # tensor-core heads (features dropped by the producing GEMM): W6 padded to 8 output rows, bias to 8, and the
            # dgrad operand [D, 64] with W6[j, :] in columns j and 8 + j (hi / lo halves of dpose)
            for tag in ("n", "e"):
                ent["w6" + tag] = torch.empty(6, D, dtype=f32, device=device)
                ent["b6" + tag] = torch.empty(6, dtype=f32, device=device)
                ent["w6%s_p" % tag] = torch.zeros(8, D, dtype=BF16, device=device)
                ent["b6%s_p" % tag] = torch.zeros(8, dtype=f32, device=device)
                ent["w6%sT" % tag] = torch.zeros(D, 64, dtype=BF16, device=device)
            self._stack_cache[str(device)] = ent
        q = ops.PackQueue(record=True)
        W = self.proj_edge.weight.data                                     # [D, 2D]: columns (min node | max node)
        q.add(W, ent["Wmm"][:D], c0=0, cols=D)
        q.add(W, ent["Wmm"][D:], c0=D, cols=D)
        q.add(W, ent["WmmT"][:, :D], c0=0, cols=D, transpose=True)
        q.add(W, ent["WmmT"][:, D:], c0=D, cols=D, transpose=True)
        for tag, (ht, hq) in (("n", (self.fc_xyz, self.fc_wpqr)), ("e", (self.fc_xyz_R, self.fc_wpqr_R))):
            for r0, head in ((0, ht), (3, hq)):
                w, b = head.weight.data, head.bias.data.view(1, 3)
                q.add(w, ent["w6" + tag][r0:r0 + 3])
                q.add(b, ent["b6" + tag][r0:r0 + 3].view(1, 3))
                q.add(w, ent["w6%s_p" % tag][r0:r0 + 3])
                q.add(b, ent["b6%s_p" % tag][r0:r0 + 3].view(1, 3))
                q.add(w, ent["w6%sT" % tag][:, r0:r0 + 3], transpose=True)
                q.add(w, ent["w6%sT" % tag][:, 8 + r0:8 + r0 + 3], transpose=True)
        q.flush()
        ent["versions"] = versions
        ent["replay"] = q.recorded
        return ent

    def _drop_args(self, keep_x, keep_e):
        # F.dropout's default training=True keeps dropout active under eval() too (posenet.py:1073-1075)
        if self.droprate <= 0:
            return (0.0, None, None, 0)
        if (keep_x is None) != (keep_e is None):
            raise ValueError("give both keep masks or neither")
        if keep_x is not None:
            keep_x = keep_x.to(torch.uint8).contiguous()
            keep_e = keep_e.to(torch.uint8).contiguous()
            return (float(self.droprate), keep_x, keep_e, 0)
        self.dropout_seed += 2
        self.last_seed = self._mixed_seed()      # what the kernels received (tests rebuild the masks from it)
        return (float(self.droprate), None, None, self.last_seed)

    def _mixed_seed(self):
        """splitmix64 of (step counter, rank): the seed the kernels see (an even number; the edge stream uses seed + 1)."""
        rank = self.dropout_rank
        if rank is None:
            import torch.distributed as dist
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        z = (int(self.dropout_seed) * 0x9E3779B97F4A7C15 + (int(rank) + 1) * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return ((z ^ (z >> 31)) & 0x7FFFFFFFFFFFFFFE)

    def forward(self, x, edge_index, keep_x=None, keep_e=None, k=None):
        """x: node embeddings [G*N, D] (what feature_extractor returns, posenet.py:1037).  keep_x / keep_e: optional
        explicit Bernoulli keep masks for the feature dropout (parity tests); otherwise a counter-based in-kernel RNG.
        With `self.knn > 0` (or `k`) the given edge_index only tells the graph size: every graph is rewired to its k-NN
        graph in embedding space (posenet.py:1043-1050) and that edge_index is returned, as in the reference."""
        if not x.is_cuda:
            raise ValueError("RelPoseGNN needs CUDA tensors: the sm_100a kernels are the only implementation")
        if x.dim() != 2 or x.size(1) != self.node_dim:
            raise ValueError(f"x must be [rows, {self.node_dim}]")
        with ops.stream_scope(x.device):
            return self._forward_scoped(x, edge_index, keep_x, keep_e, k)

    def _forward_scoped(self, x, edge_index, keep_x, keep_e, k):
        graph = graph_mod.from_edge_index(edge_index, x.size(0))
        kk = self.knn if self.knn > 0 else k
        if kk is not None and kk > 0:
            n_graphs, n_nodes = graph.G, graph.N
            edge_index = graph_mod.knn_graph(x, kk, num_nodes_per_graph=n_nodes)
            graph = graph_mod.GraphBatch.per_graph(edge_index, n_graphs, n_nodes)    # tables built on the device
            graph_mod.attach(edge_index, graph)
        drop = self._drop_args(keep_x, keep_e)
        if self.precision == "fp32":
            pose_n, pose_e = self._forward_fp32(x, graph, drop)
            return pose_n, pose_e, edge_index
        pose_n, pose_e = _StackFn.apply(x, self, graph, drop, *self._ordered_params())
        return pose_n, pose_e, edge_index

    def _packed_stack_split(self, device, training=False):
        """fp32-mode operands of the stack: proj_edge as [W_hi | W_hi | W_lo] (and its transpose for the backward)."""
        sw = self._packed_stack(device)
        D = self.node_dim
        key = (sw["versions"], bool(training))
        if sw.get("split_key") != key:
            if "Wmm3" not in sw:
                sw["Wmm3"] = torch.zeros(2 * D, 3 * D, dtype=BF16, device=device)
            if training and "WmmT3" not in sw:
                sw["WmmT3"] = torch.zeros(D, 6 * D, dtype=BF16, device=device)
            W = self.proj_edge.weight.data
            q = ops.PackQueue()
            q.add3(W, sw["Wmm3"][:D], c0=0, cols=D)
            q.add3(W, sw["Wmm3"][D:], c0=D, cols=D)
            if training:                                   # WmmT = [W_min^T | W_max^T]  [D, 2D], thirds of 2D columns
                for i, lo in enumerate((False, False, True)):
                    for blk in range(2):
                        q.add(W, sw["WmmT3"][:, i * 2 * D + blk * D:i * 2 * D + (blk + 1) * D], c0=blk * D, cols=D,
                              transpose=True, lo=lo)
            q.flush()
            sw["split_key"] = key
        return sw

    def _forward_fp32(self, x, graph, drop):
        """fp32 mode (BASELINE config B; the reference's native training precision): split-bf16 arithmetic end to end."""
        return _StackFnSplit.apply(x, self, graph, drop, *self._ordered_params())

    @staticmethod
    def compute_RP(p, edge_index):
        """posenet.py:1021-1031 without the per-edge Python loop."""
        return p[edge_index[0]] - p[edge_index[1]]


class PoseNetCriterion(nn.Module):
    """criterion.py:33-60 with L1 losses; the L1 sums and the target gather run in one fused kernel."""

    def __init__(self, sax=0.0, saq=0.0, learn_beta=True):
        super().__init__()
        self.sax = nn.Parameter(torch.tensor([float(sax)]), requires_grad=learn_beta)
        self.saq = nn.Parameter(torch.tensor([float(saq)]), requires_grad=learn_beta)

    def forward(self, pred_edges, poses, edge_index):
        """loss(pred_R, compute_RP(poses, edge_index)) -> (loss, t_loss, q_loss), differentiable w.r.t. pred, sax, saq."""
        graph = graph_mod.from_edge_index(edge_index, poses.size(0))
        return _PoseLossFn.apply(pred_edges, poses.float().contiguous(), graph, self.sax, self.saq)


class _PoseLossFn(torch.autograd.Function):
    @staticmethod
    @ops.scoped
    def forward(ctx, pred, poses, graph, sax, saq):
        if pred.dtype != torch.float32 or pred.stride(1) != 1:      # row-pitched fp32 views (tensor-core heads) pass through
            pred = pred.float().contiguous()
        ctx.set_materialize_grads(False)
        out7, dpred = ops.pose_criterion(pred, poses, graph, sax.detach().float().contiguous(),
                                         saq.detach().float().contiguous())
        ctx.save_for_backward(dpred, out7)
        return out7[2:3], out7[3], out7[4]

    @staticmethod
    @ops.scoped
    def backward(ctx, g_loss, g_t, g_q):
        dpred, out7 = ctx.saved_tensors
        if g_loss is None:                       # only the logged t_loss / q_loss were used
            return None, None, None, None, None
        g = g_loss.reshape(())
        d_s = out7[5:7] * g
        return dpred * g, None, None, d_s[0:1], d_s[1:2]
