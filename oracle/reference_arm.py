"""TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product package.

The reference's OWN implementation of the hot path as a timed step: `PoseNetX_R2` (posenet.py:920-1091, with
`simpleConvEdge_upt`, my_gnn_layer.py:277-311, `AttentionBlock`, att.py:7-34), `compute_RP` (posenet.py:1021-1031),
`PoseNetCriterion` (criterion.py:33-60) and the loop body of train.py:236-274, imported unmodified through
oracle/pyg_shim.py from /root/reference (build container) or the staged copy under baseline/_ref (GPU box).
Only the ResNet34 feature extractor is replaced by a stub that returns synthetic embeddings (it is not the target path).
"""
import time
import types

import numpy as np
import torch

from oracle import restatement as R
from oracle import stage_reference


class _StubFE(torch.nn.Module):
    """Stands in for torchvision resnet34 (train.py:173): returns the preset node embeddings."""

    def __init__(self, d):
        super().__init__()
        self.fc = torch.nn.Linear(d, d)
        self.avgpool = torch.nn.Identity()
        self._feats = None

    def forward(self, _img):
        return self._feats


def available():
    return stage_reference.reference_python_dir() is not None


class ReferenceStep:
    """One step of the reference on `G` graphs of `N` nodes: training (forward, compute_RP, criterion, backward, Adam) or
    inference (forward under no_grad).  `vector_rp=True` replaces the reference's per-edge Python loop in compute_RP by
    the equivalent indexing expression (used only for the informative torch-eager-on-GPU column, where the loop's
    thousands of tiny launches would measure the launch path instead of the layers)."""

    def __init__(self, D, N, G, train, edge_dropout, device="cpu", seed=4242, vector_rp=False):
        mods = stage_reference.import_reference_modules()
        if mods is None:
            raise RuntimeError("reference modules unavailable (neither /root/reference nor baseline/_ref)")
        self.posenet, criterion = mods
        self.D, self.N, self.G, self.train, self.edge_dropout = D, N, G, train, edge_dropout
        self.device = torch.device(device)
        self.vector_rp = vector_rp
        torch.manual_seed(seed)
        self.fe = _StubFE(D)
        self.model = self.posenet.PoseNetX_R2(self.fe, droprate=0.5, pretrained=False, feat_dim=D, edge_feat_dim=D,
                                              node_dim=D, use_gnn=True, knn=-1, gnn_recursion=2,
                                              device=str(self.device)).to(self.device)
        self.model.train(train)
        self.crit = criterion.PoseNetCriterion(sax=0.0, saq=-2.0, learn_beta=True).to(self.device)   # train.py:68-69,198-199
        gen = torch.Generator().manual_seed(seed + 1)
        self.x = torch.randn(G * N, D, generator=gen).to(self.device)
        self.poses = (0.1 * torch.randn(G * N, 6, generator=gen)).to(self.device)
        self.full = R.fc_edge_index(N)
        self.rng = np.random.RandomState(seed + 2)
        params = [p for n, p in self.model.named_parameters() if not n.startswith("feature_extractor")]
        self.opt = torch.optim.Adam(params + list(self.crit.parameters()), lr=1e-4, weight_decay=0.0) if train else None
        self.img = torch.zeros(G * N, 3 * self.model.input_img_height * 4, device=self.device)

    def edge_index(self):
        tmpl = self.full
        if self.train and self.edge_dropout:            # train.py:238-245 (mask applied to edge_index: documented intent)
            H = self.N * (self.N - 1) // 2
            keep = R.edge_dropout_keep(H, self.rng.random_sample(H))
            tmpl = R.apply_edge_dropout(tmpl, keep)
        return R.batched_edge_index(tmpl, self.G, self.N).to(self.device)

    def __call__(self):
        ei = self.edge_index()
        data = types.SimpleNamespace(x=self.img, edge_index=ei, edge_attr=None, batch=None)
        if not self.train:
            self.fe._feats = self.x
            with torch.no_grad():
                return self.model(data)[1]
        self.fe._feats = self.x.clone().requires_grad_(True)
        self.opt.zero_grad()
        pred, pred_R, ei_out = self.model(data)
        if self.vector_rp:
            target_R = self.poses[ei_out[0]] - self.poses[ei_out[1]]
        else:
            with torch.no_grad():      # on CPU `.to(device)` returns the leaf itself (posenet.py:1023): the in-place loop needs no_grad
                target_R = self.model.compute_RP(self.poses, ei_out)
        loss = self.crit(pred_R.view(1, pred_R.size(0), pred_R.size(1)), target_R.view(1, target_R.size(0), target_R.size(1)))
        loss[0].backward()
        self.opt.step()
        return loss[0].detach()


def time_reference(D, N, G, train, edge_dropout, steps, warmup, device="cpu", threads=None, vector_rp=False):
    """Median seconds per step and graphs/s of the reference step."""
    if threads:
        torch.set_num_threads(threads)
    step = ReferenceStep(D, N, G, train, edge_dropout, device=device, vector_rp=vector_rp)
    times = []
    for it in range(warmup + steps):
        if step.device.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        if step.device.type == "cuda":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    med = float(np.median(times))
    return G / med, med
