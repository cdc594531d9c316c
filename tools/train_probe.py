"""One training step (stack fwd + loss + bwd) at a given size; prints progress so a failing call is visible.
usage: python tools/train_probe.py G N D [keepseed|-1]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relpose_gnn_b200 as rpg  # noqa: E402
from relpose_gnn_b200.graph import GraphBatch, attach, edge_dropout_keep  # noqa: E402

G, N, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 7
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = rpg.RelPoseGNN(D, D, D, droprate=0.5).to(dev)
crit = rpg.PoseNetCriterion(0.0, -2.0).to(dev)
x = torch.randn(G * N, D, device=dev).bfloat16()
poses = 0.1 * torch.randn(G * N, 6, device=dev)
rng = np.random.RandomState(seed)
for it in range(3):
    keep = edge_dropout_keep(N * (N - 1) // 2, rng) if seed >= 0 else None
    graph = GraphBatch.fully_connected(G, N, dev, keep)
    ei = attach(graph.edge_index(), graph)
    print(f"step {it}: Ep={graph.Ep} Et={graph.n_edge_rows}", flush=True)
    pn, pe, _ = model(x, ei)
    torch.cuda.synchronize(); print("  forward ok", flush=True)
    loss, _, _ = crit(pe, poses, ei)
    torch.cuda.synchronize(); print("  loss ok", loss.item(), flush=True)
    loss.backward()
    torch.cuda.synchronize(); print("  backward ok", flush=True)
    model.zero_grad()
print("DONE")
