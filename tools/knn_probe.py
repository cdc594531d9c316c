"""Training step with dynamic kNN rewiring (the reference CLI default, --knn 4) at benchmark size: ms per step.
usage: python tools/knn_probe.py [G] [N] [k]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relpose_gnn_b200 as rpg  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 9
k = int(sys.argv[3]) if len(sys.argv) > 3 else 4
D = 512
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = rpg.RelPoseGNN(D, D, D, droprate=0.5, knn=k).to(dev)
crit = rpg.PoseNetCriterion(0.0, -2.0).to(dev)
x = torch.randn(G * N, D, device=dev)
poses = 0.1 * torch.randn(G * N, 6, device=dev)
graph = rpg.GraphBatch.fully_connected(G, N, dev)
ei = rpg.attach(graph.edge_index(), graph)


def step():
    model.zero_grad(set_to_none=True)
    pn, pe, ei_knn = model(x, ei)
    loss, _, _ = crit(pe, poses, ei_knn)
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"knn={k} G={G} N={N}: {ms:.3f} ms/step = {G / ms * 1e3:.0f} graphs/s, loss {loss.item():.4f}")
