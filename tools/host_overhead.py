"""Host-side cost of one training step: run the step on a tiny batch (GPU work negligible) and time the host loop."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relpose_gnn_b200 as rpg
from relpose_gnn_b200 import parallel, _lib
from relpose_gnn_b200.graph import GraphBatch, attach, edge_dropout_keep
dev = torch.device("cuda:0"); G, N, D = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 9, 512; H = 36
model = rpg.RelPoseGNN(D, D, D, droprate=0.5).to(dev); crit = rpg.PoseNetCriterion(0.0, -2.0).to(dev)
params = list(model.parameters()) + list(crit.parameters())
bucket = parallel.FlatGradBucket(params); model.attach_grad_bucket(bucket)
x = torch.randn(G * N, D, device=dev).bfloat16(); poses = 0.1 * torch.randn(G * N, 6, device=dev)
rng = np.random.RandomState(7)
T = {}
def tick(name, t0):
    t1 = time.perf_counter(); T[name] = T.get(name, 0.0) + (t1 - t0); return t1
def step(keep):
    t = time.perf_counter()
    graph = GraphBatch.fully_connected(G, N, dev, keep); t = tick("graph", t)
    ei = attach(graph.edge_index(), graph); t = tick("edge_index", t)
    bucket.zero(); t = tick("zero", t)
    pn, pe, _ = model(x, ei); t = tick("forward", t)
    loss, _, _ = crit(pe, poses, ei); t = tick("loss", t)
    loss.backward(); t = tick("backward", t)
for _ in range(20): step(edge_dropout_keep(H, rng))
torch.cuda.synchronize(); T.clear()
n = 200; l0 = _lib.load().rpg_launch_count()
t0 = time.perf_counter()
for _ in range(n): step(edge_dropout_keep(H, rng))
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"G={G}: host {1e3 * (t1 - t0) / n:.3f} ms/step, with final sync {1e3 * (t2 - t0) / n:.3f} ms/step, "
      f"{(_lib.load().rpg_launch_count() - l0) / n:.0f} library launches/step")
print({k: round(1e3 * v / n, 3) for k, v in T.items()})
