"""Per-step timing + allocator statistics of the bench training step (diagnostic for timing outliers)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relpose_gnn_b200 as rpg
from relpose_gnn_b200 import parallel
from relpose_gnn_b200.graph import GraphBatch, attach, edge_dropout_keep
dev = torch.device("cuda:0"); G, N, D = 4096, 9, 512; H = 36
torch.manual_seed(0)
model = rpg.RelPoseGNN(D, D, D, droprate=0.5).to(dev); crit = rpg.PoseNetCriterion(0.0, -2.0).to(dev)
params = list(model.parameters()) + list(crit.parameters())
bucket = parallel.FlatGradBucket(params); model.attach_grad_bucket(bucket)
x = torch.randn(G * N, D, device=dev).bfloat16(); poses = 0.1 * torch.randn(G * N, 6, device=dev)
rng = np.random.RandomState(7)
def step(keep):
    graph = GraphBatch.fully_connected(G, N, dev, keep); ei = attach(graph.edge_index(), graph)
    bucket.zero(); pn, pe, _ = model(x, ei); loss, _, _ = crit(pe, poses, ei); loss.backward(); return graph.Ep
for _ in range(2): step(np.ones(H, bool))
torch.cuda.synchronize()
sync_each = len(sys.argv) > 1
rows = []
for it in range(160):
    keep = edge_dropout_keep(H, rng)
    st = torch.cuda.memory_stats(dev)
    a0, f0 = st["num_device_alloc"], st["num_device_free"]
    t0 = time.perf_counter()
    ep = step(keep)
    t1 = time.perf_counter()
    if sync_each: torch.cuda.synchronize()
    t2 = time.perf_counter()
    st = torch.cuda.memory_stats(dev)
    rows.append((it, ep, (t1 - t0) * 1e3, (t2 - t0) * 1e3, st["num_device_alloc"] - a0, st["num_device_free"] - f0, st["reserved_bytes.all.current"] / 2**30))
torch.cuda.synchronize()
host = np.array([r[2] for r in rows]); print("host ms/step median %.2f  p90 %.2f  max %.2f" % (np.median(host), np.percentile(host, 90), host.max()))
for r in rows:
    if r[2] > 2 * np.median(host) or r[4] or r[5]:
        print("step %3d Ep=%2d host %.2f ms total %.2f ms  cudaMalloc %d cudaFree %d reserved %.1f GiB" % r)

import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for it in range(60): step(edge_dropout_keep(H, rng))
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
