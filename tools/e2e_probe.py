"""Where does the end-to-end loop lose time against the device-resident loop?  Per-step CUDA events in both modes:
busy = first kernel of the step -> last kernel; gap = end of step i -> start of step i+1.
usage: python tools/e2e_probe.py [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relpose_gnn_b200 as rpg  # noqa: E402
from relpose_gnn_b200 import parallel  # noqa: E402
from relpose_gnn_b200.graph import GraphBatch, attach, edge_dropout_keep  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
G, N, D = 4096, 9, 512
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = rpg.RelPoseGNN(D, D, D, droprate=0.5).to(dev)
crit = rpg.PoseNetCriterion(0.0, -2.0).to(dev)
params = list(model.parameters()) + list(crit.parameters())
bucket = parallel.FlatGradBucket(params)
model.attach_grad_bucket(bucket)
x_host = torch.randn(G * N, D).bfloat16().pin_memory()
poses_host = (0.1 * torch.randn(G * N, 6)).pin_memory()
x_dev, poses_dev = x_host.to(dev), poses_host.to(dev)
H = N * (N - 1) // 2


def step(x, poses, keep):
    graph = GraphBatch.fully_connected(G, N, dev, keep)
    ei = attach(graph.edge_index(), graph)
    bucket.zero()
    pn, pe, _ = model(x, ei)
    loss, _, _ = crit(pe, poses, ei)
    loss.backward()
    return loss


def run(mode):
    rng = np.random.RandomState(7)
    feeder = rpg.DeviceFeeder(dev)
    rb = rpg.ScalarReadback(1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    if mode != "device":
        feeder.stage(x_host, poses_host)
    for i in range(steps):
        keep = edge_dropout_keep(H, rng)
        if mode == "device":
            x, poses = x_dev, poses_dev
        else:
            x, poses = feeder.take()
        ev[i][0].record()
        loss = step(x, poses, keep)
        ev[i][1].record()
        if mode != "device":
            feeder.release()
            if i + 1 < steps:
                feeder.stage(x_host, poses_host)
            if mode == "e2e":
                if rb.full():
                    rb.pop()
                rb.push(loss.float())
    torch.cuda.synchronize()
    busy = [a.elapsed_time(b) for a, b in ev]
    gaps = [ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(steps - 1)]
    total = ev[0][0].elapsed_time(ev[-1][1]) / steps
    print(f"{mode:8s} total/step {total:.3f} ms   busy median {np.median(busy):.3f}   gap median {np.median(gaps):.3f} "
          f"max {np.max(gaps):.3f}")


for _ in range(2):
    for m in ("device", "copy", "e2e"):
        run(m)
