import os, sys, torch, numpy as np
sys.path.insert(0, os.getcwd())
from relpose_gnn_b200 import graph as G, ops
dev = torch.device("cuda:0")
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(n):
        big.zero_()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3
for (N, Gn, kf) in [(9, 4096, 0.5), (9, 4096, 1.0), (17, 2048, 0.5), (17, 2048, 1.0)]:
    H = N * (N - 1) // 2
    keep = np.random.RandomState(7).rand(H) < kf if kf < 1 else np.ones(H, bool)
    graph = G.GraphBatch.fully_connected(Gn, N, dev, keep)
    Et, Nt, D = graph.n_edge_rows, graph.n_node_rows, 512
    v = torch.randn(Et, D, device=dev).bfloat16()
    o = torch.empty(Nt, 2 * D, dtype=torch.bfloat16, device=dev)
    o1 = torch.empty(Nt, D, dtype=torch.bfloat16, device=dev)
    t1 = bench(lambda: ops.segment_sum(v, graph, "in", o1))
    t2 = bench(lambda: ops.segment_sum2(v, graph, "min", o[:, :D], "max", o[:, D:]))
    b1 = (Et + Nt) * D * 2; b2 = (Et + 2 * Nt) * D * 2
    print(f"N={N} G={Gn} keep={kf} Ep={graph.Ep}: sum {t1:.1f} us {b1/t1/1e6:.2f} TB/s | sum2 {t2:.1f} us {b2/t2/1e6:.2f} TB/s  (staged={os.environ.get('RPG_SEGMENT_STAGED','1')})")
