"""Times the channel-attention kernels alone (CUDA events, L2 flushed between iterations by the tensor sizes).
usage: attn_probe.py [Et] [c]   env: RPG_ATT_SERIES, RPG_ATT_PIPE, RPG_ATT_EXACT"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relpose_gnn_b200 import ops
from relpose_gnn_b200.graph import GraphBatch
Et = int(sys.argv[1]) if len(sys.argv) > 1 else 155648
c = int(sys.argv[2]) if len(sys.argv) > 2 else 64
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 0.2
dev = torch.device("cuda:0")
g = GraphBatch.fully_connected(Et // 38 if Et % 38 == 0 else Et // 72, 9, dev, None if Et % 72 == 0 and Et % 38 else None)
Et = g.n_edge_rows
gtp = (torch.randn(Et, 3 * c, device=dev) * scale)
dyn = torch.randn(g.n_node_rows, c, device=dev)
cp, c3p = ops.pad64(c), ops.pad64(3 * c)
y = torch.zeros(Et, cp, dtype=torch.bfloat16, device=dev)
dgtp = torch.zeros(Et, c3p, dtype=torch.bfloat16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]
tf = timeit(lambda: ops.attention_fwd(gtp, c, y))
tb = timeit(lambda: ops.attention_bwd(gtp, dyn, g, c, dgtp))
fb, bb = Et * c * 14, Et * c * 18 + g.n_node_rows * c * 4
print(f"Et={Et} c={c} scale={scale} env={ {k: v for k, v in os.environ.items() if k.startswith('RPG_ATT')} }: "
      f"fwd {tf:.1f} us ({fb / tf / 1e3:.0f} GB/s)  bwd {tb:.1f} us ({bb / tb / 1e3:.0f} GB/s)")
