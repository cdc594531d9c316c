"""Host-side cost of one training step through the public API (fresh int64 edge_index every step, FusedAdam): on a tiny
batch the GPU work is negligible, so the loop time is the host's.  usage: host_overhead.py [G] [N]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relpose_gnn_b200 as rpg
from relpose_gnn_b200 import parallel, _lib
from relpose_gnn_b200.graph import edge_dropout_keep
dev = torch.device("cuda:0")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 9
D, H = 512, N * (N - 1) // 2
model = rpg.RelPoseGNN(D, D, D, droprate=0.5).to(dev); crit = rpg.PoseNetCriterion(0.0, -2.0).to(dev)
params = list(model.parameters()) + list(crit.parameters())
bucket = parallel.FlatGradBucket(params); model.attach_grad_bucket(bucket)
opt = rpg.FusedAdam(params, lr=1e-5, grad_bucket=bucket, modules=[model])
rpg.set_validation("async")
x = torch.randn(G * N, D, device=dev).bfloat16(); poses = 0.1 * torch.randn(G * N, 6, device=dev)
src, dst = rpg.fc_template(N)
ei_full = rpg.batched_edge_index(src, dst, G, N).to(dev)
rng = np.random.RandomState(7)
T = {}
def tick(name, t0):
    t1 = time.perf_counter(); T[name] = T.get(name, 0.0) + (t1 - t0); return t1
def step(keep):
    t = time.perf_counter()
    ei = rpg.mask_edge_index(ei_full, keep, G); t = tick("mask_edge_index", t)
    opt.zero_grad(); t = tick("zero_grad", t)
    pn, pe, eu = model(x, ei); t = tick("forward", t)
    loss, _, _ = crit(pe, poses, eu); t = tick("loss", t)
    loss.backward(); t = tick("backward", t)
    opt.step(); t = tick("adam+repack", t)
for _ in range(20): step(edge_dropout_keep(H, rng))
torch.cuda.synchronize(); T.clear()
n = 200; l0 = _lib.load().rpg_launch_count()
t0 = time.perf_counter()
for _ in range(n): step(edge_dropout_keep(H, rng))
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"G={G} N={N}: host {1e3 * (t1 - t0) / n:.3f} ms/step, with final sync {1e3 * (t2 - t0) / n:.3f} ms/step, "
      f"{(_lib.load().rpg_launch_count() - l0) / n:.0f} library launches/step")
print({k: round(1e3 * v / n, 3) for k, v in T.items()})
if os.environ.get("RPG_HOST_PROFILE") == "1":
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200): step(edge_dropout_keep(H, rng))
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(45)
    st.sort_stats("tottime").print_stats(25)
if os.environ.get("RPG_HOST_CSIDE") == "1":
    # time spent INSIDE the library calls (ctypes -> C -> CUDA launches) vs the Python around them
    lib = _lib.load()
    acc = {}
    class Timed:
        def __init__(self, name, fn): self.name, self.fn = name, fn
        def __call__(self, *a):
            t0 = time.perf_counter(); r = self.fn(*a); dt = time.perf_counter() - t0
            c = acc.setdefault(self.name, [0, 0.0]); c[0] += 1; c[1] += dt
            return r
    class Proxy:
        def __getattr__(self, k):
            return Timed(k, getattr(lib, k))
    _lib._lib = Proxy()
    for _ in range(20): step(edge_dropout_keep(H, rng))
    torch.cuda.synchronize(); acc.clear()
    t0 = time.perf_counter()
    for _ in range(n): step(edge_dropout_keep(H, rng))
    t1 = time.perf_counter(); torch.cuda.synchronize()
    tot = sum(v[1] for v in acc.values())
    print(f"with timing proxies: host {1e3 * (t1 - t0) / n:.3f} ms/step, inside library calls {1e3 * tot / n:.3f} ms/step")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"  {k:32s} {v[0] / n:6.1f} calls/step  {1e6 * v[1] / n:8.1f} us/step  {1e6 * v[1] / v[0]:7.1f} us/call")
