"""bench.py -- RelPose-GNN message-passing hot path on B200: graphs/s forward+backward.

    python bench.py --gpus N --steps K --warmup W [--workload train_4096x9|train_2048x17|infer_4096x9]
    python bench.py --impl reference ...      (the reference algorithm on the host CPU, oracle port)

Default workload = BASELINE.json configs[2]: training step (forward + pose loss + backward, train-time edge
dropout, feature dropout 0.5) on 4096 graphs x 9 nodes, D = 512, bf16, per GPU (weak scaling: N GPUs process
N x 4096 graphs and all-reduce the gradients once per step).  Prints ONE JSON line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (graphs per GPU, nodes, D, mode)
    "train_4096x9": (4096, 9, 512, "train"),
    "train_2048x17": (2048, 17, 512, "train"),
    "infer_4096x9": (4096, 9, 512, "infer"),
    "infer_fp32_4096x9": (4096, 9, 512, "infer_fp32"),     # BASELINE configs[1]: fp32 mode (split-bf16 arithmetic)
    "infer_8192x9": (8192, 9, 512, "infer"),               # BASELINE configs[4], weak: 65 536 graphs over 8 GPUs
    "infer_65536x9_strong": (65536, 9, 512, "infer"),      # BASELINE configs[4], strong: 65 536 graphs in total
    "train_2048x17_strong": (2048, 17, 512, "train"),      # BASELINE configs[3], strong: 2048 graphs in total
    "train_knn4_4096x9": (4096, 9, 512, "train_knn"),      # the reference CLI default: dynamic 4-NN rewiring (train.py:377)
    "train_fp32_4096x9": (4096, 9, 512, "train_fp32"),     # the reference's native precision (train.py:266-274) in fp32 mode
}
STRONG = {"infer_65536x9_strong", "train_2048x17_strong"}
METRIC = "GNN graphs/sec fwd+bwd"
_emit = print
R_ROUNDS = 2


def algorithmic_flops_per_graph(D, N, E, R=R_ROUNDS, train=True):
    """BASELINE.md section 3 / SURVEY.md 8d: reference formulation with concatenated inputs, multiply-add = 2."""
    fwd = 4 * D * D * E + R * (15 * D * D * E + 6 * D * D * N) + 12 * D * (E + N)
    return fwd * (3 if train else 1)


def load_traffic():
    """DRAM bytes per tcgen05 GEMM launch (and the tensor-pipe counters) from the committed ncu launch list of the
    training step (profiles/r2_launches_train_4096x9.csv -> tools/summarize_launches.py -> profiles/r2_gemm_traffic.json)."""
    for name in ("r2_gemm_traffic.json", "r1_gemm_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)
            d["file"] = "profiles/" + name
            return d
        except Exception:
            continue
    return None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
            except ValueError:
                pass
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference_throughput(G_sample, N, D, train, steps, warmup, threads=None):
    """The reference's own modules (oracle/reference_arm.py: PoseNetX_R2 + compute_RP + PoseNetCriterion + Adam through
    the PyG shim, staged under baseline/_ref for the GPU box) on the host CPU, fp32, all threads, `G_sample` graphs per
    step; if they are not available, the oracle port (oracle/restatement.py).  Returns (graphs/s, s/step, threads, kind)."""
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    from oracle import reference_arm
    if reference_arm.available():
        v, med = reference_arm.time_reference(D, N, G_sample, train, edge_dropout=train, steps=steps, warmup=warmup,
                                              threads=threads)
        return v, med, threads, "reference"
    from oracle import restatement as R
    case = R.synth_stack_case(D, N, G_sample, 4242, droprate=0.5, edge_dropout=train, dtype=torch.float32)
    params = {k: v.clone().requires_grad_(train) for k, v in case["params"].items()}
    sax = torch.zeros(1, requires_grad=train)
    saq = torch.full((1,), -2.0, requires_grad=train)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if train:
            loss = R.training_loss(params, case["x"].clone().requires_grad_(True), case["edge_index"], case["poses"], sax,
                                   saq, R_ROUNDS, 0.5, case["keep_x"], case["keep_e"])
            loss.backward()
            for p in params.values():
                p.grad = None
        else:
            with torch.no_grad():
                R.stack_forward(params, case["x"], case["edge_index"], R_ROUNDS, 0.5, case["keep_x"], case["keep_e"])
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    med = float(np.median(times))
    return G_sample / med, med, threads, "port"


def cpu_sample_note(kind, sample, N, mode, threads, med):
    what = ("the reference's own PoseNetX_R2 / simpleConvEdge_upt / compute_RP / PoseNetCriterion / Adam, imported unmodified "
            "through oracle/pyg_shim.py") if kind == "reference" else "oracle port (oracle/restatement.py)"
    return (f"{sample} graphs x {N} nodes per step, {mode}, fp32 torch CPU, {threads} threads, {what}; "
            f"{med * 1e3:.0f} ms per sample")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    G, N, D, mode = WORKLOADS[args.workload]
    train = mode in ("train", "train_knn")
    sample = args.cpu_sample
    value, med, threads, kind = cpu_reference_throughput(sample, N, D, train, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "graphs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "graphs_per_step": sample, "nodes": N, "D": D, "mode": mode,
                       "note": "the reference's CPU implementation of the path on the host cores; each step is a bounded "
                               "sample of the workload (the reference materialises [Et, D/8, D/8] attention tensors)"},
            "cpu_baseline": {"value": value, "unit": "graphs/s", "cores": threads, "kind": kind,
                             "sample": cpu_sample_note(kind, sample, N, mode, threads, med)},
            "e2e": {"value": value, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(json.dumps(line))


# ------------------------------------------------------------------------------------------- our arm (B200)
def hbm_roofline(records, peaks, D):
    """Per kernel class of the bandwidth-bound part of the path: algorithmic bytes / CUDA-event time vs the measured HBM peak."""
    from relpose_gnn_b200 import _lib
    out = {}
    groups = {}
    for r in records:
        name = _lib.PROF_CLASSES[r.cls]
        if name == "gemm_nt" and r.N <= 8:
            name = "pose_heads(gemm N=8)"
        elif name.startswith("gemm"):
            continue
        groups.setdefault(name, []).append(r)
    for name, rs in groups.items():
        ms = sum(r.ms for r in rs)
        by = sum(r.bytes for r in rs)
        ent = {"launches": len(rs), "ms_per_step": ms, "algorithmic_bytes_per_launch": by / len(rs),
               "achieved": by / (ms * 1e-3) / 1e9 if ms > 0 else None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
               "frac": by / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if ms > 0 else None, "bound": "hbm"}
        if name.startswith("attention"):
            ex = sum(r.flops for r in rs)             # exp evaluations
            ent["bound"] = "mufu (exp2: 16/clk/SM)"
            ent["gexp_per_s"] = ex / (ms * 1e-3) / 1e9 if ms > 0 else None
        out[name] = ent
    return out


def run_ours(args):
    import torch.distributed as dist

    import relpose_gnn_b200 as rpg
    from relpose_gnn_b200 import _lib, parallel
    from relpose_gnn_b200.graph import GraphBatch, attach, edge_dropout_keep

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line
    rank, local_rank, world = parallel.init_distributed("nccl")
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    G, N, D, mode = WORKLOADS[args.workload]
    strong = args.workload in STRONG
    if strong:
        if G % world:
            raise SystemExit("strong-scaling workloads need a GPU count that divides the graph count")
        G //= world                               # contiguous block of graphs per rank
    knn = 4 if mode == "train_knn" else -1
    if knn > 0:
        mode = "train"
    train = mode == "train"
    fp32_mode = mode in ("infer_fp32", "train_fp32")
    if mode == "train_fp32":
        train = True
    H = N * (N - 1) // 2
    peaks = load_peaks()

    torch.manual_seed(0)
    model = rpg.RelPoseGNN(D, D, D, droprate=0.5, gnn_recursion=R_ROUNDS, knn=knn).to(dev)
    if fp32_mode:
        model.precision = "fp32"
    crit = rpg.PoseNetCriterion(sax=0.0, saq=-2.0).to(dev)
    params = list(model.parameters()) + list(crit.parameters())
    if world > 1:      # identical replicas
        for p in params:
            dist.broadcast(p.data, 0)
    bucket = parallel.FlatGradBucket(params) if train else None
    opt = None
    if train:
        # diagnostic only (the printed line says so): replicas step without exchanging gradients, to separate the
        # cost of the collective from what a multi-GPU box costs by itself (clocks, host contention)
        skip_reduce = os.environ.get("RPG_BENCH_SKIP_ALLREDUCE") == "1"
        if skip_reduce:
            args.overlap_allreduce = False
        model.attach_grad_bucket(bucket, overlap=args.overlap_allreduce and world > 1)
        if args.optimizer:
            # train.py:211: Adam over the parameters of the path; one fused kernel over the flat buckets.  Its 1/world
            # folds the gradient average of the data-parallel sum, and every step re-packs the bf16 operands.
            opt = rpg.FusedAdam(params, lr=1e-5, grad_bucket=bucket, modules=[model])
    rpg.set_validation("async" if args.validation == "async" else "sync")

    # synthetic inputs of the named shape: ResNet34 embeddings ~ N(0,1) in bf16, poses ~ N(0, 0.1), and the
    # PyG-batched int64 edge_index of the FULL templates as the loader produces it (train.py:24,132)
    gen = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(G * N, D, generator=gen)
    x_host = (x_host if fp32_mode else x_host.bfloat16()).pin_memory()
    poses_host = (0.1 * torch.randn(G * N, 6, generator=gen)).pin_memory()
    src_t, dst_t = rpg.fc_template(N)
    ei_host = rpg.batched_edge_index(src_t, dst_t, G, N).pin_memory()
    x_dev, poses_dev, ei_dev = x_host.to(dev), poses_host.to(dev), ei_host.to(dev)
    mask_rng = np.random.RandomState(7)           # same mask sequence on every rank (one mask per global batch)
    use_mask = train and knn <= 0
    masks = [edge_dropout_keep(H, mask_rng) if use_mask else np.ones(H, bool)
             for _ in range(2 * args.trials * args.steps + 2 * max(args.warmup, 3) + 16)]
    mask_iter = iter(masks)

    def step(x, poses, ei_full):
        keep = next(mask_iter)
        if args.boundary == "attached":           # round-1 shortcut: the caller builds the GraphBatch itself
            graph = GraphBatch.fully_connected(G, N, dev, keep)
            ei = attach(graph.edge_index(), graph)
        elif use_mask:                            # train.py:238-245 on the device, then a FRESH un-annotated edge_index
            ei = rpg.mask_edge_index(ei_full, keep, G)
        else:
            ei = ei_full
        if train:
            if opt is not None:
                opt.zero_grad()
            else:
                bucket.zero()
            pn, pe, ei_used = model(x, ei)                        # ei_used: the rewired graph when knn > 0
            loss, t_loss, q_loss = crit(pe, poses, ei_used)
            loss.backward()
            if opt is not None:
                if skip_reduce:
                    pass
                elif args.overlap_allreduce and world > 1:
                    bucket.finish_allreduce(average=False)
                else:
                    bucket.allreduce(average=False)
                opt.step(grad_scale=1.0 / world)
            else:
                bucket.allreduce()
            return loss
        with torch.no_grad():
            pn, pe, _ = model(x, ei)
        return pe

    # e2e leg: the loop a user writes with the package's own feed (relpose_gnn_b200.feed): every step copies its
    # inputs -- node embeddings, poses AND the int64 edge_index -- host -> device from pinned memory (on a copy stream,
    # overlapping the previous step's kernels), the model receives that edge_index un-annotated (validated on the
    # device, rpg.graph.from_edge_index) and the step's result is read back (training: the loss, through a pinned slot
    # read one step later; inference: the edge poses).
    feeder = rpg.DeviceFeeder(dev)
    readback = rpg.ScalarReadback(1 if train else G * N * (N - 1) * 6)

    def e2e_loop(n_steps):
        results = []
        feeder.stage(x_host, poses_host, ei_host)
        for i in range(n_steps):
            x, poses, ei_full = feeder.take()
            out = step(x, poses, ei_full)
            feeder.release()
            if i + 1 < n_steps:          # staged after this step's own small uploads are queued: they go first on the copy engine
                feeder.stage(x_host, poses_host, ei_host)
            if readback.full():
                results.append(readback.pop())
            readback.push(out.float())
        while readback.pending:
            results.append(readback.pop())
        rpg.check_pending(block=True)
        assert len(results) == n_steps and all(bool(torch.isfinite(r).all()) for r in results)

    def timed(n_steps, e2e):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        if e2e:
            e2e_loop(n_steps)
        else:
            for _ in range(n_steps):
                step(x_dev, poses_dev, ei_dev)
        ev1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return ms.item() / n_steps

    # untimed warm-up: first two steps with the full template (largest activation arenas, so the caching allocator
    # holds blocks big enough for every later edge-dropout mask), then W steps of the real mask sequence
    mask_iter = iter([np.ones(H, bool)] * 2 + masks)
    for _ in range(2 + max(args.warmup, 3)):
        step(x_dev, poses_dev, ei_dev)
    torch.cuda.synchronize()
    rpg.check_pending(block=True)
    sampler = ClockSampler(local_rank)
    if rank == 0 and args.clocks:
        sampler.start()
    launches0 = lib.rpg_launch_count()
    trials = [timed(args.steps, e2e=False) for _ in range(args.trials)]      # each trial: exactly K steps, max over ranks
    launches = (lib.rpg_launch_count() - launches0) / (args.steps * args.trials)
    clocks = sampler.stop() if (rank == 0 and args.clocks) else None
    ms_step = float(np.median(trials))
    e2e_loop(max(args.warmup, 3))             # untimed: allocates the feeder's device slots, touches the pinned buffers
    torch.cuda.synchronize()
    trials_e2e = [timed(args.steps, e2e=True) for _ in range(args.trials)]
    ms_e2e = float(np.median(trials_e2e))

    # roofline leg: per-launch CUDA events on EVERY kernel class during one more step (same stream, no PDL overlap)
    keep_prof = masks[0]
    mask_iter = iter([keep_prof] + masks)
    Ep_prof = N * knn if knn > 0 else 2 * int(keep_prof.sum())
    lib.rpg_profile_begin()
    step(x_dev, poses_dev, ei_dev)
    torch.cuda.synchronize()
    recs = (_lib.ProfRec * 1024)()
    n_rec = C.c_int()
    lib.rpg_profile_records(recs, 1024, C.byref(n_rec))
    recs = [recs[i] for i in range(min(n_rec.value, 1024))]
    if os.environ.get("RPG_BENCH_DUMP_RECORDS") == "1" and rank == 0:      # development aid: the per-launch list on stderr
        for r in recs:
            print(f"rec {_lib.PROF_CLASSES[r.cls]:14s} {r.ms * 1e3:8.1f} us  {r.flops / 1e9:8.2f} GF  {r.bytes / 1e6:8.1f} MB  "
                  f"M={r.M} N={r.N} K={r.K}", file=sys.stderr)
    gemms = [r for r in recs if r.cls <= 1]
    gemm_ms = sum(r.ms for r in gemms)
    gemm_fl = sum(r.flops for r in gemms)
    edge_gemms = [r for r in gemms if r.N > 8 and r.M >= G * Ep_prof]          # the edge-level (MLP) GEMMs of the path
    alg_flops = algorithmic_flops_per_graph(D, N, Ep_prof, train=train) * G
    mean_Ep = float(N * knn) if knn > 0 else float(np.mean([2 * m.sum() for m in masks[:args.steps]]))

    if rank == 0:
        traffic = load_traffic() or {}
        value = world * G / (ms_step * 1e-3)
        e2e_value = world * G / (ms_e2e * 1e-3)
        ei_bytes = ei_host.numel() * 8
        line = {
            "metric": METRIC, "value": value, "unit": "graphs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32 (split-bf16 x3 on the bf16 tensor pipe)" if fp32_mode else "bf16",
            "data": "synthetic",
            "trials_ms_per_step": trials, "trials_e2e_ms_per_step": trials_e2e,
            "config": {"workload": args.workload, "graphs_per_gpu": G, "nodes_per_graph": N, "D": D, "mode": mode,
                       "gnn_recursion": R_ROUNDS, "edge_dropout_keep": 0.5 if use_mask else 1.0, "knn": knn,
                       "mean_edges_per_graph": mean_Ep, "feature_dropout": 0.5,
                       "step": ("edge mask (train.py:238-245) + forward + compute_RP/L1 criterion + backward"
                                + (" + 1 NCCL all-reduce" if world > 1 else "")
                                + (" + Adam step + re-pack of the bf16 operands" if opt is not None else "")) if train else "forward",
                       "boundary": ("fresh int64 edge_index every step -> rpg.graph.from_edge_index (device-side validation + "
                                    f"template tables, validation={args.validation})") if args.boundary == "fresh"
                                   else "GraphBatch attached by the caller (no validation)",
                       "parallelism": f"dp{world} over graphs", "l2": "activations per step (>1 GB) exceed the 126 MB L2; no flush needed"},
            "e2e": {"value": e2e_value, "unit": "graphs/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": x_host.numel() * x_host.element_size() + poses_host.numel() * 4 + ei_bytes,
                    "d2h_bytes_per_step": 4 if train else G * 2 * int(np.mean([m.sum() for m in masks[:4]])) * 6 * 4,
                    "pipeline": "relpose_gnn_b200.DeviceFeeder: pinned H2D of step i+1 (embeddings, poses, int64 edge_index) on a "
                                "copy stream under step i; result read back through a pinned slot one step later"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel<NT|TN> (tcgen05)",
                         "achieved": alg_flops / (gemm_ms * 1e-3) / 1e12, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": alg_flops / (gemm_ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
                         "traffic": traffic.get("dram_bytes_per_launch") if args.workload == "train_4096x9" else None,
                         "traffic_source": f"{traffic.get('file')} (ncu dram__bytes_read+write over the GEMM launches of one train_4096x9 step / launches)",
                         "ncu_tensor_pipe_active_pct_gemm_time": traffic.get("tensor_pipe_active_pct_gemm_time") if args.workload == "train_4096x9" else None,
                         "ncu_hw_counted_tflops_gemm_time": traffic.get("hw_counted_tflops_gemm_time") if args.workload == "train_4096x9" else None,
                         "peak_source": peaks["source"] + ", sustained bf16",
                         "algorithmic_flops_per_launch": alg_flops / max(len(gemms), 1),
                         "avg_launch_ms": gemm_ms / max(len(gemms), 1),
                         "launches_per_step": len(gemms),
                         "executed_tflops": gemm_fl / (gemm_ms * 1e-3) / 1e12,
                         "executed_frac": gemm_fl / (gemm_ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
                         "edge_mlp_gemms": {"launches": len(edge_gemms), "ms": sum(r.ms for r in edge_gemms),
                                            "executed_tflops": sum(r.flops for r in edge_gemms) / max(sum(r.ms for r in edge_gemms), 1e-9) / 1e9,
                                            "executed_frac": sum(r.flops for r in edge_gemms) / max(sum(r.ms for r in edge_gemms), 1e-9) / 1e9 / peaks["bf16_sustained"]},
                         "gemm_ms_per_step": gemm_ms, "gemm_share_of_step": gemm_ms / ms_step,
                         "note": "per-launch events with programmatic dependent launch OFF (kernels serialised); in-step the GEMMs overlap their neighbours' prologues",
                         "edges_per_graph_profiled_step": Ep_prof},
            "roofline_hbm": hbm_roofline(recs, peaks, D),
        }
        if train and skip_reduce:
            line["diagnostic"] = "RPG_BENCH_SKIP_ALLREDUCE=1: replicas did not exchange gradients; not a benchmark value"
        if args.cpu_baseline and world == 1:
            v, med, threads, kind = cpu_reference_throughput(args.cpu_sample, N, D, train, steps=3, warmup=1)
            line["cpu_baseline"] = {"value": v, "unit": "graphs/s", "cores": threads, "kind": kind,
                                    "sample": cpu_sample_note(kind, args.cpu_sample, N, mode, threads, med) + ", median of 3"}
            if args.ref_eager:
                try:
                    from oracle import reference_arm
                    if reference_arm.available():
                        ve, mede = reference_arm.time_reference(D, N, 256, train, edge_dropout=train, steps=3, warmup=2,
                                                                device=str(dev), vector_rp=True)
                        line["reference_eager_b200"] = {
                            "value": ve, "unit": "graphs/s", "ms_per_sample": mede * 1e3,
                            "sample": "the reference modules in torch eager (cuBLAS) on the same B200, fp32, 256 graphs per step "
                                      "(the [Et, D/8, D/8] attention tensors bound the chunk), compute_RP vectorised; informative"}
                except Exception as exc:                     # informative column only
                    line["reference_eager_b200"] = {"unavailable": repr(exc)[:200]}
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--trials", type=int, default=5, help="timed regions of K steps each; the median is reported")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train_4096x9", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=128, help="graphs per CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-clocks", dest="clocks", action="store_false", help="(experiments) skip the nvidia-smi sampler")
    ap.add_argument("--no-optimizer", dest="optimizer", action="store_false",
                    help="(A/B) leave the Adam step + operand re-pack out of the training step")
    ap.add_argument("--boundary", default="fresh", choices=["fresh", "attached"],
                    help="fresh: the model receives an un-annotated int64 edge_index every step (the drop-in boundary); "
                         "attached: the caller builds the GraphBatch and annotates the tensor (A/B)")
    ap.add_argument("--validation", default="async", choices=["async", "sync"],
                    help="edge_index validation read-back: async (checked one step late) or sync (one event wait per step)")
    ap.add_argument("--no-overlap-allreduce", dest="overlap_allreduce", action="store_false",
                    help="(A/B) one all-reduce after the backward instead of starting it from inside the backward")
    ap.add_argument("--no-ref-eager", dest="ref_eager", action="store_false",
                    help="skip the informative reference-in-torch-eager-on-the-GPU column")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON): native libraries that write to file descriptor 1 on their own (NCCL
    # prints its version banner there) are pointed at stderr for the duration of the run.
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    global _emit
    def _emit(text):
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        print(text, flush=True)
        os.dup2(2, 1)
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)


if __name__ == "__main__":
    main()
