"""Thin tensor-level wrappers over the C ABI (include/rpg.h).  Every function enqueues on the current
CUDA stream of the tensors' device and returns immediately; there is no CPU or PyTorch fallback."""
import ctypes as C
import threading

import torch

from . import _lib
from ._lib import Gemm, check, ptr

BF16 = torch.bfloat16


_ELEMENT_SIZE = {BF16: 2, torch.float16: 2, torch.float32: 4, torch.float64: 8, torch.uint8: 1, torch.int8: 1,
                 torch.int32: 4, torch.int64: 8}


class Arena:
    """One caching-allocator block per forward (or backward) carved into 256-byte-aligned views.  A training step
    otherwise makes ~60 allocations whose sizes change with the edge-dropout mask; with the host running ahead of
    the device that fragments the caching allocator and ends in synchronous cudaMalloc/cudaFree calls."""

    def __init__(self, device, nbytes):
        self.device = device
        self.buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        self.base = self.buf.data_ptr()
        self.off = 0
        self.spill = []               # keeps overflow blocks of take_ptr alive as long as the arena

    def take_ptr(self, rows, cols, dtype=BF16):
        """Address of a scratch block the caller only ever hands to the library (no tensor view is created)."""
        n = rows * cols * _ELEMENT_SIZE[dtype]
        if self.off + n > self.buf.numel():
            t = torch.empty(rows, cols, dtype=dtype, device=self.device)
            self.spill.append(t)
            return t.data_ptr()
        p = self.base + self.off
        self.off = (self.off + n + 255) & ~255
        return p

    def take(self, rows, cols, dtype=BF16, zero=False):
        n = rows * cols * _ELEMENT_SIZE[dtype]
        if self.off + n > self.buf.numel():                       # estimate too small: fall back to the allocator
            t = torch.empty(rows, cols, dtype=dtype, device=self.device)
        else:
            t = self.buf[self.off:self.off + n].view(dtype).view(rows, cols)
            self.off = (self.off + n + 255) & ~255
        return t.zero_() if zero else t


class LazyActs(dict):
    """Activation dictionary of a layer forward: name -> tensor.  Buffers the library alone touches are recorded as
    (offset, shape, dtype) inside the step's arena and only become tensor views when somebody asks for them (tests,
    debugging, the few entries the next layer consumes): creating ~20 views per layer call was a measurable part of the
    host time of a step."""

    def __init__(self, arena):
        super().__init__()
        self.arena = arena
        self.lazy = {}

    def take(self, name, rows, cols, dtype=BF16):
        """Reserves the block, remembers how to view it, returns its address."""
        n = rows * cols * _ELEMENT_SIZE[dtype]
        ar = self.arena
        if ar.off + n > ar.buf.numel():
            t = torch.empty(rows, cols, dtype=dtype, device=ar.device)
            self[name] = t
            return t.data_ptr()
        self.lazy[name] = (ar.off, rows, cols, dtype)
        p = ar.base + ar.off
        ar.off = (ar.off + n + 255) & ~255
        return p

    def __missing__(self, name):
        off, rows, cols, dtype = self.lazy[name]                 # KeyError for unknown names, as a dict
        t = self.arena.buf[off:off + rows * cols * _ELEMENT_SIZE[dtype]].view(dtype).view(rows, cols)
        self[name] = t
        return t

    def get(self, name, default=None):
        try:
            return self[name]
        except KeyError:
            return default

    def __contains__(self, name):
        return dict.__contains__(self, name) or name in self.lazy


def layer_fwd_bytes(D, Nt, Et):
    c = D // 8
    return (Et * (5 * D * 2 + 3 * c * 4 + 4 * c * 4 + pad64(c) * 2 + 3 * (D // 8)) +
            Nt * (4 * D * 2 + 5 * D * 2 + pad64(c) * 2 + 2 * (D // 8)) + 64 * 256)


def layer_bwd_bytes(D, Nt, Et):
    c = D // 8
    lib = _lib.load()
    return (Et * (5 * D * 2 + pad64(3 * c) * 2) + Nt * (3 * D * 2 + 4 * D * 2 + c * 4 + pad64(c) * 2) +
            4 * lib.rpg_layer_bwd_ws_floats(D, 0, 0) +
            4 * lib.rpg_colsum_scratch_floats(max(Et, Nt), max(D, pad64(3 * c))) + 64 * 256)


class _Tls(threading.local):
    scoped = None           # (device index, c_void_p) while a stream_scope is open on THIS thread


_tls = _Tls()


class stream_scope:
    """Pins the stream handle the wrappers pass to the library for the duration of one top-level call (a stack forward,
    its backward, an optimizer step): `torch.cuda.current_stream()` costs ~5 us and a training step asks ~80 times.  The
    scope is opened by the entry points only, so a caller switching streams between calls is still honoured."""

    def __init__(self, device):
        self.device = device

    def __enter__(self):
        self.prev = _tls.scoped
        _tls.scoped = (self.device.index, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        return self

    def __exit__(self, *exc):
        _tls.scoped = self.prev
        return False


def scoped(fn):
    """Decorator for autograd.Function forward / backward: one stream_scope per call (device of the first CUDA tensor
    argument, remembered on ctx for the backward, whose gradient arguments may be None)."""
    def wrapper(ctx, *args):
        dev = getattr(ctx, "rpg_device", None)
        if dev is None:
            for t in args:
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    dev = t.device
                    break
            ctx.rpg_device = dev
        if dev is None:
            return fn(ctx, *args)
        with stream_scope(dev):
            return fn(ctx, *args)
    wrapper.__name__, wrapper.__doc__ = fn.__name__, fn.__doc__
    return wrapper


def stream_of(device):
    """c_void_p handle of the stream to enqueue on for `device` (the open stream_scope's, else torch's current one)."""
    sc = _tls.scoped
    if sc is not None and device.index == sc[0]:
        return sc[1]
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _stream(t):
    sc = _tls.scoped
    if sc is not None and t.device.index == sc[0]:
        return sc[1]
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def pad64(v):
    return (v + 63) // 64 * 64


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("relpose_gnn_b200 ops need CUDA tensors: the sm_100a kernels are the only implementation")


def gemm_nt(A, B, *, M=None, N=None, segs=None, bias=None, gadd=(), resid=None, row_scale=None, mask=None,
            relu=False, out=None, out_relu=None, out_f32=None, graph=None, block_n=0, mask_bits=None, out_bits=None,
            gpanel=(), resid_lo=None, out_lo=None):
    """C = epilogue(sum_s A_s @ B^T).  A: bf16 [M, K] (or `segs`: list of up to 3 such tensors concatenated along
    K); B: bf16 [N, sum K].  Row pitches are taken from stride(0), so column-sliced views are fine.
    gadd: up to two (tensor [rows, >=N] bf16, 'src'|'dst') pairs added through the graph template."""
    lib = _lib.load()
    segs = segs if segs is not None else [A]
    _require_cuda(B, *segs)
    g = Gemm()
    g.mode = 0
    g.M = M if M is not None else segs[0].size(0)
    g.N = N if N is not None else B.size(0)
    g.n_seg = len(segs)
    for i, a in enumerate(segs):
        g.A[i] = a.data_ptr()
        g.K[i] = a.size(1)
        g.lda[i] = a.stride(0)
    g.B = B.data_ptr()
    g.ldb = B.stride(0)
    g.block_n = block_n
    g.bias = ptr(bias)
    for i, (t, which) in enumerate(gadd):
        g.gadd[i] = t.data_ptr()
        g.gmap[i] = getattr(graph.struct, which)
        g.gadd_ld[i] = t.stride(0)
    if gadd:
        g.Ep, g.Nn = graph.Ep, graph.N
    if resid is not None:
        g.resid, g.resid_ld = resid.data_ptr(), resid.stride(0)
    if row_scale is not None:
        g.row_scale, g.row_scale_mod = row_scale.data_ptr(), row_scale.numel()
    if mask is not None:
        g.mask, g.mask_ld = mask.data_ptr(), mask.stride(0)
    g.relu = int(relu)
    ldo = None
    for t in (out, out_relu):
        if t is not None:
            ldo = t.stride(0) if ldo is None else ldo
            if t.stride(0) != ldo:
                raise ValueError("out and out_relu must share a row pitch")
    g.out, g.out_relu, g.ldo = ptr(out), ptr(out_relu), ldo or 0
    if out_f32 is not None:
        g.out_f32, g.ldo_f32 = out_f32.data_ptr(), out_f32.stride(0)
    for i, (t, which) in enumerate(gpanel):        # gathered adds as one-hot K panels (needs graph selection patterns)
        sel = getattr(graph.struct, "sel_" + which)
        if not sel:
            raise ValueError("this graph template has no selection patterns (row blocks reference > 64 nodes)")
        g.gsel[i], g.gsrc[i], g.gsrc_ld[i] = sel, t.data_ptr(), t.stride(0)
    if gpanel:
        g.n_gseg = len(gpanel)
        g.gsel_patterns, g.gsel_div = graph.struct.sel_patterns, graph.struct.sel_div
        g.gsrc_rows, g.Ep, g.Nn = graph.n_node_rows, graph.Ep, graph.N
    if resid_lo is not None:
        g.resid_lo = resid_lo.data_ptr()
    if out_lo is not None:
        g.out_lo = out_lo.data_ptr()
    if mask_bits is not None:
        g.mask_bits, g.mask_bits_ld = mask_bits.data_ptr(), mask_bits.stride(0)
    if out_bits is not None:
        g.out_bits, g.out_bits_ld = out_bits.data_ptr(), out_bits.stride(0)
    check(lib.rpg_gemm(C.byref(g), _stream(B)), "rpg_gemm")


def gemm_tn_partials(A, B, ws, M, N, splits, block_n=0):
    """Raw split-R TN GEMM (unit tests): ws[s] = A[rows_s, :M]^T @ B[rows_s, :N]."""
    lib = _lib.load()
    g = Gemm()
    g.mode = 1
    g.M, g.N = M, N
    g.A[0], g.lda[0] = A.data_ptr(), A.stride(0)
    g.B, g.ldb = B.data_ptr(), B.stride(0)
    g.R = A.size(0)
    g.splits, g.split_stride = splits, M * N
    g.block_n = block_n
    g.out_f32, g.ldo_f32 = ws.data_ptr(), N
    check(lib.rpg_gemm(C.byref(g), _stream(A)), "rpg_gemm(TN)")


def wgrad_ws(D, device):
    n = _lib.load().rpg_layer_bwd_ws_floats(D, 0, 0)
    return torch.empty(n, dtype=torch.float32, device=device)


def wgrad(A, B, out, ws, M=None, N=None, bias=None):
    """out[M, N] += A[:, :M]^T @ B[:, :N]  (fp32 `out`, pitch from stride(0); deterministic).
    bias (optional, fp32 [M]) += column sums of A, accumulated inside the same kernel."""
    lib = _lib.load()
    M = M if M is not None else A.size(1)
    N = N if N is not None else B.size(1)
    check(lib.rpg_wgrad_bias(A.data_ptr(), A.stride(0), M, B.data_ptr(), B.stride(0), N, A.size(0), ws.data_ptr(),
                             out.data_ptr(), out.stride(0), ptr(bias), _stream(A)), "rpg_wgrad")


def wgrad_blocks(A, B, outs, ws, bias=None):
    """outs[i][M, N] += A[:, i*M:(i+1)*M]^T @ B for the adjacent column blocks of A in ONE product (M = outs[i].size(0));
    bias (optional) += column sums of block 0."""
    lib = _lib.load()
    M, N = outs[0].size(0), B.size(1)
    arr = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
    check(lib.rpg_wgrad_blocks(A.data_ptr(), A.stride(0), M, len(outs), B.data_ptr(), B.stride(0), N, A.size(0), ws.data_ptr(),
                               arr, outs[0].stride(0), ptr(bias), _stream(A)), "rpg_wgrad_blocks")


def pack_weight(src, dst, r0=0, c0=0, rows=None, cols=None, transpose=False):
    lib = _lib.load()
    rows = rows if rows is not None else src.size(0) - r0
    cols = cols if cols is not None else src.size(1) - c0
    check(lib.rpg_pack_weight(src.data_ptr(), src.stride(0), r0, c0, rows, cols, dst.data_ptr(), dst.stride(0),
                              int(transpose), _stream(src)), "rpg_pack_weight")


class PackQueue:
    """Collects rpg_pack_weight windows and converts them in ONE launch (rpg_pack_weights_batch): a training step
    re-packs ~35 operand windows of the layer after every optimizer step.  dst dtype bf16 (operands) or fp32 (plain
    strided copies: head matrices, concatenated biases); lo=True writes the low bf16 plane of the fp32 mode."""

    def __init__(self, record=False):
        self.batch = _lib.PackBatch()
        self.batch.n = 0
        self.stream = None
        self.keep = []
        # record=True keeps every flushed batch: the windows of a re-pack are the same every optimizer step as long as
        # the parameters stay where they are, so the caller replays the recorded descriptors instead of rebuilding them
        self.recorded = [] if record else None

    def add(self, src, dst, r0=0, c0=0, rows=None, cols=None, transpose=False, lo=False):
        _require_cuda(src, dst)
        if src.dtype != torch.float32 or src.dim() != 2 or src.stride(1) != 1:
            raise TypeError("pack source must be a float32 matrix with unit column stride")
        if dst.dtype not in (BF16, torch.float32) or dst.dim() != 2 or dst.stride(1) != 1:
            raise TypeError("pack destination must be a bf16 / float32 matrix with unit column stride")
        if self.batch.n == _lib.PACK_BATCH_MAX:
            self.flush()
        d = self.batch.d[self.batch.n]
        d.src, d.dst = src.data_ptr(), dst.data_ptr()
        d.ld_src, d.r0, d.c0 = src.stride(0), r0, c0
        d.rows = rows if rows is not None else src.size(0) - r0
        d.cols = cols if cols is not None else src.size(1) - c0
        d.ld_dst = dst.stride(0)
        d.transpose, d.dst_f32, d.lo_plane, d.pad_ = int(transpose), int(dst.dtype == torch.float32), int(lo), 0
        self.batch.n += 1
        self.stream = _stream(src)
        self.keep.append((src, dst))

    def add3(self, src, dst, c0=0, cols=None, r0=0, rows=None, transpose=False):
        """dst[:, 0:3*kp] = [W_hi | W_hi | W_lo] of the window (r0, c0, rows, cols) of `src`, or of its transpose (fp32 mode
        operands); kp = dst.size(1) // 3 >= the window's K extent (extra columns stay zero)."""
        cols = cols if cols is not None else src.size(1) - c0
        rows = rows if rows is not None else src.size(0) - r0
        kp = dst.size(1) // 3
        for i, lo in enumerate((False, False, True)):
            self.add(src, dst[:, i * kp:(i + 1) * kp], r0=r0, c0=c0, rows=rows, cols=cols, transpose=transpose, lo=lo)

    def flush(self):
        if self.batch.n:
            check(_lib.load().rpg_pack_weights_batch(C.byref(self.batch), self.stream), "rpg_pack_weights_batch")
            if self.recorded is not None:
                self.recorded.append(self.batch)
                self.batch = _lib.PackBatch()
            self.batch.n = 0
            self.keep = []


def replay_packs(batches, like):
    """Re-runs recorded rpg_pack_weights_batch descriptors on the current stream of `like`'s device."""
    lib, st = _lib.load(), _stream(like)
    for b in batches:
        check(lib.rpg_pack_weights_batch(C.byref(b), st), "rpg_pack_weights_batch")


def to_bf16(t):
    if t.dtype == BF16:
        return t.contiguous()
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32 or bfloat16, got {t.dtype}")
    _require_cuda(t)
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=BF16, device=t.device)
    check(_lib.load().rpg_cast_f32_to_bf16(t.data_ptr(), out.data_ptr(), t.numel(), _stream(t)), "cast")
    return out


def to_f32(t):
    if t.dtype == torch.float32:
        return t
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
    check(_lib.load().rpg_cast_bf16_to_f32(t.data_ptr(), out.data_ptr(), t.numel(), _stream(t)), "cast")
    return out


def attention_fwd(gtp, c, y, y_lo=None, aux=None):
    check(_lib.load().rpg_attention_fwd(gtp.data_ptr(), gtp.size(0), c, y.data_ptr(), y.stride(0), ptr(y_lo), ptr(aux),
                                        _stream(gtp)), "rpg_attention_fwd")


def attention_fwd_bf16(gtp16, c, y):
    """Series attention on bf16 projections [Et, 3c] (dense rows)."""
    check(_lib.load().rpg_attention_fwd_bf16(gtp16.data_ptr(), gtp16.size(0), c, y.data_ptr(), y.stride(0), _stream(gtp16)),
          "rpg_attention_fwd_bf16")


def attention_bwd_bf16(gtp16, dyn, graph, c, dgtp):
    check(_lib.load().rpg_attention_bwd_bf16(gtp16.data_ptr(), dyn.data_ptr(), dyn.stride(0), graph.byref(), gtp16.size(0), c,
                                             dgtp.data_ptr(), dgtp.stride(0), _stream(gtp16)), "rpg_attention_bwd_bf16")


def to_split(t):
    """fp32 [rows, C] -> (hi, lo) bf16 planes with t = hi + lo to ~2^-17 relative."""
    t = t.contiguous()
    if t.dtype != torch.float32:
        raise TypeError("to_split expects float32")
    hi = torch.empty(t.shape, dtype=BF16, device=t.device)
    lo = torch.empty(t.shape, dtype=BF16, device=t.device)
    check(_lib.load().rpg_cast_f32_to_split(t.data_ptr(), hi.data_ptr(), lo.data_ptr(), t.numel(), _stream(t)), "to_split")
    return hi, lo


def from_split(hi, lo):
    out = torch.empty(hi.shape, dtype=torch.float32, device=hi.device)
    check(_lib.load().rpg_split_to_f32(hi.data_ptr(), lo.data_ptr(), out.data_ptr(), hi.numel(), _stream(hi)), "from_split")
    return out


def pack_weight3(src, dst, c0=0, cols=None):
    """dst[:, 0:3*cols] = [W_hi | W_hi | W_lo] of the column window [c0, c0+cols) of the fp32 weight `src`."""
    lib = _lib.load()
    rows = src.size(0)
    cols = cols if cols is not None else src.size(1) - c0
    kp = dst.size(1) // 3
    pack_weight(src, dst[:, 0:kp], c0=c0, cols=cols)
    pack_weight(src, dst[:, kp:2 * kp], c0=c0, cols=cols)
    check(lib.rpg_pack_weight_lo(src.data_ptr(), src.stride(0), 0, c0, rows, cols, dst[:, 2 * kp:].data_ptr(), dst.stride(0),
                                 _stream(src)), "rpg_pack_weight_lo")


def wgrad_split(A, B, out, ws, bias=None):
    """out[M, N] += (A_hi + A_lo)^T (B_hi + B_lo) for plane pairs A = (hi, lo) [R, M], B = (hi, lo) [R, N] (fp32 mode)."""
    lib = _lib.load()
    check(lib.rpg_wgrad_split(A[0].data_ptr(), A[1].data_ptr(), A[0].stride(0), A[0].size(1), B[0].data_ptr(), B[1].data_ptr(),
                              B[0].stride(0), B[0].size(1), A[0].size(0), ws.data_ptr(), out.data_ptr(), out.stride(0), ptr(bias),
                              _stream(out)), "rpg_wgrad_split")


def segment_sum_split(v, graph, which, out):
    """(hi, lo) plane form of segment_sum."""
    s = graph.struct
    check(_lib.load().rpg_segment_sum_split(v[0].data_ptr(), v[1].data_ptr(), v[0].stride(0), getattr(s, which + "_ptr"),
                                            getattr(s, which + "_idx"), None, graph.byref(), v[0].size(1), out[0].data_ptr(),
                                            out[1].data_ptr(), out[0].stride(0), _stream(v[0])), "rpg_segment_sum_split")


def head_bwd_split(dpose, feat, w6, dw_t, dw_q, db_t, db_q, keep=None, seed=0, p_drop=0.0, mask_relu=True):
    """Backward of head_fwd on (hi, lo) features; returns dfeat as a (hi, lo) pair."""
    lib = _lib.load()
    rows, D = feat[0].shape
    dev = feat[0].device
    ws = torch.empty(lib.rpg_head_bwd_ws_floats(rows, D), dtype=torch.float32, device=dev)
    dh = torch.empty(rows, D, dtype=BF16, device=dev)
    dl = torch.empty(rows, D, dtype=BF16, device=dev)
    check(lib.rpg_head_bwd_split(dpose.data_ptr(), feat[0].data_ptr(), feat[1].data_ptr(), feat[0].stride(0), rows, D, ptr(keep),
                                 seed, p_drop, w6.data_ptr(), int(mask_relu), dh.data_ptr(), dl.data_ptr(), D, dw_t.data_ptr(),
                                 dw_q.data_ptr(), db_t.data_ptr(), db_q.data_ptr(), 1, ws.data_ptr(), _stream(dh)),
          "rpg_head_bwd_split")
    return dh, dl


def edge_init_fwd_split(pmm, bias, graph, D, e_hi, e_lo, e_bits=None):
    check(_lib.load().rpg_edge_init_fwd_split(pmm.data_ptr(), pmm.stride(0), bias.data_ptr(), graph.byref(), D, e_hi.data_ptr(),
                                              e_lo.data_ptr(), e_hi.stride(0), ptr(e_bits), _stream(pmm)),
          "rpg_edge_init_fwd_split")


def attention_bwd(gtp, dyn, graph, c, dgtp, aux=None):
    check(_lib.load().rpg_attention_bwd(gtp.data_ptr(), dyn.data_ptr(), dyn.stride(0), graph.byref(), gtp.size(0), c,
                                        dgtp.data_ptr(), dgtp.stride(0), ptr(aux), _stream(gtp)), "rpg_attention_bwd")


def aggregate_mean(z, graph, a):
    check(_lib.load().rpg_aggregate_mean(z.data_ptr(), z.stride(0), graph.byref(), z.size(1), a.data_ptr(),
                                         a.stride(0), _stream(z)), "rpg_aggregate_mean")


def segment_sum(v, graph, which, out, mask=None, scale=None, D=None):
    """out[node] = scale * sum over the template CSR `which` in {'in','out','min','max'} of v rows."""
    s = graph.struct
    D = D if D is not None else v.size(1)
    check(_lib.load().rpg_segment_sum(v.data_ptr(), v.stride(0), ptr(mask), mask.stride(0) if mask is not None else 0,
                                      getattr(s, which + "_ptr"), getattr(s, which + "_idx"), ptr(scale),
                                      graph.byref(), D, out.data_ptr(), out.stride(0), _stream(v)), "rpg_segment_sum")


def segment_sum2(v, graph, which_a, out_a, which_b, out_b, D=None):
    """Two segment sums of the same edge tensor in one pass: which_* in {'in','out','min','max'}."""
    code = {"in": 0, "out": 1, "min": 2, "max": 3}
    D = D if D is not None else v.size(1)
    check(_lib.load().rpg_segment_sum2(v.data_ptr(), v.stride(0), graph.byref(), code[which_a], code[which_b], D,
                                       out_a.data_ptr(), out_a.stride(0), out_b.data_ptr(), out_b.stride(0), _stream(v)),
          "rpg_segment_sum2")


def edge_gather(pa, which_a, graph, out, pb=None, which_b="src", bias=None, relu=False, mask_bits=None, out_bits=None):
    """out[e] = act(pa[node_a(e)] + pb[node_b(e)] + bias) * bit(e); which_x in {'src', 'dst'}."""
    code = {"src": 0, "dst": 1}
    check(_lib.load().rpg_edge_gather(pa.data_ptr(), pa.stride(0), code[which_a], ptr(pb), pb.stride(0) if pb is not None else 0,
                                      code[which_b], ptr(bias), graph.byref(), out.size(1), int(relu), ptr(mask_bits),
                                      out.data_ptr(), out.stride(0), ptr(out_bits), _stream(pa)), "rpg_edge_gather")


def scale_rows(v, scale, mod, out):
    check(_lib.load().rpg_scale_rows(v.data_ptr(), v.stride(0), v.size(0), v.size(1), scale.data_ptr(), mod, out.data_ptr(),
                                     out.stride(0), _stream(v)), "rpg_scale_rows")


def edge_init_fwd(pmm, bias, graph, D, e0, e0_bits=None):
    check(_lib.load().rpg_edge_init_fwd(pmm.data_ptr(), pmm.stride(0), bias.data_ptr(), graph.byref(), D,
                                        e0.data_ptr(), e0.stride(0), ptr(e0_bits), _stream(pmm)), "rpg_edge_init_fwd")


def edge_init_fwd_f32(pmm, bias, graph, D, e_hi, e_lo):
    check(_lib.load().rpg_edge_init_fwd_f32(pmm.data_ptr(), pmm.stride(0), bias.data_ptr(), graph.byref(), D,
                                            e_hi.data_ptr(), e_lo.data_ptr(), e_hi.stride(0), _stream(pmm)),
          "rpg_edge_init_fwd_f32")


def dropout_mask(seed, p_drop, rows, D, device):
    keep = torch.empty(rows, D, dtype=torch.uint8, device=device)
    check(_lib.load().rpg_dropout_mask(seed, p_drop, rows, D, keep.data_ptr(), _stream(keep)), "rpg_dropout_mask")
    return keep


def head_fwd(feat, w6, b6, keep=None, seed=0, p_drop=0.0, feat_lo=None):
    pose = torch.empty(feat.size(0), 6, dtype=torch.float32, device=feat.device)
    check(_lib.load().rpg_head_fwd(feat.data_ptr(), ptr(feat_lo), feat.stride(0), feat.size(0), feat.size(1), ptr(keep), seed,
                                   p_drop, w6.data_ptr(), b6.data_ptr(), pose.data_ptr(), _stream(feat)), "rpg_head_fwd")
    return pose


def head_bwd_tc(dpose, feat_d, bits, w6T_ext, scale, dw_t, dw_q, db_t, db_q, want_dfeat=True):
    """Backward of the tensor-core pose heads (features dropped + rescaled by the producing GEMM); see rpg.h."""
    lib = _lib.load()
    rows, D = feat_d.shape
    dev = feat_d.device
    ws = torch.empty(lib.rpg_head_bwd_tc_ws_floats(D), dtype=torch.float32, device=dev)
    dp16 = torch.empty(rows, 64, dtype=BF16, device=dev)
    dfeat = torch.empty(rows, D, dtype=BF16, device=dev) if want_dfeat else None
    check(lib.rpg_head_bwd_tc(dpose.data_ptr(), feat_d.data_ptr(), feat_d.stride(0), ptr(bits), rows, D, float(scale),
                              w6T_ext.data_ptr(), dp16.data_ptr(), ptr(dfeat), D, dw_t.data_ptr(), dw_q.data_ptr(),
                              db_t.data_ptr(), db_q.data_ptr(), ws.data_ptr(), _stream(feat_d)), "rpg_head_bwd_tc")
    return dfeat


def head_bwd(dpose, feat, w6, dw_t, dw_q, db_t, db_q, keep=None, seed=0, p_drop=0.0, mask_relu=True, want_dfeat=True,
             accumulate=True):
    """Backward of head_fwd; dw_t/db_t (translation head) and dw_q/db_q (rotation head) are accumulated in place."""
    lib = _lib.load()
    rows, D = feat.shape
    ws = torch.empty(lib.rpg_head_bwd_ws_floats(rows, D), dtype=torch.float32, device=feat.device)
    dfeat = torch.empty(rows, D, dtype=BF16, device=feat.device) if want_dfeat else None
    check(lib.rpg_head_bwd(dpose.data_ptr(), feat.data_ptr(), feat.stride(0), rows, D, ptr(keep), seed, p_drop,
                           w6.data_ptr(), int(mask_relu), ptr(dfeat), D, dw_t.data_ptr(), dw_q.data_ptr(),
                           db_t.data_ptr(), db_q.data_ptr(), int(accumulate), ws.data_ptr(), _stream(feat)), "rpg_head_bwd")
    return dfeat


def pose_loss(pred, poses, graph, grad_scale=None, want_target=False, want_grad=True):
    """Returns (sums[2] = L1 sums over (t, q) columns, target or None, dpred or None)."""
    lib = _lib.load()
    Et = pred.size(0)
    ws = torch.empty(lib.rpg_pose_loss_ws_floats(Et), dtype=torch.float32, device=pred.device)
    sums = torch.empty(2, dtype=torch.float32, device=pred.device)
    target = torch.empty(Et, 6, dtype=torch.float32, device=pred.device) if want_target else None
    dpred = torch.empty(Et, 6, dtype=torch.float32, device=pred.device) if want_grad else None
    check(lib.rpg_pose_loss(pred.data_ptr(), poses.data_ptr(), graph.byref(), Et, ptr(grad_scale), ptr(target),
                            sums.data_ptr(), ptr(dpred), ws.data_ptr(), _stream(pred)), "rpg_pose_loss")
    return sums, target, dpred


def pose_criterion(pred, poses, graph, sax, saq):
    """criterion.py:42-60 in two launches: returns (out7 = [sum_t, sum_q, loss, t_loss, q_loss, dloss/dsax, dloss/dsaq],
    dpred = dloss/dpred); sax/saq are 1-element fp32 device tensors."""
    lib = _lib.load()
    Et = pred.size(0)
    ws = torch.empty(lib.rpg_pose_loss_ws_floats(Et), dtype=torch.float32, device=pred.device)
    out7 = torch.empty(7, dtype=torch.float32, device=pred.device)
    dpred = torch.empty(Et, 6, dtype=torch.float32, device=pred.device)
    check(lib.rpg_pose_criterion(pred.data_ptr(), pred.stride(0), poses.data_ptr(), graph.byref(), Et, sax.data_ptr(), saq.data_ptr(), None,
                                 out7.data_ptr(), dpred.data_ptr(), ws.data_ptr(), _stream(pred)), "rpg_pose_criterion")
    return out7, dpred


def colsum(v, out, cols=None, row_w=None, accumulate=True):
    lib = _lib.load()
    rows = v.size(0)
    cols = cols if cols is not None else v.size(1)
    scratch = torch.empty(lib.rpg_colsum_scratch_floats(rows, cols), dtype=torch.float32, device=v.device)
    check(lib.rpg_colsum_bf16(v.data_ptr(), v.stride(0), rows, cols, ptr(row_w), row_w.numel() if row_w is not None else 0,
                              out.data_ptr(), int(accumulate), scratch.data_ptr(), _stream(v)), "rpg_colsum_bf16")
