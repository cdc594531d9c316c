"""Host-side graph structure: the implicit per-graph edge template that replaces PyG's generic
gather/scatter indexing.

The reference batches G graphs of N nodes PyG-style (train.py:24,132): node row g*N+n, edge row
g*Ep+k, edge_index[:, g*Ep+k] = template[:, k] + g*N, where the template is the fully connected
enumeration of dataset_7Scenes_multi.py:377-385,418-422 (dataset_Cambridge_multi.py:240-248,273-278),
optionally thinned by the batch-shared edge-dropout mask of train.py:238-242.  `GraphBatch` validates
that property on the device (rpg_validate_edge_index) and holds the small per-template tables
(CSR by destination / source / min / max endpoint, degrees) the kernels index with.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _stream_of(device):
    from . import ops                      # ops does not import this module; resolved lazily to keep it that way
    return ops.stream_of(device)


def fc_template(n_nodes):
    """Closed form of the reference FC enumeration: forward half by offset d,
    then the same list with source and destination swapped.  Returns (src, dst) int32 arrays."""
    src, dst = [], []
    for d in range(1, n_nodes):
        s = np.arange(0, n_nodes - d, dtype=np.int32)
        src.append(s)
        dst.append(s + d)
    src = np.concatenate(src) if src else np.zeros(0, np.int32)
    dst = np.concatenate(dst) if dst else np.zeros(0, np.int32)
    return np.concatenate([src, dst]), np.concatenate([dst, src])


def edge_dropout_keep(n_undirected, rng, keep_factor=0.5):
    """Batch-shared edge-dropout mask, train.py:238-242: undirected edge u survives iff its uniform draw
    is < keep_factor; if none survive, all are kept.  `rng` is a numpy RandomState/Generator-like with
    .random_sample or .random.  Directed rows u and u + n_undirected share bit u."""
    draws = rng.random_sample(n_undirected) if hasattr(rng, "random_sample") else rng.random(n_undirected)
    keep = draws < keep_factor
    if keep.sum() == 0:
        keep = np.ones_like(keep)
    return keep.astype(bool)


def thin_template(src, dst, keep_undirected):
    keep = np.concatenate([keep_undirected, keep_undirected])
    return src[keep], dst[keep]


def batched_edge_index(src, dst, n_graphs, n_nodes, device=None):
    """int64 [2, G*Ep] edge_index of G template copies (what PyG's Batch produces)."""
    t = torch.from_numpy(np.stack([src, dst]).astype(np.int64))
    if device is not None:
        t = t.to(device)
    offs = torch.arange(n_graphs, dtype=torch.long, device=t.device) * n_nodes
    return (t.unsqueeze(1) + offs.view(1, -1, 1)).reshape(2, -1).contiguous()


class _TableUploader:
    """Host -> device upload of the small int32 template tables through rpg_upload_words (the words ride along as
    kernel parameters): asynchronous, no pinned staging, and independent of the copy engines, so a new template every
    step (edge dropout) never waits behind a large input copy."""

    def upload(self, packed_np, device):
        device = torch.device(device)
        if device.type != "cuda" or packed_np.size > 64 * 1024:      # big tables (one-graph fallback): a plain copy
            return torch.from_numpy(packed_np).to(device)
        packed_np = np.ascontiguousarray(packed_np, dtype=np.int32)
        dev = torch.empty(packed_np.size, dtype=torch.int32, device=device)
        lib = _lib.load()
        rc = lib.rpg_upload_words(dev.data_ptr(), packed_np.ctypes.data, packed_np.size,
                                  C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
        if rc:
            raise RuntimeError("rpg_upload_words: " + lib.rpg_last_error_string().decode())
        return dev


_uploader = _TableUploader()


def _csr(keys, n):
    order = np.argsort(keys, kind="stable").astype(np.int32)
    counts = np.bincount(keys, minlength=n)
    ptr = np.zeros(n + 1, np.int32)
    ptr[1:] = np.cumsum(counts)
    return ptr, order


class GraphBatch:
    """G graphs x N nodes sharing one edge template; owns the device tables and the rpg_graph_t struct."""

    _table_cache = {}

    def __init__(self, src, dst, n_graphs, n_nodes, device):
        src = np.ascontiguousarray(src, dtype=np.int32)
        dst = np.ascontiguousarray(dst, dtype=np.int32)
        if src.shape != dst.shape or src.ndim != 1 or src.size == 0:
            raise ValueError("template must be two equal-length 1-D index arrays with at least one edge")
        if src.min() < 0 or dst.min() < 0 or src.max() >= n_nodes or dst.max() >= n_nodes:
            raise ValueError("template node index out of range")
        self.G, self.N, self.Ep = int(n_graphs), int(n_nodes), int(src.size)
        self.device = torch.device(device)
        self.src_np, self.dst_np = src, dst
        key = (self.N, src.tobytes(), dst.tobytes(), str(self.device))
        tables = GraphBatch._table_cache.get(key)
        if tables is None:
            tables = self._build_tables(src, dst, self.N, self.device)
            if len(GraphBatch._table_cache) >= 8:                   # a new dropout mask every step: keep the cache small so
                                                                    # that its buffers recycle inside the caching allocator
                GraphBatch._table_cache.pop(next(iter(GraphBatch._table_cache)))
            GraphBatch._table_cache[key] = tables
        self._tables = tables
        s = _lib.Graph()
        s.G, s.N, s.Ep = self.G, self.N, self.Ep
        for name, t in tables.items():
            if name in ("sel_src", "sel_dst", "_buf"):
                continue
            setattr(s, name, t.data_ptr())
        if "sel_src" in tables:
            s.sel_src, s.sel_dst = tables["sel_src"].data_ptr(), tables["sel_dst"].data_ptr()
            s.sel_patterns, s.sel_div = tables["sel_src"].size(0) // 128, int(np.gcd(128, self.Ep))
        self.struct = s

    @staticmethod
    def _build_tables(src, dst, n, device):
        """All template tables in ONE host buffer / one upload (a training loop sees a new dropout mask every step)."""
        in_ptr, in_idx = _csr(dst, n)
        out_ptr, out_idx = _csr(src, n)
        min_ptr, min_idx = _csr(np.minimum(src, dst), n)
        max_ptr, max_idx = _csr(np.maximum(src, dst), n)
        deg = np.bincount(dst, minlength=n).astype(np.float32)
        host = {"src": src, "dst": dst, "in_ptr": in_ptr, "in_idx": in_idx, "out_ptr": out_ptr,
                "out_idx": out_idx, "inv_deg": (1.0 / np.maximum(deg, 1.0)).astype(np.float32), "deg": deg,
                "has_in": (deg > 0).astype(np.float32),
                "min_ptr": min_ptr, "min_idx": min_idx, "max_ptr": max_ptr, "max_idx": max_idx}
        offs, total = {}, 0
        for k, v in host.items():
            offs[k] = total
            total += (v.size + 3) // 4 * 4                          # keep every table 16-byte aligned
        packed = np.zeros(total, np.int32)
        for k, v in host.items():
            packed[offs[k]:offs[k] + v.size] = np.ascontiguousarray(v).view(np.int32)
        dev_buf = _uploader.upload(packed, device)
        out = {}
        for k, v in host.items():
            t = dev_buf[offs[k]:offs[k] + v.size]
            out[k] = t.view(torch.float32) if v.dtype == np.float32 else t
        out["_buf"] = dev_buf
        # one-hot selection tiles for the K-panel gathers (rpg_gemm_t.gsel), generated on the device.  A 128-row block
        # starting at local edge offset o (a multiple of gcd(128, Ep)) touches graphs 0 .. (o+127)//Ep of its window;
        # the panel holds 64 node rows, so every referenced node column must stay below 64.
        Ep = src.size
        g = int(np.gcd(128, Ep))
        npat = Ep // g
        worst = ((Ep - g + 127) // Ep) * n + n                      # upper bound of (graph index) * N + node + 1
        if worst <= 64 and torch.device(device).type == "cuda":
            lib = _lib.load()
            stream = torch.cuda.current_stream(device).cuda_stream
            for name, tab in (("sel_src", "src"), ("sel_dst", "dst")):
                sel = torch.empty(npat * 128, 64, dtype=torch.bfloat16, device=device)
                _lib.check(lib.rpg_selection_patterns(out[tab].data_ptr(), Ep, n, g, npat, sel.data_ptr(), stream),
                           "rpg_selection_patterns")
                out[name] = sel
        return out

    @classmethod
    def per_graph(cls, edge_index, n_graphs, n_nodes, check=False):
        """Batch whose graphs have different edge sets of equal size (dynamic kNN rewiring, posenet.py:1043-1050): graph g
        owns edge columns [g*Ep, (g+1)*Ep) and nodes [g*N, (g+1)*N).  The kernels see ONE graph of G*N nodes; its tables
        are built on the device (rpg_per_graph_tables), so nothing synchronises unless `check` asks for the validation
        result (edge endpoints inside their graph)."""
        if not edge_index.is_cuda or edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise TypeError("edge_index must be a CUDA int64 tensor of shape [2, E]")
        Et = edge_index.size(1)
        if Et == 0 or Et % n_graphs:
            raise ValueError("per-graph batches need the same number of edges in every graph")
        Ep, dev = Et // n_graphs, edge_index.device
        lib = _lib.load()
        ei = edge_index.contiguous()
        words = lib.rpg_per_graph_tables_words(n_graphs, n_nodes, Ep)
        buf = torch.empty(words + 4, dtype=torch.int32, device=dev)
        buf[words:].zero_()
        bad = buf[words:words + 1]
        _lib.check(lib.rpg_per_graph_tables(ei.data_ptr(), n_graphs, n_nodes, Ep, buf.data_ptr(), bad.data_ptr(),
                                            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "rpg_per_graph_tables")
        if check and int(bad.item()):
            raise ValueError(f"{int(bad.item())} edges connect nodes of different graphs")
        self = cls.__new__(cls)
        Nt = n_graphs * n_nodes
        self.G, self.N, self.Ep, self.device = 1, Nt, Et, dev
        self.src_np = self.dst_np = None
        self._edge_index = ei
        al = lambda v: (v + 3) // 4 * 4                                     # noqa: E731
        tables, off = {"_buf": buf}, 0
        for name, size, is_f in (("src", Et, 0), ("dst", Et, 0), ("in_ptr", Nt + 1, 0), ("in_idx", Et, 0),
                                 ("out_ptr", Nt + 1, 0), ("out_idx", Et, 0), ("min_ptr", Nt + 1, 0), ("min_idx", Et, 0),
                                 ("max_ptr", Nt + 1, 0), ("max_idx", Et, 0), ("inv_deg", Nt, 1), ("deg", Nt, 1),
                                 ("has_in", Nt, 1)):
            t = buf[off:off + size]
            tables[name] = t.view(torch.float32) if is_f else t
            off += al(size)
        self._tables = tables
        s = _lib.Graph()
        s.G, s.N, s.Ep = 1, Nt, Et
        for name, t in tables.items():
            if name != "_buf":
                setattr(s, name, t.data_ptr())
        # One-hot selection tiles, one per 128-row block: a block touches at most ceil(127 / Ep) + 1 member graphs, whose
        # nodes must fit the 64-row panel.  (Edges were validated to stay inside their graph, so no column can leave it.)
        if ((127 + Ep - 1) // Ep + 1) * n_nodes <= 64:
            nblk = (Et + 127) // 128
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for name, tab in (("sel_src", "src"), ("sel_dst", "dst")):
                sel = torch.empty(nblk * 128, 64, dtype=torch.bfloat16, device=dev)
                _lib.check(lib.rpg_selection_patterns_rows(tables[tab].data_ptr(), Et, Ep, n_nodes, sel.data_ptr(),
                                                           buf[words + 1:words + 2].data_ptr(), stream),
                           "rpg_selection_patterns_rows")
                tables[name] = sel
            s.sel_src, s.sel_dst = tables["sel_src"].data_ptr(), tables["sel_dst"].data_ptr()
            s.sel_patterns, s.sel_div = nblk, 0
            s.pg_Ep, s.pg_N = Ep, n_nodes
        self.struct = s
        return self

    @property
    def n_node_rows(self):
        return self.G * self.N

    @property
    def n_edge_rows(self):
        return self.G * self.Ep

    def byref(self):
        return C.byref(self.struct)

    def host_template(self):
        """(src, dst) int32 numpy arrays of the template; device-built batches read them back once (synchronises)."""
        if self.src_np is None:
            self.src_np = self._tables["src"].cpu().numpy().copy()
            self.dst_np = self._tables["dst"].cpu().numpy().copy()
        return self.src_np, self.dst_np

    def edge_index(self):
        """int64 [2, G*Ep] PyG-style batched edge_index, built on the device from the template tables."""
        if getattr(self, "_edge_index", None) is not None:
            return self._edge_index
        if self.device.type != "cuda":
            return batched_edge_index(self.src_np, self.dst_np, self.G, self.N, self.device)
        ei = torch.empty(2, self.G * self.Ep, dtype=torch.int64, device=self.device)
        _lib.check(_lib.load().rpg_build_edge_index(self.byref(), ei.data_ptr(),
                                                    C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                   "rpg_build_edge_index")
        return ei

    def with_graphs(self, n_graphs):
        src, dst = self.host_template()
        return GraphBatch(src, dst, n_graphs, self.N, self.device)

    @classmethod
    def fully_connected(cls, n_graphs, n_nodes, device, keep_undirected=None):
        src, dst = fc_template(n_nodes)
        if keep_undirected is not None:
            src, dst = thin_template(src, dst, np.asarray(keep_undirected, bool))
        return cls(src, dst, n_graphs, n_nodes, device)


def attach(edge_index, graph):
    """Annotates an edge_index tensor with its (already validated) GraphBatch so that modules skip the
    device-side validation and its host read-back."""
    edge_index.rpg_graph = graph
    edge_index.rpg_graph_version = edge_index._version
    return edge_index


def apply_edge_mask(t, keep_undirected, n_graphs):
    """train.py:242-247: `t[surviving_edges]` for a per-edge tensor t [G*E, ...] (edge labels, edge features) with the
    batch-shared undirected keep mask (edge_dropout_keep); both directions of an edge follow the same bit.  CUDA tensor,
    element size 4 bytes or rows that are a multiple of 4 bytes."""
    if not t.is_cuda:
        raise ValueError("apply_edge_mask needs a CUDA tensor: the sm_100a kernels are the only implementation")
    keep = np.asarray(keep_undirected, dtype=bool)
    keep2 = np.concatenate([keep, keep])
    Ep_full = keep2.size
    if t.size(0) != n_graphs * Ep_full:
        raise ValueError("apply_edge_mask: tensor rows must equal graphs x directed edges of the full template")
    idx = np.flatnonzero(keep2).astype(np.int32)
    tc = t.contiguous()
    row_bytes = tc[0].numel() * tc.element_size()
    out = torch.empty((n_graphs * idx.size,) + tuple(tc.shape[1:]), dtype=tc.dtype, device=tc.device)
    idx_dev = _uploader.upload(idx, tc.device)
    _lib.check(_lib.load().rpg_edge_mask_apply(tc.data_ptr(), n_graphs, Ep_full, int(idx.size), idx_dev.data_ptr(), row_bytes,
                                               out.data_ptr(), C.c_void_p(torch.cuda.current_stream(tc.device).cuda_stream)),
               "rpg_edge_mask_apply")
    return out


def mask_edge_index(edge_index, keep_undirected, n_graphs):
    """train.py:238-245 (documented intent: `data.edge_index[:, surviving_edges]`) on the device: the batched
    edge_index [2, G*E] of full templates thinned by the batch-shared undirected keep mask -> [2, G*E_kept].
    One row-selection launch; the mask travels as kernel parameters (no copy engine, no synchronisation)."""
    if not edge_index.is_cuda or edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise TypeError("edge_index must be a CUDA int64 tensor of shape [2, E]")
    keep = np.asarray(keep_undirected, dtype=bool)
    keep2 = np.concatenate([keep, keep])
    Ep_full = keep2.size
    if edge_index.size(1) != n_graphs * Ep_full:
        raise ValueError("mask_edge_index: columns must equal graphs x directed edges of the full template")
    idx = np.flatnonzero(keep2).astype(np.int32)
    ei = edge_index.contiguous()
    out = torch.empty(2, n_graphs * idx.size, dtype=torch.int64, device=ei.device)
    idx_dev = _uploader.upload(idx, ei.device)
    lib = _lib.load()
    stream = _stream_of(ei.device)
    # both rows in ONE launch: [2, G*E] contiguous is 2G "graphs" of E entries, and so is the output
    _lib.check(lib.rpg_edge_mask_apply(ei.data_ptr(), 2 * n_graphs, Ep_full, int(idx.size), idx_dev.data_ptr(), 8,
                                       out.data_ptr(), stream), "rpg_edge_mask_apply")
    return out


def knn_graph(x, k, batch=None, loop=False, num_nodes_per_graph=None):
    """torch_cluster.knn_graph as the reference calls it (posenet.py:1043-1050: `knn_graph(x, k, batch=data.batch,
    loop=False)`): int64 edge_index, edges (neighbour -> centre) grouped by centre, nearest first, equal distances to
    the lower node index.  x: float32 CUDA [n, D].  The graph sizes come from `num_nodes_per_graph` (equally sized
    graphs, no host synchronisation: the model's path) or from the sorted PyG `batch` vector, which may describe graphs
    of different sizes (<= 64 nodes each; a node with fewer than k candidates gets all of them)."""
    if loop:
        raise NotImplementedError("loop=True is never used by the reference")
    if not x.is_cuda:
        raise ValueError("knn_graph needs a CUDA tensor: the sm_100a kernels are the only implementation")
    if k < 1:
        raise ValueError("knn_graph: k must be >= 1")
    xf = x.float().contiguous()
    stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    if num_nodes_per_graph is None and batch is not None:
        if batch.numel() != x.size(0):
            raise ValueError("knn_graph: batch must hold one graph id per row of x")
        b = batch.to(device=x.device, dtype=torch.int64)
        if bool((b[1:] < b[:-1]).any()):
            raise ValueError("knn_graph: batch must be sorted (PyG batches are)")
        counts = torch.bincount(b)
        counts = counts[counts > 0]
        n_max = int(counts.max())
        if not bool((counts == n_max).all()):                      # graphs of different sizes
            if n_max > 64:
                raise ValueError("knn_graph: graphs of more than 64 nodes are not supported")
            node_ptr = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=x.device)
            node_ptr[1:] = counts.cumsum(0)
            edge_ptr = torch.zeros_like(node_ptr)
            edge_ptr[1:] = (counts * torch.clamp(counts - 1, max=k)).cumsum(0)
            n_edges = int(edge_ptr[-1])
            ei = torch.empty(2, n_edges, dtype=torch.int64, device=x.device)
            _lib.check(_lib.load().rpg_knn_graph_ragged(xf.data_ptr(), xf.stride(0), counts.numel(), node_ptr.data_ptr(),
                                                        edge_ptr.data_ptr(), n_max, xf.size(1), k, n_edges, ei.data_ptr(), stream),
                       "rpg_knn_graph_ragged")
            return ei
        num_nodes_per_graph = n_max
    N = int(num_nodes_per_graph) if num_nodes_per_graph is not None else x.size(0)
    if x.size(0) % N:
        raise ValueError("knn_graph: the batch must consist of equally sized graphs")
    G = xf.size(0) // N
    ei = torch.empty(2, G * N * k, dtype=torch.int64, device=x.device)
    _lib.check(_lib.load().rpg_knn_graph(xf.data_ptr(), xf.stride(0), G, N, xf.size(1), k, ei.data_ptr(), stream), "rpg_knn_graph")
    return ei


# ---------------------------------------------------------------------------------------------------------------------
# edge_index -> GraphBatch.  A fresh edge_index (what the PyG loader + train.py:238-245 hand the model every step) takes
# the DEVICE path: the batch shape (G, N) of the previous batch with the same number of node rows is the guess, one
# kernel validates every column against it and extracts the template, a second builds the template tables, two more the
# one-hot selection tiles -- no eager torch ops and no host round trip.  The only thing the host needs back is the
# 4-byte count of violating columns:
#   validation "sync"  (default): read at once through a pinned slot (ONE small event wait); a failed guess falls back to
#                                 the full inference below, so the semantics are those of a cold call.
#   validation "async" (training loops): the count is read when its copy has completed -- at the next call or at
#                                 check_pending() -- and a violation raises ValueError THEN (one step late).
# The first call for a shape infers (G, N) from the tensor itself (eager ops + read-backs; once per shape).
_validation_mode = "sync"
_shape_cache = {}            # (node rows, device) -> (G, N)
_pending = []                # async validation: (event, pinned count, description, keep-alive)
_pinned_counts = []          # recycled pinned 4-byte slots


def set_validation(mode):
    """'sync' or 'async' (see above).  Returns the previous mode."""
    global _validation_mode
    if mode not in ("sync", "async"):
        raise ValueError("validation mode must be 'sync' or 'async'")
    prev, _validation_mode = _validation_mode, mode
    return prev


def check_pending(block=False):
    """Raises ValueError if an asynchronously validated edge_index turned out not to be a batched uniform template.
    block=True waits for every outstanding check (call it before trusting the results of the last steps)."""
    keep = []
    err = None
    for ev, host, what, alive in _pending:
        if block:
            ev.synchronize()
        if ev.query():
            if int(host[0]) and err is None:
                err = f"{what}: {int(host[0])} edge_index columns violate the batched-template property (detected asynchronously)"
            _pinned_counts.append(host)
        else:
            keep.append((ev, host, what, alive))
    _pending[:] = keep
    if err:
        raise ValueError(err)


def _device_graph(ei, G, N, Ep):
    """GraphBatch of G copies of the template found in columns [0, Ep) of `ei`, tables built on the device.
    Returns (graph, bad) with bad = device int32 [1] count of violating columns (not yet read)."""
    lib = _lib.load()
    dev = ei.device
    Et = ei.size(1)
    words = lib.rpg_template_tables_words(N, Ep)
    al = lambda v: (v + 3) // 4 * 4                                       # noqa: E731
    g = int(np.gcd(128, Ep))
    npat = Ep // g
    panels = ((Ep - g + 127) // Ep) * N + N <= 64                         # same feasibility rule as GraphBatch._build_tables
    sel_words = npat * 128 * 32 if panels else 0                          # one [npat*128, 64] bf16 selection matrix
    # ONE allocation per batch: tables | template src | template dst | violation count | two selection matrices.
    # Everything the kernels need is an address inside it (no per-table tensor views on the hot path).
    o_src, o_dst = words, words + al(Ep)
    o_bad = words + 2 * al(Ep)
    o_sel = o_bad + 4
    buf = torch.empty(o_sel + 2 * sel_words, dtype=torch.int32, device=dev)
    base = buf.data_ptr()
    bad = buf[o_bad:o_bad + 1]
    stream = _stream_of(dev)
    _lib.check(lib.rpg_validate_edge_index(ei.data_ptr(), Et, G, N, Ep, base + 4 * o_src, base + 4 * o_dst, bad.data_ptr(),
                                           stream), "rpg_validate_edge_index")
    _lib.check(lib.rpg_template_tables(base + 4 * o_src, base + 4 * o_dst, N, Ep, base, stream), "rpg_template_tables")
    self = GraphBatch.__new__(GraphBatch)
    self.G, self.N, self.Ep, self.device = G, N, Ep, dev
    self.src_np = self.dst_np = None
    self._edge_index = ei
    s = _lib.Graph()
    s.G, s.N, s.Ep = G, N, Ep
    offs, off = {}, 0
    for name, size in (("src", Ep), ("dst", Ep), ("in_ptr", N + 1), ("in_idx", Ep), ("out_ptr", N + 1), ("out_idx", Ep),
                       ("min_ptr", N + 1), ("min_idx", Ep), ("max_ptr", N + 1), ("max_idx", Ep), ("inv_deg", N), ("deg", N),
                       ("has_in", N)):
        setattr(s, name, base + 4 * off)
        offs[name] = (off, size)
        off += al(size)
    if panels:
        for i, (name, tab) in enumerate((("sel_src", "src"), ("sel_dst", "dst"))):
            sel_ptr = base + 4 * (o_sel + i * sel_words)
            _lib.check(lib.rpg_selection_patterns(base + 4 * offs[tab][0], Ep, N, g, npat, sel_ptr, stream), "rpg_selection_patterns")
            setattr(s, name, sel_ptr)
            offs[name] = (o_sel + i * sel_words, sel_words)
        s.sel_patterns, s.sel_div = npat, g
    self._tables = _LazyTables(buf, offs)
    self.struct = s
    return self, bad


class _LazyTables:
    """Tensor views of the device-built template tables, made on demand (host_template, debugging): the hot path only
    uses addresses inside the one buffer that holds them."""

    def __init__(self, buf, offs):
        self.buf, self.offs = buf, offs

    def __getitem__(self, name):
        off, size = self.offs[name]
        t = self.buf[off:off + size]
        if name.startswith("sel_"):
            return t.view(torch.bfloat16).view(-1, 64)            # one-hot selection matrix [patterns * 128, 64]
        return t.view(torch.float32) if name in ("inv_deg", "deg", "has_in") else t


def _read_count_async(bad):
    host = _pinned_counts.pop() if _pinned_counts else torch.empty(1, dtype=torch.int32).pin_memory()
    host.copy_(bad, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(bad.device))
    return ev, host


def from_edge_index(edge_index, n_node_rows):
    """Returns the GraphBatch of a batched edge_index (see the block comment above); raises ValueError if it is not a
    batch of G copies of one per-graph template (no PyG-scatter fallback).  The result is cached as an
    attribute of the tensor."""
    g = getattr(edge_index, "rpg_graph", None)
    if g is not None and getattr(edge_index, "rpg_graph_version", edge_index._version) == edge_index._version:
        if g.n_node_rows != n_node_rows or g.n_edge_rows != edge_index.size(1):
            raise ValueError("attached GraphBatch does not match x / edge_index shapes")
        return g
    if edge_index.dim() != 2 or edge_index.size(0) != 2 or edge_index.dtype != torch.int64:
        raise TypeError("edge_index must be an int64 tensor of shape [2, E]")
    if not edge_index.is_cuda:
        raise ValueError("edge_index must live on the CUDA device (no CPU path exists)")
    if edge_index.size(1) == 0:
        raise ValueError("empty edge_index: the layer needs at least one edge per graph")
    ei = edge_index.contiguous()
    Et = ei.size(1)
    if _pending:
        check_pending()
    key = (int(n_node_rows), str(ei.device))
    guess = _shape_cache.get(key)
    g = None
    if guess is not None and Et % guess[0] == 0 and guess[1] <= 1024 and guess[1] * (Et // guess[0]) <= (1 << 22):
        G_, N_ = guess
        g, bad = _device_graph(ei, G_, N_, Et // G_)
        ev, host = _read_count_async(bad)
        if _validation_mode == "async":
            _pending.append((ev, host, f"edge_index [2, {Et}] as {G_} graphs x {N_} nodes", g))
        else:
            ev.synchronize()
            nbad = int(host[0])
            _pinned_counts.append(host)
            if nbad:
                g = None                              # not that shape: infer from scratch
    if g is None:
        g = _infer_graph(ei, n_node_rows)
        if g.G > 1:
            _shape_cache[key] = (g.G, g.N)
    edge_index.rpg_graph = g                      # cached on the tensor object itself
    edge_index.rpg_graph_version = edge_index._version
    return g


def _infer_graph(ei, n_node_rows):
    """Cold path: infers (G, N, Ep) from the tensor and validates on the device (read-backs; once per batch shape)."""
    Et = ei.size(1)
    # Column Ep of a batched template equals column 0 shifted by N on both rows.  Candidates are the columns
    # with an equal, positive shift on both rows; the first one consistent with (Et, n_node_rows) is validated
    # in full on the device.
    d0 = ei[0] - ei[0, 0]
    d1 = ei[1] - ei[1, 0]
    cand = ((d0 == d1) & (d0 > 0)).nonzero().flatten()
    cand_host = torch.stack([cand, d0[cand]]).cpu().numpy() if cand.numel() else np.zeros((2, 0), np.int64)
    options = []
    for j, shift in zip(cand_host[0].tolist(), cand_host[1].tolist()):
        if Et % j == 0 and n_node_rows % (Et // j) == 0 and n_node_rows // (Et // j) == shift:
            options.append((j, Et // j, shift))
        if len(options) == 4:
            break
    options.append((Et, 1, n_node_rows))          # a single graph
    lib = _lib.load()
    stream = torch.cuda.current_stream(ei.device).cuda_stream
    g, bad = None, None
    for Ep, G, N in options:
        buf = torch.empty(2 * Ep + 1, dtype=torch.int32, device=ei.device)
        _lib.check(lib.rpg_validate_edge_index(ei.data_ptr(), Et, G, N, Ep, buf.data_ptr(), buf[Ep:].data_ptr(),
                                               buf[2 * Ep:].data_ptr(), stream), "rpg_validate_edge_index")
        host = buf.cpu().numpy()
        bad = int(host[2 * Ep])
        if bad == 0:
            g = GraphBatch(host[:Ep].copy(), host[Ep:2 * Ep].copy(), G, N, ei.device)
            break
    if g is None:
        raise ValueError("edge_index is not a batch of G copies of one per-graph edge template with node offset "
                         f"g*N ({bad} violating columns for the last candidate); per-graph edge sets are not "
                         "supported (no PyG-scatter fallback exists)")
    return g
