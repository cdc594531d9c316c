"""CPU-only checks of the host logic and of the C-ABI library's surface (no compute calls, no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import relpose_gnn_b200 as rpg
from oracle import restatement as R
from relpose_gnn_b200 import _lib, graph as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "rpg.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rpg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/rpg.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.rpg_version() >= 100
    assert isinstance(lib.rpg_last_error_string(), bytes)


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    rc = lib.rpg_gemm(None, None)
    assert rc == -1 and b"null" in lib.rpg_last_error_string()
    with pytest.raises(_lib.RpgError):
        _lib.check(rc, "rpg_gemm")


@pytest.mark.parametrize("n", [2, 3, 4, 8, 9, 17])
def test_fc_template_is_the_reference_enumeration(n):
    fx = np.load(os.path.join(GOLD, "fc_enumeration.npz"))
    src, dst = G.fc_template(n)
    assert np.array_equal(np.stack([src, dst]), fx[f"fc_N{n}"])
    ei = G.batched_edge_index(src, dst, 3, n)
    assert torch.equal(ei, R.batched_edge_index(R.fc_edge_index(n), 3, n))


def test_edge_dropout_mask_matches_reference():
    fx = np.load(os.path.join(GOLD, "edge_dropout.npz"))
    for case in range(4):
        n, batch, seed = [int(v) for v in fx[f"case{case}_meta"]]
        keep = G.edge_dropout_keep(n * (n - 1) // 2, np.random.RandomState(seed))
        assert np.array_equal(np.tile(keep, 2 * batch).astype(np.int64), fx[f"case{case}_tiled"])
    class AllOnes:
        def random_sample(self, k):
            return np.ones(k)
    assert G.edge_dropout_keep(5, AllOnes()).all()        # nothing survives -> keep everything (train.py:240-241)


def test_graph_tables():
    src, dst = G.fc_template(5)
    keep = np.array([1, 0, 1, 1, 0, 0, 1, 0, 1, 1], bool)
    s, d = G.thin_template(src, dst, keep)
    g = G.GraphBatch(s, d, 4, 5, "cpu")
    t = g._tables
    for n in range(5):
        ins = t["in_idx"][t["in_ptr"][n]:t["in_ptr"][n + 1]].numpy()
        assert sorted(ins) == sorted(np.nonzero(d == n)[0]) and list(ins) == sorted(ins)   # fixed (ascending) order
        outs = t["out_idx"][t["out_ptr"][n]:t["out_ptr"][n + 1]].numpy()
        assert sorted(outs) == sorted(np.nonzero(s == n)[0])
        lo = t["min_idx"][t["min_ptr"][n]:t["min_ptr"][n + 1]].numpy()
        assert sorted(lo) == sorted(np.nonzero(np.minimum(s, d) == n)[0])
        assert t["deg"][n].item() == (d == n).sum()
        assert t["inv_deg"][n].item() == pytest.approx(1.0 / max(1, (d == n).sum()))
    assert g.n_edge_rows == 4 * s.size and g.n_node_rows == 20
    assert torch.equal(g.edge_index(), R.batched_edge_index(torch.from_numpy(np.stack([s, d]).astype(np.int64)), 4, 5))
    with pytest.raises(ValueError):
        G.GraphBatch(np.array([0, 7]), np.array([1, 2]), 1, 5, "cpu")


def test_state_dict_layout_matches_reference():
    m = rpg.simpleConvEdge_upt(128, 128, 128)
    sd = m.state_dict()
    want = R.LAYER_SHAPES(128)
    assert list(sd.keys()) == list(want.keys())          # same names, same order
    for k, shp in want.items():
        assert tuple(sd[k].shape) == shp, k
    m.load_state_dict(R.synth_params(want, 3, torch.float32))
    s = rpg.RelPoseGNN(128, 128, 128)
    names = set(s.state_dict().keys())
    assert set(R.stack_shapes(128)) <= names


def test_default_init_consumes_rng_like_the_reference_constructor():
    """Same construction order => same default-initialised weights as torch builds for the reference class."""
    from torch.nn import Linear
    torch.manual_seed(11)
    m = rpg.simpleConvEdge_upt(128, 128, 128)
    torch.manual_seed(11)
    D = 128
    ref = [Linear(2 * D, D), Linear(D, D), Linear(2 * D, D), Linear(D, D), Linear(3 * D, D), Linear(D, D),
           Linear(D, D // 8), Linear(D, D // 8), Linear(D, D // 8), Linear(D // 8, D)]
    got = [m.mlp[0], m.mlp[2], m.mlp_updating[0], m.mlp_updating[2], m.edge_model.edge_mlp[0],
           m.edge_model.edge_mlp[2], m.att.g, m.att.theta, m.att.phi, m.att.W]
    for a, b in zip(got, ref):
        assert torch.equal(a.weight, b.weight) and torch.equal(a.bias, b.bias)


def test_constructor_and_input_errors():
    with pytest.raises(AttributeError):
        rpg.simpleConvEdge_upt(128, 128, 128, use_attention=False)
    with pytest.raises(ValueError):
        rpg.simpleConvEdge_upt(128, 256, 128)
    with pytest.raises(ValueError):
        rpg.simpleConvEdge_upt(96, 96, 96)
    m = rpg.simpleConvEdge_upt(128, 128, 128)
    ei = R.batched_edge_index(R.fc_edge_index(4), 2, 4)
    with pytest.raises(ValueError, match="CUDA"):        # no CPU fallback: fails loudly
        m(torch.zeros(8, 128), ei, torch.zeros(24, 128))
    with pytest.raises(ValueError):
        m(torch.zeros(8, 64), ei, torch.zeros(24, 128))
    with pytest.raises(ValueError, match="CUDA"):
        rpg.RelPoseGNN(128, 128, 128)(torch.zeros(8, 128), ei)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "relpose_gnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text, f


def test_evaluation_api_refuses_cpu_tensors():
    """No CPU fallback: the evaluation helpers raise on host tensors."""
    import relpose_gnn_b200 as rpg
    with pytest.raises(ValueError):
        rpg.qexp(torch.zeros(4, 3))
    with pytest.raises(ValueError):
        rpg.compose_query_pose(torch.zeros(72, 6), torch.zeros(9, 6), torch.zeros(2, 72, dtype=torch.long))


def test_sibling_layer_state_dict_layout_matches_reference():
    """simpleConvEdge / simpleConv keep the reference's parameter names and shapes
    (my_gnn_layer.py:242-252, 394-399) and refuse host tensors."""
    D = 128
    m = rpg.simpleConvEdge(D, D, D)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == R.CONV_EDGE_SHAPES(D)
    assert list(m.state_dict().keys())[:4] == ["mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias"]
    c = rpg.simpleConv(D, D)
    assert {k: tuple(v.shape) for k, v in c.state_dict().items()} == R.CONV_SHAPES(D)
    with pytest.raises(ValueError):
        c(torch.zeros(9, D), torch.zeros(2, 72, dtype=torch.long))
    with pytest.raises(ValueError):
        m(torch.zeros(9, D), torch.zeros(2, 72, dtype=torch.long), torch.zeros(72, D))
    with pytest.raises(AttributeError):
        rpg.simpleConvEdge(D, D, D, use_attention=False)


def test_knn_graph_api():
    """the restated torch_cluster semantics (edges neighbour -> centre, grouped by centre, nearest
    first, no self loops) and the no-CPU-fallback rule."""
    x = torch.tensor([[0.0, 0.0], [1.0, 0.0], [3.0, 0.0], [7.0, 0.0]])
    ei = R.knn_graph(x, 2, 1, 4)
    assert ei.tolist() == [[1, 2, 0, 2, 1, 0, 2, 1], [0, 0, 1, 1, 2, 2, 3, 3]]
    with pytest.raises(ValueError):
        rpg.knn_graph(torch.zeros(8, 64), 2, num_nodes_per_graph=4)
    # general batch vectors: every graph on its own, short graphs give all their other nodes, ties to the lower index
    xb = torch.tensor([[0.0], [1.0], [3.0], [7.0], [5.0], [5.0], [9.0]])
    eb = R.knn_graph_batch(xb, 2, torch.tensor([0, 0, 0, 0, 1, 1, 2]))
    assert eb.tolist() == [[1, 2, 0, 2, 1, 0, 2, 1, 5, 4], [0, 0, 1, 1, 2, 2, 3, 3, 4, 5]]
    tie = R.knn_graph(torch.tensor([[0.0], [1.0], [2.0], [1.0]]), 2, 1, 4)
    assert tie[0].view(4, 2).tolist() == [[1, 3], [3, 0], [1, 3], [1, 0]]


def test_save_poses_writes_the_reference_npz_layout(tmp_path):
    """test.py:38-42: keys rel_path, abs_t, abs_q, targ_t, targ_q."""
    pred, targ = np.arange(14.0).reshape(2, 7), np.arange(14.0, 28.0).reshape(2, 7)
    out = tmp_path / "res.npz"
    rpg.save_poses(torch.from_numpy(pred), ["a/0.png", "a/1.png"], out, targ)
    z = np.load(out)
    assert sorted(z.files) == ["abs_q", "abs_t", "rel_path", "targ_q", "targ_t"]
    assert np.array_equal(z["abs_t"], pred[:, :3]) and np.array_equal(z["targ_q"], targ[:, 3:])
    with pytest.raises(ValueError):
        rpg.save_poses(pred, ["only one"], out, targ)


def test_reference_arm_runs_the_unmodified_reference_modules():
    """bench.py --impl reference / cpu_baseline: the reference's own PoseNetX_R2 + compute_RP + criterion (+ Adam) through
    the shim, from /root/reference here or from the staged baseline/_ref on the GPU box."""
    from oracle import reference_arm, stage_reference
    if not reference_arm.available():
        pytest.skip("no reference tree and nothing staged under baseline/_ref")
    step = reference_arm.ReferenceStep(128, 4, 2, train=True, edge_dropout=True)
    assert type(step.model).__name__ == "PoseNetX_R2" and type(step.model.gnn1).__name__ == "simpleConvEdge_upt"
    assert step.model.__class__.__module__ == "niantic.modules.posenet"
    l0 = step()
    assert torch.isfinite(l0).all()
    infer = reference_arm.ReferenceStep(128, 4, 2, train=False, edge_dropout=False)
    pe = infer()
    assert pe.shape == (2 * 12, 6)
    if os.path.isdir("/root/reference/python"):
        assert stage_reference.stage() is not None and stage_reference.staged()
        for f in stage_reference.FILES:
            src = os.path.join("/root/reference/python", f)
            if os.path.isfile(src):
                assert open(src, "rb").read() == open(os.path.join(stage_reference.DST, f), "rb").read()   # unmodified


def test_packed_weight_caches_are_not_pickled_and_can_be_invalidated():
    import copy
    m = rpg.RelPoseGNN(128, 128, 128)
    m._stack_cache["x"] = object()
    m.gnn1._pack_cache["x"] = object()
    m2 = copy.deepcopy(m)
    assert m2._stack_cache == {} and m2.gnn1._pack_cache == {}
    e0, l0 = m._pack_epoch, m.gnn1._pack_epoch
    m.invalidate_packed()
    assert m._pack_epoch == e0 + 1 and m.gnn1._pack_epoch == l0 + 1
    m.load_state_dict(m2.state_dict())
    assert m._pack_epoch == e0 + 2 and m.gnn1._pack_epoch > l0 + 1
    s1 = m._mixed_seed()
    m.dropout_rank = 1
    assert m._mixed_seed() != s1 and m._mixed_seed() % 2 == 0      # replicas draw different dropout masks


def _fake_pyg_classes():
    """Stand-ins carrying torch_geometric 2.0.1's module / class names, so that torch.save writes the pickle a real
    `Data` object would produce (class references + state dicts)."""
    import sys
    import types
    import weakref

    mods = {}
    for name in ("torch_geometric", "torch_geometric.data", "torch_geometric.data.data", "torch_geometric.data.storage"):
        mods[name] = types.ModuleType(name)

    class BaseStorage:                                       # storage.py: state = __dict__ with _parent dereferenced
        def __init__(self, mapping, parent):
            self.__dict__["_mapping"] = dict(mapping)
            self.__dict__["_parent"] = weakref.ref(parent)

        def __getstate__(self):
            out = self.__dict__.copy()
            out["_parent"] = out["_parent"]()
            return out

        def __setstate__(self, mapping):
            self.__dict__.update(mapping)

    GlobalStorage = type("GlobalStorage", (BaseStorage,), {"__module__": "torch_geometric.data.storage"})
    BaseStorage.__module__ = "torch_geometric.data.storage"

    class Data:
        def __init__(self, **kw):
            self.__dict__["_store"] = GlobalStorage(kw, self)
    Data.__module__ = "torch_geometric.data.data"

    class DataV1:                                            # PyG 1.x: plain attributes
        def __init__(self, **kw):
            self.__dict__.update(kw)
    DataV1.__module__ = "torch_geometric.data.data"
    DataV1.__name__ = DataV1.__qualname__ = "Data"
    for cls, nm in ((BaseStorage, "BaseStorage"), (GlobalStorage, "GlobalStorage"), (Data, "Data"), (DataV1, "Data")):
        cls.__name__ = cls.__qualname__ = nm                 # pickle looks classes up by (module, qualified name)
    mods["torch_geometric.data.storage"].GlobalStorage = GlobalStorage
    mods["torch_geometric.data.storage"].BaseStorage = BaseStorage
    return mods, Data, DataV1


def test_reference_graph_files_load_and_batch(tmp_path):
    """dataset_7Scenes_multi.py:437-446 writes Data(x, edge_index, y, edge_attr) pickles; both PyG layouts read back
    without torch_geometric installed, and collate() reproduces PyG batching (train.py:24,132)."""
    import sys
    from relpose_gnn_b200 import graph_io
    mods, Data, DataV1 = _fake_pyg_classes()
    n = 4
    src, dst = G.fc_template(n)
    ei = torch.from_numpy(np.stack([src, dst]).astype(np.int64))
    graphs = []
    for i, cls in enumerate((Data, DataV1, Data)):
        x = torch.randn(n, 12)
        y = torch.randn(n, 6)
        ea = y[ei[1]] - y[ei[0]]                              # dataset_7Scenes_multi.py:425-429
        path = str(tmp_path / f"data_{i:06d}.pt")
        sys.modules.update(mods)
        mods["torch_geometric.data.data"].Data = cls
        try:
            torch.save(cls(x=x, edge_index=ei, y=y, edge_attr=ea), path)
        finally:
            for k in mods:
                sys.modules.pop(k, None)
        assert "torch_geometric" not in sys.modules
        g = graph_io.load_graph(path)
        assert torch.equal(g.x, x) and torch.equal(g.edge_index, ei) and torch.equal(g.y, y) and torch.equal(g.edge_attr, ea)
        graphs.append(g)
    b = graph_io.collate(graphs)
    assert b.x.shape == (3 * n, 12) and b.num_graphs == 3
    assert torch.equal(b.edge_index, G.batched_edge_index(src, dst, 3, n))      # == the template batch the kernels expect
    assert torch.equal(b.batch, torch.arange(3).repeat_interleave(n))
    with pytest.raises(ValueError):
        torch.save({"foo": 1}, str(tmp_path / "bad.pt"))
        graph_io.load_graph(str(tmp_path / "bad.pt"))


def test_fast_params_sees_reassigned_parameters_and_survives_deepcopy():
    """layers.fast_params caches the owning leaf modules (not the parameters): a re-assigned nn.Parameter is picked up at
    once, the cache does not travel through deepcopy / pickle, and invalidate_packed() drops it."""
    import copy
    import pickle
    from relpose_gnn_b200.layers import PARAM_ORDER, fast_params
    m = rpg.simpleConvEdge_upt(128, 128, 128)
    p = fast_params(m, PARAM_ORDER)
    assert list(p) == list(PARAM_ORDER) and all(p[n] is m.get_parameter(n) for n in PARAM_ORDER)
    new_w = torch.nn.Parameter(torch.zeros_like(m.mlp[0].weight))
    m.mlp[0].weight = new_w
    assert fast_params(m, PARAM_ORDER)["mlp.0.weight"] is new_w
    m2 = copy.deepcopy(m)
    assert "_rpg_leaf_cache" not in m2.__dict__ or all(
        leaf is not m.mlp[0] for _, leaf, _ in m2.__dict__["_rpg_leaf_cache"].get(PARAM_ORDER, []))
    assert fast_params(m2, PARAM_ORDER)["mlp.0.weight"] is m2.mlp[0].weight
    m3 = pickle.loads(pickle.dumps(m))
    assert fast_params(m3, PARAM_ORDER)["mlp.2.bias"] is m3.mlp[2].bias
    m.invalidate_packed()
    assert "_rpg_leaf_cache" not in m.__dict__
    e0, v0 = m._pack_epoch, m._value_epoch
    m.mark_values_changed()                      # optimizer-step notification: values only, the structure epoch stays
    assert (m._pack_epoch, m._value_epoch) == (e0, v0 + 1)
    g = rpg.RelPoseGNN(128, 128, 128)
    assert [q is r for q, r in zip(g._ordered_params(), [g.get_parameter(n) for n in g._param_names()])] == [True] * len(g._param_names())
    g.mark_values_changed()
    assert g.gnn1._value_epoch == 1 and g._value_epoch == 1


def test_arena_and_lazy_activation_views():
    """ops.Arena / ops.LazyActs (host logic of the per-step allocations): blocks are 256-byte aligned slices of ONE buffer,
    addresses taken without a tensor become views of the right shape and dtype on first access only, and a request the
    arena cannot hold falls back to its own tensor (kept alive by the arena / the dictionary)."""
    from relpose_gnn_b200 import ops
    ar = ops.Arena(torch.device("cpu"), 4096)
    t = ar.take(3, 5)                                   # bf16 [3, 5] = 30 bytes
    assert t.dtype == torch.bfloat16 and tuple(t.shape) == (3, 5) and t.data_ptr() == ar.base
    p = ar.take_ptr(2, 8, torch.float32)                # next block starts on the next 256-byte boundary
    assert p == ar.base + 256 and ar.off == 512
    acts = ops.LazyActs(ar)
    q = acts.take("h1", 4, 16)                          # 128 bytes at offset 512
    r = acts.take("bits", 4, 2, torch.uint8)
    assert q == ar.base + 512 and r == ar.base + 768
    assert "h1" in acts and "nope" not in acts and acts.get("nope") is None and not dict.__contains__(acts, "h1")
    h1 = acts["h1"]
    assert tuple(h1.shape) == (4, 16) and h1.dtype == torch.bfloat16 and h1.data_ptr() == q
    assert acts["h1"] is h1 and dict.__contains__(acts, "h1")        # materialised once
    assert acts["bits"].dtype == torch.uint8 and acts["bits"].data_ptr() == r
    h1.fill_(1.0)                                       # a view of the arena buffer, not a copy
    assert int(ar.buf[512:512 + 128].view(torch.bfloat16).float().sum()) == 64
    big = acts.take("big", 64, 64, torch.float32)       # 16 KB does not fit the 4 KB arena: own tensor
    assert dict.__contains__(acts, "big") and acts["big"].data_ptr() == big and tuple(acts["big"].shape) == (64, 64)
    spill = ar.take_ptr(64, 64, torch.float32)
    assert len(ar.spill) == 1 and ar.spill[0].data_ptr() == spill
    with pytest.raises(KeyError):
        acts["missing"]
