"""Reader for the reference's pre-generated graph files and PyG-style batching of them.

The reference datasets write one graph per file, `processed/data_{idx:06d}.pt`, with
`torch.save(Data(x=..., edge_index=..., y=..., edge_attr=...), path)` (dataset_7Scenes_multi.py:437-446,
dataset_Cambridge_multi.py) and read it back with `torch.load` (:453-456): a pickled
`torch_geometric.data.Data` object (torch-geometric 2.0.1, requirements-cu111.txt:9).  torch_geometric is not
installable here, and unpickling needs its classes -- so the reader maps every `torch_geometric.*` class in the pickle to
a plain attribute bag and extracts the four tensors from the two layouts PyG has used:

  * PyG >= 2.0: `Data.__dict__ = {'_store': GlobalStorage}`, `GlobalStorage.__dict__ = {'_mapping': {x, edge_index, y,
    edge_attr}, '_parent': Data}`;
  * PyG 1.x:    `Data.__dict__ = {x, edge_index, y, edge_attr, ...}` directly.

`collate` is PyG's `Batch.from_data_list` for these graphs (train.py:24,132): node tensors concatenated, `edge_index`
offset by the running node count, plus the `batch` vector.  Host-side plumbing only (no arithmetic): the tensors then go
to the device (DeviceFeeder) and `edge_index` straight into the model.

PARITY NOTE: the byte layout is restated from the published PyG sources, not pinned to a file written by the real
package (it cannot be installed here); tests/test_host_cpu.py writes both layouts with stand-in classes carrying PyG's
module / class names and reads them back.
"""
import io
import pickle
import types

import torch


class _Bag:
    """Stands in for any torch_geometric class found in a pickle: keeps whatever state the pickle sets."""

    def __init__(self, *args, **kwargs):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in ("torch_geometric", "torch_sparse", "torch_scatter", "torch_cluster"):
            return type(name, (_Bag,), {"__module__": module})
        return super().find_class(module, name)


_pickle_module = types.ModuleType("relpose_gnn_b200._pyg_pickle")
_pickle_module.Unpickler = _Unpickler
_pickle_module.load = lambda f, **kw: _Unpickler(f, **kw).load()
_pickle_module.loads = lambda b, **kw: _Unpickler(io.BytesIO(b), **kw).load()
_pickle_module.__name__ = "pickle"


class GraphData(types.SimpleNamespace):
    """x [N, F], edge_index [2, E] int64, y [N, 6], edge_attr [E, 6] or None -- the fields the reference stores."""


_FIELDS = ("x", "edge_index", "y", "edge_attr")


def _fields_of(obj):
    d = getattr(obj, "__dict__", {})
    store = d.get("_store")
    if store is not None:                               # PyG >= 2.0
        mapping = getattr(store, "__dict__", {}).get("_mapping")
        if mapping is None and isinstance(store, dict):
            mapping = store
        if mapping is None:
            raise ValueError("unrecognised torch_geometric Data layout: _store without _mapping")
        return {k: mapping.get(k) for k in _FIELDS}
    if any(k in d for k in _FIELDS):                    # PyG 1.x
        return {k: d.get(k) for k in _FIELDS}
    if isinstance(obj, dict) and any(k in obj for k in _FIELDS):
        return {k: obj.get(k) for k in _FIELDS}
    raise ValueError(f"not a torch_geometric Data pickle: {type(obj).__name__}")


def load_graph(path):
    """One `processed/data_*.pt` file of the reference -> GraphData (CPU tensors)."""
    obj = torch.load(path, map_location="cpu", pickle_module=_pickle_module, weights_only=False)
    f = _fields_of(obj)
    if f["x"] is None or f["edge_index"] is None:
        raise ValueError(f"{path}: graph file without x / edge_index")
    ei = f["edge_index"]
    if ei.dim() != 2 or ei.size(0) != 2:
        raise ValueError(f"{path}: edge_index must be [2, E]")
    return GraphData(x=f["x"], edge_index=ei.long(), y=f["y"], edge_attr=f["edge_attr"])


def collate(graphs, pin_memory=False):
    """PyG `Batch.from_data_list` for GraphData objects: x / y / edge_attr concatenated, edge_index offset by the running
    node count, `batch` [sum N] = graph id per node.  With pin_memory the result is ready for DeviceFeeder.stage."""
    if not graphs:
        raise ValueError("collate: empty list")
    xs, ys, eas, eis, batch = [], [], [], [], []
    off = 0
    for i, g in enumerate(graphs):
        n = g.x.size(0)
        xs.append(g.x)
        eis.append(g.edge_index + off)
        if g.y is not None:
            ys.append(g.y)
        if g.edge_attr is not None:
            eas.append(g.edge_attr)
        batch.append(torch.full((n,), i, dtype=torch.long))
        off += n
    out = GraphData(x=torch.cat(xs), edge_index=torch.cat(eis, dim=1).contiguous(),
                    y=torch.cat(ys) if len(ys) == len(graphs) else None,
                    edge_attr=torch.cat(eas) if len(eas) == len(graphs) else None,
                    batch=torch.cat(batch), num_graphs=len(graphs))
    if pin_memory:
        for k in ("x", "edge_index", "y", "edge_attr", "batch"):
            t = getattr(out, k)
            if t is not None:
                setattr(out, k, t.pin_memory())
    return out
