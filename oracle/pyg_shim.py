"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Stand-ins for the three third-party packages the reference imports but which are
not installable in this image (no network):

  * torch_geometric==2.0.1   (requirements-cu111.txt:9)  -> ``MessagePassing``, ``knn_graph``
  * torch_scatter==2.0.8     (requirements-cu111.txt:2)  -> scatter(..., reduce='mean')
  * torch_cluster==1.5.9     (requirements-cu111.txt:6)  -> ``knn_graph``

With these stubs registered, the reference's own first-party files
(`python/niantic/modules/my_gnn_layer.py`, `att.py`, `posenet.py`) import and run
UNMODIFIED from /root/reference, so every first-party line of the hot path executes
as the authors wrote it.  Only ``MessagePassing.propagate`` is restated here, from the
published PyG 2.0.1 behaviour for the dense ``edge_index`` path:

  flow='source_to_target'  =>  (i, j) = (1, 0)
  message arg ``foo_i``  <- foo.index_select(0, edge_index[1])   (destination rows)
  message arg ``foo_j``  <- foo.index_select(0, edge_index[0])   (source rows)
  any other message arg is passed through unchanged
  aggregate('mean'): out[n] = sum_{e: edge_index[1][e]==n} msg[e] / max(1, count[n]),
                     with dim_size = size[1]
  update(aggr_out, **kwargs named in its signature); the default update is the identity.

Call sites in the reference: my_gnn_layer.py:7 (import), :262/:301 (propagate).

This file exists so that `oracle/make_golden.py` can produce fixtures from the real
reference in THIS container.  It cannot travel to the GPU box (no /root/reference there).
"""
import inspect
import os
import sys
import types

import torch

REFERENCE_PYTHON = os.environ.get("RPG_REFERENCE_PYTHON", "/root/reference/python")


class MessagePassing(torch.nn.Module):
    """Dense-edge_index restatement of torch_geometric.nn.conv.MessagePassing (2.0.1)."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=0):
        super().__init__()
        if aggr not in ("add", "mean", "max"):
            raise ValueError(aggr)
        if flow != "source_to_target" or node_dim != 0:
            raise NotImplementedError("shim covers the reference's usage only")
        self.aggr = aggr
        self._msg_params = [p for p in inspect.signature(self.message).parameters]
        self._upd_params = [p for p in inspect.signature(self.update).parameters][1:]

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        msg_kwargs = {}
        for name in self._msg_params:
            if name.endswith("_i"):
                msg_kwargs[name] = kwargs[name[:-2]].index_select(0, dst)
            elif name.endswith("_j"):
                msg_kwargs[name] = kwargs[name[:-2]].index_select(0, src)
            else:
                msg_kwargs[name] = kwargs[name]
        msg = self.message(**msg_kwargs)
        if size is not None:
            dim_size = size[1]
        else:
            dim_size = next(v for v in kwargs.values() if torch.is_tensor(v)).size(0)
        out = self.aggregate(msg, dst, dim_size)
        upd_kwargs = {k: kwargs[k] for k in self._upd_params if k in kwargs}
        return self.update(out, **upd_kwargs)

    def aggregate(self, inputs, index, dim_size):
        out = inputs.new_zeros((dim_size,) + tuple(inputs.shape[1:]))
        if self.aggr == "max":
            idx = index.view(-1, *([1] * (inputs.dim() - 1))).expand_as(inputs)
            return out.scatter_reduce(0, idx, inputs, reduce="amax", include_self=False)
        out = out.index_add(0, index, inputs)
        if self.aggr == "mean":
            count = torch.zeros(dim_size, dtype=inputs.dtype, device=inputs.device)
            count = count.index_add(0, index, torch.ones_like(index, dtype=inputs.dtype))
            out = out / count.clamp(min=1).view(-1, *([1] * (inputs.dim() - 1)))
        return out

    def message(self, x_j):  # PyG default
        return x_j

    def update(self, aggr_out):  # PyG default
        return aggr_out


def knn_graph(x, k, batch=None, loop=False, flow="source_to_target"):
    """torch_cluster.knn_graph restated: edges (neighbour -> centre), grouped by centre,
    nearest first, self excluded unless ``loop``.  Only used by the 'next' kNN row."""
    n = x.size(0)
    if batch is None:
        batch = x.new_zeros(n, dtype=torch.long)
    d = torch.cdist(x, x)
    d = d.masked_fill(batch.view(-1, 1) != batch.view(1, -1), float("inf"))
    if not loop:
        d = d + torch.diag(x.new_full((n,), float("inf")))
    nbr = d.topk(k, dim=1, largest=False).indices            # [n, k]
    centre = torch.arange(n, device=x.device).view(-1, 1).expand_as(nbr)
    return torch.stack([nbr.reshape(-1), centre.reshape(-1)], 0)


def install():
    """Register the stub packages and put the reference's python/ on sys.path."""
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tg_nn = types.ModuleType("torch_geometric.nn")
        tg_conv = types.ModuleType("torch_geometric.nn.conv")
        tg_conv.MessagePassing = MessagePassing
        tg_nn.conv = tg_conv
        tg_nn.knn_graph = knn_graph
        tg_nn.MessagePassing = MessagePassing
        tg.nn = tg_nn
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.nn"] = tg_nn
        sys.modules["torch_geometric.nn.conv"] = tg_conv
    if "torch_cluster" not in sys.modules:
        tc = types.ModuleType("torch_cluster")
        tc.knn_graph = knn_graph
        sys.modules["torch_cluster"] = tc
    if REFERENCE_PYTHON not in sys.path:
        sys.path.insert(0, REFERENCE_PYTHON)


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_PYTHON, "niantic", "modules", "my_gnn_layer.py"))


def import_reference():
    """Returns (my_gnn_layer, att, posenet) modules of the unmodified reference."""
    install()
    from niantic.modules import att, my_gnn_layer, posenet  # noqa: E402
    return my_gnn_layer, att, posenet
