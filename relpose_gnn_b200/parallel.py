"""Data parallelism over graphs: one process per GPU, contiguous blocks of graphs per rank, replicated weights.

Graphs are independent units (PyG batching only offsets indices), so inference needs no
communication at all and training needs exactly ONE all-reduce per step over a single flat fp32 bucket that
holds every gradient of the path (gnn1 + proj_edge + heads; 3.29 M parameters = 13.1 MB at D=512).
"""
import os

import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous block [lo, hi) of `n_total` graphs owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_distributed(backend=None):
    """Rendezvous from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns
    (rank, local_rank, world).  A single process without the environment is world size 1 (no process group)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


class FlatGradBucket:
    """All gradients of `params` as views of one flat fp32 buffer, so a training step issues a single collective.

    `param.grad` is pointed at its slice once; autograd then accumulates in place, `zero()` is one memset and
    `allreduce()` one `all_reduce(SUM)` followed by the 1/world scaling (equal shards + mean-reduced loss =>
    the single-GPU gradient up to summation order)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("FlatGradBucket expects float32 parameters on one device")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce(self, group=None, average=True):
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return self.flat
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.div_(dist.get_world_size(group))
        return self.flat

    def nbytes(self):
        return self.numel * 4
