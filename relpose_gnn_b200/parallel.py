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

    `param.grad` is pointed at its slice; autograd (and the fused weight-gradient kernels, RelPoseGNN.attach_grad_bucket)
    then accumulate in place, `zero()` is one memset and `allreduce()` one `all_reduce(SUM)` (+ the 1/world scaling:
    equal shards + mean-reduced loss => the single-GPU gradient up to summation order).

    `optimizer.zero_grad()` / `model.zero_grad()` with the torch >= 2 default `set_to_none=True` (the reference loop
    calls it, train.py:252) detach `param.grad` from the bucket; autograd then allocates fresh gradient tensors.  Both
    `zero()` and `allreduce()` therefore re-attach: a gradient that no longer aliases its slice is copied into it
    (allreduce) or discarded (zero) and `param.grad` is pointed back, so the collective never reduces a stale buffer.
    """

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("FlatGradBucket expects float32 parameters on one device")
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.reattach(keep=False)

    def reattach(self, keep):
        """Points every `param.grad` at its slice again.  keep=True first copies gradients that live elsewhere (fresh
        tensors autograd allocated after a `zero_grad(set_to_none=True)`) into the slice; keep=False drops them.
        Returns the number of parameters that had to be re-pointed."""
        moved = 0
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is not None and g.data_ptr() == v.data_ptr() and g.shape == v.shape:
                continue
            if keep:
                if g is not None:
                    v.copy_(g)
                else:
                    v.zero_()                     # no gradient reached this parameter in this step
            p.grad = v
            moved += 1
        return moved

    def zero(self):
        self.reattach(keep=False)
        self.flat.zero_()

    def allreduce(self, group=None, average=True, async_op=False):
        """SUM over the ranks of `group` (then / world if `average`).  async_op=True returns the work handle of the
        collective (the caller waits, and folds the 1/world into its optimizer step: no extra kernel)."""
        self.reattach(keep=True)
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None if async_op else self.flat
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return work
        if average:
            self.flat.div_(dist.get_world_size(group))
        return self.flat

    # ---- overlapped reduction: the part of the bucket that is final before the backward has finished goes first
    def ranges_excluding(self, late_params):
        """Contiguous [lo, hi) ranges of the flat buffer that do NOT hold gradients of `late_params`."""
        late = {id(p) for p in late_params}
        ranges, off, lo = [], 0, None
        for p in self.params:
            if id(p) in late:
                if lo is not None:
                    ranges.append((lo, off))
                    lo = None
            elif lo is None:
                lo = off
            off += p.numel()
        if lo is not None:
            ranges.append((lo, off))
        return ranges

    def begin_allreduce(self, ranges, group=None):
        """Starts asynchronous all-reduces (SUM) of flat[lo:hi] for the given ranges: the collectives run on NCCL's own
        stream behind everything queued so far, while the caller keeps enqueueing the rest of the backward.  The
        remaining ranges are reduced by the next `allreduce()` call, which also waits for these."""
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        self._pending = [(lo, hi, dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True))
                         for lo, hi in ranges]

    def finish_allreduce(self, group=None, average=True):
        """Reduces whatever `begin_allreduce` left out, waits for everything, scales by 1 / world if `average`."""
        self.reattach(keep=True)
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            self._pending = []
            return self.flat
        done = sorted((lo, hi) for lo, hi, _ in getattr(self, "_pending", []))
        pos = 0
        for lo, hi in done + [(self.numel, self.numel)]:
            if lo > pos:
                dist.all_reduce(self.flat[pos:lo], op=dist.ReduceOp.SUM, group=group)
            pos = max(pos, hi)
        for _, _, work in getattr(self, "_pending", []):
            work.wait()
        self._pending = []
        if average:
            self.flat.div_(dist.get_world_size(group))
        return self.flat

    def nbytes(self):
        return self.numel * 4
