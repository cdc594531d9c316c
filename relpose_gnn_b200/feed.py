"""Host -> device input feed for the training / inference loop.

The reference hands the GNN a PyG batch that the DataLoader already moved to the GPU with a blocking `.to(device)`
(train.py:132-135).  Here the copy of batch i+1 runs on its own CUDA stream from pinned memory while batch i is being
computed, and the scalar loss comes back through a pinned slot that is read one step later, so neither direction of the
PCIe traffic ever stalls the compute stream.  Plumbing only: no arithmetic happens here.
"""
import torch


class DeviceFeeder:
    """Double-buffered pinned-host -> device copies on a side stream.

        feeder.stage(x_host, poses_host)            # enqueue the copy of the first batch
        loop:
            x, poses = feeder.take()                # compute stream waits for that copy (no host sync)
            ... launch the step on x, poses ...
            feeder.release()                        # the slot may be overwritten once the step has consumed it
            feeder.stage(next_x, next_poses)        # AFTER the step is queued, so the step's own small uploads (graph
                                                    # tables) are ahead of the big copy on the copy engine
    """

    def __init__(self, device, depth=2):
        device = torch.device(device)
        if device.type != "cuda":
            raise ValueError("DeviceFeeder needs a CUDA device")
        self.device = device
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device)
        self.slots = [None] * depth                  # device tensors per slot
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [None] * depth                   # recorded on the compute stream when the consumer is done
        self.head = 0                                # next slot to stage
        self.tail = 0                                # next slot to take
        self.pending = 0
        self.taken = None
        self.h2d_bytes = 0

    def stage(self, *host_tensors):
        if self.pending == self.depth:
            raise RuntimeError("DeviceFeeder: all slots are staged; take() one first")
        k = self.head
        for t in host_tensors:
            if not t.is_pinned():
                raise ValueError("DeviceFeeder: host tensors must be pinned (a pageable copy synchronises the stream)")
        bufs = self.slots[k]
        if bufs is None or len(bufs) != len(host_tensors) or any(
                b.shape != t.shape or b.dtype != t.dtype for b, t in zip(bufs, host_tensors)):
            bufs = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host_tensors]
            self.slots[k] = bufs
            # A fresh block of the caching allocator may be a recycled one whose previous user (a kernel still queued on
            # the compute stream) has not run yet: the copy stream must not write into it before that work is done.
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.copy_stream):
            if self.free[k] is not None:
                self.copy_stream.wait_event(self.free[k])
            for b, t in zip(bufs, host_tensors):
                # chunks of <= 8 MB: small uploads of the compute stream (graph tables) can slip in between on the copy engine
                rows = t.size(0) if t.dim() else 1
                per = max(1, (8 << 20) // max(1, t[0].numel() * t.element_size())) if t.dim() else 1
                if t.dim() == 0 or rows <= per:
                    b.copy_(t, non_blocking=True)
                else:
                    for r0 in range(0, rows, per):
                        b[r0:r0 + per].copy_(t[r0:r0 + per], non_blocking=True)
                self.h2d_bytes += t.numel() * t.element_size()
            self.ready[k].record(self.copy_stream)
        self.head = (k + 1) % self.depth
        self.pending += 1

    def take(self):
        if self.pending == 0:
            raise RuntimeError("DeviceFeeder: nothing staged")
        k = self.tail
        torch.cuda.current_stream(self.device).wait_event(self.ready[k])
        self.tail = (k + 1) % self.depth
        self.pending -= 1
        self.taken = k
        return tuple(self.slots[k])

    def release(self):
        if self.taken is None:
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[self.taken] = ev
        self.taken = None


class ScalarReadback:
    """Pipelined device -> host read of one small tensor per step (the loss): `push` enqueues the copy into a pinned
    slot, `pop` returns the oldest value once its copy has finished (blocking only on that copy)."""

    def __init__(self, numel=1, depth=2, dtype=torch.float32):
        self.host = [torch.empty(numel, dtype=dtype).pin_memory() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.depth = depth
        self.head = self.tail = self.pending = 0
        self.d2h_bytes = 0
        self.copy_stream = None                      # created on first use (per device)

    def push(self, t):
        """The copy runs on a side stream behind everything queued on the current stream so far: large results
        (inference poses) leave the device while the next step computes."""
        if self.pending == self.depth:
            raise RuntimeError("ScalarReadback: pop() before pushing more")
        k = self.head
        src = t.detach().reshape(-1)
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(t.device)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(t.device))
        src.record_stream(self.copy_stream)          # the caching allocator must not recycle it under the copy
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            self.host[k].copy_(src, non_blocking=True)
            self.done[k].record(self.copy_stream)
        self.d2h_bytes += self.host[k].numel() * self.host[k].element_size()
        self.head = (k + 1) % self.depth
        self.pending += 1

    def pop(self):
        if self.pending == 0:
            raise RuntimeError("ScalarReadback: nothing pending")
        k = self.tail
        self.done[k].synchronize()
        self.tail = (k + 1) % self.depth
        self.pending -= 1
        return self.host[k].clone()

    def full(self):
        return self.pending == self.depth
