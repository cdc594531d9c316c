"""Optimizer step of the training loop (train.py:211,274: `torch.optim.Adam(param_list, lr, weight_decay)`,
`optimizer.step()`) over flat buffers: ONE kernel per step instead of a multi-tensor sweep over ~50 parameters.

    bucket = parallel.FlatGradBucket(params)
    opt = FusedAdam(params, lr=1e-4, weight_decay=..., grad_bucket=bucket, modules=[model])
    loop:  opt.zero_grad(); loss.backward(); bucket.allreduce(average=False); opt.step(grad_scale=1 / world)

The parameters are re-pointed at views of one flat fp32 buffer (their values are preserved), so `param.data` and
`state_dict()` keep working.  The kernel writes the parameters without touching their autograd version counters; the
modules listed in `modules` are told to rebuild their packed bf16 operands (`invalidate_packed()`).
"""
import ctypes as C

import torch

from . import _lib


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_bucket=None, modules=()):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise ValueError("FusedAdam needs CUDA parameters: the sm_100a kernel is the only implementation")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.empty(self.numel, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("FusedAdam expects float32 parameters on one device")
            v = self.flat[off:off + p.numel()].view_as(p)
            v.copy_(p.data)
            p.data = v                                   # the parameter now lives in the flat buffer
            off += p.numel()
        if grad_bucket is None:
            from .parallel import FlatGradBucket
            grad_bucket = FlatGradBucket(self.params)
        if [id(p) for p in grad_bucket.params] != [id(p) for p in self.params]:
            raise ValueError("grad_bucket must hold exactly the optimizer's parameters, in the same order")
        self.bucket = grad_bucket
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.t = 0
        self.modules = list(modules)
        for m in self.modules:                           # data pointers changed
            m.invalidate_packed()

    def zero_grad(self, set_to_none=False):
        self.bucket.zero()

    def step(self, grad_scale=1.0):
        self.bucket.reattach(keep=True)
        self.t += 1
        lib = _lib.load()
        stream = C.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
        _lib.check(lib.rpg_adam_step(self.flat.data_ptr(), self.bucket.flat.data_ptr(), self.exp_avg.data_ptr(),
                                     self.exp_avg_sq.data_ptr(), self.numel, self.lr, self.betas[0], self.betas[1], self.eps,
                                     self.weight_decay, float(grad_scale), self.t, stream), "rpg_adam_step")
        for m in self.modules:                           # same storage, new values: the recorded re-pack is replayed
            if hasattr(m, "mark_values_changed"):
                m.mark_values_changed()
            else:
                m.invalidate_packed()

    def state_dict(self):
        return {"t": self.t, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "lr": self.lr,
                "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        self.t = int(sd["t"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr, self.betas, self.eps, self.weight_decay = sd["lr"], tuple(sd["betas"]), sd["eps"], sd["weight_decay"]
