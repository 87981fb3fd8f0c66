"""Flat-buffer training state: the trainable parameters of a module and their gradients are
re-laid as views of ONE fp32 buffer each, so that

* the data-parallel gradient exchange is one all-reduce of the gradient buffer, with no gather / scatter copies
  (reference: DistributedDataParallel buckets, run/trainer.py:138-139);
* gradient clipping is one norm over that buffer;
* the optimiser step -- Adam with the reference's settings plus the exponential moving average of the parameters
  (torch_ema in the reference) -- is ONE kernel (``e3b_adam_ema_step``) instead of ~10 multi-tensor launches, and
  skipping a step with non-finite gradients (run/sde_utils.py:240-246) needs no host round trip.
"""
import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import check, count_launch, ptr, stream


class FlatAdam:
    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, ema_decay=None,
                 ema_use_num_updates=True):
        self.params = [p for p in module.parameters() if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        assert all(p.dtype == torch.float32 and p.device == dev for p in self.params), "flat state needs fp32 parameters on one device"
        total = sum(p.numel() for p in self.params)
        self.param = torch.empty(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            n = p.numel()
            self.param[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.param[o:o + n].view(p.shape)                     # the module now computes on the flat buffer
            p.grad = self.grad[o:o + n].view(p.shape)                      # autograd accumulates in place into the views
            o += n
        self.exp_avg = torch.zeros_like(self.param)
        self.exp_avg_sq = torch.zeros_like(self.param)
        self.ema = self.param.clone() if ema_decay is not None else None
        self.ema_decay, self.ema_use_num_updates = ema_decay, ema_use_num_updates
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        self.n_steps = 0
        self._skip = torch.zeros(1, dtype=torch.int32, device=dev)
        self._scale = torch.ones(1, dtype=torch.float32, device=dev)
        self._backup = None
        ops.WEIGHTS_EPOCH += 1

    def zero_grad(self):
        self.grad.zero_()
        for p, view in zip(self.params, self._grad_views()):               # a backward with set_to_none semantics may have
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():      # replaced the view: restore the aliasing
                p.grad = view

    def _grad_views(self):
        o = 0
        for p in self.params:
            n = p.numel()
            yield self.grad[o:o + n].view(p.shape)
            o += n

    def all_reduce(self):
        """averages the gradient buffer over the ranks (NCCL over NVLink on the GPU box, gloo in the CPU tests)"""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)
            self.grad /= dist.get_world_size()

    def step(self, max_grad_norm=None, skip_nonfinite=False):
        """one Adam (+ EMA) update; everything stays on the device"""
        lib = _lib.load()
        _lib.require_cuda(self.param)
        scale = skip = None
        if max_grad_norm is not None or skip_nonfinite:
            norm = torch.linalg.vector_norm(self.grad)
            if max_grad_norm is not None:                                  # torch.nn.utils.clip_grad_norm_ semantics
                torch.clamp(max_grad_norm / (norm + 1e-6), max=1.0, out=self._scale[0])
                scale = self._scale
            if skip_nonfinite:
                self._skip.copy_((~torch.isfinite(norm)).to(torch.int32).reshape(1))
                skip = self._skip
        self.n_steps += 1
        decay = 0.0
        if self.ema is not None:
            decay = self.ema_decay
            if self.ema_use_num_updates:                                   # torch_ema: min(decay, (1 + n) / (10 + n))
                decay = min(decay, (1 + self.n_steps) / (10 + self.n_steps))
        check(lib.e3b_adam_ema_step(ptr(self.param), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self.ema),
                                    self.param.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                    self.n_steps, float(decay), ptr(scale), ptr(skip), stream()))
        count_launch()
        ops.WEIGHTS_EPOCH += 1                                             # packed tensor-core weights must be rebuilt

    # -- EMA weights for evaluation (torch_ema store / copy_to / restore) ----------------------------------
    def ema_swap_in(self):
        assert self.ema is not None and self._backup is None
        self._backup = self.param.clone()
        self.param.copy_(self.ema)
        ops.WEIGHTS_EPOCH += 1

    def ema_swap_out(self):
        self.param.copy_(self._backup)
        self._backup = None
        ops.WEIGHTS_EPOCH += 1
