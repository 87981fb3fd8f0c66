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
        self.n_steps = 0                                                   # calls of step() = updates of the moving average
        # number of Adam updates actually APPLIED (a skipped step does not count, as in torch.optim.Adam): lives on
        # the device because the skip flag does; two slots, read from one and written to the other each call
        self._applied = torch.zeros(2, dtype=torch.int64, device=dev)
        self._skip = torch.zeros(1, dtype=torch.int32, device=dev)
        self._scale = torch.ones(1, dtype=torch.float32, device=dev)
        self._backup = None
        self._buckets = None
        ops.WEIGHTS_EPOCH += 1

    @property
    def applied_steps(self):
        """number of updates applied so far (host synchronisation; for logging / checkpoints)"""
        return int(self._applied[self.n_steps % 2])

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

    def enable_overlap(self, bucket_bytes=2 << 20):
        """Issue the gradient all-reduce bucket by bucket DURING the backward pass: the flat buffer is cut into
        contiguous ranges of about `bucket_bytes` (parameters are laid out in forward order, the backward finalises them
        roughly back to front); a post-accumulate hook per parameter counts arrivals and, when a bucket is complete,
        launches its asynchronous all-reduce -- NCCL runs it on its own stream, ordered after the gradient kernels
        issued so far, while the rest of the backward keeps the compute stream busy.  `all_reduce()` then only issues
        what is left, waits and scales.  (reference: DistributedDataParallel's bucketed hooks, run/trainer.py:138-139)"""
        if self._buckets is not None:
            return
        self._buckets, self._bucket_of = [], {}
        o = start = 0
        members = []
        for i, p in enumerate(self.params):
            members.append(i)
            o += p.numel()
            if (o - start) * 4 >= bucket_bytes or i == len(self.params) - 1:
                self._buckets.append({"lo": start, "hi": o, "n": len(members), "seen": 0, "work": None})
                for m in members:
                    self._bucket_of[m] = len(self._buckets) - 1
                start, members = o, []
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(lambda _p, i=i: self._arrived(i))

    def _arrived(self, i):
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        b = self._buckets[self._bucket_of[i]]
        b["seen"] += 1
        if b["seen"] == b["n"] and b["work"] is None:
            b["work"] = dist.all_reduce(self.grad[b["lo"]:b["hi"]], op=dist.ReduceOp.SUM, async_op=True)

    def all_reduce(self):
        """averages the gradient buffer over the ranks (NCCL over NVLink on the GPU box, gloo in the CPU tests)"""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            if self._buckets is None:
                dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)
            else:
                for b in self._buckets:                   # same order on every rank: buckets whose hooks did not all fire
                    if b["work"] is None:                 # (parameters without a gradient this step) go now
                        b["work"] = dist.all_reduce(self.grad[b["lo"]:b["hi"]], op=dist.ReduceOp.SUM, async_op=True)
                for b in self._buckets:
                    b["work"].wait()
                    b["work"], b["seen"] = None, 0
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
        src, dst = self._applied[self.n_steps % 2:], self._applied[(self.n_steps + 1) % 2:]
        self.n_steps += 1
        decay = 0.0
        if self.ema is not None:
            decay = self.ema_decay
            if self.ema_use_num_updates:                                   # torch_ema: min(decay, (1 + n) / (10 + n))
                decay = min(decay, (1 + self.n_steps) / (10 + self.n_steps))
        check(lib.e3b_adam_ema_step(ptr(self.param), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self.ema),
                                    self.param.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                    0, float(decay), ptr(scale), ptr(skip), ptr(src), ptr(dst), stream()))
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

    # -- checkpoints (reference Trainer.save: optimiser + EMA + progress; run/trainer.py:632-763) ----------------
    def state_dict(self):
        return {"param": self.param.clone(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "ema": None if self.ema is None else self.ema.clone(), "n_steps": self.n_steps,
                "applied_steps": self.applied_steps, "lr": self.lr}

    def load_state_dict(self, state):
        self.param.copy_(state["param"])
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])
        if self.ema is not None and state.get("ema") is not None:
            self.ema.copy_(state["ema"])
        self.n_steps = int(state["n_steps"])
        self._applied.fill_(int(state["applied_steps"]))
        self.lr = float(state.get("lr", self.lr))
        ops.WEIGHTS_EPOCH += 1

    def ema_state_dict(self, module):
        """the module's state_dict with the averaged weights in place of the raw ones: the deployable model the
        reference saves (Trainer.save_ema_model)"""
        sd = {k: v.clone() for k, v in module.state_dict().items()}
        if self.ema is None:
            return sd
        by_ptr = {}
        o = 0
        for p in self.params:
            by_ptr[p.data_ptr()] = (o, p.numel(), p.shape)
            o += p.numel()
        for name, p in module.named_parameters():
            if p.data_ptr() in by_ptr:
                o, n, shape = by_ptr[p.data_ptr()]
                sd[name] = self.ema[o:o + n].view(shape).clone()
        return sd
