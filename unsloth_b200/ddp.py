"""Data-parallel plumbing for QLoRA on an 8 x B200 NVLink/NVSwitch box.

The reference has no collective of its own: HF Trainer wraps the model in torch DDP and only the
LoRA A/B tensors carry gradients (loader_utils.py:91-106, 849-865; the studio design note
fsdp2_design_notes.md:21-23 states "per-rank replicated bases, DDP on the LoRA params only").
Here the LoRA parameters, their gradients and the AdamW moments live in ONE flat fp32 bucket each:
  * autograd accumulates straight into views of the flat gradient bucket,
  * the exchange is a single NCCL all-reduce of that bucket (168 MB for Llama-3-8B r=16) -- one
    launch sized for latency, not link count (NVSwitch: every peer at full bandwidth),
  * the optimiser is a single streaming launch over the bucket (ub200_adamw_flat).
Frozen NF4 bases are replicated per rank and never communicated.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import _lib as L


def init_distributed():
    """One process per GPU (torchrun env).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)          # loader_utils.py:91-106: rank -> cuda:{LOCAL_RANK}
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


class FlatLoRABucket:
    """Flat fp32 storage for LoRA params / grads / Adam moments + all-reduce + fused AdamW."""

    def __init__(self, params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable LoRA parameters"
        dev = self.params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]     # 16-byte aligned slices
        total = sum(sizes)
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p, sz in zip(self.params, sizes):
            assert p.dtype == torch.float32, "LoRA params are kept in fp32 (models/_utils.py:2482-2496)"
            n = p.numel()
            self.flat_p[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + n].view(p.shape)
            p.grad = self.flat_g[off:off + n].view(p.shape)
            off += sz
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.t = 0
        self.plan = None

    def numel(self):
        return self.flat_p.numel()

    def zero_grad(self):
        self.flat_g.zero_()

    # -- batched step housekeeping (kernels/utils.py::StepPlan) ---------------------------------
    def begin_step(self):
        """zero_grad + (from the second step on) ONE batched refresh of every LoRA cast the step will use;
        until `end_backward()` the projections' d_A / d_B bypass autograd's 448 AccumulateGrad launches."""
        from .kernels import utils as KU
        self.zero_grad()
        if os.environ.get("UB200_STEP_PLAN", "1") in ("0", "off", "false"):     # A/B: per-tensor launches + autograd
            return
        if self.plan is None:
            self.plan = KU.StepPlan(self.params)
        KU.ACTIVE_PLAN = self.plan
        self.plan.in_step = True
        self.plan.refresh()

    def end_backward(self):
        """After `loss.backward()`: add the collected LoRA gradients into the bucket (batched launches)."""
        from .kernels import utils as KU
        if self.plan is not None:
            self.plan.flush_grads()
            self.plan.in_step = False
        if KU.ACTIVE_PLAN is self.plan:
            KU.ACTIVE_PLAN = None

    def broadcast_params(self, src=0):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(self.flat_p, src)

    def all_reduce_grads(self):
        """Sum over ranks (the mean is folded into the optimiser's grad_scale)."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)

    def step(self, grad_scale=None):
        self.t += 1
        world = dist.get_world_size() if dist.is_initialized() else 1
        gs = (1.0 / world) if grad_scale is None else grad_scale
        b1, b2 = self.betas
        if self.flat_p.is_cuda:
            L.call("ub200_adamw_flat", L.ptr(self.flat_p), L.ptr(self.flat_g), L.ptr(self.m),
                   L.ptr(self.v), self.flat_p.numel(), float(self.lr), float(b1), float(b2),
                   float(self.eps), float(self.wd), 1.0 - b1 ** self.t, 1.0 - b2 ** self.t, float(gs),
                   L.stream())
            from .kernels.utils import bump_param_epoch
            bump_param_epoch()      # invalidate the per-step LoRA cast cache
        else:
            raise RuntimeError("unsloth_b200: the optimiser step runs only on CUDA (no CPU fallback)")
