"""Activation-memory policies of the training step: the recompute / offload /
tiling options the reference exposes through `use_gradient_checkpointing` and `unsloth_tiled_mlp`
(models/_utils.py:360-386 `apply_unsloth_gradient_checkpointing`; models/llama.py:1169-1192 per-layer
`torch.utils.checkpoint`; loader.py:1106-1111 `patch_tiled_mlp`; the "unsloth" offloaded checkpointer
and the tiled MLP themselves live in the external unsloth_zoo and are restated here from their
published behaviour).

  * `True`       recompute every decoder layer in the backward (torch.utils.checkpoint, non-reentrant);
  * `"unsloth"`  the same recompute, but the ONE tensor kept per layer -- the layer input [B,S,H] --
                 is copied to pinned host memory on a side stream during the forward and brought back
                 just before its layer's backward (NVLink-C2C / PCIe copy overlapped with compute), so
                 device memory holds O(1) layer inputs instead of n_layers;
  * tiled MLP    the MLP runs over `n` token shards under no_grad and is recomputed shard by shard in
                 the backward: the [T, I] gate / up activations (the largest tensors of a layer) never
                 exist for more than T/n tokens.
On a 180 GB B200 none of this is needed for the BASELINE configs (cfg2 peaks at 47 GiB); it is what
lets sequence lengths grow beyond them (bench.py --gradient-checkpointing unsloth --seq ...).
"""
from __future__ import annotations

import torch

_SIDE = {}


def _side_stream(device):
    st = _SIDE.get(device.index)
    if st is None:
        st = _SIDE[device.index] = torch.cuda.Stream(device=device)
    return st


class OffloadedCheckpoint(torch.autograd.Function):
    """Run `fn(hidden, *args)` without saving activations; keep only `hidden`, on the host."""

    @staticmethod
    def forward(ctx, fn, anchor, hidden, *args):
        # `anchor`: an empty tensor that requires grad, so that this node is part of the graph even
        # when `hidden` itself does not require grad (first layer: the trainable LoRA parameters are
        # reached through `fn`'s closure, not through the inputs)
        ctx.fn, ctx.args = fn, args
        ctx.device = hidden.device
        if hidden.is_cuda:
            side = _side_stream(hidden.device)
            side.wait_stream(torch.cuda.current_stream(hidden.device))
            with torch.cuda.stream(side):
                host = torch.empty(hidden.shape, dtype=hidden.dtype, device="cpu", pin_memory=True)
                host.copy_(hidden, non_blocking=True)
            hidden.record_stream(side)
            ctx.event = torch.cuda.Event()
            ctx.event.record(side)
        else:                                   # CPU (tests through the ABI emulator)
            host = hidden.detach().clone()
            ctx.event = None
        ctx.host = host
        with torch.no_grad():
            out = fn(hidden, *args)
        return out

    @staticmethod
    def backward(ctx, *grads):
        if ctx.event is not None:
            side = _side_stream(ctx.device)
            with torch.cuda.stream(side):
                side.wait_event(ctx.event)
                hidden = ctx.host.to(ctx.device, non_blocking=True)
            torch.cuda.current_stream(ctx.device).wait_stream(side)
        else:
            hidden = ctx.host.clone()
        ctx.host = None
        hidden.requires_grad_(True)
        with torch.enable_grad():
            out = ctx.fn(hidden, *ctx.args)
        outs = out if isinstance(out, (tuple, list)) else (out,)
        pairs = [(o, g) for o, g in zip(outs, grads) if g is not None and o.requires_grad]
        torch.autograd.backward([o for o, _ in pairs], [g for _, g in pairs])
        return (None, None, hidden.grad) + (None,) * len(ctx.args)


def _anchor(device):
    return torch.empty(0, device=device, requires_grad=True)


def offloaded_checkpoint(fn, hidden, *args):
    return OffloadedCheckpoint.apply(fn, _anchor(hidden.device), hidden, *args)


class TiledMLP(torch.autograd.Function):
    """mlp_fn(X) over `n_shards` token shards; nothing but X is kept for the backward."""

    @staticmethod
    def forward(ctx, mlp_fn, anchor, X, n_shards):
        ctx.mlp_fn, ctx.n_shards = mlp_fn, n_shards
        shape = X.shape
        X2 = X.reshape(-1, shape[-1])
        ctx.save_for_backward(X2)
        ctx.shape = shape
        outs = []
        with torch.no_grad():
            for xs in torch.chunk(X2, n_shards, dim=0):
                outs.append(mlp_fn(xs.unsqueeze(0)).squeeze(0))
        return torch.cat(outs, 0).view(*shape[:-1], -1)

    @staticmethod
    def backward(ctx, dY):
        (X2,) = ctx.saved_tensors
        dY2 = dY.reshape(-1, dY.shape[-1])
        dX = torch.empty_like(X2)
        r0 = 0
        for xs, gs in zip(torch.chunk(X2, ctx.n_shards, dim=0), torch.chunk(dY2, ctx.n_shards, dim=0)):
            x = xs.detach().clone().requires_grad_(True)
            with torch.enable_grad():
                y = ctx.mlp_fn(x.unsqueeze(0)).squeeze(0)
            y.backward(gs)                      # LoRA parameter gradients accumulate across shards
            dX[r0:r0 + x.shape[0]] = x.grad
            r0 += x.shape[0]
        return None, None, dX.view(ctx.shape), None


def tiled_mlp_forward(mlp_fn, n_shards):
    """Wrap a bound `mlp.forward` (e.g. apply_lora_mlp_swiglu bound to the module)."""
    def forward(X):
        if n_shards <= 1 or not torch.is_grad_enabled():
            return mlp_fn(X)
        return TiledMLP.apply(mlp_fn, _anchor(X.device), X, n_shards)
    return forward
