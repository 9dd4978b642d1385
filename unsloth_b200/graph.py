"""CUDA-graph capture of the training step (B200-first: streams and graphs, no tracing compiler).

A QLoRA step is ~2,400 short launches (ctypes calls into libunsloth_b200.so, flash/cuDNN attention,
a few torch elementwise ops).  Shapes are static for a fixed (batch, seq), so forward + backward
are captured ONCE into a CUDA graph and replayed; the LoRA-gradient all-reduce and the AdamW launch
stay outside the graph (their scalars change every step).

Capture rules honoured by the kernels: every launch goes to torch's CURRENT stream, nothing
synchronises, all scratch comes from torch's caching allocator (graph-private pool during
capture), TMA descriptors travel by value in the kernel parameters.
"""
from __future__ import annotations

import torch

from . import _lib as L
from .kernels import utils as KU


class GraphedTrainStep:
    def __init__(self, model, bucket, batch_size, seq_len, device, warmup=2):
        self.model, self.bucket = model, bucket
        self.ids = torch.zeros(batch_size, seq_len, dtype=torch.int64, device=device)
        self.labels = torch.zeros(batch_size, seq_len, dtype=torch.int64, device=device)
        self.loss = None
        self.launches_per_replay = 0
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):            # warm-up off the default stream, as capture requires
            for _ in range(warmup):
                self._fwd_bwd()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        KU.bump_param_epoch()                    # the per-step LoRA casts must be IN the graph
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count
        with torch.cuda.graph(self.graph):
            self._fwd_bwd()
        self.launches_per_replay = L.launch_count - n0

    def _fwd_bwd(self):
        self.bucket.begin_step()             # zero_grad + batched LoRA cast refresh
        try:
            out = self.model(input_ids=self.ids, labels=self.labels)
            out.loss.backward()
        finally:
            self.bucket.end_backward()       # batched accumulation of the LoRA gradients
        self.loss = out.loss.detach()

    def step(self, input_ids, labels):
        """input_ids / labels: device or pinned-host int64 [batch, seq].  Returns the loss tensor
        (device, valid after the replay completes)."""
        self.ids.copy_(input_ids, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.graph.replay()
        self.bucket.all_reduce_grads()
        self.bucket.step()
        return self.loss
