"""Whole-step CUDA-graph capture for the SegMamba training step.

One training step of the default model enqueues ~4 500 kernels (cuDNN / cuBLAS / ATen / native); at ~100 ms of GPU time the
host needs ~65 ms just to launch them (bench.py: host_enqueue_ms_per_step), so the step is one kernel-speedup away from
being launch-bound.  ``GraphedTrainStep`` captures forward + loss + backward + gradient clipping + optimizer step once and
replays it with a single ``cudaGraphLaunch`` per step (static input / label buffers, private memory pool).  All native
kernels are capture-safe by construction: they only enqueue work on the current stream and allocate through the PyTorch
caching allocator.

This replaces nothing in the reference (its trainer launches eagerly, light_training/trainer.py:422-483); it is the
"CUDA graphs instead of a tracing compiler" part of the B200-first design.
"""
from __future__ import annotations

from typing import Callable

import torch


class GraphedTrainStep:
    """step = GraphedTrainStep(model, optimizer, loss_fn, x_example, y_example); loss = step(x, y)

    ``model`` may be a DistributedDataParallel wrapper (construct it, run the warm-up and capture on the same side stream,
    as PyTorch's CUDA-graph notes require).  ``loss_fn(logits, y) -> scalar``.  The returned loss is a static device tensor
    that is overwritten by the next replay.
    """

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn: Callable, x_example: torch.Tensor,
                 y_example: torch.Tensor, autocast_dtype: torch.dtype | None = torch.bfloat16, clip_grad_norm: float | None = 12.0,
                 warmup_iters: int = 3, master_weights=None):
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.autocast_dtype, self.clip = autocast_dtype, clip_grad_norm
        self.static_x = x_example.clone()
        self.static_y = y_example.clone()
        self.master_weights = master_weights                  # optional MasterWeights (bf16 parameters, fp32 masters)
        self.params = (master_weights.optimizer_parameters() if master_weights is not None
                       else [p for p in model.parameters() if p.requires_grad])
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup_iters):                # warm-up on the capture stream (cuDNN autotune, lazy inits, allocator)
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self._zero()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._eager_step(zero=False)

    def _zero(self):
        self.optimizer.zero_grad(set_to_none=True)
        if self.master_weights is not None:
            self.master_weights.zero_grad()

    def _eager_step(self, zero: bool = True):
        if zero:
            self._zero()
        with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            logits = self.model(self.static_x)
            loss = self.loss_fn(logits.float(), self.static_y)
        loss.backward()
        if self.master_weights is not None:
            self.master_weights.grads_to_master()
        if self.clip is not None:
            torch.nn.utils.clip_grad_norm_(self.params, self.clip, foreach=True)
        self.optimizer.step()
        if self.master_weights is not None:
            self.master_weights.master_to_model()
        return loss.detach()

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.static_x.copy_(x, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.static_loss
