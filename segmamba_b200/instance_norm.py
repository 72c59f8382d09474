"""Fused InstanceNorm3d (+ second operand) (+ ReLU / LeakyReLU) on channels-last activations (smb_instnorm_fwd / _bwd).

Replaces the nn.InstanceNorm3d + activation + residual-add chains of the reference's GSC and UnetResBlock
(model_segmamba/segmamba.py:111-130,147,171; monai/networks/blocks/dynunet_block.py:98-111).  ``fused_instance_norm`` keeps
the semantics of ``act(instance_norm(x, eps=1e-5) [+ instance_norm(x2) | + x2])`` with affine=False and batch statistics
(track_running_stats=False), which is what the reference modules compute in both train and eval mode.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_ACT = {None: 0, "none": 0, "relu": 1, "leaky_relu": 2}


def _as_cl(t):
    """(B, C, D, H, W) -> channels_last_3d storage (a no-op when it already is)."""
    return t.contiguous(memory_format=torch.channels_last_3d)


def _dims(x):
    if x.dim() != 5:
        raise RuntimeError("fused_instance_norm expects a 5-D (B, C, D, H, W) tensor")
    B, C = x.shape[:2]
    return B, C, x.shape[2] * x.shape[3] * x.shape[4]


def _fwd(x, x2, act, slope, mode2, eps):
    _lib.require_cuda(x, x2)
    B, C, S = _dims(x)
    dev = x.device
    with torch.cuda.device(dev):
        y = torch.empty_like(x, memory_format=torch.channels_last_3d)
        stats = torch.empty(B, C, 2, dtype=torch.float32, device=dev)
        stats2 = torch.empty(B, C, 2, dtype=torch.float32, device=dev) if mode2 == 2 else None
        l = _lib.lib()
        dt = _lib.dtype_code(x.dtype)
        wsb = l.smb_instnorm_workspace_bytes(B, C, S, dt)
        if wsb == 0:
            raise RuntimeError(f"fused_instance_norm: channels={C} must be a multiple of {16 // x.element_size()} for {x.dtype}")
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        a = _lib.InstNormArgs()
        a.batch, a.channels, a.dtype, a.act, a.mode2 = B, C, dt, act, mode2
        a.slope, a.eps, a.spatial = float(slope), float(eps), S
        a.x, a.x2, a.y = _lib.ptr(x), _lib.ptr(x2), _lib.ptr(y)
        a.stats, a.stats2 = _lib.ptr(stats), _lib.ptr(stats2)
        a.workspace, a.workspace_bytes = ws.data_ptr(), wsb
        sp = _lib.stream_ptr(dev)
        _lib.call("instnorm_fwd", (B, C, S, x.element_size(), mode2), lambda: l.smb_instnorm_fwd(ctypes.byref(a), sp), dev)
    return y, stats, stats2


def _bwd(x, x2, dy, stats, stats2, act, slope, mode2, eps, need_dx2):
    B, C, S = _dims(x)
    dev = x.device
    with torch.cuda.device(dev):
        dx = torch.empty_like(x, memory_format=torch.channels_last_3d)
        dx2 = torch.empty_like(x, memory_format=torch.channels_last_3d) if (mode2 and need_dx2) else None
        l = _lib.lib()
        dt = _lib.dtype_code(x.dtype)
        wsb = l.smb_instnorm_workspace_bytes(B, C, S, dt)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        a = _lib.InstNormBwdArgs()
        a.batch, a.channels, a.dtype, a.act, a.mode2 = B, C, dt, act, mode2
        a.slope, a.eps, a.spatial = float(slope), float(eps), S
        a.x, a.x2, a.dy = _lib.ptr(x), _lib.ptr(x2), _lib.ptr(dy)
        a.stats, a.stats2 = _lib.ptr(stats), _lib.ptr(stats2)
        a.dx, a.dx2 = _lib.ptr(dx), _lib.ptr(dx2)
        a.workspace, a.workspace_bytes = ws.data_ptr(), wsb
        sp = _lib.stream_ptr(dev)
        _lib.call("instnorm_bwd", (B, C, S, x.element_size(), mode2), lambda: l.smb_instnorm_bwd(ctypes.byref(a), sp), dev)
    return dx, dx2


class _FusedInstanceNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, x2, act, slope, mode2, eps):
        x = _as_cl(x)
        if x2 is not None:
            x2 = _as_cl(x2.to(x.dtype))
        y, stats, stats2 = _fwd(x, x2, act, slope, mode2, eps)
        ctx.save_for_backward(x, x2, stats, stats2)
        ctx.cfg = (act, slope, mode2, eps)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, x2, stats, stats2 = ctx.saved_tensors
        act, slope, mode2, eps = ctx.cfg
        dy = _as_cl(dy.to(x.dtype))
        need2 = mode2 != 0 and ctx.needs_input_grad[1]
        dx, dx2 = _bwd(x, x2, dy, stats, stats2, act, slope, mode2, eps, need2)
        return dx, dx2, None, None, None, None


def fused_instance_norm(x, act=None, negative_slope=0.01, add=None, add_norm=False, eps=1e-5):
    """act(IN(x)), act(IN(x) + add) or act(IN(x) + IN(add)); x, add: (B, C, D, H, W)."""
    if act not in _ACT:
        raise ValueError(f"unknown activation {act!r}")
    mode2 = 0 if add is None else (2 if add_norm else 1)
    return _FusedInstanceNorm.apply(x, add, _ACT[act], negative_slope, mode2, eps)
