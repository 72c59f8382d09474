"""Drop-in for the reference's pybind module ``causal_conv1d_cuda`` (causal-conv1d/csrc/causal_conv1d.cpp:329-333).

``causal_conv1d_fwd`` / ``causal_conv1d_bwd`` keep the reference signatures and allocation rules and forward to
``smb_conv1d_fwd`` / ``smb_conv1d_bwd``.  The channel-last layout and the decode-time ``causal_conv1d_update``
are out of scope (never reached by SegMamba) and raise RuntimeError.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _validate(x, weight, bias_):
    _lib.require_cuda(x, weight, bias_)
    _check(x.dtype in (torch.float32, torch.float16, torch.bfloat16), "causal_conv1d: x must be float32/float16/bfloat16")
    _check(weight.dtype in (torch.float32, torch.float16, torch.bfloat16), "causal_conv1d: bad weight dtype")
    _check(x.dim() == 3, "causal_conv1d: x must be (batch, dim, seqlen)")
    batch, dim, L = x.shape
    _check(weight.dim() == 2 and weight.shape[0] == dim, "causal_conv1d: weight must be (dim, width)")
    width = weight.shape[1]
    _check(2 <= width <= 4, "causal_conv1d only supports width between 2 and 4")            # causal_conv1d.cpp:158
    _check(x.stride(2) == 1, "causal_conv1d: segmamba_b200 supports the (batch, dim, seqlen) layout with stride(2) == 1 only "
                             "(channel-last is out of scope)")
    if bias_ is not None:
        _check(bias_.dtype == weight.dtype and tuple(bias_.shape) == (dim,) and bias_.stride(-1) == 1,
               "causal_conv1d: bias must be contiguous (dim,) with the weight dtype")         # causal_conv1d.cpp:160-166
    return batch, dim, L, width


def causal_conv1d_fwd_ex(x, weight, bias_, silu_activation, *, direction=0, out=None):
    batch, dim, L, width = _validate(x, weight, bias_)
    dev = x.device
    w32 = weight if weight.dtype == torch.float32 else weight.float()
    b32 = None if bias_ is None else (bias_ if bias_.dtype == torch.float32 else bias_.float())
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty_like(x)                                                        # causal_conv1d.cpp:168
        a = _lib.Conv1dArgs()
        a.batch, a.dim, a.seqlen, a.width = batch, dim, L, width
        a.dtype = _lib.dtype_code(x.dtype)
        a.silu = int(bool(silu_activation))
        a.direction = int(direction)
        a.x, a.weight, a.bias, a.out = _lib.ptr(x), _lib.ptr(w32), _lib.ptr(b32), _lib.ptr(out)
        a.x_bs, a.x_ds = x.stride(0), x.stride(1)
        a.out_bs, a.out_ds = out.stride(0), out.stride(1)
        a.w_ds, a.w_ws = w32.stride(0), w32.stride(1)
        sp = _lib.stream_ptr(dev)
        _lib.call("conv1d_fwd", (batch, dim, L, x.element_size()), lambda: _lib.lib().smb_conv1d_fwd(ctypes.byref(a), sp), dev)
    return out


def causal_conv1d_fwd(x, weight, bias_, silu_activation):
    """causal_conv1d_cuda.causal_conv1d_fwd(x, weight, bias_, silu) -> out   (causal_conv1d.cpp:130-189)."""
    return causal_conv1d_fwd_ex(x, weight, bias_, silu_activation)


def causal_conv1d_bwd_ex(x, weight, bias_, dout, dx_, silu_activation, *, direction=0):
    batch, dim, L, width = _validate(x, weight, bias_)
    _lib.require_cuda(dout, dx_)
    _check(dout.dtype == x.dtype and tuple(dout.shape) == (batch, dim, L) and dout.stride(2) == 1,
           "causal_conv1d_bwd: dout must match x with stride(2) == 1")
    dev = x.device
    w32 = weight if weight.dtype == torch.float32 else weight.float()
    b32 = None if bias_ is None else (bias_ if bias_.dtype == torch.float32 else bias_.float())
    with torch.cuda.device(dev):
        if dx_ is not None:
            _check(dx_.dtype == x.dtype and tuple(dx_.shape) == (batch, dim, L) and dx_.stride(2) == 1,
                   "causal_conv1d_bwd: dx must match x with stride(2) == 1")                  # causal_conv1d.cpp:226-236
            dx = dx_
        else:
            dx = torch.empty_like(x)
        dweight = torch.zeros(dim, width, dtype=torch.float32, device=dev)                   # causal_conv1d.cpp:247
        dbias = torch.zeros(dim, dtype=torch.float32, device=dev) if bias_ is not None else None
        a = _lib.Conv1dBwdArgs()
        a.batch, a.dim, a.seqlen, a.width = batch, dim, L, width
        a.dtype = _lib.dtype_code(x.dtype)
        a.silu = int(bool(silu_activation))
        a.direction = int(direction)
        a.x, a.dout, a.weight, a.bias = _lib.ptr(x), _lib.ptr(dout), _lib.ptr(w32), _lib.ptr(b32)
        a.dx, a.dweight, a.dbias = _lib.ptr(dx), _lib.ptr(dweight), _lib.ptr(dbias)
        a.x_bs, a.x_ds = x.stride(0), x.stride(1)
        a.dout_bs, a.dout_ds = dout.stride(0), dout.stride(1)
        a.dx_bs, a.dx_ds = dx.stride(0), dx.stride(1)
        a.w_ds, a.w_ws = w32.stride(0), w32.stride(1)
        sp = _lib.stream_ptr(dev)
        _lib.call("conv1d_bwd", (batch, dim, L, x.element_size()), lambda: _lib.lib().smb_conv1d_bwd(ctypes.byref(a), sp), dev)
    return dx, dweight, dbias


def causal_conv1d_bwd(x, weight, bias_, dout, dx_, silu_activation):
    """causal_conv1d_cuda.causal_conv1d_bwd -> [dx, dweight, dbias]   (causal_conv1d.cpp:191-268)."""
    dx, dweight, dbias = causal_conv1d_bwd_ex(x, weight, bias_, dout, dx_, silu_activation)
    dweight = dweight.to(weight.dtype)                                                       # causal_conv1d.cpp:267
    if dbias is not None:
        dbias = dbias.to(bias_.dtype)
    return [dx, dweight, dbias]


def causal_conv1d_update(x, conv_state, weight, bias_, silu_activation):
    raise RuntimeError("causal_conv1d_update (decode-time step) is out of scope of segmamba_b200: SegMamba never decodes")


def seq_permute(src, nslices, inverse=False, out=None, accumulate=False):
    """Inter-slice re-ordering of the last axis (smb_seq_permute).  src: (..., L) with unit last stride and a
    single row stride when flattened; returns a contiguous tensor unless ``out`` is given."""
    _lib.require_cuda(src, out)
    L = src.shape[-1]
    _check(L % nslices == 0, "seq_permute: L must be divisible by nslices")
    if out is None and src.dim() == 3 and not src.is_contiguous() and src.permute(1, 0, 2).is_contiguous():
        # the "HBL" layout of xz (channel-major, mamba_simple.py:204-208): permute rows in place of a copy
        return seq_permute(src.permute(1, 0, 2), nslices, inverse=inverse).permute(1, 0, 2)
    s2 = src.reshape(-1, L) if src.is_contiguous() else src.contiguous().reshape(-1, L)
    if out is None:
        out = torch.empty(src.shape, dtype=src.dtype, device=src.device)
        accumulate = False
    _check(out.is_contiguous() and out.shape == src.shape and out.dtype == src.dtype, "seq_permute: bad out")
    a = _lib.SeqPermuteArgs()
    a.rows, a.seqlen, a.nslices = s2.shape[0], L, int(nslices)
    a.dtype = _lib.dtype_code(src.dtype)
    a.inverse, a.accumulate = int(bool(inverse)), int(bool(accumulate))
    a.src, a.dst = _lib.ptr(s2), _lib.ptr(out)
    a.src_rs, a.dst_rs = s2.stride(0), L
    with torch.cuda.device(src.device):
        sp = _lib.stream_ptr(src.device)
        _lib.call("seq_permute", (int(s2.shape[0]), L, src.element_size()), lambda: _lib.lib().smb_seq_permute(ctypes.byref(a), sp), src.device)
    return out
