"""CPU oracle for the SegMamba hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  ``segmamba_b200`` never does.

It wraps ``oracle/segmamba_oracle.c`` (plain-C restatement of the reference scan / conv1d algorithms)
with ctypes and restates, in plain CPU PyTorch, the Python orchestration of the reference:

* ``mamba_inner_no_out_proj``  <- MambaInnerFnNoOutProj.forward,
  mamba/mamba_ssm/ops/selective_scan_interface.py:159-224
* ``mamba_v3_forward``         <- Mamba.forward (bimamba_type="v3"),
  mamba/mamba_ssm/modules/mamba_simple.py:188-264
* ``segmamba_forward``         <- model_segmamba/segmamba.py:49-193,327-343 and the MONAI blocks
  monai/networks/blocks/dynunet_block.py:25-111,247-267, unetr_block.py:22-86,209-259

Parity pinning: checked against golden vectors generated from the reference's own pure-PyTorch
functions by ``oracle/gen_golden.py`` (committed under ``tests/golden/``); see tests/test_oracle_golden.py.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, ctypes.CDLL] = {}
_PRECISION = "f64"     # arithmetic of the autograd wrappers: "f64" = checker, "f32" = the timed CPU baseline


def set_precision(p: str) -> None:
    global _PRECISION
    assert p in ("f64", "f32")
    _PRECISION = p

_fp = ctypes.POINTER(ctypes.c_float)


def build(force: bool = False) -> None:
    """Compile the C oracle (gcc, seconds).  Building the checker is not using it."""
    so = os.path.join(_HERE, "libsegmamba_oracle.so")
    so32 = os.path.join(_HERE, "libsegmamba_oracle_f32.so")
    src = os.path.join(_HERE, "segmamba_oracle.c")
    if (not force and os.path.exists(so) and os.path.exists(so32)
            and os.path.getmtime(so) >= os.path.getmtime(src)
            and os.path.getmtime(so32) >= os.path.getmtime(src)):
        return
    subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)


def _lib(precision: str = "f64") -> ctypes.CDLL:
    name = "libsegmamba_oracle.so" if precision == "f64" else "libsegmamba_oracle_f32.so"
    if name not in _LIBS:
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.orc_num_threads.restype = ctypes.c_int
        lib.orc_real_bytes.restype = ctypes.c_int
        _LIBS[name] = lib
    return _LIBS[name]


def num_threads() -> int:
    return int(_lib().orc_num_threads())


def _f32(t):
    """contiguous fp32 CPU tensor (values of fp16/bf16 inputs are preserved exactly)."""
    if t is None:
        return None
    return t.detach().to(device="cpu", dtype=torch.float32).contiguous()


def _p(t):
    if t is None:
        return ctypes.cast(None, _fp)
    return ctypes.cast(t.data_ptr(), _fp)


# ----------------------------------------------------------------------------------------------
# raw (non-autograd) entry points
# ----------------------------------------------------------------------------------------------
def selective_scan_fwd_raw(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                           chunk=2048, precision="f64"):
    """returns (y, out_z or None, last_state, xchunks), all fp32.  B, C: (batch, [G,] N, L)."""
    u32, d32, A32 = _f32(u), _f32(delta), _f32(A)
    B32, C32 = _f32(B), _f32(C)
    if B32.dim() == 3:
        B32 = B32.unsqueeze(1)
    if C32.dim() == 3:
        C32 = C32.unsqueeze(1)
    batch, dim, L = u32.shape
    N = A32.shape[1]
    G = B32.shape[1]
    D32, z32, b32 = _f32(D), _f32(z), _f32(delta_bias)
    y = torch.empty_like(u32)
    oz = torch.empty_like(u32) if z is not None else None
    last = torch.empty(batch, dim, N)
    nch = (L + chunk - 1) // chunk
    xc = torch.empty(batch, dim, nch, 2 * N)
    _lib(precision).orc_selective_scan_fwd(
        _p(u32), _p(d32), _p(A32), _p(B32), _p(C32), _p(D32), _p(z32), _p(b32),
        ctypes.c_int(int(bool(delta_softplus))),
        ctypes.c_int(batch), ctypes.c_int(dim), ctypes.c_int(L), ctypes.c_int(N), ctypes.c_int(G),
        ctypes.c_int(chunk), _p(y), _p(oz), _p(last), _p(xc))
    return y, oz, last, xc


def selective_scan_bwd_raw(u, delta, A, B, C, D, z, delta_bias, delta_softplus, dout, precision="f64"):
    """returns dict of fp32 grads: du, ddelta, dA, dB, dC (batch,G,N,L), dD, ddelta_bias, dz."""
    u32, d32, A32 = _f32(u), _f32(delta), _f32(A)
    B32, C32 = _f32(B), _f32(C)
    if B32.dim() == 3:
        B32 = B32.unsqueeze(1)
    if C32.dim() == 3:
        C32 = C32.unsqueeze(1)
    batch, dim, L = u32.shape
    N = A32.shape[1]
    G = B32.shape[1]
    D32, z32, b32, g32 = _f32(D), _f32(z), _f32(delta_bias), _f32(dout)
    out = dict(du=torch.empty_like(u32), ddelta=torch.empty_like(u32), dA=torch.empty_like(A32),
               dB=torch.empty_like(B32), dC=torch.empty_like(C32), dD=torch.empty(dim),
               ddelta_bias=torch.empty(dim), dz=torch.empty_like(u32) if z is not None else None)
    _lib(precision).orc_selective_scan_bwd(
        _p(u32), _p(d32), _p(A32), _p(B32), _p(C32), _p(D32), _p(z32), _p(b32),
        ctypes.c_int(int(bool(delta_softplus))), _p(g32),
        ctypes.c_int(batch), ctypes.c_int(dim), ctypes.c_int(L), ctypes.c_int(N), ctypes.c_int(G),
        _p(out["du"]), _p(out["ddelta"]), _p(out["dA"]), _p(out["dB"]), _p(out["dC"]), _p(out["dD"]),
        _p(out["ddelta_bias"]), _p(out["dz"]))
    return out


def causal_conv1d_fwd_raw(x, weight, bias=None, silu=False, precision="f64"):
    x32, w32, b32 = _f32(x), _f32(weight), _f32(bias)
    batch, dim, L = x32.shape
    out = torch.empty_like(x32)
    _lib(precision).orc_causal_conv1d_fwd(_p(x32), _p(w32), _p(b32), ctypes.c_int(int(bool(silu))),
                                          ctypes.c_int(batch), ctypes.c_int(dim), ctypes.c_int(L),
                                          ctypes.c_int(w32.shape[1]), _p(out))
    return out


def causal_conv1d_bwd_raw(x, weight, bias, dout, silu=False, precision="f64"):
    x32, w32, b32, g32 = _f32(x), _f32(weight), _f32(bias), _f32(dout)
    batch, dim, L = x32.shape
    dx, dw, db = torch.empty_like(x32), torch.empty_like(w32), torch.empty(dim)
    _lib(precision).orc_causal_conv1d_bwd(_p(x32), _p(w32), _p(b32), _p(g32), ctypes.c_int(int(bool(silu))),
                                          ctypes.c_int(batch), ctypes.c_int(dim), ctypes.c_int(L),
                                          ctypes.c_int(w32.shape[1]), _p(dx), _p(dw), _p(db))
    return dx, dw, db


# ----------------------------------------------------------------------------------------------
# autograd wrappers (CPU), mirroring selective_scan_fn / causal_conv1d_fn
# ----------------------------------------------------------------------------------------------
class _SelectiveScanOracle(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, delta, A, B, C, D, z, delta_bias, delta_softplus):
        y, oz, _, _ = selective_scan_fwd_raw(u, delta, A, B, C, D, z, delta_bias, delta_softplus, precision=_PRECISION)
        ctx.save_for_backward(u, delta, A, B, C, D, z, delta_bias)
        ctx.delta_softplus = delta_softplus
        return (oz if z is not None else y).to(u.dtype)

    @staticmethod
    def backward(ctx, dout):
        u, delta, A, B, C, D, z, delta_bias = ctx.saved_tensors
        g = selective_scan_bwd_raw(u, delta, A, B, C, D, z, delta_bias, ctx.delta_softplus, dout, precision=_PRECISION)
        dB = g["dB"] if B.dim() == 4 else g["dB"].squeeze(1)
        dC = g["dC"] if C.dim() == 4 else g["dC"].squeeze(1)
        return (g["du"].to(u.dtype), g["ddelta"].to(delta.dtype), g["dA"], dB.to(B.dtype), dC.to(C.dtype),
                g["dD"] if D is not None else None,
                g["dz"].to(z.dtype) if z is not None else None,
                g["ddelta_bias"] if delta_bias is not None else None, None)


def selective_scan(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False):
    """Same contract as selective_scan_fn (ssi.py:77-83) without return_last_state."""
    return _SelectiveScanOracle.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus)


class _CausalConv1dOracle(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, silu):
        ctx.save_for_backward(x, weight, bias)
        ctx.silu = silu
        return causal_conv1d_fwd_raw(x, weight, bias, silu, precision=_PRECISION).to(x.dtype)

    @staticmethod
    def backward(ctx, dout):
        x, weight, bias = ctx.saved_tensors
        dx, dw, db = causal_conv1d_bwd_raw(x, weight, bias, dout, ctx.silu, precision=_PRECISION)
        return dx.to(x.dtype), dw.to(weight.dtype), (db.to(bias.dtype) if bias is not None else None), None


def causal_conv1d(x, weight, bias=None, activation=None):
    """Same contract as causal_conv1d_fn (causal_conv1d_interface.py:37-46)."""
    return _CausalConv1dOracle.apply(x, weight, bias, activation in ("silu", "swish"))


# ----------------------------------------------------------------------------------------------
# Python orchestration restated on CPU
# ----------------------------------------------------------------------------------------------
def mamba_inner_no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                            A, D, delta_bias, delta_softplus=True):
    """MambaInnerFnNoOutProj.forward, ssi.py:159-224 (variable B, C; real A).  xz: (B, 2*d_inner, L)."""
    L = xz.shape[-1]
    delta_rank = delta_proj_weight.shape[1]
    d_state = A.shape[-1]
    w = conv1d_weight.reshape(conv1d_weight.shape[0], conv1d_weight.shape[-1])      # "d 1 w -> d w" :174
    x, z = xz.chunk(2, dim=1)                                                        # :175
    conv1d_out = causal_conv1d(x, w, conv1d_bias, "silu")                            # :177
    bsz, d_inner, _ = conv1d_out.shape
    x_dbl = F.linear(conv1d_out.transpose(1, 2).reshape(bsz * L, d_inner), x_proj_weight)   # :181
    delta = (delta_proj_weight @ x_dbl[:, :delta_rank].t()).reshape(d_inner, bsz, L).transpose(0, 1)  # :182
    Bm = x_dbl[:, delta_rank:delta_rank + d_state].reshape(bsz, L, d_state).transpose(1, 2).unsqueeze(1)  # :187-195
    Cm = x_dbl[:, -d_state:].reshape(bsz, L, d_state).transpose(1, 2).unsqueeze(1)                        # :199-207
    return selective_scan(conv1d_out, delta.contiguous(), A, Bm.contiguous(), Cm.contiguous(), D, z.contiguous(),
                          delta_bias, delta_softplus)                                                      # :213-215


def mamba_v3_forward(p: dict, hidden_states, nslices: int, prefix: str = ""):
    """Mamba.forward with bimamba_type="v3", mamba_simple.py:188-264.  ``p`` maps the reference's
    parameter names (A_log, D, in_proj.weight, conv1d.weight, ... ) to tensors."""
    g = lambda k: p[prefix + k]
    batch, seqlen, _ = hidden_states.shape
    d_inner = g("A_log").shape[0]
    xz = (g("in_proj.weight") @ hidden_states.reshape(batch * seqlen, -1).t())            # :204-208
    xz = xz.reshape(2 * d_inner, batch, seqlen).transpose(0, 1)

    def inner(xz_dir, sfx):
        A = -torch.exp(g(f"A{sfx}_log").float())                                          # :212,216,243
        return mamba_inner_no_out_proj(
            xz_dir, g(f"conv1d{sfx}.weight"), g(f"conv1d{sfx}.bias"), g(f"x_proj{sfx}.weight"),
            g(f"dt_proj{sfx}.weight"), A, g(f"D{sfx}").float(), g(f"dt_proj{sfx}.bias").float(), True)

    out = inner(xz, "")                                                                   # :217-229
    out_b = inner(xz.flip([-1]), "_b")                                                    # :230-242
    xz_s = torch.stack(xz.chunk(nslices, dim=-1), dim=-1).flatten(-2)                     # :245-247
    out_s = inner(xz_s, "_s")                                                             # :248-260
    out_s = out_s.reshape(batch, d_inner, seqlen // nslices, nslices).permute(0, 1, 3, 2).flatten(-2)  # :261
    y = out + out_b.flip([-1]) + out_s
    return F.linear(y.transpose(1, 2), g("out_proj.weight"))                              # :264


def _inorm(x):
    return F.instance_norm(x, eps=1e-5)


def segmamba_forward(sd: dict, x_in, depths=(2, 2, 2, 2), nslices=(64, 32, 16, 8)):
    """SegMamba.forward restated functionally from a reference state_dict (model_segmamba/segmamba.py:327-343)."""
    def conv(x, key, stride=1, padding=0):
        return F.conv3d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride=stride, padding=padding)

    def gsc(x, pre):                                   # segmamba.py:111-132
        r = x
        x1 = F.relu(_inorm(conv(x, pre + "proj", padding=1)))
        x1 = F.relu(_inorm(conv(x1, pre + "proj2", padding=1)))
        x2 = F.relu(_inorm(conv(x, pre + "proj3")))
        x = F.relu(_inorm(conv(x1 + x2, pre + "proj4")))
        return x + r

    def mamba_layer(x, pre, ns):                       # segmamba.py:63-76
        Bsz, Cc = x.shape[:2]
        dims = x.shape[2:]
        xf = x.reshape(Bsz, Cc, -1).transpose(-1, -2)
        xn = F.layer_norm(xf, (Cc,), sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-5)
        xm = mamba_v3_forward(sd, xn, ns, prefix=pre + "mamba.")
        return xm.transpose(-1, -2).reshape(Bsz, Cc, *dims) + x

    def res_block(x, pre):                             # dynunet_block.py:98-111 (UnetResBlock)
        residual = x
        out = F.leaky_relu(_inorm(conv(x, pre + "conv1.conv", padding=1)), 0.01)
        out = _inorm(conv(out, pre + "conv2.conv", padding=1))
        if (pre + "conv3.conv.weight") in sd:
            residual = _inorm(conv(residual, pre + "conv3.conv"))
        return F.leaky_relu(out + residual, 0.01)

    def up_block(inp, skip, pre):                      # unetr_block.py:81-86 (UnetrUpBlock)
        out = F.conv_transpose3d(inp, sd[pre + "transp_conv.conv.weight"], stride=2)
        return res_block(torch.cat((out, skip), dim=1), pre + "conv_block.")

    outs = []
    x = x_in
    for i in range(4):                                 # segmamba.py:176-189
        if i == 0:
            x = conv(x, "vit.downsample_layers.0.0", stride=2, padding=3)
        else:
            x = conv(_inorm(x), f"vit.downsample_layers.{i}.1", stride=2)
        x = gsc(x, f"vit.gscs.{i}.")
        for j in range(depths[i]):
            x = mamba_layer(x, f"vit.stages.{i}.{j}.", nslices[i])
        xo = _inorm(x)
        xo = conv(F.gelu(conv(xo, f"vit.mlps.{i}.fc1")), f"vit.mlps.{i}.fc2")
        outs.append(xo)
    enc1 = res_block(x_in, "encoder1.layer.")
    enc2 = res_block(outs[0], "encoder2.layer.")
    enc3 = res_block(outs[1], "encoder3.layer.")
    enc4 = res_block(outs[2], "encoder4.layer.")
    enc_hidden = res_block(outs[3], "encoder5.layer.")
    dec3 = up_block(enc_hidden, enc4, "decoder5.")
    dec2 = up_block(dec3, enc3, "decoder4.")
    dec1 = up_block(dec2, enc2, "decoder3.")
    dec0 = up_block(dec1, enc1, "decoder2.")
    out = res_block(dec0, "decoder1.layer.")
    return conv(out, "out.conv.conv")


# ----------------------------------------------------------------------------------------------
# sliding-window inference restated (monai/inferers/utils.py:138-321, monai/data/utils.py:171-211,1088-1137)
# ----------------------------------------------------------------------------------------------
def sliding_window_starts(image_size, roi_size, overlap=0.5):
    """dense_patch_slices start indices (monai/data/utils.py:171-211) with scan_interval from
    _get_scan_interval (monai/inferers/utils.py:363-384)."""
    starts = []
    for dim_sz, roi in zip(image_size, roi_size):
        interval = roi if roi == dim_sz else max(int(roi * (1 - overlap)), 1)
        num = 1 if dim_sz <= roi else int(math.ceil(float(dim_sz - roi) / interval)) + 1
        starts.append([min(i * interval, dim_sz - roi) for i in range(num)])
    out = []
    for a in starts[0]:
        for b in starts[1]:
            for c in starts[2]:
                out.append((a, b, c))
    return out


def gaussian_importance_map(roi_size, sigma_scale=0.125):
    """compute_importance_map(mode="gaussian") (monai/data/utils.py:1088-1137)."""
    m = None
    for i, roi in enumerate(roi_size):
        sigma = sigma_scale * roi
        center = (roi - 1) / 2.0
        x = torch.arange(roi, dtype=torch.float32) - center
        gi = torch.exp(x ** 2 / (-2 * sigma ** 2))
        shape = [1, 1, 1]
        shape[i] = roi
        gi = gi.reshape(shape)
        m = gi if m is None else m * gi
    min_non_zero = max(float(m.min()), 1e-3)
    return torch.clamp(m, min=min_non_zero)
