"""The native CUDA kernels, run WITHOUT a GPU: csrc/*.cu compiled for the host by tools/simt_emu (one fiber per CUDA thread,
SIMT barrier / shuffle semantics, shared memory poisoned with NaN on block entry) and driven through the unmodified Python
shims and the C ABI.  Parity against the oracle and the reference's golden vectors at small sizes.

This is a functional check of what the kernel code computes (indexing, tails, strides, barriers, reductions) that the CPU-only
suite can run; it is not a product path (tests/emu.py) and says nothing about performance.  The `-m gpu` suite runs the
same comparisons on the real device at full sizes.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import emu
import golden_inputs as gi
import test_gpu_scan as tg
from util import GRAD_TOL, TOL, assert_close, rand_scan_inputs


@pytest.fixture(autouse=True)
def _emulated():
    with emu.emulated():
        emu.emu_lib().smb_emu_set_reverse(0)
        yield


def _oracle():
    from oracle import oracle as orc
    return orc


# ---------------------------------------------------------------------------------------------------------- selective scan
@pytest.mark.parametrize("case", gi.SCAN_CASES, ids=lambda c: c[0])
def test_emu_scan_vs_golden_and_oracle(case):
    name, seed, batch, dim, L, N, G, tl, has_D, has_z, has_b, sp = case
    d = gi.scan_inputs(seed, batch, dim, L, N, G, tl)
    res = tg._run_fwd_bwd(d, has_D, has_z, has_b, sp)
    ref = tg._oracle_fwd_bwd(d, has_D, has_z, has_b, sp)
    tg._compare(res, ref, torch.float32, has_z)
    gold = gi.load("scan_" + name)
    assert_close(res[2] if has_z else res[0], gold["out"], 1e-3, "out vs reference golden")
    assert_close(res[1][:, :, -1, 1::2], gold["last_state"], 1e-3, "last_state vs reference golden")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16], ids=["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 40, 700, 16, 1), (1, 33, 31, 16, 1), (2, 48, 600, 8, 2), (1, 64, 2300, 16, 1)],
                         ids=lambda s: "b%d_d%d_L%d_n%d_g%d" % s)
@pytest.mark.parametrize("direction", [0, 1], ids=["fwd", "rev"])
def test_emu_scan_random_vs_oracle(dtype, shape, direction):
    batch, dim, L, N, G = shape
    d = rand_scan_inputs(100 + L, batch, dim, L, N, G, dtype, device="cpu")
    res = tg._run_fwd_bwd(d, direction=direction, use_hstates=(L % 2 == 0))
    ref = tg._oracle_fwd_bwd(d, flip=bool(direction))
    tg._compare(res, ref, dtype, True)


@pytest.mark.parametrize("order", [1, 2, 7], ids=["descending", "random2", "random7"])
def test_emu_scan_other_thread_orders(order):
    """same kernels with the threads of every block resumed in descending or pseudo-random order: a cross-lane shared-memory
    dependency that lacks a barrier produces a different (wrong) result under some order."""
    emu.emu_lib().smb_emu_set_reverse(order)
    d = rand_scan_inputs(9, 2, 40, 900, 16, 1, torch.float32, device="cpu")
    for direction in (0, 1):
        res = tg._run_fwd_bwd(d, direction=direction, use_hstates=True)
        ref = tg._oracle_fwd_bwd(d, flip=bool(direction))
        tg._compare(res, ref, torch.float32, True)


def test_emu_scan_strided_hbl_layout():
    batch, dim, L, N = 2, 64, 500, 16
    d = rand_scan_inputs(7, batch, dim, L, N, device="cpu")
    hbl = lambda t: t.permute(1, 0, 2).contiguous().permute(1, 0, 2)
    d2 = dict(d)
    for k in ("u", "delta", "z", "dout"):
        d2[k] = hbl(d[k])
        assert d2[k].stride() == (L, batch * L, 1)
    tg._compare(tg._run_fwd_bwd(d2), tg._oracle_fwd_bwd(d), torch.float32, True)


@pytest.mark.parametrize("order", [1, 5], ids=["descending", "random5"])
def test_emu_norms_and_conv_other_thread_orders(order):
    from segmamba_b200 import causal_conv1d_cuda as cc
    from segmamba_b200.instance_norm import fused_instance_norm
    from segmamba_b200.layer_norm import fused_layer_norm
    emu.emu_lib().smb_emu_set_reverse(order)
    orc = _oracle()
    d = gi.conv_inputs(77, 2, 40, 1300, 4)
    out = cc.causal_conv1d_fwd(d["x"], d["weight"], d["bias"], True)
    dx, dw, db = cc.causal_conv1d_bwd(d["x"], d["weight"], d["bias"], d["dout"], None, True)
    assert_close(out, orc.causal_conv1d_fwd_raw(d["x"], d["weight"], d["bias"], True), 1e-5, "conv out")
    odx, odw, odb = orc.causal_conv1d_bwd_raw(d["x"], d["weight"], d["bias"], d["dout"], True)
    assert_close(dx, odx, 1e-4, "conv dx"); assert_close(dw, odw, 1e-4, "conv dw"); assert_close(db, odb, 1e-4, "conv db")
    x = (torch.randn(2, 48, 8, 8, 8) + 0.5).contiguous(memory_format=torch.channels_last_3d).requires_grad_()
    a = torch.randn(2, 48, 8, 8, 8).requires_grad_()
    y = fused_instance_norm(x, "leaky_relu", 0.01, add=a, add_norm=True)
    g = torch.autograd.grad(y, [x, a], torch.ones_like(y))
    xr, ar = x.detach().clone().requires_grad_(), a.detach().clone().requires_grad_()
    yr = F.leaky_relu(F.instance_norm(xr) + F.instance_norm(ar), 0.01)
    gr = torch.autograd.grad(yr, [xr, ar], torch.ones_like(yr))
    assert_close(y, yr, 1e-4, "instnorm y"); assert_close(g[0], gr[0], 1e-3, "instnorm dx"); assert_close(g[1], gr[1], 1e-3, "instnorm dx2")
    t = torch.randn(700, 96).requires_grad_()
    w, b = torch.rand(96).requires_grad_(), torch.randn(96).requires_grad_()
    yl = fused_layer_norm(t, w, b)
    gl = torch.autograd.grad(yl, [t, w, b], torch.ones_like(yl) * 0.3)
    tr, wr, br = (v.detach().clone().requires_grad_() for v in (t, w, b))
    ylr = F.layer_norm(tr, (96,), wr, br)
    glr = torch.autograd.grad(ylr, [tr, wr, br], torch.ones_like(ylr) * 0.3)
    assert_close(yl, ylr, 1e-5, "layernorm y")
    for u, v, n in zip(gl, glr, ("dx", "dw", "db")):
        assert_close(u, v, 2e-4, "layernorm " + n)


# ------------------------------------------------------------------------------------------------- conv1d and seq permute
@pytest.mark.parametrize("case", gi.CONV_CASES, ids=lambda c: c[0])
def test_emu_conv_vs_golden(case):
    from segmamba_b200 import causal_conv1d_cuda as cc
    name, seed, batch, dim, L, width, has_b, silu = case
    d = gi.conv_inputs(seed, batch, dim, L, width)
    b = d["bias"] if has_b else None
    out = cc.causal_conv1d_fwd(d["x"], d["weight"], b, silu)
    dx, dw, db = cc.causal_conv1d_bwd(d["x"], d["weight"], b, d["dout"], None, silu)
    gold = gi.load("conv_" + name)
    assert_close(out, gold["out"], 1e-3, "out vs golden")
    assert_close(dx, gold["dx"], 1e-3, "dx vs golden")
    assert_close(dw, gold["dweight"], 1e-3, "dweight vs golden")
    if has_b:
        assert_close(db, gold["dbias"], 1e-3, "dbias vs golden")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 40, 1300, 4), (1, 33, 1031, 3), (2, 8, 7, 2)], ids=lambda s: "b%d_d%d_L%d_w%d" % s)
@pytest.mark.parametrize("direction", [0, 1], ids=["fwd", "rev"])
def test_emu_conv_random_vs_oracle(dtype, shape, direction):
    from segmamba_b200 import causal_conv1d_cuda as cc
    orc = _oracle()
    batch, dim, L, width = shape
    d = gi.conv_inputs(200 + L, batch, dim, L, width)
    x, dout, w, b = d["x"].to(dtype), d["dout"].to(dtype), d["weight"], d["bias"]
    xz = torch.empty(2 * dim, batch, L, dtype=dtype).permute(1, 0, 2)      # channel-major views, dx in place (ssi.py:244-245,281)
    xz[:, :dim] = x
    xv = xz[:, :dim]
    dxz = torch.zeros_like(xz)
    out = cc.causal_conv1d_fwd_ex(xv, w, b, True, direction=direction)
    dx, dw, db = cc.causal_conv1d_bwd_ex(xv, w, b, dout, dxz[:, :dim], True, direction=direction)
    f = (lambda t: t.flip(-1)) if direction else (lambda t: t)
    o = f(orc.causal_conv1d_fwd_raw(f(x.float()), w, b, True))
    odx, odw, odb = orc.causal_conv1d_bwd_raw(f(x.float()), w, b, f(dout.float()), True)
    assert_close(out, o, TOL[dtype], "out")
    assert_close(dxz[:, :dim], f(odx), GRAD_TOL[dtype], "dx (in place)")
    assert float(dxz[:, dim:].abs().max()) == 0.0
    assert_close(dw, odw, GRAD_TOL[dtype], "dweight")
    assert_close(db, odb, GRAD_TOL[dtype], "dbias")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_emu_seq_permute(dtype):
    from segmamba_b200 import causal_conv1d_cuda as cc
    for (b, d, L, ns) in [(2, 12, 1024, 32), (1, 10, 512, 8), (2, 6, 96, 16)]:
        x = torch.randn(b, d, L).to(dtype)
        ref = torch.stack(x.chunk(ns, dim=-1), dim=-1).flatten(-2)
        got = cc.seq_permute(x, ns)
        assert torch.equal(got, ref)
        assert torch.equal(cc.seq_permute(got, ns, inverse=True), x)
        hbl = x.permute(1, 0, 2).contiguous().permute(1, 0, 2)
        assert torch.equal(cc.seq_permute(hbl, ns), ref)


# ----------------------------------------------------------------------------------------------------- fused instance norm
def _in_ref(x, add, add_norm, act, slope):
    v = F.instance_norm(x.float(), eps=1e-5)
    if add is not None:
        v = v + (F.instance_norm(add.float(), eps=1e-5) if add_norm else add.float())
    if act == "relu":
        v = F.relu(v)
    elif act == "leaky_relu":
        v = F.leaky_relu(v, slope)
    return v


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 48, 8, 8, 8), (1, 96, 9, 7, 5), (2, 768, 2, 2, 2), (1, 32, 20, 13, 11)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("mode,act", [("plain", None), ("plain", "relu"), ("add", "leaky_relu"), ("addnorm", "leaky_relu")])
def test_emu_fused_instance_norm(dtype, shape, mode, act):
    from segmamba_b200.instance_norm import fused_instance_norm
    torch.manual_seed(sum(shape))
    x = (torch.randn(shape) * 2.0 + 0.7).to(dtype).contiguous(memory_format=torch.channels_last_3d).requires_grad_()
    add = (torch.randn(shape) * 0.5 - 0.3).to(dtype).requires_grad_() if mode != "plain" else None
    dy = torch.randn(shape).to(dtype)
    y = fused_instance_norm(x, act, 0.01, add=add, add_norm=(mode == "addnorm"))
    assert y.is_contiguous(memory_format=torch.channels_last_3d) and y.dtype == dtype
    gx = torch.autograd.grad(y, [x] + ([add] if add is not None else []), dy)
    xr = x.detach().clone().requires_grad_()
    ar = add.detach().clone().requires_grad_() if add is not None else None
    yr = _in_ref(xr, ar, mode == "addnorm", act, 0.01)
    gr = torch.autograd.grad(yr, [xr] + ([ar] if ar is not None else []), dy.float())
    assert_close(y, yr, 1e-4 if dtype == torch.float32 else 1e-2, "y")
    for a, b, n in zip(gx, gr, ("dx", "dadd")):
        assert_close(a, b, 2e-4 if dtype == torch.float32 else 3e-2, n)


def test_emu_fused_instance_norm_large_mean():
    from segmamba_b200.instance_norm import fused_instance_norm
    x = (torch.randn(1, 16, 16, 16, 16) * 0.01 + 50.0).contiguous(memory_format=torch.channels_last_3d)
    assert_close(fused_instance_norm(x), F.instance_norm(x.double(), eps=1e-5).float(), 2e-3, "y (mean 50, std 0.01)")


# --------------------------------------------------------------------------------------- fused inner op, mixer, full module
@pytest.mark.parametrize("case", gi.INNER_CASES, ids=lambda c: c[0])
def test_emu_inner_vs_reference_golden(case):
    from segmamba_b200.selective_scan_interface import mamba_inner_fn_no_out_proj
    name, seed, batch, d_model, L = case
    d = gi.inner_inputs(seed, batch, d_model, L)
    gold = gi.load("inner_" + name)
    keys = ["xz", "conv1d_weight", "conv1d_bias", "x_proj_weight", "delta_proj_weight", "A", "D", "delta_bias"]
    lv = {k: d[k].clone().requires_grad_() for k in keys}
    out = mamba_inner_fn_no_out_proj(lv["xz"], lv["conv1d_weight"], lv["conv1d_bias"], lv["x_proj_weight"],
                                     lv["delta_proj_weight"], lv["A"], None, None, lv["D"], delta_bias=lv["delta_bias"],
                                     delta_softplus=True)
    assert_close(out, gold["out"], 1e-3, "out")
    grads = torch.autograd.grad(out, [lv[k] for k in keys], d["dout"])
    for k, g in zip(keys, grads):
        assert_close(g, gold["d" + k], 2e-3, "d" + k)


@pytest.mark.parametrize("case", gi.MAMBA_CASES, ids=lambda c: c[0])
def test_emu_mamba_v3_vs_reference_golden(case):
    from segmamba_b200.mamba_simple import Mamba
    name, seed, batch, d_model, L, ns = case
    gold = gi.load("mamba_" + name)
    m = Mamba(d_model=d_model, d_state=16, d_conv=4, expand=2, bimamba_type="v3", nslices=ns)
    m.load_state_dict({k[len("param."):]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("param.")}, strict=True)
    r = np.random.RandomState(seed + 1000)
    x = torch.from_numpy(r.standard_normal((batch, L, d_model)).astype(np.float32)).requires_grad_()
    dout = torch.from_numpy(r.standard_normal((batch, L, d_model)).astype(np.float32))
    out = m(x)
    assert_close(out, gold["out"], 1e-3, "out")
    names = [n for n, _ in m.named_parameters()]
    grads = torch.autograd.grad(out, [x] + [p for _, p in m.named_parameters()], dout)
    assert_close(grads[0], gold["dx"], 2e-3, "dx")
    for n, g in zip(names, grads[1:]):
        assert_close(g, gold["grad." + n], 3e-3, "grad." + n)


def test_emu_segmamba_vs_reference_golden():
    """full module forward + backward on the tiny fixture: every native kernel (scan, conv1d, permute, instance norm) in its
    place inside the model, dense convs / GEMMs by CPU PyTorch, against the reference SegMamba's golden outputs."""
    from segmamba_b200.segmamba import SegMamba
    c = gi.MODEL_CASE
    gold = gi.load("model_" + c["name"])
    m = SegMamba(in_chans=c["in_chans"], out_chans=c["out_chans"], depths=c["depths"], feat_size=c["feat_size"],
                 hidden_size=c["hidden_size"])
    keys = [str(k) for k in gold["state_dict_keys"]]
    shapes = [tuple(int(s) for s in str(x).split(",")) if str(x) else () for x in gold["state_dict_shapes"]]
    m.load_state_dict(gi.randomize_state_dict(gi.reference_like_init(keys, shapes), c["seed"]), strict=True)
    m.train()
    x = gi.model_input(c["seed"] + 1, (c["batch"], c["in_chans"], c["spatial"], c["spatial"], c["spatial"]))
    out = m(x)
    assert_close(out, gold["out"], 1e-3, "logits")
    r = np.random.RandomState(c["seed"] + 2)
    dout = torch.from_numpy(r.standard_normal(tuple(out.shape)).astype(np.float32)) / out.numel() ** 0.5
    names = [n for n, _ in m.named_parameters()]
    grads = torch.autograd.grad(out, [p for _, p in m.named_parameters()], dout)
    norms = np.array([float(g.double().norm()) for g in grads])
    ref_norms = gold["grad_norms"]
    floor = 1e-4 * float(ref_norms.max())
    bad = [(n, a, b) for n, a, b in zip(names, norms, ref_norms) if abs(a - b) > 5e-3 * b + floor]
    assert not bad, f"grad norm mismatch: {bad[:5]}"
    for n, g, rn in zip(names, grads, ref_norms):
        if "grad." + n in gold.files and rn > 10 * floor:
            assert_close(g, gold["grad." + n], 2e-2, "grad." + n)


# --------------------------------------------------------------------------------------------------------- fused layer norm
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["f32", "bf16", "f16"])
@pytest.mark.parametrize("rows,C", [(4096, 48), (1000, 96), (513, 192), (77, 384), (130, 32), (9, 768), (64, 8)],
                         ids=lambda v: str(v))
def test_emu_fused_layer_norm(dtype, rows, C):
    from segmamba_b200.layer_norm import fused_layer_norm, supported
    torch.manual_seed(rows + C)
    x = (torch.randn(2, rows, C) * 1.7 + 0.4).to(dtype).requires_grad_()
    w = (torch.rand(C) + 0.5).requires_grad_()
    b = (torch.randn(C) * 0.3).requires_grad_()
    dy = torch.randn(2, rows, C).to(dtype)
    if not supported(x, C):
        pytest.skip("shape outside the kernel's vector constraints for this dtype")
    y = fused_layer_norm(x, w, b, 1e-5)
    assert y.dtype == dtype and y.shape == x.shape
    gx, gw, gb = torch.autograd.grad(y, [x, w, b], dy)
    xr = x.detach().float().requires_grad_()
    wr, br = w.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
    yr = F.layer_norm(xr, (C,), wr, br, 1e-5)
    rx, rw, rb = torch.autograd.grad(yr, [xr, wr, br], dy.float())
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert_close(y, yr, tol, "y")
    assert_close(gx, rx, 1e-4 if dtype == torch.float32 else 2e-2, "dx")
    assert_close(gw, rw, 1e-4 if dtype == torch.float32 else 2e-2, "dweight")
    assert_close(gb, rb, 1e-4 if dtype == torch.float32 else 2e-2, "dbias")


def test_emu_fused_layer_norm_large_mean_and_no_bias():
    from segmamba_b200.layer_norm import fused_layer_norm
    x = (torch.randn(300, 48) * 0.01 + 30.0)
    w = torch.ones(48)
    assert_close(fused_layer_norm(x, w, None, 1e-5), F.layer_norm(x.double(), (48,), eps=1e-5).float(), 2e-3, "mean 30, std 0.01")


def test_emu_segmamba_with_fused_layer_norm(monkeypatch):
    """the full module with the fused LayerNorm switched on equals the module with nn.LayerNorm (forward and parameter grads)."""
    from segmamba_b200 import layer_norm
    from segmamba_b200.segmamba import SegMamba
    c = gi.MODEL_CASE
    torch.manual_seed(3)
    m = SegMamba(in_chans=c["in_chans"], out_chans=c["out_chans"], depths=c["depths"], feat_size=c["feat_size"],
                 hidden_size=c["hidden_size"]).train()
    x = gi.model_input(c["seed"] + 1, (c["batch"], c["in_chans"], c["spatial"], c["spatial"], c["spatial"]))
    outs, grads = [], []
    for on in (False, True):
        monkeypatch.setattr(layer_norm, "ENABLED", on)
        out = m(x)
        g = torch.autograd.grad(out.square().mean(), [p for p in m.parameters()])
        outs.append(out.detach())
        grads.append(g)
    assert_close(outs[1], outs[0], 1e-4, "logits, fused vs nn.LayerNorm")
    num = sum(float((a.double() - b.double()).pow(2).sum()) for a, b in zip(grads[1], grads[0]))
    den = sum(float(b.double().pow(2).sum()) for b in grads[0])
    assert (num / den) ** 0.5 < 1e-3


@pytest.mark.parametrize("seg_min", ["32", "64", "128", "256"])
def test_emu_scan_short_segments(monkeypatch, seg_min):
    """SMB_SEG_MIN tuning knob: segments shorter than the 256-position checkpoint interval (more, shorter serial chains for the
    small late-stage problems) give the same results, chunk states and checkpoints included."""
    monkeypatch.setenv("SMB_SEG_MIN", seg_min)
    for shape, direction in (((2, 40, 700, 16, 1), 0), ((1, 33, 2300, 16, 1), 1), ((2, 48, 512, 8, 2), 0)):
        batch, dim, L, N, G = shape
        d = rand_scan_inputs(300 + L, batch, dim, L, N, G, torch.float32, device="cpu")
        res = tg._run_fwd_bwd(d, direction=direction, use_hstates=True)
        ref = tg._oracle_fwd_bwd(d, flip=bool(direction))
        tg._compare(res, ref, torch.float32, True)


# ----------------------------------------------------------------------------- software-pipelined forward scan (SMB_FWD_V2=1)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("shape", [(2, 40, 700, 16, 1), (1, 33, 31, 16, 1), (2, 48, 600, 8, 2), (1, 64, 2304, 16, 1), (1, 32, 256, 16, 1)],
                         ids=lambda s: "b%d_d%d_L%d_n%d_g%d" % s)
@pytest.mark.parametrize("direction", [0, 1], ids=["fwd", "rev"])
def test_emu_scan_fwd_v2_vs_oracle(monkeypatch, dtype, shape, direction):
    """cp.async double-buffered forward kernels: same results as the oracle (and hence as the default kernels); the emulator
    defers every cp.async until the issuing thread waits for it, so a missing wait or barrier shows up as NaN poison."""
    monkeypatch.setenv("SMB_FWD_V2", "1")
    batch, dim, L, N, G = shape
    d = rand_scan_inputs(100 + L, batch, dim, L, N, G, dtype, device="cpu")
    res = tg._run_fwd_bwd(d, direction=direction, use_hstates=True)     # the backward consumes the v2 forward's hstates
    ref = tg._oracle_fwd_bwd(d, flip=bool(direction))
    tg._compare(res, ref, dtype, True)


@pytest.mark.parametrize("order", [1, 3], ids=["descending", "random3"])
def test_emu_scan_fwd_v2_thread_orders_and_layouts(monkeypatch, order):
    from segmamba_b200 import selective_scan_cuda as ssc
    monkeypatch.setenv("SMB_FWD_V2", "1")
    emu.emu_lib().smb_emu_set_reverse(order)
    batch, dim, L, N = 2, 64, 1504, 16
    d = rand_scan_inputs(7, batch, dim, L, N, 1, torch.bfloat16, device="cpu")
    hbl = lambda t: t.permute(1, 0, 2).contiguous().permute(1, 0, 2)            # the mixer's channel-major layout
    d2 = dict(d)
    for k in ("u", "delta", "z", "dout"):
        d2[k] = hbl(d[k])
    for direction in (0, 1):
        res = tg._run_fwd_bwd(d2, direction=direction, use_hstates=True)
        tg._compare(res, tg._oracle_fwd_bwd(d, flip=bool(direction)), torch.bfloat16, True)
    # same arithmetic as the default kernels (up to FMA contraction choices of the compiler), with and without z / out
    B, C = d["B"].unsqueeze(1), d["C"].unsqueeze(1)
    v2 = ssc.fwd_ex(d["u"], d["delta"], d["A"], B, C, d["D"], d["z"], d["delta_bias"], True, want_out=True, want_x=True, want_hstates=True)
    v2n = ssc.fwd_ex(d["u"], d["delta"], d["A"], B, C, d["D"], None, d["delta_bias"], True, want_out=True, want_x=False)
    monkeypatch.setenv("SMB_FWD_V2", "0")
    v1 = ssc.fwd_ex(d["u"], d["delta"], d["A"], B, C, d["D"], d["z"], d["delta_bias"], True, want_out=True, want_x=True, want_hstates=True)
    v1n = ssc.fwd_ex(d["u"], d["delta"], d["A"], B, C, d["D"], None, d["delta_bias"], True, want_out=True, want_x=False)
    for a, b, tol in zip(v2, v1, (8e-3, 1e-5, 8e-3, 1e-5)):          # (out, x, out_z, hstates): one bf16 ulp / fp32 round-off
        assert_close(a, b, tol, "pipelined vs default kernels")
    assert_close(v2n[0], v1n[0], 8e-3, "pipelined vs default kernels, no z")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("shape", [(2, 40, 700, 16, 1), (1, 33, 31, 16, 1), (2, 48, 600, 8, 2), (1, 64, 2304, 16, 1)],
                         ids=lambda s: "b%d_d%d_L%d_n%d_g%d" % s)
@pytest.mark.parametrize("direction", [0, 1], ids=["fwd", "rev"])
def test_emu_scan_ragg_v2_vs_oracle(monkeypatch, dtype, shape, direction):
    """software-pipelined R1 (SMB_RAGG_V2=1) together with the pipelined forward: full forward + both backward paths."""
    monkeypatch.setenv("SMB_RAGG_V2", "1")
    monkeypatch.setenv("SMB_FWD_V2", "1")
    batch, dim, L, N, G = shape
    d = rand_scan_inputs(100 + L, batch, dim, L, N, G, dtype, device="cpu")
    for has_z in (True, False):
        res = tg._run_fwd_bwd(d, has_z=has_z, direction=direction, use_hstates=True)
        ref = tg._oracle_fwd_bwd(d, has_z=has_z, flip=bool(direction))
        tg._compare(res, ref, dtype, has_z)


def test_emu_scan_ragg_v2_thread_orders(monkeypatch):
    monkeypatch.setenv("SMB_RAGG_V2", "1")
    d = rand_scan_inputs(9, 2, 40, 900, 16, 1, torch.bfloat16, device="cpu")
    for order in (1, 4):
        emu.emu_lib().smb_emu_set_reverse(order)
        for direction in (0, 1):
            res = tg._run_fwd_bwd(d, direction=direction, use_hstates=True)
            tg._compare(res, tg._oracle_fwd_bwd(d, flip=bool(direction)), torch.bfloat16, True)


# ----------------------------------------------------------------------------- second-generation R3 (SMB_R3_V2=1)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["f32", "bf16", "f16"])
@pytest.mark.parametrize("shape", [(2, 40, 700, 16, 1), (1, 33, 31, 16, 1), (2, 48, 600, 8, 2), (1, 64, 2304, 16, 1), (1, 7, 1021, 16, 1)],
                         ids=lambda s: "b%d_d%d_L%d_n%d_g%d" % s)
@pytest.mark.parametrize("direction", [0, 1], ids=["fwd", "rev"])
def test_emu_scan_r3_v2_vs_oracle(monkeypatch, dtype, shape, direction):
    """scan_bwd_main2_kernel (predicated Kogge-Stone steps, folded chunk seeds, two states per barrier round, vector dB / dC
    reductions, shared-memory dA partials): all gradients against the oracle, with and without z, ragged lengths (scalar
    atomics path: L % 4 != 0), partial channel octets and groups."""
    monkeypatch.setenv("SMB_R3_V2", "1")
    batch, dim, L, N, G = shape
    d = rand_scan_inputs(100 + L, batch, dim, L, N, G, dtype, device="cpu")
    for has_z in (True, False):
        res = tg._run_fwd_bwd(d, has_z=has_z, direction=direction, use_hstates=True)
        ref = tg._oracle_fwd_bwd(d, has_z=has_z, flip=bool(direction))
        tg._compare(res, ref, dtype, has_z)


def test_emu_scan_r3_v2_matches_default_and_thread_orders(monkeypatch):
    """same inputs through both R3 generations: gradients agree to fp32 round-off (same arithmetic up to the order of the
    fp32 sums), also with the threads of a block resumed in descending / pseudo-random order (missing barrier => mismatch)."""
    d = rand_scan_inputs(11, 2, 40, 900, 16, 1, torch.float32, device="cpu")
    for direction in (0, 1):
        monkeypatch.setenv("SMB_R3_V2", "0")
        emu.emu_lib().smb_emu_set_reverse(0)
        ref = tg._run_fwd_bwd(d, direction=direction, use_hstates=True)
        monkeypatch.setenv("SMB_R3_V2", "1")
        for order in (0, 1, 5):
            emu.emu_lib().smb_emu_set_reverse(order)
            got = tg._run_fwd_bwd(d, direction=direction, use_hstates=True)
            for a, b in zip(got[4], ref[4]):
                if a is not None and b is not None:
                    assert_close(a, b, 2e-5, "R3 v2 vs default R3")


# ------------------------------------------------------------------------------------------- properties and inference path
def test_emu_scan_linearity_and_reverse_equals_flip():
    """size-independent properties: y is linear in u for fixed delta, B, C; direction=1 equals the op on flipped operands."""
    from segmamba_b200 import selective_scan_cuda as ssc
    d = rand_scan_inputs(5, 2, 40, 1000, 16, 1, torch.float32, device="cpu")
    B, C = d["B"].unsqueeze(1), d["C"].unsqueeze(1)
    f = lambda u, direction=0: ssc.fwd_ex(u, d["delta"], d["A"], B, C, None, None, d["delta_bias"], True, direction=direction, want_x=False)[0]
    u2 = torch.randn_like(d["u"])
    assert_close(f(0.7 * d["u"] - 1.3 * u2), 0.7 * f(d["u"]) - 1.3 * f(u2), 1e-5, "linearity in u")
    fl = lambda t: t.flip(-1).contiguous()
    rev = f(d["u"], direction=1)
    ref = ssc.fwd_ex(fl(d["u"]), fl(d["delta"]), d["A"], fl(B), fl(C), None, None, d["delta_bias"], True, want_x=False)[0].flip(-1)
    assert_close(rev, ref, 1e-6, "reverse walk vs flipped operands")


def test_emu_sliding_window_with_segmamba():
    """the GPU-resident sliding-window driver with the native model as predictor (eval mode, no_grad), against a window-by-window
    evaluation with explicit gaussian accumulation; mirror TTA keeps shape and finiteness."""
    from segmamba_b200 import sliding_window as sw
    from segmamba_b200.segmamba import SegMamba
    c = gi.MODEL_CASE
    torch.manual_seed(0)
    m = SegMamba(in_chans=4, out_chans=4, depths=c["depths"], feat_size=c["feat_size"], hidden_size=c["hidden_size"]).eval()
    x = torch.rand(1, 4, 40, 48, 33)
    with torch.no_grad():
        out = sw.sliding_window_inference(x, (32, 32, 32), 2, m, overlap=0.5, mode="gaussian")
        assert out.shape == (1, 4, 40, 48, 33)
        starts = sw.window_starts((40, 48, 33), (32, 32, 32), 0.5)
        w = sw.gaussian_importance_map((32, 32, 32))[None, None]
        acc, cnt = torch.zeros_like(out), torch.zeros(1, 1, 40, 48, 33)
        for (a, b, cc) in starts:
            sl = (slice(None), slice(None), slice(a, a + 32), slice(b, b + 32), slice(cc, cc + 32))
            acc[sl] += m(x[sl].contiguous()) * w
            cnt[sl] += w
        assert torch.allclose(out, acc / cnt, rtol=1e-3, atol=1e-4)
        tta = sw.sliding_window_inference(x, (32, 32, 32), 2, m, mirror_axes=(0,))
        assert tta.shape == out.shape and torch.isfinite(tta).all()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("shape", [(2, 40, 1300, 4), (1, 33, 1031, 3), (2, 8, 7, 2), (1, 5, 4096, 4), (1, 3, 17, 4)], ids=lambda s: "b%d_d%d_L%d_w%d" % s)
@pytest.mark.parametrize("direction", [0, 1], ids=["fwd", "rev"])
def test_emu_conv_v2_vs_oracle(monkeypatch, dtype, shape, direction):
    """16-positions-per-thread conv1d kernels for 16-bit activations (SMB_CONV_V2=1): forward, dx (in place), dweight, dbias."""
    from segmamba_b200 import causal_conv1d_cuda as cc
    monkeypatch.setenv("SMB_CONV_V2", "1")
    orc = _oracle()
    batch, dim, L, width = shape
    d = gi.conv_inputs(200 + L, batch, dim, L, width)
    x, dout, w, b = d["x"].to(dtype), d["dout"].to(dtype), d["weight"], d["bias"]
    xz = torch.empty(2 * dim, batch, L, dtype=dtype).permute(1, 0, 2)
    xz[:, :dim] = x
    xv = xz[:, :dim]
    dxz = torch.zeros_like(xz)
    for silu in (True, False):
        out = cc.causal_conv1d_fwd_ex(xv, w, b, silu, direction=direction)
        dx, dw, db = cc.causal_conv1d_bwd_ex(xv, w, b, dout, dxz[:, :dim], silu, direction=direction)
        f = (lambda t: t.flip(-1)) if direction else (lambda t: t)
        o = f(orc.causal_conv1d_fwd_raw(f(x.float()), w, b, silu))
        odx, odw, odb = orc.causal_conv1d_bwd_raw(f(x.float()), w, b, f(dout.float()), silu)
        assert_close(out, o, TOL[dtype], "out")
        assert_close(dxz[:, :dim], f(odx), GRAD_TOL[dtype], "dx (in place)")
        assert float(dxz[:, dim:].abs().max()) == 0.0
        assert_close(dw, odw, GRAD_TOL[dtype], "dweight")
        assert_close(db, odb, GRAD_TOL[dtype], "dbias")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("shape", [(2, 40, 704, 16, 1), (1, 33, 31, 16, 1), (2, 48, 600, 8, 1), (1, 64, 2304, 16, 1), (1, 16, 264, 16, 1)],
                         ids=lambda s: "b%d_d%d_L%d_n%d_g%d" % s)
@pytest.mark.parametrize("direction", [0, 1], ids=["fwd", "rev"])
def test_emu_scan_fwd_tma_vs_oracle(monkeypatch, dtype, shape, direction):
    """TMA mode of the pipelined forward kernels (SMB_FWD_V2=2): tiles staged by cp.async.bulk.tensor with the 64-byte swizzle and
    hardware zero fill of ragged tiles, completion through an mbarrier.  The emulator performs the box copies when a thread waits
    on the barrier and aborts if the byte count differs from expect_tx."""
    monkeypatch.setenv("SMB_FWD_V2", "2")
    batch, dim, L, N, G = shape
    d = rand_scan_inputs(100 + L, batch, dim, L, N, G, dtype, device="cpu")
    for has_z in (True, False):
        res = tg._run_fwd_bwd(d, has_z=has_z, direction=direction, use_hstates=True)
        ref = tg._oracle_fwd_bwd(d, has_z=has_z, flip=bool(direction))
        tg._compare(res, ref, dtype, has_z)


def test_emu_scan_fwd_tma_layouts_and_orders(monkeypatch):
    monkeypatch.setenv("SMB_FWD_V2", "2")
    batch, dim, L, N = 2, 64, 1504, 16
    d = rand_scan_inputs(7, batch, dim, L, N, 1, torch.bfloat16, device="cpu")
    hbl = lambda t: t.permute(1, 0, 2).contiguous().permute(1, 0, 2)            # channel-major: batch stride < channel stride
    d2 = dict(d)
    for k in ("u", "delta", "z", "dout"):
        d2[k] = hbl(d[k])
    for order in (0, 1, 6):
        emu.emu_lib().smb_emu_set_reverse(order)
        for direction in (0, 1):
            res = tg._run_fwd_bwd(d2, direction=direction, use_hstates=True)
            tg._compare(res, tg._oracle_fwd_bwd(d, flip=bool(direction)), torch.bfloat16, True)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_emu_seq_permute_into_out_and_accumulate(dtype):
    """the `out=` / `accumulate=` form of smb_seq_permute: dst (+)= permuted src."""
    from segmamba_b200 import causal_conv1d_cuda as cc
    x = torch.randn(2, 12, 512).to(dtype)
    ref = x.reshape(2, 12, 512 // 16, 16).permute(0, 1, 3, 2).flatten(-2)          # inverse permutation (mamba_simple.py:261)
    y0 = torch.randn(2, 12, 512).to(dtype)
    y = y0.clone()
    got = cc.seq_permute(x, 16, inverse=True, out=y, accumulate=True)
    assert got is y
    assert_close(y, (y0.float() + ref.float()), 1e-6 if dtype == torch.float32 else 1e-2, "accumulate")
    z = torch.empty_like(x)
    cc.seq_permute(x, 16, inverse=True, out=z)
    assert torch.equal(z, ref)


@pytest.mark.parametrize("mode", ["1", "2"], ids=["cp_async", "tma"])
def test_emu_scan_fwd_v2_mixer_layout(monkeypatch, mode):
    """the operand layout the mixer really produces (selective_scan_interface.py: MambaInnerFnNoOutProj): u, delta, z channel-major
    and B, C row slices of one (R + 2N, b * l) matrix, i.e. state stride b * l > batch stride l -- the tensor maps then take the
    batch axis before the channel / state axis."""
    monkeypatch.setenv("SMB_FWD_V2", mode)
    monkeypatch.setenv("SMB_RAGG_V2", "1")
    batch, dim, L, N = 2, 40, 1000, 16
    d = rand_scan_inputs(13, batch, dim, L, N, 1, torch.bfloat16, device="cpu")
    hbl = lambda t: t.permute(1, 0, 2).contiguous().permute(1, 0, 2)
    x_dblT = torch.zeros(3 + 2 * N, batch * L, dtype=torch.bfloat16)
    x_dblT[3:3 + N] = d["B"].permute(1, 0, 2).reshape(N, batch * L)
    x_dblT[3 + N:] = d["C"].permute(1, 0, 2).reshape(N, batch * L)
    Bm = x_dblT[3:3 + N].view(N, batch, L).permute(1, 0, 2).unsqueeze(1)
    Cm = x_dblT[-N:].view(N, batch, L).permute(1, 0, 2).unsqueeze(1)
    assert Bm.stride() == (L, N * batch * L, batch * L, 1) or Bm.stride(2) == batch * L
    d2 = dict(d)
    for k in ("u", "delta", "z", "dout"):
        d2[k] = hbl(d[k])
    d2["B"], d2["C"] = Bm, Cm
    for direction in (0, 1):
        res = tg._run_fwd_bwd(d2, direction=direction, use_hstates=True)
        tg._compare(res, tg._oracle_fwd_bwd(d, flip=bool(direction)), torch.bfloat16, True)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("L,ns", [(512, 16), (4096, 64), (600, 6), (1000, 10), (96, 8), (510, 5)], ids=lambda v: str(v))
def test_emu_seq_permute_v2(monkeypatch, dtype, L, ns):
    """4-byte-access permutation kernel (SMB_PERMUTE_V2=1): bit-identical to the default kernel in both directions, on strided
    (channel-major) views, with out= / accumulate=; odd shapes fall back to the default kernel."""
    from segmamba_b200 import causal_conv1d_cuda as cc
    torch.manual_seed(L + ns)
    base = torch.randn(12, 2, L).to(dtype)
    x = base.permute(1, 0, 2)                                  # (batch, dim, L) view with the mixer's channel-major strides
    y0 = torch.randn(2, 12, L).to(dtype)
    res = {}
    for v2 in ("0", "1"):
        monkeypatch.setenv("SMB_PERMUTE_V2", v2)
        for order in (0, 3):
            emu.emu_lib().smb_emu_set_reverse(order)
            f = cc.seq_permute(x, ns)
            b = cc.seq_permute(f, ns, inverse=True)
            acc = cc.seq_permute(x, ns, inverse=True, out=y0.clone(), accumulate=True)
            res[(v2, order)] = (f, b, acc)
    ref = x.reshape(2, 12, ns, L // ns).transpose(-1, -2).flatten(-2)              # mamba_simple.py:245-247
    for key, (f, b, acc) in res.items():
        assert torch.equal(f, ref), key
        assert torch.equal(b, x), key
        assert torch.equal(acc, res[("0", 0)][2]), key


@pytest.mark.parametrize("opc", ["1", "3"], ids=["opc1", "opc3"])
@pytest.mark.parametrize("order", [0, 3], ids=["ascending", "random3"])
def test_emu_scan_dense_checkpoints_and_octet_walk(monkeypatch, opc, order):
    """scan-free main backward pass on the dense checkpoints (state / local adjoint at every 8th position), one CTA walking 1 or
    3 channel octets, under two thread orders: all gradients against the oracle (ragged length, partial octet, two B/C groups,
    both walk directions)."""
    monkeypatch.setenv("SMB_R3_OPC", opc)
    emu.emu_lib().smb_emu_set_reverse(order)
    for shape, direction in (((2, 44, 1003, 16, 1), 0), ((1, 48, 520, 8, 2), 1)):
        batch, dim, L, N, G = shape
        d = rand_scan_inputs(700 + L, batch, dim, L, N, G, torch.bfloat16, device="cpu")
        res = tg._run_fwd_bwd(d, direction=direction, use_hstates=True)
        ref = tg._oracle_fwd_bwd(d, flip=bool(direction))
        tg._compare(res, ref, torch.bfloat16, True)
    emu.emu_lib().smb_emu_set_reverse(0)
