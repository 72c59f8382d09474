"""Pins the CPU oracle (oracle/segmamba_oracle.c + oracle/oracle.py) against golden vectors produced by the
REFERENCE's own pure-PyTorch functions (oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import golden_inputs as gi
from oracle import oracle as orc


def _close(a, b, rtol, atol, what):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    assert bool((err <= tol).all()), f"{what}: max err {err.max().item():.3e} (|ref| max {b.abs().max().item():.3e})"


@pytest.mark.parametrize("case", gi.SCAN_CASES + [gi.CONFIG1], ids=lambda c: c[0])
def test_scan_oracle_vs_reference(case):
    name, seed, batch, dim, L, N, G, tl, has_D, has_z, has_b, sp = case
    d = gi.scan_inputs(seed, batch, dim, L, N, G, tl)
    gold = gi.load("scan_" + name)
    assert np.allclose(gold["input_checksum"], [float(d["u"].double().sum()), float(d["B"].double().sum())])
    D = d["D"] if has_D else None
    z = d["z"] if has_z else None
    bias = d["delta_bias"] if has_b else None
    y, oz, last, xc = orc.selective_scan_fwd_raw(d["u"], d["delta"], d["A"], d["B"], d["C"], D, z, bias, sp)
    out = oz if has_z else y
    # the reference ref runs in fp32; the oracle in fp64 -> fp32-rounding-level agreement
    _close(out, gold["out"], 2e-5, 2e-5 * float(np.abs(gold["out"]).max()), "out")
    _close(last, gold["last_state"], 2e-5, 2e-5 * float(np.abs(gold["last_state"]).max()), "last_state")
    # last chunk entry of xchunks is the last state (ssi.py:40)
    _close(xc[:, :, -1, 1::2], gold["last_state"], 2e-5, 2e-5 * float(np.abs(gold["last_state"]).max()), "x[-1]")
    if name == gi.CONFIG1[0]:
        return
    g = orc.selective_scan_bwd_raw(d["u"], d["delta"], d["A"], d["B"], d["C"], D, z, bias, sp, d["dout"])
    keymap = {"du": "du", "ddelta": "ddelta", "dA": "dA", "dB": "dB", "dC": "dC", "dD": "dD", "dz": "dz",
              "ddelta_bias": "ddelta_bias"}
    for gk, ok in keymap.items():
        if gk not in gold.files:
            continue
        ref = gold[gk]
        got = g[ok]
        if got.dim() == 4 and ref.ndim == 3:
            got = got.squeeze(1)
        _close(got, ref, 1e-4, 1e-4 * float(np.abs(ref).max()), gk)


@pytest.mark.parametrize("case", gi.CONV_CASES, ids=lambda c: c[0])
def test_conv_oracle_vs_reference(case):
    name, seed, batch, dim, L, width, has_b, silu = case
    d = gi.conv_inputs(seed, batch, dim, L, width)
    gold = gi.load("conv_" + name)
    b = d["bias"] if has_b else None
    out = orc.causal_conv1d_fwd_raw(d["x"], d["weight"], b, silu)
    _close(out, gold["out"], 1e-5, 1e-5, "out")
    dx, dw, db = orc.causal_conv1d_bwd_raw(d["x"], d["weight"], b, d["dout"], silu)
    _close(dx, gold["dx"], 1e-5, 1e-5, "dx")
    _close(dw, gold["dweight"], 1e-4, 1e-4, "dweight")
    if has_b:
        _close(db, gold["dbias"], 1e-4, 1e-4, "dbias")


@pytest.mark.parametrize("case", gi.INNER_CASES, ids=lambda c: c[0])
def test_inner_oracle_vs_reference(case):
    name, seed, batch, d_model, L = case
    d = gi.inner_inputs(seed, batch, d_model, L)
    gold = gi.load("inner_" + name)
    keys = ["xz", "conv1d_weight", "conv1d_bias", "x_proj_weight", "delta_proj_weight", "A", "D", "delta_bias"]
    lv = {k: d[k].clone().requires_grad_() for k in keys}
    out = orc.mamba_inner_no_out_proj(lv["xz"], lv["conv1d_weight"], lv["conv1d_bias"], lv["x_proj_weight"],
                                      lv["delta_proj_weight"], lv["A"], lv["D"], lv["delta_bias"], True)
    _close(out.detach(), gold["out"], 1e-4, 1e-4, "out")
    grads = torch.autograd.grad(out, [lv[k] for k in keys], d["dout"])
    for k, g in zip(keys, grads):
        ref = gold["d" + k]
        _close(g, ref, 2e-4, 2e-4 * float(np.abs(ref).max()), "d" + k)


@pytest.mark.parametrize("case", gi.MAMBA_CASES, ids=lambda c: c[0])
def test_mamba_v3_oracle_vs_reference(case):
    name, seed, batch, d_model, L, ns = case
    gold = gi.load("mamba_" + name)
    p = {k[len("param."):]: torch.from_numpy(gold[k]).requires_grad_() for k in gold.files if k.startswith("param.")}
    r = np.random.RandomState(seed + 1000)
    x = torch.from_numpy(r.standard_normal((batch, L, d_model)).astype(np.float32)).requires_grad_()
    dout = torch.from_numpy(r.standard_normal((batch, L, d_model)).astype(np.float32))
    out = orc.mamba_v3_forward(p, x, ns)
    _close(out.detach(), gold["out"], 1e-4, 1e-4, "out")
    names = list(p.keys())
    grads = torch.autograd.grad(out, [x] + [p[n] for n in names], dout)
    _close(grads[0], gold["dx"], 2e-4, 2e-4 * float(np.abs(gold["dx"]).max()), "dx")
    for n, g in zip(names, grads[1:]):
        ref = gold["grad." + n]
        _close(g, ref, 5e-4, 5e-4 * float(np.abs(ref).max()), "grad." + n)


def test_model_oracle_vs_reference():
    c = gi.MODEL_CASE
    gold = gi.load("model_" + c["name"])
    # rebuild the randomised reference state_dict from the committed surface (names + shapes)
    keys = [str(k) for k in gold["state_dict_keys"]]
    shapes = [tuple(int(s) for s in str(x).split(",")) if str(x) else () for x in gold["state_dict_shapes"]]
    sd0 = gi.reference_like_init(keys, shapes)
    sd = gi.randomize_state_dict(sd0, c["seed"])
    x = gi.model_input(c["seed"] + 1, (c["batch"], c["in_chans"], c["spatial"], c["spatial"], c["spatial"]))
    out = orc.segmamba_forward(sd, x, depths=c["depths"])
    _close(out, gold["out"], 5e-4, 5e-4 * float(np.abs(gold["out"]).max()), "logits")


def test_sliding_window_oracle_vs_monai():
    gold = gi.load("sliding_window")
    st = orc.sliding_window_starts((155, 240, 240), (128, 128, 128), 0.5)
    assert np.array_equal(np.array(st), gold["starts_brats"])
    st = orc.sliding_window_starts((40, 50, 33), (32, 32, 32), 0.5)
    assert np.array_equal(np.array(st), gold["starts_small"])
    g32 = orc.gaussian_importance_map((32, 32, 32))
    _close(g32, gold["gauss32"], 1e-6, 1e-7, "gauss32")
    g128 = orc.gaussian_importance_map((128, 128, 128))
    idx = np.arange(128)
    _close(g128.numpy()[idx, idx, idx], gold["gauss128_diag"], 1e-6, 1e-7, "gauss128")
