"""GPU parity of the native causal conv1d and the inter-slice permutation."""
import pytest
import torch

import golden_inputs as gi
from util import GRAD_TOL, TOL, assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", gi.CONV_CASES, ids=lambda c: c[0])
def test_conv_vs_golden_and_oracle(case):
    from oracle import oracle as orc
    from segmamba_b200 import causal_conv1d_cuda as cc
    name, seed, batch, dim, L, width, has_b, silu = case
    d = {k: v.cuda() for k, v in gi.conv_inputs(seed, batch, dim, L, width).items()}
    b = d["bias"] if has_b else None
    out = cc.causal_conv1d_fwd(d["x"], d["weight"], b, silu)
    dx, dw, db = cc.causal_conv1d_bwd(d["x"], d["weight"], b, d["dout"], None, silu)
    gold = gi.load("conv_" + name)
    assert_close(out, gold["out"], 1e-3, "out vs golden")
    assert_close(dx, gold["dx"], 1e-3, "dx vs golden")
    assert_close(dw, gold["dweight"], 1e-3, "dweight vs golden")
    if has_b:
        assert_close(db, gold["dbias"], 1e-3, "dbias vs golden")
    o = orc.causal_conv1d_fwd_raw(d["x"], d["weight"], b, silu)
    assert_close(out, o, 1e-5, "out vs oracle")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16], ids=["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 96, 5000, 4), (1, 33, 1031, 3), (2, 8, 7, 2), (1, 192, 32768, 4)], ids=lambda s: "b%d_d%d_L%d_w%d" % s)
@pytest.mark.parametrize("direction", [0, 1], ids=["fwd", "rev"])
def test_conv_random_vs_oracle(dtype, shape, direction):
    from oracle import oracle as orc
    from segmamba_b200 import causal_conv1d_cuda as cc
    batch, dim, L, width = shape
    d = gi.conv_inputs(200 + L, batch, dim, L, width)
    x = d["x"].to(dtype).cuda()
    dout = d["dout"].to(dtype).cuda()
    w, b = d["weight"].cuda(), d["bias"].cuda()
    # channel-major ("HBL") strided views, and dx written in place into the left half of a dxz buffer (ssi.py:244-245,281)
    xz = torch.empty(2 * dim, batch, L, dtype=dtype, device="cuda").permute(1, 0, 2)
    xz[:, :dim] = x
    xv = xz[:, :dim]
    dxz = torch.zeros_like(xz)
    out = cc.causal_conv1d_fwd_ex(xv, w, b, True, direction=direction)
    dx, dw, db = cc.causal_conv1d_bwd_ex(xv, w, b, dout, dxz[:, :dim], True, direction=direction)
    f = (lambda t: t.flip(-1)) if direction else (lambda t: t)
    o = f(orc.causal_conv1d_fwd_raw(f(x.float().cpu()), w.cpu(), b.cpu(), True))
    odx, odw, odb = orc.causal_conv1d_bwd_raw(f(x.float().cpu()), w.cpu(), b.cpu(), f(dout.float().cpu()), True)
    assert_close(out, o, TOL[dtype], "out")
    assert_close(dxz[:, :dim], f(odx), GRAD_TOL[dtype], "dx (in place)")
    assert dx.data_ptr() == dxz[:, :dim].data_ptr()
    assert float(dxz[:, dim:].abs().max()) == 0.0
    assert_close(dw, odw, GRAD_TOL[dtype], "dweight")
    assert_close(db, odb, GRAD_TOL[dtype], "dbias")


def test_conv_race_determinism():
    """the reference's stress test (test_causal_conv1d.py:117-173), shortened: out/dx bit-exact across repeats."""
    from segmamba_b200 import causal_conv1d_cuda as cc
    d = {k: v.cuda() for k, v in gi.conv_inputs(3, 2, 64, 4096, 4).items()}
    out0 = cc.causal_conv1d_fwd(d["x"], d["weight"], d["bias"], True)
    dx0, dw0, db0 = cc.causal_conv1d_bwd(d["x"], d["weight"], d["bias"], d["dout"], None, True)
    for _ in range(200):
        out = cc.causal_conv1d_fwd(d["x"], d["weight"], d["bias"], True)
        dx, dw, db = cc.causal_conv1d_bwd(d["x"], d["weight"], d["bias"], d["dout"], None, True)
        assert torch.equal(out, out0) and torch.equal(dx, dx0)
        assert torch.allclose(dw, dw0, rtol=1e-4, atol=1e-4) and torch.allclose(db, db0, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_seq_permute(dtype):
    """bit-exact against the reference's stack(chunk).flatten / reshape.permute.flatten (mamba_simple.py:245-247,261)."""
    from segmamba_b200 import causal_conv1d_cuda as cc
    for (b, d, L, ns) in [(2, 24, 4096, 64), (1, 10, 512, 8), (2, 6, 96, 16)]:
        x = torch.randn(b, d, L, device="cuda").to(dtype)
        ref = torch.stack(x.chunk(ns, dim=-1), dim=-1).flatten(-2)
        got = cc.seq_permute(x, ns)
        assert torch.equal(got, ref)
        back = cc.seq_permute(got, ns, inverse=True)
        ref_back = ref.reshape(b, d, L // ns, ns).permute(0, 1, 3, 2).flatten(-2)
        assert torch.equal(back, ref_back) and torch.equal(back, x)
        hbl = x.permute(1, 0, 2).contiguous().permute(1, 0, 2)
        assert torch.equal(cc.seq_permute(hbl, ns), ref)
