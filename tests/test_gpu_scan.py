"""GPU parity of the native selective scan (through the selective_scan_cuda drop-in -> C ABI) against the CPU
oracle and the committed golden vectors (reference outputs)."""
import pytest
import torch

import golden_inputs as gi
from util import GRAD_TOL, TOL, assert_close, rand_scan_inputs

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import oracle as orc
    return orc


def _run_fwd_bwd(d, has_D=True, has_z=True, has_b=True, softplus=True, direction=0, use_hstates=False):
    from segmamba_b200 import selective_scan_cuda as ssc
    D = d["D"] if has_D else None
    z = d["z"] if has_z else None
    bias = d["delta_bias"] if has_b else None
    B = d["B"] if d["B"].dim() == 4 else d["B"].unsqueeze(1)
    C = d["C"] if d["C"].dim() == 4 else d["C"].unsqueeze(1)
    out, x, out_z, hst, hd = ssc.fwd_ex(d["u"], d["delta"], d["A"], B, C, D, z, bias, softplus, direction=direction,
                                        want_out=True, want_x=True, want_hstates=True, want_hdense=True)
    # g: chunk-parallel path (also returns the recomputed out_z) -- with the saved states it is the scan-free main pass on the
    # dense checkpoints, without them the recompute + warp-scan one; g2: the state-stash sweep path
    g = ssc.bwd_ex(d["u"], d["delta"], d["A"], B, C, D, z, bias, d["dout"], None, softplus, True, direction=direction,
                   hstates=hst if use_hstates else None, hdense=hd if use_hstates else None, low_memory=True)
    g2 = ssc.bwd_ex(d["u"], d["delta"], d["A"], B, C, D, z, bias, d["dout"], None, softplus, False, direction=direction,
                    hstates=hst if use_hstates else None, low_memory=False)
    torch.cuda.synchronize()
    return out, x, out_z, hst, g, g2


def _oracle_fwd_bwd(d, has_D=True, has_z=True, has_b=True, softplus=True, flip=False):
    orc = _oracle()
    f = (lambda t: t.flip(-1)) if flip else (lambda t: t)
    c = {k: (f(v.float().cpu()) if k in ("u", "delta", "z", "B", "C", "dout") else v.float().cpu()) for k, v in d.items()}
    D = c["D"] if has_D else None
    z = c["z"] if has_z else None
    bias = c["delta_bias"] if has_b else None
    y, oz, last, xc = orc.selective_scan_fwd_raw(c["u"], c["delta"], c["A"], c["B"], c["C"], D, z, bias, softplus)
    g = orc.selective_scan_bwd_raw(c["u"], c["delta"], c["A"], c["B"], c["C"], D, z, bias, softplus, c["dout"])
    if flip:
        y = y.flip(-1)
        oz = oz.flip(-1) if oz is not None else None
        for k in ("du", "ddelta", "dz", "dB", "dC"):
            if g[k] is not None:
                g[k] = g[k].flip(-1)
    return y, oz, last, xc, g


def _compare(res, ref, dtype, has_z):
    out, x, out_z, hst, g = res[:5]
    if len(res) > 5:
        _compare_grads(res[5], ref[4], dtype, has_z, "sweep:")
    y, oz, last, xc, go = ref
    tol, gtol = TOL[dtype], GRAD_TOL[dtype]
    assert_close(out, y, tol, "out")
    if has_z:
        assert_close(out_z, oz, tol, "out_z")
        assert_close(g[8], oz, tol, "recomputed out_z")
    assert_close(x[:, :, -1, 1::2], last, tol, "last_state")
    assert_close(x[..., 1::2], xc[..., 1::2], tol, "x (chunk states)")
    _compare_grads(g, go, dtype, has_z, "recompute:")


def _compare_grads(g, go, dtype, has_z, tag):
    gtol = GRAD_TOL[dtype]
    du, ddelta, dA, dB, dC, dD, dbias, dz, _ = g
    assert_close(du, go["du"], gtol, tag + "du")
    assert_close(ddelta, go["ddelta"], gtol, tag + "ddelta")
    assert_close(dA, go["dA"], gtol, tag + "dA")
    assert_close(dB, go["dB"], gtol, tag + "dB")
    assert_close(dC, go["dC"], gtol, tag + "dC")
    if dD is not None:
        assert_close(dD, go["dD"], gtol, tag + "dD")
    if dbias is not None:
        assert_close(dbias, go["ddelta_bias"], gtol, tag + "ddelta_bias")
    if has_z:
        assert_close(dz, go["dz"], gtol, tag + "dz")


@pytest.mark.parametrize("case", gi.SCAN_CASES, ids=lambda c: c[0])
def test_scan_vs_golden_and_oracle(case):
    """the reference's own outputs (golden) and the oracle, on the committed cases (fp32)."""
    name, seed, batch, dim, L, N, G, tl, has_D, has_z, has_b, sp = case
    d = {k: v.cuda() for k, v in gi.scan_inputs(seed, batch, dim, L, N, G, tl).items()}
    res = _run_fwd_bwd(d, has_D, has_z, has_b, sp)
    ref = _oracle_fwd_bwd(d, has_D, has_z, has_b, sp)
    _compare(res, ref, torch.float32, has_z)
    gold = gi.load("scan_" + name)
    assert_close(res[2] if has_z else res[0], gold["out"], 1e-3, "out vs reference golden")
    assert_close(res[1][:, :, -1, 1::2], gold["last_state"], 1e-3, "last_state vs reference golden")
    names = ["du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias", "dz"]
    for i, k in enumerate(names):
        if k in gold.files and res[4][i] is not None:
            got = res[5][i]
            if got.dim() == 4 and gold[k].ndim == 3:
                got = got.squeeze(1)
            assert_close(got, gold[k], 2e-3, k + " vs reference golden")


def test_scan_config1_golden():
    """BASELINE.json configs[0]: B=1, L=4096, D=16, N=16."""
    name, seed, batch, dim, L, N, G, tl, has_D, has_z, has_b, sp = gi.CONFIG1
    d = {k: v.cuda() for k, v in gi.scan_inputs(seed, batch, dim, L, N, G, tl).items()}
    from segmamba_b200.selective_scan_interface import selective_scan_fn
    out, last = selective_scan_fn(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["z"], d["delta_bias"], True, True)
    gold = gi.load("scan_" + name)
    assert_close(out, gold["out"], 1e-3, "config1 out")
    assert_close(last, gold["last_state"], 1e-3, "config1 last_state")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16], ids=["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 96, 5000, 16, 1), (1, 40, 777, 16, 1), (2, 48, 2048, 8, 2), (1, 33, 31, 16, 1),
                                   (3, 64, 8192, 16, 1)], ids=lambda s: "b%d_d%d_L%d_n%d_g%d" % s)
@pytest.mark.parametrize("direction", [0, 1], ids=["fwd", "rev"])
def test_scan_random_vs_oracle(dtype, shape, direction):
    """ragged lengths, dim not a multiple of 32, groups, both walk directions, all three I/O dtypes."""
    batch, dim, L, N, G = shape
    d = rand_scan_inputs(100 + L, batch, dim, L, N, G, dtype)
    res = _run_fwd_bwd(d, direction=direction, use_hstates=(L % 2 == 0))
    ref = _oracle_fwd_bwd(d, flip=bool(direction))
    if direction:   # x / last_state are in walk order: the oracle ran on flipped inputs, so they already match
        pass
    _compare(res, ref, dtype, True)


def test_scan_strided_hbl_layout():
    """the reference's channel-major "HBL" layout: u/delta/z/dout are views with strides (L, B*L, 1)."""
    batch, dim, L, N = 2, 64, 1500, 16
    d = rand_scan_inputs(7, batch, dim, L, N)
    hbl = lambda t: t.permute(1, 0, 2).contiguous().permute(1, 0, 2)
    d2 = dict(d)
    for k in ("u", "delta", "z", "dout"):
        d2[k] = hbl(d[k])
        assert d2[k].stride() == (L, batch * L, 1)
    res = _run_fwd_bwd(d2)
    ref = _oracle_fwd_bwd(d)
    _compare(res, ref, torch.float32, True)


def test_scan_hstates_equals_recompute():
    d = rand_scan_inputs(11, 2, 96, 3000, 16)
    ra, rb = _run_fwd_bwd(d, use_hstates=True), _run_fwd_bwd(d, use_hstates=False)
    for a, b in ((ra[4], rb[4]), (ra[5], rb[5])):
        for x, y, n in zip(a[:2], b[:2], ("du", "ddelta")):
            assert_close(x, y, 1e-5, n)


def test_scan_full_size_stage0():
    """BASELINE size (stage 0: D=96, L=262144, N=16): parity against the oracle on a channel subset, plus the
    size-independent linearity property y(a*u1 + b*u2) = a*y(u1) + b*y(u2) for fixed delta, B, C."""
    from segmamba_b200 import selective_scan_cuda as ssc
    batch, dim, L, N = 1, 96, 262144, 16
    d = rand_scan_inputs(5, batch, dim, L, N)
    B, C = d["B"].unsqueeze(1), d["C"].unsqueeze(1)
    out, x, out_z, _ = ssc.fwd_ex(d["u"], d["delta"], d["A"], B, C, d["D"], d["z"], d["delta_bias"], True)
    sel = [0, 31, 32, 95]
    orc = _oracle()
    y, oz, last, xc = orc.selective_scan_fwd_raw(d["u"][:, sel].cpu(), d["delta"][:, sel].cpu(), d["A"][sel].cpu(), d["B"].cpu(),
                                                 d["C"].cpu(), d["D"][sel].cpu(), d["z"][:, sel].cpu(), d["delta_bias"][sel].cpu(), True)
    assert_close(out[:, sel], y, 1e-3, "full-size out")
    assert_close(out_z[:, sel], oz, 1e-3, "full-size out_z")
    assert_close(x[:, sel][..., 1::2], xc[..., 1::2], 1e-3, "full-size chunk states")
    u2 = torch.randn_like(d["u"])
    f = lambda u: ssc.fwd_ex(u, d["delta"], d["A"], B, C, None, None, d["delta_bias"], True, want_x=False)[0]
    lhs = f(0.7 * d["u"] - 1.3 * u2)
    rhs = 0.7 * f(d["u"]) - 1.3 * f(u2)
    assert_close(lhs, rhs, 1e-4, "linearity in u")


def test_scan_errors():
    """error behaviour mirrors the reference binding: RuntimeError on dtype / shape / device violations."""
    from segmamba_b200 import selective_scan_cuda as ssc
    d = rand_scan_inputs(1, 1, 8, 64, 16)
    B, C = d["B"].unsqueeze(1), d["C"].unsqueeze(1)
    with pytest.raises(RuntimeError):
        ssc.fwd(d["u"].cpu(), d["delta"], d["A"], B, C, None, None, None, False)
    with pytest.raises(RuntimeError):
        ssc.fwd(d["u"], d["delta"].half(), d["A"], B, C, None, None, None, False)
    with pytest.raises(RuntimeError):
        ssc.fwd(d["u"], d["delta"], d["A"][:, :5].contiguous(), B[:, :, :5].contiguous(), C[:, :, :5].contiguous(), None, None, None, False)
    with pytest.raises(RuntimeError):
        ssc.fwd(d["u"], d["delta"][:, :4], d["A"], B, C, None, None, None, False)
    res = ssc.fwd(d["u"], d["delta"], d["A"], B, C, d["D"], d["z"], d["delta_bias"], True)
    assert len(res) == 3 and res[1].shape == (1, 8, 1, 32)
