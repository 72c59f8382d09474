"""Every scan kernel a run-time switch can select (SMB_FWD_V2 = 2: TMA-staged forward [default], 1: cp.async-pipelined, 0:
single-buffered; SMB_RAGG_V2: pipelined reverse aggregate; SMB_R3_V2: second-generation main backward pass; SMB_SEG_MIN: shortest
segment) against the CPU ORACLE --
not against the default kernels -- at a ragged small size (partial tiles, partial channel octets, scalar tails) and at the
BASELINE stage-0 size (batch 1, D=96, L=262144, N=16: every segment / carry / vector-reduction path at full length).
A switch whose kernel fails here is removed from the tree, not masked."""
import pytest
import torch

from test_gpu_scan import _compare_grads, _oracle_fwd_bwd
from util import GRAD_TOL, TOL, assert_close, rand_scan_inputs

pytestmark = pytest.mark.gpu

# "default" = what ships (TMA-staged forward, pipelined R1, second-generation R3 for 16-bit activations; the single-buffered
# kernels for fp32); the other entries select the A/B alternatives one at a time, "legacy" all of them
VARIANTS = {
    "default": {},
    "fwd_cp_async": {"SMB_FWD_V2": "1"},
    "fwd_single_buffered": {"SMB_FWD_V2": "0"},
    "ragg_v1": {"SMB_RAGG_V2": "0"},

    "seg256": {"SMB_SEG_MIN": "256"},
    "r3_opc1": {"SMB_R3_OPC": "1"},                   # one CTA per channel octet (no octet walk)
    "r3_opc3": {"SMB_R3_OPC": "3"},                   # three octets per CTA whatever the problem size
    "no_dense": {"DENSE": "0"},                      # 256-position checkpoints only: main backward pass with warp scans
    "no_dense_r3_v1": {"DENSE": "0", "SMB_R3_V2": "0"},
    "legacy": {"SMB_FWD_V2": "0", "SMB_RAGG_V2": "0", "SMB_R3_V2": "0", "SMB_SEG_MIN": "256", "DENSE": "0"},
}
ALL_SWITCHES = ("SMB_FWD_V2", "SMB_RAGG_V2", "SMB_R3_V2", "SMB_SEG_MIN", "SMB_R3_OPC")
_oracle_cache = {}


def _oracle_cached(key, d, flip):
    if key not in _oracle_cache:
        _oracle_cache.clear()                     # one full-size oracle result (a few hundred MB) alive at a time
        _oracle_cache[key] = _oracle_fwd_bwd(d, flip=flip)
    return _oracle_cache[key]


def _run(d, direction, monkeypatch, env):
    from segmamba_b200 import selective_scan_cuda as ssc
    for k in ALL_SWITCHES:
        if k in env:
            monkeypatch.setenv(k, env[k])
        else:
            monkeypatch.delenv(k, raising=False)
    B, C = d["B"].unsqueeze(1), d["C"].unsqueeze(1)
    monkeypatch.setattr(ssc, "DENSE_STATES", env.get("DENSE", "1") == "1")
    out, x, out_z, hst, hd = ssc.fwd_ex(d["u"], d["delta"], d["A"], B, C, d["D"], d["z"], d["delta_bias"], True, direction=direction,
                                        want_out=True, want_x=True, want_hstates=True, want_hdense=True)
    g = ssc.bwd_ex(d["u"], d["delta"], d["A"], B, C, d["D"], d["z"], d["delta_bias"], d["dout"], None, True, True,
                   direction=direction, hstates=hst, hdense=hd, low_memory=True)
    torch.cuda.synchronize()
    return out, x, out_z, g


@pytest.mark.parametrize("variant", list(VARIANTS), ids=list(VARIANTS))
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32], ids=["bf16", "f16", "f32"])
@pytest.mark.parametrize("direction", [0, 1], ids=["fwd", "rev"])
@pytest.mark.parametrize("shape", [(2, 44, 5003), (1, 96, 262144)], ids=["ragged", "stage0"])
def test_scan_variant_vs_oracle(monkeypatch, variant, dtype, direction, shape):
    env = VARIANTS[variant]
    if dtype == torch.float32 and variant in ("fwd_cp_async", "fwd_single_buffered", "ragg_v1"):
        pytest.skip("switch only affects 16-bit activations")
    if shape[2] > 100000 and dtype == torch.float16:
        pytest.skip("full size: bf16 and fp32 only")
    batch, dim, L = shape
    d = rand_scan_inputs(31 + L, batch, dim, L, 16, 1, dtype)
    out, x, out_z, g = _run(d, direction, monkeypatch, env)
    y, oz, last, xc, go = _oracle_cached((shape, dtype, direction), d, bool(direction))
    tol = TOL[dtype]
    assert_close(out, y, tol, "out")
    assert_close(out_z, oz, tol, "out_z")
    assert_close(g[8], oz, tol, "recomputed out_z")
    assert_close(x[..., 1::2], xc[..., 1::2], tol, "x (chunk states)")
    _compare_grads(g, go, dtype, True, variant + ":")
