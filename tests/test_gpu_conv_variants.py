"""Alternative causal-conv1d / inter-slice-permutation kernels (SMB_CONV_V2: 16 positions per thread for 16-bit types;
SMB_PERMUTE_V2: 4-byte accesses) against the CPU oracle (causal_conv1d_ref restated, oracle/segmamba_oracle.c) and against the
reference's own index arithmetic (mamba_simple.py:245-247,261), at the model's stage-0 shape and at ragged shapes."""
import pytest
import torch

import golden_inputs as gi
from util import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("v2", ["0", "1"], ids=["default", "conv_v2"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("direction", [0, 1], ids=["fwd", "rev"])
@pytest.mark.parametrize("shape", [(2, 96, 262144), (3, 37, 1001)], ids=["stage0", "ragged"])
def test_conv_variant_vs_oracle(monkeypatch, v2, dtype, direction, shape):
    from oracle import oracle as orc
    from segmamba_b200 import causal_conv1d_cuda as cc
    batch, dim, L = shape
    d = gi.conv_inputs(23, batch, dim, L, 4)
    xq, doq = d["x"].to(dtype), d["dout"].to(dtype)
    x = torch.empty(dim, batch, L, dtype=dtype, device="cuda").permute(1, 0, 2)      # channel-major view, as in the mixer
    x.copy_(xq.cuda())
    dout = doq.cuda()
    w, b = d["weight"].cuda(), d["bias"].cuda()
    monkeypatch.setenv("SMB_CONV_V2", v2)
    out = cc.causal_conv1d_fwd_ex(x, w, b, True, direction=direction)
    dx, dw, db = cc.causal_conv1d_bwd_ex(x, w, b, dout, None, True, direction=direction)
    torch.cuda.synchronize()
    f = (lambda t: t.flip(-1)) if direction else (lambda t: t)
    ro = orc.causal_conv1d_fwd_raw(f(xq.float()), d["weight"], d["bias"], True)
    rdx, rdw, rdb = orc.causal_conv1d_bwd_raw(f(xq.float()), d["weight"], d["bias"], f(doq.float()), True)
    assert_close(out, f(ro), 1e-2, "out")
    assert_close(dx, f(rdx), 2e-2, "dx")
    assert_close(dw, rdw, 2e-2, "dweight")
    assert_close(db, rdb, 2e-2, "dbias")


@pytest.mark.parametrize("v2", ["0", "1"], ids=["default", "permute_v2"])
@pytest.mark.parametrize("rows,L,ns", [(384, 262144, 64), (768, 32768, 32), (3072, 512, 8), (10, 1000, 10)], ids=lambda v: str(v))
def test_seq_permute_variant_vs_reference_indexing(monkeypatch, v2, rows, L, ns):
    from segmamba_b200 import causal_conv1d_cuda as cc
    torch.manual_seed(rows)
    x = torch.randn(rows // 2, 2, L, device="cuda").bfloat16().permute(1, 0, 2)
    monkeypatch.setenv("SMB_PERMUTE_V2", v2)
    f = cc.seq_permute(x, ns)
    back = cc.seq_permute(f, ns, inverse=True)
    torch.cuda.synchronize()
    ref = torch.stack(x.chunk(ns, dim=-1), dim=-1).flatten(-2)                        # mamba_simple.py:245-247
    assert torch.equal(f, ref)
    ref_back = ref.reshape(2, rows // 2, L // ns, ns).permute(0, 1, 3, 2).flatten(-2)   # mamba_simple.py:261
    assert torch.equal(back, ref_back) and torch.equal(back, x)
