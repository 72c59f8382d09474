"""Fused LayerNorm (smb_layernorm_fwd / _bwd; MambaLayer.norm, segmamba.py:53,70) against torch's fp32 LayerNorm at the model's
(tokens, C) shapes and ragged ones, all three I/O dtypes, and inside the model."""
import pytest
import torch
import torch.nn.functional as F

from util import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["f32", "bf16", "f16"])
@pytest.mark.parametrize("rows,C", [(262144, 48), (32768, 96), (4096, 192), (512, 384), (1000, 32), (77, 768)], ids=lambda v: str(v))
def test_fused_layer_norm(dtype, rows, C):
    from segmamba_b200.layer_norm import fused_layer_norm, supported
    torch.manual_seed(rows + C)
    x = (torch.randn(2, rows, C, device="cuda") * 1.7 + 0.4).to(dtype).requires_grad_()
    if not supported(x, C):
        pytest.skip("more than 128 sixteen-byte vectors per row for this dtype")
    w = (torch.rand(C, device="cuda") + 0.5).requires_grad_()
    b = (torch.randn(C, device="cuda") * 0.3).requires_grad_()
    dy = torch.randn(2, rows, C, device="cuda").to(dtype)
    y = fused_layer_norm(x, w, b, 1e-5)
    gx, gw, gb = torch.autograd.grad(y, [x, w, b], dy)
    xr = x.detach().float().requires_grad_()
    wr, br = w.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
    yr = F.layer_norm(xr, (C,), wr, br, 1e-5)
    rx, rw, rb = torch.autograd.grad(yr, [xr, wr, br], dy.float())
    lo = dtype == torch.float32
    assert_close(y, yr, 1e-5 if lo else 1e-2, "y")
    assert_close(gx, rx, 1e-4 if lo else 2e-2, "dx")
    assert_close(gw, rw, 2e-4 if lo else 2e-2, "dweight")
    assert_close(gb, rb, 2e-4 if lo else 2e-2, "dbias")


def test_segmamba_step_with_fused_layer_norm(monkeypatch):
    """the model with the fused kernel against the same model with nn.LayerNorm (ATen), fp32 (TF32 off) and bf16 autocast."""
    import golden_inputs as gi
    from segmamba_b200 import layer_norm
    from segmamba_b200.segmamba import SegMamba
    c = gi.MODEL_CASE
    torch.manual_seed(3)
    m = SegMamba(in_chans=4, out_chans=4, depths=c["depths"], feat_size=c["feat_size"], hidden_size=c["hidden_size"]).cuda().train()
    x = torch.rand(2, 4, 32, 32, 32, device="cuda")
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        outs, outs32 = [], []
        for on in (False, True):
            monkeypatch.setattr(layer_norm, "ENABLED", on)
            outs32.append(m(x))
            with torch.autocast("cuda", dtype=torch.bfloat16):
                outs.append(m(x).float())
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    assert_close(outs32[1], outs32[0], 1e-4, "fp32 logits, fused vs nn.LayerNorm")
    assert_close(outs[1], outs[0], 2e-2, "bf16 logits, fused vs nn.LayerNorm")
