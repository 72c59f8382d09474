"""Sliding-window inference on the GPU with the native SegMamba as predictor, checked against the CPU oracle: the oracle's
window enumeration + gaussian map (pinned to vendored MONAI by tests/test_oracle_golden.py) blending the oracle's own
SegMamba forward (pinned to the reference model by the same file) window by window.  The product is never compared with
itself.  TF32 is pinned off, because an fp32 forward through cuDNN/cuBLAS TF32 kernels is accurate to ~1e-3 only and is not
batch-invariant (round 1's red test: batch-2 vs batch-1 windows differed by 2e-4 abs); the batch-1/batch-2 agreement is
asserted separately so that the cause stays on record."""
import itertools

import numpy as np
import pytest
import torch

import golden_inputs as gi
from util import assert_close, rel_err

pytestmark = pytest.mark.gpu

ROI = (32, 32, 32)
IMG = (40, 48, 33)


@pytest.fixture(autouse=True)
def _exact_fp32():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _model_and_sd():
    from segmamba_b200.segmamba import SegMamba
    c = gi.MODEL_CASE
    gold = gi.load("model_" + c["name"])
    keys = [str(k) for k in gold["state_dict_keys"]]
    shapes = [tuple(int(s) for s in str(x).split(",")) if str(x) else () for x in gold["state_dict_shapes"]]
    sd = gi.randomize_state_dict(gi.reference_like_init(keys, shapes), c["seed"])
    m = SegMamba(in_chans=4, out_chans=4, depths=c["depths"], feat_size=c["feat_size"], hidden_size=c["hidden_size"])
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd, c


def _oracle_sliding_window(sd, c, x_cpu, flips=((),)):
    """monai/inferers/utils.py:138-321 + prediction.py:128-155 restated with the oracle's pieces, fp32 CPU."""
    from oracle import oracle as orc
    starts = orc.sliding_window_starts(tuple(x_cpu.shape[2:]), ROI, 0.5)
    w = orc.gaussian_importance_map(ROI)[None, None]
    total = None
    for fl in flips:
        xin = torch.flip(x_cpu, [2 + a for a in fl]) if fl else x_cpu
        acc = torch.zeros(1, 4, *x_cpu.shape[2:])
        cnt = torch.zeros(1, 1, *x_cpu.shape[2:])
        for st in starts:
            sl = (slice(None), slice(None)) + tuple(slice(s, s + r) for s, r in zip(st, ROI))
            with torch.no_grad():
                acc[sl] += orc.segmamba_forward(sd, xin[sl].contiguous(), depths=c["depths"]) * w
            cnt[sl] += w
        res = acc / cnt
        if fl:
            res = torch.flip(res, [2 + a for a in fl])
        total = res if total is None else total + res
    return total / len(flips)


def test_sliding_window_with_segmamba_vs_oracle():
    from segmamba_b200 import sliding_window as sw
    m, sd, c = _model_and_sd()
    x = gi.model_input(60, (1, 4) + IMG)
    ref = _oracle_sliding_window(sd, c, x)
    with torch.no_grad():
        for bs in (1, 2, 3):
            out = sw.sliding_window_inference(x.cuda(), ROI, bs, m, overlap=0.5, mode="gaussian")
            assert out.shape == (1, 4) + IMG and out.device == next(m.parameters()).device
            assert_close(out, ref, 1e-3, f"blended logits, sw_batch_size={bs}")


def test_sliding_window_mirror_tta_vs_oracle():
    from segmamba_b200 import sliding_window as sw
    m, sd, c = _model_and_sd()
    x = gi.model_input(61, (1, 4) + IMG)
    flips = [f for k in range(4) for f in itertools.combinations((0, 1, 2), k)]
    ref = _oracle_sliding_window(sd, c, x, flips)
    with torch.no_grad():
        out = sw.sliding_window_inference(x.cuda(), ROI, 2, m, mirror_axes=(0, 1, 2))
        out5 = sw.sliding_window_inference(x.cuda(), ROI, 5, m, mirror_axes=(0, 1, 2))
    assert_close(out, ref, 1e-3, "TTA logits")
    assert_close(out5, ref, 1e-3, "TTA logits, sw_batch_size=5")


def test_forward_batch_invariance_fp32():
    """the fp32 forward of a window must not depend on what it is batched with (TF32 off: only reduction-order noise of
    the per-(batch, channel) statistics kernels may differ, their CTA plan depends on the batch size)."""
    m, sd, c = _model_and_sd()
    x = gi.model_input(62, (2, 4) + ROI).cuda()
    with torch.no_grad():
        both = m(x)
        one = torch.cat([m(x[:1]), m(x[1:])])
    e = rel_err(both, one)
    assert e <= 1e-4, f"batch-2 vs batch-1 forward differ by {e:.2e} with TF32 off"
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
    try:
        with torch.no_grad():
            e_tf32 = rel_err(m(x), one)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    print(f"batch invariance: fp32 {e:.2e}, TF32 library kernels {e_tf32:.2e}")
    assert e_tf32 <= 2e-2, f"TF32 forward off by {e_tf32:.2e}"
