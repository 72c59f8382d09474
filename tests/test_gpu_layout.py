"""Channel concatenation of channels-last activations on the native strided-copy kernel (smb_copy2d) against torch.cat, forward and
backward, at the decoder shapes of the default model (unetr_block.py:81-86) and a ragged one."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("shape", [(2, 48, 48, 128), (2, 96, 96, 64), (1, 384, 384, 16), (3, 8, 40, 7)], ids=lambda s: "b%d_ca%d_cb%d_s%d" % s)
def test_cat_channels_vs_torch(dtype, shape):
    from segmamba_b200 import layout
    Bz, Ca, Cb, S = shape
    torch.manual_seed(S)
    mk = lambda c: torch.randn(Bz, c, S, S, max(S - 1, 1), device="cuda").to(dtype).contiguous(memory_format=torch.channels_last_3d).requires_grad_()
    a, b = mk(Ca), mk(Cb)
    assert layout.supported(a, b)
    out = layout.cat_channels(a, b)
    ref = torch.cat((a, b), dim=1)
    assert out.is_contiguous(memory_format=torch.channels_last_3d) and torch.equal(out, ref)
    g = torch.randn_like(ref)
    ga, gb = torch.autograd.grad(out, [a, b], g)
    ra, rb = torch.autograd.grad(ref, [a, b], g)
    assert torch.equal(ga, ra) and torch.equal(gb, rb)
    assert ga.is_contiguous(memory_format=torch.channels_last_3d)


def test_cat_channels_rejects_other_layouts():
    from segmamba_b200 import layout
    a = torch.randn(1, 8, 4, 4, 4, device="cuda")          # NCDHW storage
    with pytest.raises(RuntimeError):
        layout.cat_channels(a, a)
