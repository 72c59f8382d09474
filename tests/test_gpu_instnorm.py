"""GPU parity of the fused InstanceNorm(+add)(+activation) kernels against plain PyTorch fp32 (F.instance_norm is the
third-party arithmetic the reference calls here; parity for it is pinned on torch itself)."""
import pytest
import torch
import torch.nn.functional as F

from util import assert_close

pytestmark = pytest.mark.gpu


def _ref(x, add, add_norm, act, slope):
    v = F.instance_norm(x.float(), eps=1e-5)
    if add is not None:
        v = v + (F.instance_norm(add.float(), eps=1e-5) if add_norm else add.float())
    if act == "relu":
        v = F.relu(v)
    elif act == "leaky_relu":
        v = F.leaky_relu(v, slope)
    return v


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["f32", "bf16", "f16"])
@pytest.mark.parametrize("shape", [(2, 48, 16, 16, 16), (1, 96, 9, 7, 5), (2, 768, 4, 4, 4), (1, 32, 40, 33, 21), (3, 8, 2, 2, 2)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("mode", ["plain", "add", "addnorm"])
@pytest.mark.parametrize("act", [None, "relu", "leaky_relu"])
def test_fused_instance_norm(dtype, shape, mode, act):
    from segmamba_b200.instance_norm import fused_instance_norm
    torch.manual_seed(sum(shape))
    x = (torch.randn(shape, device="cuda") * 2.0 + 0.7).to(dtype)          # non-zero mean: exercises the centred variance
    x = x.contiguous(memory_format=torch.channels_last_3d).requires_grad_()
    add = None
    if mode != "plain":
        add = (torch.randn(shape, device="cuda") * 0.5 - 0.3).to(dtype).requires_grad_()
    dy = torch.randn(shape, device="cuda").to(dtype)
    y = fused_instance_norm(x, act, 0.01, add=add, add_norm=(mode == "addnorm"))
    assert y.is_contiguous(memory_format=torch.channels_last_3d) and y.dtype == dtype
    gx = torch.autograd.grad(y, [x] + ([add] if add is not None else []), dy)
    xr = x.detach().clone().requires_grad_()
    ar = add.detach().clone().requires_grad_() if add is not None else None
    yr = _ref(xr, ar, mode == "addnorm", act, 0.01)
    gr = torch.autograd.grad(yr, [xr] + ([ar] if ar is not None else []), dy.float())
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    assert_close(y, yr, tol, "y")
    # activation kinks: a bf16-rounded pre-activation can land on the other side of 0; compare gradients where it cannot
    gtol = 2e-4 if dtype == torch.float32 else 3e-2
    for a, b, n in zip(gx, gr, ("dx", "dadd")):
        assert_close(a, b, gtol, n)


def test_fused_instance_norm_large_mean():
    """|mean| >> std: the centred per-CTA sums + Chan merge keep the variance accurate (no E[x^2]-E[x]^2 cancellation)."""
    from segmamba_b200.instance_norm import fused_instance_norm
    x = (torch.randn(1, 16, 32, 32, 32, device="cuda") * 0.01 + 50.0).contiguous(memory_format=torch.channels_last_3d)
    y = fused_instance_norm(x)
    yr = F.instance_norm(x.double(), eps=1e-5).float()
    assert_close(y, yr, 2e-3, "y (mean 50, std 0.01)")


def test_fused_instance_norm_nchw_input_and_errors():
    from segmamba_b200.instance_norm import fused_instance_norm
    x = torch.randn(2, 16, 4, 6, 8, device="cuda")                          # NCDHW input is converted, not rejected
    y = fused_instance_norm(x, "relu")
    assert_close(y, F.relu(F.instance_norm(x)), 1e-4, "nchw input")
    with pytest.raises(RuntimeError):
        fused_instance_norm(torch.randn(2, 6, 4, 4, 4, device="cuda"))     # channels not a multiple of 4 (fp32)
    with pytest.raises(RuntimeError):
        fused_instance_norm(torch.randn(2, 8, 4, 4, 4))                    # CPU tensor: no fallback
