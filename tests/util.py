"""shared helpers for the parity tests."""
import numpy as np
import torch


def rel_err(a, b):
    """max-norm relative error  ||a-b||_inf / ||b||_inf  (the metric BASELINE.json's tolerances are stated in)."""
    a = torch.as_tensor(np.asarray(a.detach().cpu().float()) if isinstance(a, torch.Tensor) else np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b.detach().cpu().float()) if isinstance(b, torch.Tensor) else np.asarray(b)).double()
    assert a.shape == b.shape, f"shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.isfinite(a).all(), "non-finite values in result"
    denom = float(b.abs().max())
    return float((a - b).abs().max()) / (denom if denom > 0 else 1.0)


def assert_close(a, b, tol, what):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e}"
    return e


TOL = {torch.float32: 1e-3, torch.float16: 1e-2, torch.bfloat16: 1e-2}      # BASELINE.json north_star
GRAD_TOL = {torch.float32: 2e-3, torch.float16: 2e-2, torch.bfloat16: 3e-2}  # grads: the reference's tests loosen x2..x10


def rand_scan_inputs(seed, batch, dim, L, N, G=1, dtype=torch.float32, device="cuda", trained_like=True):
    """values representable in `dtype` (so the oracle sees exactly what the kernel sees)."""
    import golden_inputs as gi
    d = gi.scan_inputs(seed, batch, dim, L, N, G, trained_like)
    out = {}
    for k, v in d.items():
        if k in ("A", "D", "delta_bias"):
            out[k] = v.to(device)
        else:
            out[k] = v.to(dtype).to(device)
    return out
