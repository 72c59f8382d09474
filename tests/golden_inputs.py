"""Deterministic inputs + case tables for the golden fixtures (numpy legacy RandomState: stable across
numpy/torch versions).  Used by oracle/gen_golden.py (which runs the REFERENCE on them, in the build
container) and by the tests (which run the oracle / the CUDA path on the same inputs)."""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def scan_inputs(seed, batch, dim, L, N, G=1, trained_like=False):
    """distributions of mamba/tests/ops/test_selective_scan.py:58-88; trained_like follows
    mamba_simple.py:99-116 (A = -(1..N), dt bias = inv_softplus(loguniform[1e-3, 1e-1]))."""
    r = np.random.RandomState(seed)
    f = lambda *s: torch.from_numpy(r.standard_normal(s).astype(np.float32))
    uni = lambda *s: torch.from_numpy(r.random_sample(s).astype(np.float32))
    d = {}
    if trained_like:
        d["A"] = -torch.arange(1, N + 1, dtype=torch.float32).repeat(dim, 1).contiguous()
        dt = torch.exp(uni(dim) * (np.log(0.1) - np.log(0.001)) + np.log(0.001)).clamp(min=1e-4)
        d["delta_bias"] = dt + torch.log(-torch.expm1(-dt))
        d["delta"] = 0.5 * f(batch, dim, L)
    else:
        d["A"] = -0.5 * uni(dim, N)
        d["delta_bias"] = 0.5 * uni(dim)
        d["delta"] = 0.5 * uni(batch, dim, L)
    d["B"] = f(batch, G, N, L) if G > 1 else f(batch, N, L)
    d["C"] = f(batch, G, N, L) if G > 1 else f(batch, N, L)
    d["D"] = f(dim)
    d["z"] = f(batch, dim, L)
    d["u"] = f(batch, dim, L)
    d["dout"] = f(batch, dim, L)
    return d


SCAN_CASES = [
    # name, seed, batch, dim, L, N, G, trained_like, has_D, has_z, has_bias, softplus
    ("testdist_L256", 0, 2, 4, 256, 8, 1, False, True, True, True, True),
    ("testdist_L1000_g2", 1, 2, 4, 1000, 8, 2, False, True, True, True, True),      # ragged length, 2 groups
    ("testdist_noz_nobias", 2, 2, 4, 128, 8, 1, False, False, False, False, False),
    ("trained_L2304", 3, 1, 8, 2304, 16, 1, True, True, True, True, True),         # > one 2048 chunk
    ("trained_L33_d40", 4, 2, 40, 33, 16, 1, True, True, True, True, True),        # odd length, dim > 32
]
CONFIG1 = ("config1_L4096", 0, 1, 16, 4096, 16, 1, False, True, True, True, True)    # BASELINE.json configs[0]


def conv_inputs(seed, batch, dim, L, width):
    r = np.random.RandomState(seed)
    f = lambda *s: torch.from_numpy(r.standard_normal(s).astype(np.float32))
    return dict(x=f(batch, dim, L), weight=f(dim, width), bias=f(dim), dout=f(batch, dim, L))


CONV_CASES = [
    # name, seed, batch, dim, L, width, has_bias, silu   (cc1d/tests/test_causal_conv1d.py:14-27)
    ("w4_silu_L512", 10, 2, 8, 512, 4, True, True),
    ("w4_silu_L151", 11, 2, 8, 151, 4, True, True),
    ("w3_nosilu_L8", 12, 2, 8, 8, 3, False, False),
    ("w2_silu_nobias_L372", 13, 1, 40, 372, 2, False, True),
    ("w4_nosilu_L1134", 14, 2, 4, 1134, 4, True, False),
]


def inner_inputs(seed, batch, d_model, L, N=16):
    d_inner = 2 * d_model
    R = -(-d_model // 16)
    r = np.random.RandomState(seed)
    f = lambda *s: torch.from_numpy(r.standard_normal(s).astype(np.float32))
    uni = lambda *s: torch.from_numpy(r.random_sample(s).astype(np.float32))
    return dict(
        xz=f(batch, 2 * d_inner, L), conv1d_weight=0.5 * f(d_inner, 1, 4), conv1d_bias=0.1 * f(d_inner),
        x_proj_weight=f(R + 2 * N, d_inner) / np.sqrt(d_inner), delta_proj_weight=f(d_inner, R) / np.sqrt(R),
        A=-torch.arange(1, N + 1, dtype=torch.float32).repeat(d_inner, 1).contiguous(), D=torch.ones(d_inner) + 0.1 * f(d_inner),
        delta_bias=torch.log(torch.expm1(0.001 + 0.1 * uni(d_inner))), dout=f(batch, d_inner, L))


INNER_CASES = [("dm16_L320", 20, 2, 16, 320), ("dm24_L97", 21, 1, 24, 97)]


def model_input(seed, shape):
    r = np.random.RandomState(seed)
    return torch.from_numpy(r.random_sample(shape).astype(np.float32))


def randomize_state_dict(sd, seed):
    """deterministic, non-degenerate parameters (numpy RandomState): keeps the reference's init for the
    structured tensors (A_log, D, dt bias) but perturbs them so directional parameter mix-ups show."""
    r = np.random.RandomState(seed)
    out = {}
    for k, v in sd.items():
        noise = torch.from_numpy(r.standard_normal(tuple(v.shape)).astype(np.float32))
        if k.endswith("A_log") or k.endswith("A_b_log") or k.endswith("A_s_log"):
            out[k] = v + 0.05 * noise
        elif ".mamba.D" in k:
            out[k] = v + 0.1 * noise
        elif "dt_proj" in k and k.endswith("bias"):
            out[k] = torch.log(torch.expm1(torch.full_like(v, 0.01) + 0.05 * noise.abs()))
        elif k.endswith("norm.weight"):
            out[k] = 1.0 + 0.1 * noise
        elif v.dim() >= 2:
            fan_in = int(np.prod(v.shape[1:]))
            out[k] = noise / np.sqrt(fan_in)
        else:
            out[k] = 0.1 * noise
    return out


MAMBA_CASES = [("dm16_L512_ns8", 30, 2, 16, 512, 8)]


MODEL_CASE = dict(name="tiny32", seed=40, in_chans=4, out_chans=4, depths=[1, 1, 1, 1], feat_size=[48, 32, 32, 64],
                  hidden_size=64, spatial=32, batch=1)


def reference_like_init(keys, shapes):
    """The only entries of the reference's initial state_dict that randomize_state_dict() reads are the
    structured ones: A*_log = log(1..N) rows (mamba_simple.py:111-118) and D* = ones (:121)."""
    sd = {}
    for k, s in zip(keys, shapes):
        if k.endswith("A_log") or k.endswith("A_b_log") or k.endswith("A_s_log"):
            sd[k] = torch.log(torch.arange(1, s[1] + 1, dtype=torch.float32)).repeat(s[0], 1).contiguous()
        elif ".mamba.D" in k:
            sd[k] = torch.ones(s)
        else:
            sd[k] = torch.zeros(s)
    return sd


def train_toy_model():
    """stand-in network of the training-loop fixture (oracle/gen_golden.py: train_toy_model)."""
    import torch
    torch.manual_seed(7)
    return torch.nn.Sequential(torch.nn.Conv3d(2, 6, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv3d(6, 3, 1))
