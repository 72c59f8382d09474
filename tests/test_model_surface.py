"""CPU checks of the nn.Module surface: state_dict keys / shapes / order identical to the reference's (fixture written by
oracle/gen_golden.py from the reference class), constructor defaults, and the sliding-window host logic."""
import torch

import golden_inputs as gi


def test_state_dict_surface_default_model():
    from segmamba_b200.segmamba import SegMamba
    g = gi.load("state_dict_surface_default")
    m = SegMamba(in_chans=4, out_chans=4, depths=[2, 2, 2, 2], feat_size=[48, 96, 192, 384])
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert [",".join(map(str, v.shape)) for v in sd.values()] == [str(s) for s in g["shapes"]]
    assert sum(v.numel() for v in sd.values()) == int(g["n_params"][0]) == 67416196


def test_mamba_init_matches_reference_rules():
    from segmamba_b200.mamba_simple import Mamba
    torch.manual_seed(0)
    m = Mamba(d_model=48, nslices=64)
    assert m.d_inner == 96 and m.dt_rank == 3
    A = torch.exp(m.A_log)
    assert torch.allclose(A, torch.arange(1, 17, dtype=torch.float32).repeat(96, 1))
    assert torch.equal(m.D, torch.ones(96)) and torch.equal(m.D_b, torch.ones(96)) and torch.equal(m.D_s, torch.ones(96))
    dt = torch.nn.functional.softplus(m.dt_proj.bias)
    assert float(dt.min()) >= 1e-4 * 0.99 and float(dt.max()) <= 0.1 * 1.01
    assert m.conv1d.weight.shape == (96, 1, 4) and m.x_proj.weight.shape == (35, 96) and m.out_proj.weight.shape == (48, 96)
    assert getattr(m.A_log, "_no_weight_decay") and getattr(m.D_s, "_no_weight_decay")
