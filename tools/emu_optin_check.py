"""Integrated check of the opt-in kernels without a GPU: the default SegMamba (depths [2,2,2,2], dims [48,96,192,384]) on one
32^3 patch under bf16 autocast, forward + backward on the CPU SIMT emulator, once with the default kernels and once per opt-in
switch; prints the loss and the relative L2 distance of the gradients of all mixer parameters to the default run.

    python tools/emu_optin_check.py            (about 15 minutes on 8 cores; test infrastructure, like tests/emu.py)

Dense-convolution weights are left out of the comparison: PyTorch's CPU bf16 convolution backward is not reliable at the 2^3-voxel
resolution the deepest stage has on a 32^3 patch (it returned 1e33 for finite inputs in some runs), which has nothing to do with
the kernels under test.  Switches that change a bf16 rounding of the forward (the fused LayerNorm: a few output roundings differ
from ATen's) are not comparable this way on a 32^3 patch -- instance norms over 2^3 .. 8^3 voxels amplify a 2e-6 perturbation of
the first mixer input to 2 % at the last one -- and are checked per layer instead (tests/test_emu_kernels.py).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import emu
    import torch.amp.autocast_mode as am
    from segmamba_b200 import segmamba as sm
    # CPU autocast stands in for CUDA autocast (what tools/dry_run_bench.py does)
    real_enabled, real_dtype, real_ac = torch.is_autocast_enabled, torch.get_autocast_dtype, torch.autocast
    torch.is_autocast_enabled = lambda d="cpu": real_enabled("cpu")
    torch.get_autocast_dtype = lambda d="cpu": real_dtype("cpu")

    class AC(real_ac):
        def __init__(self, device_type, *a, **k):
            super().__init__("cpu", *a, **k)
    am.autocast = AC
    torch.cuda.is_bf16_supported = lambda *a, **k: True

    torch.manual_seed(0)
    m = sm.SegMamba(in_chans=4, out_chans=4, depths=[2, 2, 2, 2], feat_size=[48, 96, 192, 384]).train()
    x = torch.rand(1, 4, 32, 32, 32)
    y = torch.randint(0, 4, (1, 32, 32, 32))
    sel = [p for n, p in m.named_parameters() if "mamba" in n]

    def run():
        for p in m.parameters():
            p.grad = None
        with torch.autocast("cpu", dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(m(x).float(), y)
        loss.backward()
        return float(loss.detach()), [p.grad.clone() for p in sel]

    def dist(a, b):
        num = sum(float(((u.double() - v.double()) ** 2).sum()) for u, v in zip(a[1], b[1]))
        den = sum(float((v.double() ** 2).sum()) for v in b[1])
        return (num / den) ** 0.5

    with emu.emulated():
        base = run()
        print("default            loss %.6f" % base[0], flush=True)
        print("default, repeated  loss %.6f  distance %.3e   (noise floor)" % ((lambda r: (r[0], dist(r, base)))(run())), flush=True)
        for k, v in (("SMB_R3_V2", "1"), ("SMB_FWD_V2", "1"), ("SMB_FWD_V2", "2"), ("SMB_RAGG_V2", "1"), ("SMB_CONV_V2", "1"),
                     ("SMB_PERMUTE_V2", "1"), ("SMB_SEG_MIN", "64"), ("SMB_ALIGN_GEMMS", None), ("SMB_RECOMPUTE", None)):
            if v is None:                                       # module-level switches of the mixer op (default: aligned, kept)
                from segmamba_b200 import selective_scan_interface as ssi
                attr = "ALIGN_GEMMS" if k == "SMB_ALIGN_GEMMS" else "KEEP_CONV_DELTA"
                setattr(ssi, attr, False)
                r = run()
                setattr(ssi, attr, True)
                label = k + ("=0" if k == "SMB_ALIGN_GEMMS" else "=1")
            else:
                os.environ[k] = v
                r = run()
                del os.environ[k]
                label = f"{k}={v}"
            print("%-18s loss %.6f  distance %.3e" % (label, r[0], dist(r, base)), flush=True)
        sm.PAD_CIN = True
        r = run()
        sm.PAD_CIN = False
        print("%-18s loss %.6f  distance %.3e" % ("SMB_PAD_CIN=1", r[0], dist(r, base)), flush=True)


if __name__ == "__main__":
    main()
