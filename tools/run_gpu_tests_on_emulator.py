"""Dry-run `-m gpu` test files on the CPU SIMT emulator: every "cuda" placement is redirected to the CPU and the native
library to tools/simt_emu, so the TEST CODE itself (argument plumbing, tolerances, shapes) is exercised before a GPU slot is
spent on it.  Development aid only; the results say nothing about the hardware.

    python tools/run_gpu_tests_on_emulator.py tests/test_gpu_zz_layernorm.py [-k expr] [pytest args...]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _redirect_cuda_to_cpu():
    def strip(kwargs):
        dev = kwargs.get("device")
        if dev is not None and "cuda" in str(dev):
            kwargs["device"] = "cpu"
        return kwargs

    for name in ("randn", "rand", "randint", "zeros", "ones", "empty", "tensor", "arange", "full", "randn_like", "zeros_like", "empty_like"):
        orig = getattr(torch, name)
        setattr(torch, name, (lambda f: lambda *a, **k: f(*a, **strip(k)))(orig))
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    orig_to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple("cpu" if isinstance(x, (str, torch.device)) and "cuda" in str(x) else x for x in a)
        return orig_to(self, *a, **strip(k))
    torch.Tensor.to = to
    orig_mto = torch.nn.Module.to

    def mto(self, *a, **k):
        a = tuple("cpu" if isinstance(x, (str, torch.device)) and "cuda" in str(x) else x for x in a)
        return orig_mto(self, *a, **strip(k))
    torch.nn.Module.to = mto
    torch.cuda.is_available = lambda: True


def main():
    import pytest
    import emu
    _redirect_cuda_to_cpu()
    with emu.emulated():
        return pytest.main(["-q", "-m", "gpu", "-p", "no:cacheprovider"] + sys.argv[1:])


if __name__ == "__main__":
    sys.exit(main())
