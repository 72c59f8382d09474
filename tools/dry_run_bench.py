"""Dry-run bench.py's NATIVE arm without a GPU: every "cuda" placement goes to the CPU, the native library to tools/simt_emu, CUDA
events / streams / pinned memory are stubbed.  It exercises the control flow of bench.py (argument handling, the step, the e2e
leg, the roofline / JSON assembly) before a GPU slot is spent on it; the numbers it prints are meaningless.

    python tools/dry_run_bench.py --patch 32 --batch 1 --steps 1 --warmup 1 --no-cpu-baseline [--bf16-params]

Development aid only (test infrastructure, like tests/emu.py); nothing in the product refers to it.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a, **k):
        pass

    def elapsed_time(self, other):
        return 1.0


def main():
    import emu
    import run_gpu_tests_on_emulator as redirect
    redirect._redirect_cuda_to_cpu()
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.Event = _Event
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    real_device = torch.device

    class _Dev:                                            # torch.device("cuda", i) -> cpu
        def __new__(cls, *a, **k):
            if a and "cuda" in str(a[0]):
                return real_device("cpu")
            return real_device(*a, **k)
    torch.device = _Dev
    # autocast("cuda") -> autocast("cpu"), including the queries the mixer op makes (selective_scan_interface.py)
    real_autocast, real_enabled, real_dtype = torch.autocast, torch.is_autocast_enabled, torch.get_autocast_dtype
    cpu = lambda d: "cpu" if "cuda" in str(d) else d

    class _Autocast(real_autocast):
        def __init__(self, device_type, *a, **k):
            super().__init__(cpu(device_type), *a, **k)
    torch.autocast = _Autocast
    torch.amp.autocast_mode.autocast = _Autocast              # what custom_fwd / custom_bwd instantiate
    torch.cuda.is_bf16_supported = lambda *a, **k: True
    torch.is_autocast_enabled = lambda device_type="cpu": real_enabled(cpu(device_type))
    torch.get_autocast_dtype = lambda device_type="cpu": real_dtype(cpu(device_type))
    import bench
    with emu.emulated():
        sys.argv = ["bench.py"] + sys.argv[1:]
        args = bench.parse_args()
        bench.main_native(args)


if __name__ == "__main__":
    main()
