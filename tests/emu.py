"""Run the native kernels WITHOUT a GPU: the same csrc/*.cu sources compiled for the host by tools/simt_emu (every CUDA thread
a fiber, barriers and shuffles with SIMT semantics) and driven through the unmodified Python shims.

Test infrastructure only: `emulated()` temporarily points segmamba_b200._lib at the emulated library and neutralises the
three CUDA-runtime touch points of the shims (device guard, stream handle, CUDA-tensor check).  Nothing in the product
refers to this module; outside the context manager the product still refuses CPU tensors.
"""
import contextlib
import ctypes
import importlib.util
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_emu_lib() -> str:
    spec = importlib.util.spec_from_file_location("simt_emu_build", os.path.join(ROOT, "tools", "simt_emu", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(asan=os.environ.get("SMB_EMU_ASAN") == "1")     # memcheck mode: see tools/simt_emu/build.py


_EMU = None


def emu_lib() -> ctypes.CDLL:
    global _EMU
    if _EMU is None:
        from segmamba_b200 import _lib
        l = ctypes.CDLL(build_emu_lib())
        for name, (res, args) in _lib.EXPORTS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _EMU = l
    return _EMU


class _NullDevice:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


@contextlib.contextmanager
def emulated():
    from segmamba_b200 import _lib
    saved = (_lib._lib, _lib.require_cuda, _lib.stream_ptr, _lib.call, torch.cuda.device, torch.cuda.synchronize)
    _lib._lib = emu_lib()
    _lib.require_cuda = lambda *tensors: None
    _lib.stream_ptr = lambda device: 0
    _lib.call = lambda op, meta, fn, device: _lib.check(fn())
    torch.cuda.device = _NullDevice
    torch.cuda.synchronize = lambda *a, **k: None
    try:
        yield
    finally:
        (_lib._lib, _lib.require_cuda, _lib.stream_ptr, _lib.call, torch.cuda.device, torch.cuda.synchronize) = saved
