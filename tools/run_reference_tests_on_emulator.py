"""Run the REFERENCE's own pytest files against the native kernels, without a GPU: install_dropin() puts the C-ABI shims under the
reference's import names, the kernels execute on the CPU SIMT emulator, and every "cuda" placement in the test code is
redirected to the CPU.  Build-container tool (needs /root/reference); nothing here is product code.

    python tools/run_reference_tests_on_emulator.py /root/reference/mamba/tests/ops/test_selective_scan.py -k test_selective_scan
    python tools/run_reference_tests_on_emulator.py /root/reference/causal-conv1d/tests/test_causal_conv1d.py -k "test_causal_conv1d and not update"
"""
import importlib.util
import os
import re
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import pytest
    import emu
    spec = importlib.util.spec_from_file_location("emu_redirect", os.path.join(ROOT, "tools", "run_gpu_tests_on_emulator.py"))
    red = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(red)
    import segmamba_b200
    segmamba_b200.install_dropin(force=True)
    sys.path.insert(0, os.path.join(REF, "causal-conv1d"))
    gen = types.ModuleType("mamba_ssm.utils.generation")
    gen.GenerationMixin = type("GenerationMixin", (), {})
    hf = types.ModuleType("mamba_ssm.utils.hf")
    hf.load_config_hf = hf.load_state_dict_hf = lambda *a, **k: None
    sys.path.insert(0, os.path.join(REF, "mamba"))
    pkg = types.ModuleType("mamba_ssm")
    pkg.__path__ = [os.path.join(REF, "mamba", "mamba_ssm")]
    sys.modules.update({"mamba_ssm": pkg, "mamba_ssm.utils.generation": gen, "mamba_ssm.utils.hf": hf})
    red._redirect_cuda_to_cpu()
    # Some of the reference's test files end with a stray module-level call of one test (runs at import).  Test files are
    # staged, minus such lines, in the git-ignored emulator build directory for the duration of the run -- never committed.
    stage = os.path.join(ROOT, "tools", "simt_emu", "_build", "ref_tests")
    os.makedirs(stage, exist_ok=True)
    args = []
    for a in sys.argv[1:]:
        if a.endswith(".py") and os.path.isfile(a):
            text = "".join(l for l in open(a) if not re.match(r"^test_\w+\(.*\)\s*$", l))
            dst = os.path.join(stage, os.path.basename(a))
            open(dst, "w").write(text)
            a = dst
        args.append(a)
    try:
        with emu.emulated():
            return pytest.main(["-q", "-p", "no:cacheprovider", "--rootdir", stage] + args)
    finally:
        for f in os.listdir(stage):
            os.remove(os.path.join(stage, f))


if __name__ == "__main__":
    sys.exit(main())
