"""Worker of tests/test_emu_dropin_reference.py (own process: it rebinds sys.modules entries).

Level-1 drop-in check: the REFERENCE's own, unmodified Python -- mamba_ssm/ops/selective_scan_interface.py
(SelectiveScanFn, MambaInnerFnNoOutProj), mamba_ssm/modules/mamba_simple.py (Mamba v3) and model_segmamba/segmamba.py (SegMamba) --
imported from /root/reference after `segmamba_b200.install_dropin()`, so that its `import selective_scan_cuda` /
`import causal_conv1d_cuda` resolve to the native C-ABI shims.  The kernels run on the CPU SIMT emulator (tests/emu.py); outputs
and gradients are compared with the golden vectors that the same reference code produced with its own *_ref kernels.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import emu  # noqa: E402
import golden_inputs as gi  # noqa: E402
from util import assert_close  # noqa: E402


def load_reference_on_dropin():
    import segmamba_b200
    segmamba_b200.install_dropin(force=True)
    sys.path.insert(0, os.path.join(REF, "causal-conv1d"))
    gen = types.ModuleType("mamba_ssm.utils.generation")          # the package __init__ pulls an incompatible `transformers`
    gen.GenerationMixin = type("GenerationMixin", (), {})
    hf = types.ModuleType("mamba_ssm.utils.hf")
    hf.load_config_hf = hf.load_state_dict_hf = lambda *a, **k: None
    sys.path.insert(0, os.path.join(REF, "mamba"))
    sys.path.insert(0, REF)
    pkg = types.ModuleType("mamba_ssm")
    pkg.__path__ = [os.path.join(REF, "mamba", "mamba_ssm")]
    sys.modules["mamba_ssm"] = pkg
    sys.modules["mamba_ssm.utils.generation"] = gen
    sys.modules["mamba_ssm.utils.hf"] = hf
    import mamba_ssm.modules.mamba_simple as ms                   # reference module, unmodified
    import mamba_ssm.ops.selective_scan_interface as ssi          # reference module, unmodified
    import selective_scan_cuda
    assert selective_scan_cuda.__name__.startswith("segmamba_b200"), "the reference did not pick up the drop-in"
    assert ssi.selective_scan_cuda is selective_scan_cuda
    sys.modules["mamba_ssm"].Mamba = ms.Mamba
    spec = importlib.util.spec_from_file_location("ref_segmamba", os.path.join(REF, "model_segmamba", "segmamba.py"))
    seg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(seg)
    return ssi, ms, seg


def main():
    with emu.emulated():
        ssi, ms, seg = load_reference_on_dropin()
        # 1. reference selective_scan_fn (BASELINE.json configs[0]) on the native kernels
        name, seed, batch, dim, L, N, G, tl, has_D, has_z, has_b, sp = gi.CONFIG1
        d = gi.scan_inputs(seed, batch, dim, L, N, G, tl)
        out, last = ssi.selective_scan_fn(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["z"], d["delta_bias"], True, True)
        gold = gi.load("scan_" + name)
        assert_close(out, gold["out"], 1e-3, "reference selective_scan_fn on the drop-in: out")
        assert_close(last, gold["last_state"], 1e-3, "last_state")
        # 2. reference Mamba (v3) forward + backward
        for name, seed, batch, d_model, L, ns in gi.MAMBA_CASES:
            gold = gi.load("mamba_" + name)
            m = ms.Mamba(d_model=d_model, d_state=16, d_conv=4, expand=2, bimamba_type="v3", nslices=ns)
            m.load_state_dict({k[len("param."):]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("param.")}, strict=True)
            r = np.random.RandomState(seed + 1000)
            x = torch.from_numpy(r.standard_normal((batch, L, d_model)).astype(np.float32)).requires_grad_()
            dout = torch.from_numpy(r.standard_normal((batch, L, d_model)).astype(np.float32))
            out = m(x)
            assert_close(out, gold["out"], 1e-3, f"reference Mamba on the drop-in ({name}): out")
            names = [n for n, _ in m.named_parameters()]
            grads = torch.autograd.grad(out, [x] + [p for _, p in m.named_parameters()], dout)
            assert_close(grads[0], gold["dx"], 2e-3, "dx")
            for n, g in zip(names, grads[1:]):
                assert_close(g, gold["grad." + n], 3e-3, "grad." + n)
        # 3. reference SegMamba forward
        c = gi.MODEL_CASE
        gold = gi.load("model_" + c["name"])
        torch.manual_seed(0)
        model = seg.SegMamba(in_chans=c["in_chans"], out_chans=c["out_chans"], depths=c["depths"], feat_size=c["feat_size"],
                             hidden_size=c["hidden_size"])
        model.load_state_dict(gi.randomize_state_dict(model.state_dict(), c["seed"]))
        model.train()
        x = gi.model_input(c["seed"] + 1, (c["batch"], c["in_chans"], c["spatial"], c["spatial"], c["spatial"]))
        out = model(x)
        assert_close(out, gold["out"], 1e-3, "reference SegMamba on the drop-in: logits")
    print("DROPIN_OK")


if __name__ == "__main__":
    main()
