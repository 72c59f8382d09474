"""Level-1 drop-in: the reference's own unmodified Python on the native kernels (through install_dropin, the ctypes shims and
the C ABI), executed on the CPU SIMT emulator and compared with the reference's golden outputs.  Needs /root/reference, i.e.
runs in the build container only."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/mamba"), reason="the reference checkout is not present")
def test_reference_python_runs_on_the_dropin_modules():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_dropin_reference_worker.py")], capture_output=True, text=True,
                         timeout=1200, cwd=ROOT)
    assert out.returncode == 0 and "DROPIN_OK" in out.stdout, (out.stdout[-2000:] + "\n" + out.stderr[-4000:])
