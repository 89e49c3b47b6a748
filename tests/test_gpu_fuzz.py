"""GPU: a short run of tools/fuzz_parity.py (random shapes / pixel types / option combinations against the oracle)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("args", [("15", "101"), ("10", "102", "big")], ids=["small", "long-rows"])
def test_randomised_parity(args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), *args], capture_output=True, text=True)
    assert out.returncode == 0 and "fuzz ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
