"""A short run of the randomised parity sweep (tests/fuzz_parity.py): random layers, grids, methods and batch sizes."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [11, 12])
def test_randomised_parity_sweep(seed):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_parity.py"), "--cases", "60", "--seed", str(seed)],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "0 above their bound" in res.stdout
