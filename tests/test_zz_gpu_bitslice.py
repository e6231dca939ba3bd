"""The bit-sliced K-A variant on the GPU (ka_variant = 2), in a process of its own: see
tests/bitslice_gpu_check.py.  Named zz so that it runs after the parity tests of the default path."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_bitslice_variant_matches_oracle_and_classic():
    r = subprocess.run([sys.executable, os.path.join(HERE, "bitslice_gpu_check.py"), "--quick"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["ok"]
    assert out["large"]["dirty_tiles"] == 0
