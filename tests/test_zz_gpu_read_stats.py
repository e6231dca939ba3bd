"""mdbg_read_stats / `--read-stats` on the GPU, in a process of its own: see tests/read_stats_gpu_check.py.
The entry point was written at the end of round 1 after the GPU budget was spent: its oracle side and
bindings are tested on the CPU, the kernels have not run on hardware yet, so a failure here is reported
(xfail) instead of hiding the results of the verified path."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.xfail(strict=False, reason="new entry point, first hardware run")
def test_read_stats_matches_oracle():
    r = subprocess.run([sys.executable, os.path.join(HERE, "read_stats_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    assert json.loads(r.stdout.strip().splitlines()[-1])["ok"]
