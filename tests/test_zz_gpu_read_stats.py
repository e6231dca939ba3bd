"""mdbg_read_stats / `--read-stats` on the GPU, in a process of its own: see tests/read_stats_gpu_check.py
(passed on the B200 at the end of round 1; a failure here is a regression)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_read_stats_matches_oracle():
    r = subprocess.run([sys.executable, os.path.join(HERE, "read_stats_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    assert json.loads(r.stdout.strip().splitlines()[-1])["ok"]
