"""The scripts under examples/ run to completion on the GPU box."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("script", ["quick_start.py", "custom_channel.py", "fresh_channel_per_frame.py", "noise_bump_experiment.py"])
def test_example_runs(script, tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script)], cwd=tmp_path, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "BER" in r.stdout or "dB" in r.stdout
