"""The reference's OWN example scripts, unmodified, driven against this package (north_star: "the existing examples
and tests drive it unchanged").  The scripts are the byte-for-byte copies oracle/build_ref.py leaves in
oracle/_ref/examples (git-ignored, made in the build container where /root/reference is mounted; they travel to the GPU
box with the snapshot).  Each one runs in a scratch working directory with the shipped ``config/`` linked in - the
scripts use CWD-relative paths and write results/, images/ and docs/figures/ - with only this package on PYTHONPATH.
matplotlib is not part of this image, so the scripts that import it (overview.py, the *_demo plots) are not in the list."""
import os
import subprocess
import sys

import pytest

from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu

EXAMPLES = os.path.join(ROOT, "oracle", "_ref", "examples")
# script, something it must print, files it must leave behind (relative to its working directory)
CASES = [
    ("test_new_naming.py", "Test completed successfully", ["results/ber_results.csv", "images/severe_multipath"]),
    ("test_new_naming2.py", "Generated Files in images/severe_multipath", ["results/ber_results.csv", "images/severe_multipath"]),
    ("configurable_simulation_demo.py", "", []),
    # its second half loads simulation_settings_custom_channel.json (num_symbols = 100000, not a multiple of 64): the live
    # reference raises this ValueError from to_parallel and the script prints it with a traceback - same behaviour here
    ("custom_channel_demo.py", "Error: Length of data must be divisible by number of streams.", []),
    ("quick_start_adaptive.py", "", []),
]


@pytest.mark.parametrize("script,expect,files", CASES, ids=[c[0] for c in CASES])
def test_reference_example_runs_unchanged(script, expect, files, tmp_path):
    path = os.path.join(EXAMPLES, script)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/examples is absent (oracle/build_ref.py runs where /root/reference is mounted)")
    os.symlink(os.path.join(ROOT, "config"), tmp_path / "config")
    env = dict(os.environ, PYTHONPATH=PKG)
    env.pop("OFDM_B200_FORCE_GENERAL", None)
    run = subprocess.run([sys.executable, path], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    tail = (run.stdout[-3000:] + "\n" + run.stderr[-3000:])
    assert run.returncode == 0, tail
    if not expect.startswith("Error:"):
        assert "Traceback" not in run.stderr, tail
    assert expect in run.stdout, tail
    for f in files:
        assert (tmp_path / f).exists(), (f, tail)
