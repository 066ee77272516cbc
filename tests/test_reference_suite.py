"""API-parity gate: the reference's OWN 334 tests, unmodified and in place (never copied), are run
against this repo's ``ofdm_based_systems`` package.  Only possible where /root/reference is mounted
(the build container); skipped on the GPU box.

Expected non-passes:
* 4 tests that are stale against the reference's own code (they expect ``transmit`` to renormalise power
  and the water level to be flat despite the extra 1/N), and
* on a machine without a GPU, the 3 tests that call ``Simulation.run()`` - the hot path is CUDA-only.
  (tests/test_simulation_gpu.py runs the same assertions on the B200.)
"""
import os
import re
import subprocess
import sys

import pytest

from conftest import PKG

REF_TESTS = "/root/reference/tests"
STALE = {
    "test_waterfilling_water_level_property",
    "test_transmit_power_normalization",
    "test_transmit_with_zero_signal",
    "test_power_normalization_across_multiple_transmissions",
}
NEEDS_GPU = {"test_simulation_run_basic", "test_simulation_run_with_different_configurations",
             "test_simulation_reproducibility"}


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference checkout not mounted")
def test_reference_test_suite_passes_against_this_package(tmp_path):
    env = dict(os.environ, PYTHONPATH=PKG)
    r = subprocess.run([sys.executable, "-m", "pytest", REF_TESTS, "-o", "addopts=", "-p", "no:cacheprovider", "-q"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=900)
    tail = r.stdout[-3000:]
    failed = set(re.findall(r"FAILED \S+::(\w+)", r.stdout))
    m = re.search(r"(\d+) passed", r.stdout)
    assert m, tail
    passed = int(m.group(1))
    from ofdm_based_systems import _native
    allowed = set(STALE) | (NEEDS_GPU if _native.device_count() <= 0 else set())
    assert failed <= allowed, f"unexpected failures: {sorted(failed - allowed)}\n{tail}"
    assert passed >= 334 - len(allowed), tail
