"""A fixed-seed slice of the randomised differential test (tests/fuzz_parity.py): random link shapes through the fast
kernel, the CPU oracle and the general kernel must agree."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12])
def test_random_link_shapes_agree_across_kernels_and_oracle(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_parity.py"), "60", str(seed)], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 mismatches" in r.stdout
