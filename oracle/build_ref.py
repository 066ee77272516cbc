"""Recipe for oracle/_ref: the UNMODIFIED reference as the CPU baseline.  TEST / BENCH INFRASTRUCTURE ONLY.

    python oracle/build_ref.py            (also run by __graft_entry__.build())

The reference is pure Python, so "building" it is a file copy: when /root/reference is mounted (the build
container), its package sources and its example scripts are copied byte for byte into oracle/_ref/, which is
git-ignored (no reference source enters the history) but travels to the GPU box with the snapshot, like the
built libofdm_b200.so does.  On the GPU box /root/reference does not exist and this script is a no-op; whatever
the build container left in oracle/_ref/ is used as is.

Users: bench.py (`cpu_baseline` and `--impl reference`, via oracle/ref_pipeline.py) and
tests/test_reference_examples_gpu.py (the reference's own example scripts run against this package)."""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OFDM_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")


def build_ref(verbose: bool = False) -> str | None:
    """Returns the path of oracle/_ref (None when neither the reference nor an earlier copy is there)."""
    src_pkg = os.path.join(REF, "src", "ofdm_based_systems")
    if not os.path.isdir(src_pkg):
        return DEST if os.path.isdir(os.path.join(DEST, "src", "ofdm_based_systems")) else None
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", "*.old", "*.backup", "*.ipynb")
    shutil.copytree(src_pkg, os.path.join(DEST, "src", "ofdm_based_systems"), ignore=ignore)
    shutil.copytree(os.path.join(REF, "examples"), os.path.join(DEST, "examples"), ignore=ignore)
    with open(os.path.join(DEST, "README"), "w") as f:
        f.write("Byte-for-byte copy of /root/reference/src/ofdm_based_systems and /root/reference/examples made by\n"
                "oracle/build_ref.py in the build container.  Git-ignored; CPU baseline and example scripts only.\n")
    if verbose:
        print(f"copied the reference package and examples into {DEST}")
    return DEST


if __name__ == "__main__":
    print(build_ref(verbose=True))
    sys.exit(0)
