"""Builds libofdm_b200.so IN-TREE with plain nvcc for sm_100a (no torch, no JIT cache).

    python ofdm-based-systems_b200/build_native.py [--force] [--verbose]

One translation unit per supported FFT size (csrc/link_inst.cu with -DOFDM_INST_N=<N>) plus the C-ABI
unit (csrc/ofdm_b200.cu) and the batched water-filling unit, compiled in parallel and linked into
``ofdm-based-systems_b200/libofdm_b200.so``.  Objects are cached under ``build/`` keyed by a hash of
the sources and flags, so an unchanged tree re-links in seconds.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libofdm_b200.so")
SIZES = (8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192)
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-I", INCLUDE, "-diag-suppress", "177",
              # ~200 kernel instantiations with line tables: zstd-compressed fatbins keep the library near 10 MB
              "--compress-mode=size"]
if os.environ.get("OFDM_FAST_PHILOX_ROUNDS"):      # 10: the Random123 default instead of the crush-resistant minimum (7)
    NVCC_FLAGS.append("-DOFDM_FAST_PHILOX_ROUNDS=" + str(int(os.environ["OFDM_FAST_PHILOX_ROUNDS"])))


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libofdm_b200.so cannot be built (there is no CPU fallback)")


def _source_hash() -> str:
    h = hashlib.sha256()
    for root in (CSRC, INCLUDE):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def _units():
    units = [("ofdm_b200", os.path.join(CSRC, "ofdm_b200.cu"), [])]
    if os.path.exists(os.path.join(CSRC, "waterfill.cu")):
        units.append(("waterfill", os.path.join(CSRC, "waterfill.cu"), []))
    if os.path.exists(os.path.join(CSRC, "frames.cu")):
        units.append(("frames", os.path.join(CSRC, "frames.cu"), []))
    if os.path.exists(os.path.join(CSRC, "link_fast.cu")):
        units.append(("link_fast", os.path.join(CSRC, "link_fast.cu"), []))
    if os.path.exists(os.path.join(CSRC, "link_fast_inst.cu")):
        for e, t in ((8, 8), (8, 16), (16, 16), (16, 32), (32, 32), (32, 64), (32, 128), (32, 256)):
            units.append((f"link_fast_inst_{e}x{t}", os.path.join(CSRC, "link_fast_inst.cu"),
                          [f"-DOFDM_FAST_E={e}", f"-DOFDM_FAST_T={t}"]))
    for n in SIZES:
        units.append((f"link_inst_{n}", os.path.join(CSRC, "link_inst.cu"), [f"-DOFDM_INST_N={n}"]))
    return units


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    tag = _source_hash()
    objdir = os.path.join(BUILD, tag)
    stamp = os.path.join(objdir, "linked")
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        return LIB
    os.makedirs(objdir, exist_ok=True)

    def compile_one(unit):
        name, src, extra = unit
        obj = os.path.join(objdir, name + ".o")
        if os.path.exists(obj) and not force:
            return obj, ""
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-Xptxas", "-v", "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
        with open(os.path.join(objdir, name + ".ptxas.log"), "w") as f:
            f.write(r.stderr)
        return obj, r.stderr

    units = _units()
    with cf.ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, units))
    objs = [o for o, _ in results]
    if verbose:
        for (name, _, _), (_, log) in zip(units, results):
            for line in log.splitlines():
                if "registers" in line or "spill" in line:
                    print(f"[{name}] {line.strip()}")
    r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(tag)
    # keep only the current object cache
    for d in os.listdir(BUILD):
        if d != tag:
            shutil.rmtree(os.path.join(BUILD, d), ignore_errors=True)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(path)
