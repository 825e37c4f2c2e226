"""Build librdm_b200.so (sm_100a only) in-tree with nvcc.

    python -m md_rdm_b200.build [--force] [--verbose]

No JIT cache and no torch dependency: the library is a plain C-ABI shared object
(include/rdm_b200.h) that the Python host code loads with ctypes.  The build is
skipped when the .so is newer than every source.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.normpath(os.path.join(HERE, "..", "include"))
LIB = os.path.join(HERE, "librdm_b200.so")
SOURCES = ("rdm_api.cu", "rdm_pair.cu", "rdm_als.cu", "rdm_als_sparse.cu", "rdm_tail.cu", "rdm_dorn.cu")

# No -use_fast_math: bins must be bit-exact and the f32 pair products must not be contracted
# (the kernels use explicit __fmul_rn / __frcp_rn / __dmul_rn where it matters).
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: librdm_b200.so cannot be built")
    return exe


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "rdm_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """`defines` / `out` build an experimental variant (tools/als_variants.py); the product is the default."""
    if not force and out == LIB and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build", os.path.basename(out).replace(".so", ""))
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- {src}\n{log}", flush=True)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed (see output above)")
    tmp = out + ".tmp"
    # static cudart: the library is self-contained and shares the primary context (and therefore
    # torch's streams) with whatever runtime the host process already uses
    subprocess.check_call([nvcc, "-shared", "-o", tmp, *objs, "-cudart", "static"])
    os.replace(tmp, out)
    return out


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
