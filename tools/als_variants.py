#!/usr/bin/env python
"""Build instrumented / experimental variants of librdm_b200.so and print their ALS loop timings on the
GPU (under gpurun).  RDM_TIMING adds clock64 printouts per CTA; RDM_EXP removes pieces of the iteration
(wrong results, timing only): 1 record, 2 reciprocals, 4 barriers, 8 iterate history, 16 reduce-scatter.

    python tools/als_variants.py build          # here (nvcc), writes md_rdm_b200/variants/*.so (git-ignored)
    python tools/als_variants.py run            # on the GPU box
"""
import os
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
VAR = os.path.join(ROOT, "md_rdm_b200", "variants")
EXPS = [0, 1, 2, 4, 8, 15]


def name(x):
    return os.path.join(VAR, f"librdm_t{x}.so")


if sys.argv[1] == "build":
    from md_rdm_b200 import build
    os.makedirs(VAR, exist_ok=True)
    for x in EXPS:
        build.build(force=True, defines=("RDM_TIMING=1", f"RDM_EXP={x}"), out=name(x))
        print("built", name(x), flush=True)
else:
    for x in EXPS:
        env = dict(os.environ, RDM_B200_LIB=name(x))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_als.py"), "2"], env=env, capture_output=True, text=True)
        lines = [ln for ln in out.stdout.splitlines() if "block 37 loop" in ln or "unit 37 rows" in ln or "fuse_tail" in ln]
        print(f"== RDM_EXP={x}\n" + "\n".join(lines[-3:]), flush=True)
