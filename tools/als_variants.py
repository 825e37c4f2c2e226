#!/usr/bin/env python
"""Build experimental variants of the ALS inner loop (dot / norm strategies) as separate .so files and,
with `run`, time the dominant kernel for each on the GPU (under gpurun).

    python tools/als_variants.py build          # here (nvcc), writes md_rdm_b200/variants/*.so
    python tools/als_variants.py run            # on the GPU box
"""
import json
import os
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
VAR = os.path.join(ROOT, "md_rdm_b200", "variants")
COMBOS = [(1, 0, 0), (1, 0, 1), (1, 0, 2), (1, 0, 4), (1, 0, 7)]   # (dot, norm, experiment mask)


def name(d, n, x=0):
    return os.path.join(VAR, f"librdm_d{d}n{n}x{x}.so")


if sys.argv[1] == "build":
    from md_rdm_b200 import build
    os.makedirs(VAR, exist_ok=True)
    for d, n, x in COMBOS:
        build.build(force=True, defines=(f"RDM_DOT={d}", f"RDM_NORM={n}", f"RDM_EXP={x}"), out=name(d, n, x))
        print("built", name(d, n, x), flush=True)
else:
    for d, n, x in COMBOS:
        env = dict(os.environ, RDM_B200_LIB=name(d, n, x))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1000", "--warmup", "10", "--no-cpu-baseline"],
                             env=env, capture_output=True, text=True)
        try:
            j = json.loads(out.stdout.strip().splitlines()[-1])
            print(f"dot={d} norm={n} exp={x}: als_iterate {j['config']['kernel_ms']['als_iterate'] * 1e3:.1f} us, step(4 streams) "
                  f"{j['ms_per_step'] * 1e3:.1f} us, single-stream step {j['config']['single_stream_ms_per_step'] * 1e3:.1f} us, "
                  f"value {j['value']:.0f} maps/s", flush=True)
        except Exception as e:
            print(d, n, x, "FAILED", e, out.stderr[-500:], flush=True)
