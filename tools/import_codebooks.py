#!/usr/bin/env python
"""Import the Lloyd codebooks shipped by the reference as MATLAB v5 files into
`md_rdm_b200/data/depth_ratio_codebooks.json` (exact IEEE-754 hex strings).

The reference loads `depth_ratio_{008,016,032,064,128}_*_quant.mat` from the
current directory (network/RDM_Net.py:403-418).  Only 016..128 are shipped;
008 is listed in `.MISSING_LARGE_BLOBS`.  The four shipped tables satisfy
codebook(s) == codebook(2s)**2 elementwise, so the 008
table is DERIVED here as table(016)**2 and is labelled `"derived": true` --
it is not the authors' bytes.

Run in the build container only (needs /root/reference):
    python tools/import_codebooks.py [/root/reference]
"""
import json
import os
import sys

import numpy as np
import scipy.io


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = {"format": "float64 hex (float.hex)", "source": "az16/MD_RDM depth_ratio_*_quant.mat", "tables": {}}
    for s in (16, 32, 64, 128):
        tag = f"{s:03d}_{s:03d}"
        m = scipy.io.loadmat(os.path.join(ref, f"depth_ratio_{tag}_quant.mat"))
        q = np.asarray(m[f"depth_ratio_{tag}_quant"], dtype=np.float64).reshape(-1)
        inv = np.asarray(m[f"depth_ratio_{tag}_quant_inv"], dtype=np.float64).reshape(-1)
        assert q.shape == (40,) and inv.shape == (41,)
        out["tables"][str(s)] = {
            "derived": False,
            "thresholds": [float(v).hex() for v in q],
            "levels": [float(v).hex() for v in inv],
        }
    q16 = np.array([float.fromhex(h) for h in out["tables"]["16"]["thresholds"]])
    i16 = np.array([float.fromhex(h) for h in out["tables"]["16"]["levels"]])
    out["tables"]["8"] = {
        "derived": True,
        "derivation": "table(016)**2 elementwise (float64); stand-in for the missing depth_ratio_008_008_quant.mat",
        "thresholds": [float(v).hex() for v in q16 * q16],
        "levels": [float(v).hex() for v in i16 * i16],
    }
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "md_rdm_b200", "data",
                       "depth_ratio_codebooks.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.normpath(dst))


if __name__ == "__main__":
    main()
