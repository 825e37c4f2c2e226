"""Lloyd codebooks (40 thresholds + 41 levels per scale).

Mirrors `Quantization` (reference network/RDM_Net.py:397-442): same attribute names
(`depth_ratio_XXX_XXX_quant[_inv]`, numpy (40,1)/(41,1) float64), `get_with_id`,
`get_size_id`.  The reference loads five MATLAB files from the current directory; the
008 file is missing from its tree (`.MISSING_LARGE_BLOBS`), so the tables shipped here
(md_rdm_b200/data/depth_ratio_codebooks.json, exact IEEE-754 hex, written by
tools/import_codebooks.py) carry a DERIVED 008 table = table(016)**2.  `from_mat_dir`
loads real .mat files instead when they are available.
"""
from __future__ import annotations

import json
import warnings
import os
from typing import Dict, Tuple

import numpy as np
import torch

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "depth_ratio_codebooks.json")
SCALES = (8, 16, 32, 64, 128)


def _tag(scale: int) -> str:
    return f"depth_ratio_{scale:03d}_{scale:03d}_quant"


class Quantization:
    """Drop-in for the reference class of the same name, plus device-resident tables."""

    def __init__(self, path: str = _DATA):
        with open(path) as f:
            raw = json.load(f)
        self.derived = {}
        for key, tab in raw["tables"].items():
            s = int(key)
            q = np.array([float.fromhex(h) for h in tab["thresholds"]], dtype=np.float64).reshape(40, 1)
            inv = np.array([float.fromhex(h) for h in tab["levels"]], dtype=np.float64).reshape(41, 1)
            setattr(self, _tag(s), q)
            setattr(self, _tag(s) + "_inv", inv)
            self.derived[s] = bool(tab.get("derived", False))
        self._dev: Dict[Tuple[int, str], Tuple[torch.Tensor, torch.Tensor]] = {}

    @classmethod
    def from_mat_dir(cls, directory: str) -> "Quantization":
        """Load `depth_ratio_*_quant.mat` like RN:403-418 does (scipy needed); scales whose file
        is absent keep the packaged table."""
        import scipy.io
        self = cls()
        for s in SCALES:
            p = os.path.join(directory, _tag(s) + ".mat")
            if os.path.exists(p):
                m = scipy.io.loadmat(p)
                setattr(self, _tag(s), np.asarray(m[_tag(s)], dtype=np.float64).reshape(40, 1))
                setattr(self, _tag(s) + "_inv", np.asarray(m[_tag(s) + "_inv"], dtype=np.float64).reshape(41, 1))
                self.derived[s] = False
        self._dev.clear()
        return self

    def to_mat_dir(self, directory: str, scales=SCALES) -> None:
        """Write the tables in the reference's wire format (RN:403-418): one MATLAB v5 file per scale,
        `depth_ratio_XXX_XXX_quant.mat` with keys `<tag>` (40,1) and `<tag>_inv` (41,1), float64.  The
        reference constructor finds them in its working directory."""
        import scipy.io
        os.makedirs(directory, exist_ok=True)
        for s in scales:
            scipy.io.savemat(os.path.join(directory, _tag(s) + ".mat"),
                             {_tag(s): getattr(self, _tag(s)), _tag(s) + "_inv": getattr(self, _tag(s) + "_inv")})

    def derive(self, scale: int, from_scale: int) -> Tuple[np.ndarray, np.ndarray]:
        """Generate the table of `scale` from the table of `from_scale` with the relation the shipped
        tables satisfy to 1e-14: codebook(s) == codebook(2s)**2 elementwise, i.e. the
        log-ratio range doubles every time the resolution halves.  Installs the result (marked derived)
        and returns (thresholds (40,1), levels (41,1))."""
        if scale not in SCALES or from_scale not in SCALES:
            raise ValueError(f"scales must be in {SCALES}")
        expo = float(from_scale) / float(scale)          # 2s -> s: square; s -> 2s: square root
        q = np.power(getattr(self, _tag(from_scale)), expo)
        inv = np.power(getattr(self, _tag(from_scale) + "_inv"), expo)
        setattr(self, _tag(scale), q)
        setattr(self, _tag(scale) + "_inv", inv)
        self.derived[scale] = True
        self._dev = {k: v for k, v in self._dev.items() if k[0] != scale}
        return q, inv

    def _warn_if_derived(self, scale: int) -> None:
        """The 008 table is missing from the reference tree and is served as table(016)**2: say
        so once per process, a deployment that holds the authors' file gets different bins at that scale."""
        if self.derived.get(scale) and scale not in _warned:
            _warned.add(scale)
            warnings.warn(f"md_rdm_b200: the {scale:03d} Lloyd codebook in use is DERIVED (table of another scale to the power "
                          f"2s/s), not the reference authors' depth_ratio_{scale:03d}_{scale:03d}_quant.mat; load the real file with "
                          "Quantization.from_mat_dir() if you have it", stacklevel=3)

    # ---- reference surface (RN:420-442)
    def get_with_id(self, id):
        if 3 <= id <= 7:
            s = 1 << id
            self._warn_if_derived(s)
            return getattr(self, _tag(s)), getattr(self, _tag(s) + "_inv")

    def get_size_id(self, id):
        if 3 <= id <= 7:
            return 1 << id

    # ---- device tables for the kernels
    def device_tables(self, scale: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        """(thresholds[40], levels[41]) as f64 tensors on `device`, cached."""
        device = torch.device(device)
        key = (scale, str(device))
        hit = self._dev.get(key)
        if hit is None:
            self._warn_if_derived(scale)
            q = torch.from_numpy(getattr(self, _tag(scale)).reshape(-1).copy()).to(device)
            inv = torch.from_numpy(getattr(self, _tag(scale) + "_inv").reshape(-1).copy()).to(device)
            hit = (q, inv)
            self._dev[key] = hit
        return hit


_default = None
_warned = set()


def default_quantization() -> Quantization:
    global _default
    if _default is None:
        _default = Quantization()
    return _default
