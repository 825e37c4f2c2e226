"""md_rdm_b200: B200-native (sm_100a) implementation of the MD_RDM depth-map fusion path.

    import md_rdm_b200.computations as cp          # drop-in for network/computations.py
    from md_rdm_b200.rdm_net import Ordinal_Layer, Quantization, Weights
    from md_rdm_b200.fusion import FusionPlan, fuse_maps   # batched fast path

Importing the package does not load the CUDA library; the first op call does, and raises
if `librdm_b200.so` has not been built (`python -m md_rdm_b200.build`).  There is no CPU or
PyTorch fallback anywhere in this package.
"""
__version__ = "0.1.0"
