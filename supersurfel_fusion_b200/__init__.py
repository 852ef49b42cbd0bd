"""supersurfel_fusion_b200 -- B200-native (sm_100a) supersurfel tracking-and-fusion hot path.

The product is the CUDA library ``libssf.so`` (C-ABI in ``include/ssf.h``); this package is
its Python host mirror of the reference's ``SupersurfelFusion`` class.
"""
from .engine import (CamParam, SsfError, SupersurfelFusion, Supersurfels, lib_path, load_library)  # noqa: F401

__all__ = ["CamParam", "SsfError", "SupersurfelFusion", "Supersurfels", "lib_path", "load_library"]
