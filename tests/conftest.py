import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import orc as _orc
    _orc.build()
    return _orc


@pytest.fixture(scope="session")
def ssf_lib_path():
    from supersurfel_fusion_b200 import build as _b
    return _b.build()


TUM_PARAMS = dict(cell_size=16, lambda_pos=10.0, lambda_bound=1000.0, lambda_size=1000.0, lambda_disp=1e8,
                  thresh_disp=1e-4, seg_iter=10, seg_use_ransac=1, nb_samples=16, filter_iter=3,
                  filter_alpha=0.1, filter_beta=1.0, filter_threshold=0.05, range_min=0.2, range_max=5.0,
                  delta_t=20, conf_thresh=2560.0, nb_supersurfels_max=100000, icp_iter=10, icp_cov_thresh=0.05)
DEFAULT_PARAMS = dict(cell_size=16, lambda_pos=50.0, lambda_bound=1000.0, lambda_size=10000.0, lambda_disp=1e6,
                      thresh_disp=1e-4, seg_iter=10, seg_use_ransac=1, nb_samples=16, filter_iter=4,
                      filter_alpha=0.1, filter_beta=1.0, filter_threshold=0.05, range_min=0.2, range_max=5.0,
                      delta_t=20, conf_thresh=2500.0, nb_supersurfels_max=50000, icp_iter=10, icp_cov_thresh=0.04)


def make_pair(orc_mod, seq, params):
    """(oracle engine, CUDA engine) configured identically for a SyntheticSequence."""
    from supersurfel_fusion_b200 import CamParam, SupersurfelFusion
    cam = seq.cam_param()
    ocfg = orc_mod.default_config(cam=cam, **params)
    oeng = orc_mod.Engine(ocfg)
    kw = dict(params)
    kw["seg_use_ransac"] = bool(kw["seg_use_ransac"])
    geng = SupersurfelFusion().initialize(CamParam(*cam), **kw)
    return oeng, geng


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = max(np.abs(b).max(), 1e-30) if b.size else 1.0
    return float(np.abs(a - b).max() / den) if b.size else 0.0


def rot_angle(Ra, Rb):
    """Angle of Ra^T Rb.  atan2 of the skew part: arccos(trace) cannot resolve angles below
    ~5e-4 rad when the matrices are fp32."""
    d = np.asarray(Ra, np.float64).T @ np.asarray(Rb, np.float64)
    s = 0.5 * np.array([d[2, 1] - d[1, 2], d[0, 2] - d[2, 0], d[1, 0] - d[0, 1]])
    return float(np.arctan2(np.linalg.norm(s), (np.trace(d) - 1.0) / 2.0))


def same_floats(a, b):
    """Bit-identical fp32 arrays, except that any NaN matches any NaN (the device produces the
    canonical 0x7FFFFFFF, x86 SSE produces 0xFFC00000: NaN payloads are not arithmetic results)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    if a.shape != b.shape:
        return False
    both_nan = np.isnan(a) & np.isnan(b)
    return bool(np.all(both_nan | (a.view(np.uint32) == b.view(np.uint32))))
