"""Ingest in front of the path: bilateral depth filter, 16-bit depth
decode, grey image.  CPU: the oracle against the committed OpenCV golden vectors.  GPU: the CUDA
kernels against the oracle, and the whole frame with SSF_FLAG_BILATERAL / 16-bit depth."""
import os

import numpy as np
import pytest

from conftest import TUM_PARAMS
from supersurfel_fusion_b200.synth import SyntheticSequence

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ingest_golden.npz")


def test_oracle_ingest_matches_opencv_golden(orc):
    g = np.load(GOLDEN)
    decoded = orc.depth16_to_metres(g["depth16"], float(g["scale"]))
    assert np.array_equal(decoded, g["decoded"])
    filtered = orc.bilateral_filter(decoded)
    # OpenCV's CPU filter interpolates its colour weights from a 4096-bin table: 1e-4 m is generous
    assert np.abs(filtered - g["filtered"]).max() < 1e-4
    assert np.array_equal(filtered == 0, g["filtered"] == 0)          # holes stay holes
    gray = orc.rgb_to_gray(g["rgb"])
    # 14-bit (OpenCV 3.4 CUDA) vs 15-bit (OpenCV 4 CPU) coefficients: at most one grey level apart
    assert np.abs(gray.astype(int) - g["gray_cv2"].astype(int)).max() <= 1


def test_oracle_bilateral_properties(orc):
    rs = np.random.RandomState(3)
    flat = np.full((40, 56), 1.25, np.float32)
    assert np.allclose(orc.bilateral_filter(flat), 1.25, atol=1e-6)            # constants are fixed points
    step = flat.copy(); step[:, 28:] = 2.5                                     # a 1.25 m step is an edge: preserved
    out = orc.bilateral_filter(step)
    assert np.abs(out - step).max() < 1e-5
    noisy = (flat + rs.normal(0, 0.004, flat.shape)).astype(np.float32)
    out = orc.bilateral_filter(noisy)
    assert out.std() < 0.5 * noisy.std()                                       # in-plane noise is smoothed
    k3 = orc.bilateral_filter(noisy, kernel_size=3)                            # explicit kernel size: radius 1
    assert k3.std() > out.std()


@pytest.mark.gpu
@pytest.mark.parametrize("size,seed", [((320, 240), 11), ((640, 480), 12), ((77, 45), 13)])
def test_gpu_bilateral_matches_oracle(orc, size, seed):
    from supersurfel_fusion_b200 import CamParam, SupersurfelFusion
    w, h = size
    seq = SyntheticSequence(width=max(w, 64), height=max(h, 48), seed=seed)
    depth = np.ascontiguousarray(seq.frame(1)[1][:h, :w])
    eng = SupersurfelFusion().initialize(CamParam(525.0, 525.0, w / 2.0, h / 2.0, h, w), nb_supersurfels_max=2000)
    for args in ((-1, 0.03, 4.5), (5, 0.05, 2.0), (-1, 0.02, 8.0)):
        want = orc.bilateral_filter(depth, *args)
        got = eng.bilateralFilter(depth, *args)
        # same taps, same order, same fp32 operations; expf differs in the last ulp between libm and the GPU
        assert np.abs(got - want).max() <= 2e-6 * max(1.0, float(np.abs(want).max()))
        assert np.array_equal(got == 0, want == 0)
    eng.close()


@pytest.mark.gpu
def test_gpu_frame_with_bilateral_and_depth16(orc):
    """processFrame with the in-library filter == processFrame on depth filtered by the oracle;
    the 16-bit entry point == the float entry point on the decoded image; grey image bit-exact."""
    from supersurfel_fusion_b200 import CamParam, SupersurfelFusion
    from supersurfel_fusion_b200.engine import SSF_FLAG_BILATERAL
    seq = SyntheticSequence(width=320, height=240, seed=21)
    cam = CamParam(*seq.cam_param())
    params = dict(TUM_PARAMS, nb_supersurfels_max=20000)
    a = SupersurfelFusion().initialize(cam, **params)      # filter inside the frame graph
    b = SupersurfelFusion().initialize(cam, **params)      # oracle-filtered depth fed in
    c = SupersurfelFusion().initialize(cam, **params)      # 16-bit depth + filter
    for k in range(5):
        rgb, depth = seq.frame(k)
        d16 = np.round(depth * 5000.0).astype(np.uint16)
        dec = orc.depth16_to_metres(d16, 0.0002)
        sa = a.processFrame(rgb, dec, flags=SSF_FLAG_BILATERAL)
        sb = b.processFrame(rgb, orc.bilateral_filter(dec))
        sc = c.processFrameDepth16(rgb, d16, 0.0002, flags=SSF_FLAG_BILATERAL)
        assert np.abs(a.getFilteredDepth() - orc.bilateral_filter(dec)).max() < 1e-5
        # identical inputs up to the last ulp of expf: counts equal, poses within the parity bound
        assert sa["nb_supersurfels"] == sc["nb_supersurfels"] and sa["icp_valid"] == sc["icp_valid"]
        assert np.array_equal(a.getSegmentation()["labels"], c.getSegmentation()["labels"])
        assert np.array_equal(a.getPose()[1], c.getPose()[1])
        assert abs(sa["nb_supersurfels"] - sb["nb_supersurfels"]) <= 2
        assert np.linalg.norm(a.getPose()[1] - b.getPose()[1]) < 1e-4
        assert np.array_equal(a.getGray(), orc.rgb_to_gray(rgb))
    for e in (a, b, c):
        e.close()
