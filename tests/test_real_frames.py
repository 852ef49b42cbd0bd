"""BASELINE.json configs[0]/[1] on REAL data: three consecutive TUM fr1/xyz frames (tests/golden/tum_fr1_xyz).
CPU: the oracle tracks them and agrees with the dataset's ground-truth motion to the accuracy band of
the reference's own trajectory (ATE 0.02 m on this sequence).  GPU: the CUDA path is
bit-exact with the oracle on the label maps and within 1e-4 m / 1e-4 rad on the poses, through the
16-bit depth entry point with the in-library bilateral filter."""
import os

import numpy as np
import pytest

from conftest import TUM_PARAMS, rot_angle

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tum_fr1_xyz")
CAM = (525.0, 525.0, 319.5, 239.5, 480, 640)         # rgbd_benchmark/fr1_cam.yaml
DEPTH_SCALE = 0.0002                                  # launch/supersurfel_fusion_rgbd_benchmark.launch


def _quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def load_frames():
    import cv2
    frames = []
    for line in open(os.path.join(HERE, "associations_with_gt.txt")):
        w = line.split()
        bgr = cv2.imread(os.path.join(HERE, w[1]), cv2.IMREAD_COLOR)
        rgb = np.ascontiguousarray(bgr[:, :, ::-1])                       # the node converts BGR -> RGB (:607-608)
        d16 = cv2.imread(os.path.join(HERE, w[3]), cv2.IMREAD_UNCHANGED)
        assert rgb.shape == (480, 640, 3) and d16.dtype == np.uint16
        gt_t = np.array([float(v) for v in w[5:8]])
        gt_R = _quat_to_R([float(v) for v in w[8:12]])
        frames.append((w[0], rgb, d16, gt_R, gt_t))
    return frames


def test_oracle_tracks_real_frames(orc):
    frames = load_frames()
    assert 0.15 < float((frames[0][2] == 0).mean()) < 0.35               # a quarter of the depth is missing
    eng = orc.Engine(orc.default_config(cam=CAM, **TUM_PARAMS))
    poses = []
    for _, rgb, d16, _, _ in frames:
        depth = orc.bilateral_filter(orc.depth16_to_metres(d16, DEPTH_SCALE))
        st = eng.process_frame(rgb, depth)
        poses.append(eng.pose())
    assert st["icp_valid"] == 1 and st["nb_supersurfels"] > 500
    # relative motion frame 0 -> 2 against the ground truth (both expressed in the first camera frame)
    R0, t0 = frames[0][3], frames[0][4]
    R2, t2 = frames[2][3], frames[2][4]
    gt_rel_t = R0.T @ (t2 - t0)
    gt_rel_R = R0.T @ R2
    est_R, est_t = poses[2]
    assert np.linalg.norm(est_t - gt_rel_t) < 0.02                        # within the reference's own ATE band
    assert rot_angle(est_R, gt_rel_R) < 0.03
    assert np.linalg.norm(gt_rel_t) > 0.01                                # the camera did move


@pytest.mark.gpu
def test_gpu_matches_oracle_on_real_frames(orc):
    from supersurfel_fusion_b200 import CamParam, SupersurfelFusion
    from supersurfel_fusion_b200.engine import SSF_FLAG_BILATERAL
    frames = load_frames()
    cpu = orc.Engine(orc.default_config(cam=CAM, **TUM_PARAMS))
    gpu = SupersurfelFusion().initialize(CamParam(*CAM), **TUM_PARAMS)
    for k, (stamp, rgb, d16, _, _) in enumerate(frames):
        dec = orc.depth16_to_metres(d16, DEPTH_SCALE)
        gpu.processFrameDepth16(rgb, d16, DEPTH_SCALE, flags=SSF_FLAG_BILATERAL)
        # the two bilateral filters agree to the last ulp of expf, not bit for bit: feed the oracle the GPU's
        # filtered image so that everything after the filter can be compared exactly
        filt = gpu.getFilteredDepth()
        assert np.abs(filt - orc.bilateral_filter(dec)).max() < 1e-5
        so = cpu.process_frame(rgb, filt)
        sg = gpu.getFrameStats()
        assert np.array_equal(gpu.getSegmentation()["labels"], cpu.tps.get()["labels"]), "labels differ at frame %d" % k
        assert sg["nb_supersurfels"] == so["nb_supersurfels"] and sg["icp_valid"] == so["icp_valid"]
        Rg, tg = gpu.getPose()
        Ro, to = cpu.pose()
        assert np.linalg.norm(tg - to) < 1e-4 and rot_angle(Rg, Ro) < 1e-4
    line = gpu.formatTumPose(stamp)
    assert line.split()[0] == stamp and len(line.split()) == 8
    gpu.close()
