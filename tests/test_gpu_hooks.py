"""GPU parity of the hooks through which the out-of-scope neighbours touch the model:
  * MOD mask hook: ssf_invalidate_frame_supersurfels sets frame confidences to -1 (motion_detection.cu:573)
    between extraction and registration; fusion must neither fuse nor insert those supersurfels;
  * extractLocalPointCloud (supersurfel_fusion_kernels.cu:490-520, supersurfel_fusion.cu:884-920);
  * ssf_transform_model = applyTransformSuperSurfel (supersurfel_fusion_kernels.cu:467-488) + ssf_set_pose,
    the rigid loop-closure correction.
CUDA path (through the C-ABI) vs the CPU oracle restatements."""
import numpy as np
import pytest

from conftest import TUM_PARAMS, make_pair, rel_err, rot_angle
from supersurfel_fusion_b200.synth import SyntheticSequence

pytestmark = pytest.mark.gpu


def _row_order(a):
    return np.lexsort(np.round(np.asarray(a), 4).T[::-1])


def _rot(axis, ang):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return (np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)).astype(np.float32)


def test_mod_mask_hook_matches_oracle(orc):
    seq = SyntheticSequence(width=320, height=240, seed=17)
    # icp_cov_thresh relaxed: with the launch file's 0.05 a 320x240 frame (300 superpixels) never passes the
    # covariance gate, and the hook must be exercised on frames whose registration is applied
    params = dict(TUM_PARAMS, nb_supersurfels_max=20000, icp_cov_thresh=1.0)
    oeng, geng = make_pair(orc, seq, params)
    _, plain = make_pair(orc, seq, params)
    rng = np.random.RandomState(5)
    S = geng.nbSuperpixels
    for k in range(6):
        rgb, depth = seq.frame(k)
        mask = (rng.uniform(size=S) < 0.25).astype(np.uint8) if k >= 1 else None
        so = oeng.process_frame(rgb, depth, mask=mask)
        sg = geng.processFrameStaged(rgb, depth, dynamic_mask=mask)
        sp = plain.processFrame(rgb, depth)
        for key in ("stamp", "nb_supersurfels", "nb_visible", "nb_removed", "icp_valid", "icp_iters") + \
                (("nb_matched", "nb_inserted") if k else ()):        # (the bootstrap copy counts as S insertions here)
            assert sg[key] == so[key], (k, key, sg, so)
        if k >= 1:
            assert so["icp_valid"] == 1, k
        if mask is not None:
            fg = geng.getFrame()
            assert np.all(fg.confidences[mask != 0] == -1.0)            # motion_detection.cu:573
            assert np.array_equal(fg.confidences, oeng.frame().confidences)
        Ro, to = oeng.pose()
        Rg, tg = geng.getPose()
        assert np.linalg.norm(tg - to) < 1e-4 and rot_angle(Rg, Ro) < 1e-4, k
    # the masked supersurfels were neither fused nor inserted: fewer insertions than without a mask
    assert sg["nb_supersurfels"] < sp["nb_supersurfels"]
    n = so["nb_supersurfels"]
    mo, mg = oeng.model(), geng.getModel(n)
    assert np.array_equal(mg.confidences, mo.confidences) and np.array_equal(mg.stamps, mo.stamps)
    for name in ("positions", "colors", "orientations", "shapes", "dims"):
        assert rel_err(getattr(mg, name), getattr(mo, name)) < 1e-4, name


def test_staged_frame_without_mask_equals_process_frame(orc):
    """The stage entry points run the same kernels as ssf_process_frame."""
    seq = SyntheticSequence(width=320, height=240, seed=18)
    _, a = make_pair(orc, seq, TUM_PARAMS)
    _, b = make_pair(orc, seq, TUM_PARAMS)
    for k in range(4):
        rgb, depth = seq.frame(k)
        sa = a.processFrame(rgb, depth)
        sb = b.processFrameStaged(rgb, depth)
        for key in ("stamp", "nb_supersurfels", "nb_visible", "nb_removed", "nb_matched", "nb_inserted"):
            assert sa[key] == sb[key], (k, key)
        assert np.array_equal(a.getPose()[1], b.getPose()[1]) and np.array_equal(a.getPose()[0], b.getPose()[0])
    ma, mb = a.getModel(), b.getModel()
    assert np.array_equal(ma.positions, mb.positions) and np.array_equal(ma.confidences, mb.confidences)


def test_local_point_cloud_matches_oracle(orc):
    seq = SyntheticSequence(width=320, height=240, seed=19)
    params = dict(TUM_PARAMS, conf_thresh=300.0, nb_supersurfels_max=20000, icp_cov_thresh=1.0)
    oeng, geng = make_pair(orc, seq, params)
    for k in range(5):
        rgb, depth = seq.frame(k)
        oeng.process_frame(rgb, depth)
        geng.processFrame(rgb, depth)
    m = geng.getModel()
    stable = int((m.confidences >= 300.0).sum())
    assert 0 < stable < m.n                       # the threshold actually splits this model
    for radius in (None, 2.9, 2.0):
        r = params["range_max"] if radius is None else radius
        po, no = oeng.local_cloud(r)
        pg, ng = geng.extractLocalPointCloud(radius)
        assert len(pg) == len(po) > 0, (radius, len(pg), len(po))
        # the reference appends by atomic ticket: compare as sets of rows, paired by position
        og, oo = _row_order(pg), _row_order(po)
        assert np.abs(pg[og] - po[oo]).max() < 1e-5
        assert np.abs(ng[og] - no[oo]).max() < 1e-5
        assert np.all(np.linalg.norm(pg, axis=1) < r)
        assert np.abs(np.linalg.norm(ng, axis=1) - 1.0).max() < 1e-5
    assert len(geng.extractLocalPointCloud(None)[0]) <= stable
    assert len(geng.extractLocalPointCloud(2.0)[0]) < len(geng.extractLocalPointCloud(2.9)[0]) < len(geng.extractLocalPointCloud(None)[0])


def test_transform_model_matches_oracle_and_tracking_continues(orc):
    """Rigid loop-closure correction: model <- T model, pose <- T pose (supersurfel_fusion.cu:794-822 applies the
    same pair); tracking of the next frames must be unaffected in the moved world frame."""
    seq = SyntheticSequence(width=320, height=240, seed=23)
    params = dict(TUM_PARAMS, nb_supersurfels_max=20000, icp_cov_thresh=1.0)     # see test_mod_mask_hook_matches_oracle
    oeng, geng = make_pair(orc, seq, params)
    _, still = make_pair(orc, seq, params)
    for k in range(4):
        rgb, depth = seq.frame(k)
        oeng.process_frame(rgb, depth)
        geng.processFrame(rgb, depth)
        still.processFrame(rgb, depth)
    R = _rot((0.2, 1.0, -0.3), 0.35)
    t = np.array([0.4, -0.2, 0.7], np.float32)
    before = geng.getModel()
    oeng.transform_model(R, t)
    geng.transformModel(R, t)
    n = oeng.last["nb_supersurfels"]
    mo, mg = oeng.model(), geng.getModel(n)
    for name in ("positions", "orientations", "shapes"):
        assert rel_err(getattr(mg, name), getattr(mo, name)) < 1e-5, name
    assert np.array_equal(mg.colors, before.colors) and np.array_equal(mg.confidences, before.confidences)
    assert np.abs(mg.positions - (before.positions @ R.T + t)).max() < 1e-5
    # move the pose the same way on both sides and keep tracking
    Rp, tp = geng.getPose()
    Rn, tn = (R @ Rp).astype(np.float32), (R @ tp + t).astype(np.float32)
    oeng.set_pose(Rn, tn)
    geng.setPose(Rn, tn)
    for k in range(4, 7):
        rgb, depth = seq.frame(k)
        so = oeng.process_frame(rgb, depth)
        sg = geng.processFrame(rgb, depth)
        ss = still.processFrame(rgb, depth)
        assert sg["nb_supersurfels"] == so["nb_supersurfels"] and sg["icp_valid"] == so["icp_valid"] == 1
        Ro, to = oeng.pose()
        Rg, tg = geng.getPose()
        assert np.linalg.norm(tg - to) < 1e-4 and rot_angle(Rg, Ro) < 1e-4, k
        # same trajectory as the engine whose world was not moved, seen through T
        Rs, ts = still.getPose()
        assert np.linalg.norm(tg - (R @ ts + t)) < 2e-3 and rot_angle(Rg, R @ Rs) < 2e-3, k
        assert abs(sg["nb_supersurfels"] - ss["nb_supersurfels"]) <= 0.02 * ss["nb_supersurfels"] + 5
