"""GPU parity of the parts of the model life cycle that need MANY frames:
the age-based removal of filterModel -- `time_diff > delta_t && conf < conf_thresh && stamp > delta_t`,
core/src/supersurfel_fusion_kernels.cu:429 -- and the steady-state model turnover behind it (stable 3-way
partition with non-zero nbRemoved, core/src/supersurfel_fusion.cu:441-475), CUDA path vs CPU oracle;
and the full pipeline at 1280x960 (BASELINE configs[2]) vs the oracle."""
import numpy as np
import pytest

from conftest import TUM_PARAMS, make_pair, rel_err, rot_angle
from supersurfel_fusion_b200.synth import SyntheticSequence

pytestmark = pytest.mark.gpu

STAT_KEYS = ("stamp", "nb_supersurfels", "nb_visible", "nb_removed", "icp_valid", "icp_iters", "nb_matched", "nb_inserted")


def _compare_models(oeng, geng, tol=1e-4):
    n = oeng.last["nb_supersurfels"]
    mo, mg = oeng.model(), geng.getModel(n)
    assert np.array_equal(mg.stamps, mo.stamps)
    assert np.array_equal(mg.confidences, mo.confidences)
    for name in ("positions", "colors", "orientations", "shapes", "dims"):
        assert rel_err(getattr(mg, name), getattr(mo, name)) < tol, name


def _run(orc, seq, params, n_frames, model_every):
    oeng, geng = make_pair(orc, seq, params)
    reasons = dict(stale=0, invalid=0, occluded=0)
    worst_t = worst_r = 0.0
    for k in range(n_frames):
        rgb, depth = seq.frame(k)
        so = oeng.process_frame(rgb, depth)
        sg = geng.processFrame(rgb, depth)
        for key in STAT_KEYS[:-2 if k == 0 else None]:      # (the bootstrap copy counts as S insertions on the CUDA side)
            assert sg[key] == so[key], (k, key, sg, so)
        assert sg["icp_inliers"] == so["icp_inliers"], k
        for r in reasons:
            reasons[r] += so["nb_removed_" + r]
        Ro, to = oeng.pose()
        Rg, tg = geng.getPose()
        worst_t = max(worst_t, float(np.linalg.norm(tg - to)))
        worst_r = max(worst_r, rot_angle(Rg, Ro))
        if k % model_every == model_every - 1:
            _compare_models(oeng, geng)
            assert np.array_equal(geng.getSegmentation()["labels"], oeng.tps.get()["labels"]), k
    assert worst_t < 1e-4 and worst_r < 1e-4, (worst_t, worst_r)     # north_star: pose within 1e-4 m
    _compare_models(oeng, geng)
    return reasons


def test_stale_removal_short_window(orc):
    """delta_t = 5: supersurfels not re-observed for more than 5 frames and still below conf_thresh die; 32 frames
    of the VGA sequence remove > 1000 of them through that branch (counted by the oracle), and the CUDA path
    removes exactly the same ones: per-frame counts equal, model equal after the partition."""
    seq = SyntheticSequence(seed=1234)
    reasons = _run(orc, seq, dict(TUM_PARAMS, delta_t=5), 32, model_every=8)
    assert reasons["stale"] > 500 and reasons["occluded"] > 0, reasons


def test_stale_removal_tum_window(orc):
    """The TUM launch file's delta_t = 20 needs > 21 frames before the branch can fire at all (stamp > delta_t)."""
    seq = SyntheticSequence(seed=1234)
    reasons = _run(orc, seq, dict(TUM_PARAMS, delta_t=20), 36, model_every=12)
    assert reasons["stale"] > 50, reasons


def test_full_pipeline_1280x960_vs_oracle(orc):
    """BASELINE configs[2]: 1280x960 stream (S = 4800), the whole per-frame path against the oracle."""
    seq = SyntheticSequence(width=1280, height=960, seed=1234)
    oeng, geng = make_pair(orc, seq, TUM_PARAMS)
    for k in range(3):
        rgb, depth = seq.frame(k)
        so = oeng.process_frame(rgb, depth)
        sg = geng.processFrame(rgb, depth)
        for key in STAT_KEYS[:-2 if k == 0 else None]:
            assert sg[key] == so[key], (k, key, sg, so)
        assert sg["icp_inliers"] == so["icp_inliers"], k
        seg_g, seg_o = geng.getSegmentation(), oeng.tps.get()
        assert np.array_equal(seg_g["labels"], seg_o["labels"]), k
        assert np.array_equal(seg_g["inliers"], seg_o["inliers"]), k
        Ro, to = oeng.pose()
        Rg, tg = geng.getPose()
        assert np.linalg.norm(tg - to) < 1e-4 and rot_angle(Rg, Ro) < 1e-4, k
    _compare_models(oeng, geng)
    fo, fg = oeng.frame(), geng.getFrame()
    assert np.array_equal(fg.confidences, fo.confidences)
    assert rel_err(fg.positions, fo.positions) < 1e-4
