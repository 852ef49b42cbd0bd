"""CPU tests: the oracle against golden vectors produced by the REFERENCE'S OWN kernels
(tests/golden/ref_golden_320x240.npz, generated on a B200 by tests/golden/make_ref_golden.py
from the reference sources compiled unmodified, oracle/_ref).  This is what pins the oracle;
the reference ships no tests or vectors of its own.

Tolerances: the reference is compiled with --use_fast_math and accumulates with fp32 atomics
in scheduling order, so comparisons against it carry that noise floor: exact for counts and
labels-derived integers, 1e-5 for sums without cancellation, looser where the reference's
own fp32 cancellation dominates (covariances / normals)."""
import os

import numpy as np
import pytest
from scipy.spatial import cKDTree

from conftest import rel_err, rot_angle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden_320x240.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(G))


@pytest.fixture(scope="module")
def cfg(orc, gold):
    params = dict(eval(str(gold["params_json"])))
    return orc.default_config(cam=tuple(gold["cam"][:4]) + (int(gold["cam"][4]), int(gold["cam"][5])), **params)


def _surfels(orc, gold, prefix, n=None):
    m = len(gold[prefix + "positions"]) if n is None else n
    s = orc.Surfels(m)
    for name, _, _ in orc.Surfels.FIELDS:
        getattr(s, name)[:] = gold[prefix + name][:m]
    return s


def test_icp_system_against_reference_kernel(orc, gold, cfg):
    nv = int(gold["nv"])
    model, frame = _surfels(orc, gold, "model_"), _surfels(orc, gold, "frame_")
    got = orc.icp_system(cfg.cam, model.positions[:nv], model.colors[:nv], model.orientations[:nv], frame.colors,
                         frame.orientations, frame.confidences, gold["icp_Rv"], gold["icp_tv"], gold["seg_labels"],
                         gold["seg_slanted"])
    want = gold["icp_system"]
    assert want[28] > 100
    assert got[28] == want[28]                       # inlier count of computeSymmetricICPSystem<128>
    assert rel_err(got, want) < 1e-5


def test_icp_loop_against_reference_host_loop(orc, gold, cfg):
    nv = int(gold["nv"])
    model, frame = _surfels(orc, gold, "model_"), _surfels(orc, gold, "frame_")
    ok, R, t, st = orc.icp(cfg.cam, model.positions[:nv], model.colors[:nv], model.orientations[:nv], frame.colors,
                           frame.orientations, frame.confidences, gold["icp2_Rv"], gold["icp2_tv"],
                           gold["seg_labels"], gold["seg_slanted"], nb_iter=cfg.icp_iter,
                           cov_thresh=cfg.icp_cov_thresh)
    assert ok == bool(gold["icp2_valid"]) and ok
    assert st["iters"] >= 2
    # DenseRegistration::featureConstrainedSymmetricICP with Eigen LDLT / AngleAxis / Quaternion
    assert np.linalg.norm(t - gold["icp2_trel"]) < 1e-6
    assert rot_angle(R, gold["icp2_Rrel"]) < 1e-6


def test_extraction_against_reference_kernels(orc, gold, cfg):
    S = len(gold["frame_confidences"])
    got = orc.generate_supersurfels(cfg.cam, S, gold["rgba"], gold["seg_slanted"], gold["seg_labels"],
                                    gold["seg_inliers"], gold["seg_bound"], cfg.range_min, cfg.range_max, 4)
    want = _surfels(orc, gold, "frame_")
    assert np.array_equal(got.confidences > 0, want.confidences > 0)
    v = want.confidences > 0
    assert v.sum() > 150
    assert np.array_equal(got.confidences[v], want.confidences[v])          # pixel counts
    assert np.array_equal(got.stamps, want.stamps)
    assert rel_err(got.positions[v], want.positions[v]) < 1e-5
    assert rel_err(got.colors[v], want.colors[v]) < 1e-4
    # covariance = E[pp^T] - mu mu^T in fp32: the reference's atomic-order noise is ~1e-3 relative
    assert rel_err(got.shapes[v], want.shapes[v]) < 2e-2
    ang = np.arccos(np.clip(np.abs((got.orientations[v][:, 6:9] * want.orientations[v][:, 6:9]).sum(1)), 0, 1))
    assert np.median(ang) < 5e-3 and ang.max() < 0.1
    # invalid entries keep the raw sums, as in the reference (they are copied into the model on frame 0)
    inv = ~v & (np.abs(want.positions).sum(1) > 0)
    if inv.any():
        assert rel_err(got.positions[inv], want.positions[inv]) < 1e-4


def test_fusion_against_reference_kernels(orc, gold, cfg):
    nb, nv = int(gold["nb"]), int(gold["nv"])
    frame = _surfels(orc, gold, "frame_")
    model = orc.Surfels(cfg.nb_supersurfels_max)
    for name, _, _ in orc.Surfels.FIELDS:
        getattr(model, name)[:nb] = gold["model_" + name][:nb]
    c = orc.fuse(cfg.cam, frame, model, cfg.nb_supersurfels_max, gold["pose_R"], gold["pose_t"], gold["seg_labels"],
                 gold["seg_slanted"], cfg.range_min, cfg.range_max, 4, cfg.delta_t, cfg.conf_thresh, nb, nv)
    want = gold["fuse_counts"]
    assert [c["nb_supersurfels"], c["nb_visible"], c["nb_removed"]] == list(want)
    # the reference's model order is not reproducible (atomic insertion slots, unstable sort):
    # match supersurfels by nearest position and compare attributes
    n = int(want[0])
    ref_pos = gold["fused_positions"][:n]
    tree = cKDTree(model.positions[:n])
    d, idx = tree.query(ref_pos)
    # findBestMatches keeps its per-superpixel minimum with a non-atomic compare followed by two
    # atomicExch (supersurfel_fusion_kernels.cu:590-594): when two model supersurfels compete for
    # one frame superpixel the reference may keep the farther one.  The oracle keeps the true
    # arg-min (appendix B11), so a few pairs may legitimately differ.
    same = d < 1e-5
    assert same.mean() > 0.98, same.mean()
    idx_s = idx[same]
    assert len(np.unique(idx_s)) == same.sum()
    assert np.array_equal(model.confidences[:n][idx_s], gold["fused_confidences"][:n][same])
    assert np.array_equal(model.stamps[:n][idx_s], gold["fused_stamps"][:n][same])
    assert rel_err(model.colors[:n][idx_s], gold["fused_colors"][:n][same]) < 1e-4
    assert rel_err(model.shapes[:n][idx_s], gold["fused_shapes"][:n][same]) < 1e-4
    # total confidence mass is conserved whichever competitor won
    assert abs(model.confidences[:n].sum() - gold["fused_confidences"][:n].sum()) < 1e-3 * gold["fused_confidences"][:n].sum()


def test_stale_removal_against_reference_kernels(orc):
    """filterModel's age-based removal (`time_diff > delta_t && conf < conf_thresh && stamp > delta_t`,
    supersurfel_fusion_kernels.cu:429) pinned to the reference's own kernels: the model the reference built over 27
    frames (delta_t = 20, so the branch can only fire from frame 21 on) and what its model update did to it on frame
    27 (tests/golden/make_ref_golden_long.py), against the oracle's fusion on the same inputs."""
    import ast
    path = os.path.join(os.path.dirname(G), "ref_golden_long_320x240.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ref_golden_long_320x240.npz has not been generated (needs a GPU)")
    g = np.load(path)
    params = dict(ast.literal_eval(str(g["params_json"])))
    cfg = orc.default_config(cam=tuple(g["cam"][:4]) + (int(g["cam"][4]), int(g["cam"][5])), **params)
    assert cfg.delta_t == 20 and int(g["stamp"]) > cfg.delta_t
    nb, nv, stamp = int(g["nb"]), int(g["nv"]), int(g["stamp"])
    frame = _surfels(orc, g, "frame_")
    model = orc.Surfels(cfg.nb_supersurfels_max)
    for name, _, _ in orc.Surfels.FIELDS:
        getattr(model, name)[:nb] = g["model_" + name][:nb]
    # which model supersurfels the stale branch must take, straight from its definition
    age = stamp - g["model_stamps"][:nb, 1]
    stale = (age > cfg.delta_t) & (g["model_confidences"][:nb] < cfg.conf_thresh) & (g["model_confidences"][:nb] > 0)
    assert stale.sum() > 0, "the recorded state does not exercise the branch"
    c = orc.fuse(cfg.cam, frame, model, cfg.nb_supersurfels_max, g["pose_R"], g["pose_t"], g["seg_labels"],
                 g["seg_slanted"], cfg.range_min, cfg.range_max, stamp, cfg.delta_t, cfg.conf_thresh, nb, nv)
    want = [int(v) for v in g["fuse_counts"]]
    # a stale supersurfel inside the visible prefix may still be re-observed (its stamp refreshed) by this very
    # frame's association before filterModel looks at it, so the branch removes a subset of `stale`
    assert 0 < c["nb_removed_stale"] <= int(stale.sum())
    # model size, visible prefix and removals: the reference's (the occlusion test p.z < 0.8 z runs under
    # --use_fast_math there, so allow the odd borderline supersurfel)
    assert abs(c["nb_removed"] - want[2]) <= 2 and abs(c["nb_supersurfels"] - want[0]) <= 2, (c, want)
    assert abs(c["nb_visible"] - want[1]) <= 2, (c, want)
    # the survivors are the reference's survivors
    n = min(c["nb_supersurfels"], want[0])
    d, _ = cKDTree(model.positions[:c["nb_supersurfels"]]).query(g["fused_positions"][:want[0]])
    assert (d < 1e-5).mean() > 0.97, (d < 1e-5).mean()
    assert abs(model.confidences[:n].sum() - g["fused_confidences"][:n].sum()) < 2e-2 * g["fused_confidences"][:n].sum()


def test_tps_against_reference_kernels(orc, gold, cfg):
    """The reference's label passes race; the oracle is one legal serialisation, so a
    small, seam-localised label mismatch is expected and bounded here.  RNG streams: frame 5 of
    a fresh generator is not frame 5 of the reference's (state carries over), so RANSAC planes
    differ slightly too; inlier maps still have to agree almost everywhere."""
    from supersurfel_fusion_b200.synth import SyntheticSequence
    seq = SyntheticSequence(width=320, height=240, seed=77)
    tps = orc.Tps(cfg)
    for k in range(4):
        tps.compute(*seq.frame(k))
    rgb, depth = seq.frame(4)
    assert np.array_equal(rgb, gold["rgb"]) and np.array_equal(depth, gold["depth"])
    o = tps.compute(rgb, depth)
    mism = (o["labels"] != gold["seg_labels"]).mean()
    assert mism < 0.01, mism
    assert ((o["inliers"] > 0) != (gold["seg_inliers"] > 0)).mean() < 0.01
    both = np.isfinite(o["slanted"]) & np.isfinite(gold["seg_slanted"]) & (o["labels"] == gold["seg_labels"])
    rel = np.abs(o["slanted"][both] - gold["seg_slanted"][both]) / np.abs(gold["seg_slanted"][both])
    assert np.median(rel) < 1e-4
