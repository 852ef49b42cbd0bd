"""Consumers either side of the model: applyDeformation, the
marker geometry the node publishes, the TUM trajectory line, the model text export.
CPU: oracle properties.  GPU: CUDA kernels / host formatters against the oracle."""
import numpy as np
import pytest

from conftest import TUM_PARAMS, rel_err
from supersurfel_fusion_b200.synth import SyntheticSequence


def _rot(axis, angle):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return (np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K).astype(np.float32)


def _graph(rs, n_nodes, n_model, positions):
    node_pos = (positions[rs.randint(0, n_model, n_nodes)] + rs.normal(0, 0.05, (n_nodes, 3))).astype(np.float32)
    node_rot = np.stack([_rot(rs.normal(size=3), rs.uniform(-0.05, 0.05)) for _ in range(n_nodes)]).reshape(n_nodes, 9)
    node_trans = rs.normal(0, 0.01, (n_nodes, 3)).astype(np.float32)
    nn = np.stack([rs.choice(n_nodes, 4, replace=False) for _ in range(n_model)]).astype(np.int32)
    w = rs.uniform(0.1, 1.0, (n_model, 4)).astype(np.float32)
    w /= w.sum(1, keepdims=True)
    return node_pos, node_rot.astype(np.float32), node_trans, w.astype(np.float32), nn


@pytest.fixture(scope="module")
def tracked(orc):
    seq = SyntheticSequence(width=320, height=240, seed=5)
    cam = seq.cam_param()
    eng = orc.Engine(orc.default_config(cam=cam, **dict(TUM_PARAMS, nb_supersurfels_max=8000)))
    for k in range(4):
        eng.process_frame(*seq.frame(k))
    return dict(seq=seq, cam=cam, eng=eng)


def test_oracle_deformation_identity_and_rigid(orc, tracked):
    m = tracked["eng"].model()
    n = len(m.positions)
    rs = np.random.RandomState(1)
    node_pos, node_rot, node_trans, w, nn = _graph(rs, 12, n, m.positions)
    # identity graph: nothing moves
    ident = np.tile(np.eye(3, dtype=np.float32).reshape(1, 9), (12, 1))
    a = orc.Surfels(n)
    for name, _, _ in orc.Surfels.FIELDS:
        getattr(a, name)[:] = getattr(m, name)
    orc.apply_deformation(a, node_pos, ident, np.zeros_like(node_trans), w, nn)
    assert np.abs(a.positions - m.positions).max() < 1e-5
    assert np.abs(a.orientations - m.orientations).max() < 1e-5
    assert rel_err(a.shapes, m.shapes) < 1e-4
    # every node carries the same pure translation: the model is translated rigidly
    shift = np.array([0.02, -0.01, 0.03], np.float32)
    orc.apply_deformation(a, node_pos, ident, np.tile(shift, (12, 1)), w, nn)
    assert np.abs(a.positions - (m.positions + shift)).max() < 1e-5


def test_oracle_local_cloud_and_rigid_transform(orc, tracked):
    """extractLocalPointCloud (supersurfel_fusion_kernels.cu:490-520) and applyTransformSuperSurfel (:467-488)
    restatements against plain numpy in float64."""
    eng = tracked["eng"]
    m = eng.model()
    R, t = eng.pose()
    thr, radius = 300.0, 1.6
    pos, nrm = orc.extract_local_point_cloud(m, thr, R, t, radius)
    cam = (m.positions.astype(np.float64) - t) @ R.astype(np.float64)             # R^T (p - t), row-vector form
    keep = (m.confidences >= thr) & (np.linalg.norm(cam, axis=1) < radius)
    edge = np.abs(np.linalg.norm(cam, axis=1) - radius) < 1e-5                    # fp32 vs fp64 at the rim
    assert abs(len(pos) - int(keep.sum())) <= int(edge.sum()) and len(pos) > 0 and (~keep).any()
    if not edge.any():
        assert np.abs(pos - cam[keep]).max() < 1e-5
        want_n = m.orientations[keep][:, 6:9].astype(np.float64) @ R.astype(np.float64)
        want_n /= np.linalg.norm(want_n, axis=1, keepdims=True)
        assert np.abs(nrm - want_n).max() < 1e-5
    assert len(eng.local_cloud(radius)[0]) == len(orc.extract_local_point_cloud(m, eng.cfg.conf_thresh, R, t, radius)[0])
    # rigid motion: positions, orientation rows and shapes move; supersurfels with confidence <= 0 stay
    a = orc.Surfels(m.n)
    for name, _, _ in orc.Surfels.FIELDS:
        getattr(a, name)[:] = getattr(m, name)
    a.confidences[:5] = -1.0
    Rm, tm = _rot((0.3, -1.0, 0.2), 0.4), np.array([0.1, 0.2, -0.3], np.float32)
    orc.transform_model(a, Rm, tm)
    assert np.array_equal(a.positions[:5], m.positions[:5]) and np.array_equal(a.shapes[:5], m.shapes[:5])
    assert np.abs(a.positions[5:] - (m.positions[5:] @ Rm.T + tm)).max() < 1e-5
    O = m.orientations[5:].reshape(-1, 3, 3)
    assert np.abs(a.orientations[5:].reshape(-1, 3, 3) - O @ Rm.T).max() < 1e-5
    def full(s):
        return np.stack([s[:, 0], s[:, 1], s[:, 2], s[:, 1], s[:, 3], s[:, 4], s[:, 2], s[:, 4], s[:, 5]], 1).reshape(-1, 3, 3)
    assert rel_err(full(a.shapes[5:]), Rm.astype(np.float64) @ full(m.shapes[5:]).astype(np.float64) @ Rm.T.astype(np.float64)) < 1e-5


def test_oracle_mod_mask_blocks_fusion_and_insertion(orc):
    """MOD hook in the oracle engine: masked frame supersurfels (confidence -1) are neither fused nor inserted."""
    seq = SyntheticSequence(width=320, height=240, seed=9)
    cfg = orc.default_config(cam=seq.cam_param(), **dict(TUM_PARAMS, nb_supersurfels_max=8000))
    a, b = orc.Engine(cfg), orc.Engine(cfg)
    for k in range(3):
        rgb, depth = seq.frame(k)
        sa = a.process_frame(rgb, depth)
        mask = np.ones(a.S, np.uint8) if k == 2 else None        # everything dynamic in the last frame
        sb = b.process_frame(rgb, depth, mask=mask)
    assert sa["nb_matched"] > 0 and sa["nb_inserted"] >= 0
    assert sb["nb_matched"] == 0 and sb["nb_inserted"] == 0 and sb["icp_valid"] == 0
    assert np.all(b.frame().confidences == -1.0)


def test_oracle_markers_and_tum_line(orc, tracked):
    m = tracked["eng"].model()
    pts, col = orc.markers(m, 300.0)
    keep = m.confidences > 300.0
    assert keep.any() and (~keep).any()
    assert not pts[~keep].any() and not col[~keep][..., :3].any() and (col[..., 3] == 1).all()
    # quad centre == position, triangles share the diagonal p0-p2
    centre = 0.25 * (pts[:, 0] + pts[:, 1] + pts[:, 2] + pts[:, 5])
    assert np.abs(centre[keep] - m.positions[keep]).max() < 1e-5
    assert np.array_equal(pts[:, 0], pts[:, 3]) and np.array_equal(pts[:, 2], pts[:, 4])
    ext = np.linalg.norm(pts[:, 0] - pts[:, 1], axis=1)[keep]                  # 2 * 3 sqrt(lambda2)
    assert np.allclose(ext, 6.0 * np.sqrt(m.dims[keep, 1]), rtol=1e-4, atol=1e-6)
    R, t = tracked["eng"].pose()
    line = orc.format_tum_pose(R, t, "1305031102.175304")
    f = line.split()
    assert f[0] == "1305031102.175304" and len(f) == 8 and line.endswith("\n")
    q = np.array([float(x) for x in f[4:]])
    assert abs(np.linalg.norm(q) - 1.0) < 1e-5 and np.allclose([float(x) for x in f[1:4]], t, atol=1e-5)
    assert orc.format_tum_pose(np.eye(3), np.zeros(3), "0") == "0 0 0 0 0 0 0 1\n"


@pytest.mark.gpu
def test_gpu_consumers_match_oracle(orc, tracked, tmp_path):
    from supersurfel_fusion_b200 import CamParam, SupersurfelFusion, Supersurfels
    m = tracked["eng"].model()
    n = len(m.positions)
    R, t = tracked["eng"].pose()
    eng = SupersurfelFusion().initialize(CamParam(*tracked["cam"]), **dict(TUM_PARAMS, nb_supersurfels_max=8000))
    eng.setModel(Supersurfels.from_arrays(**m.as_dict()), n, n)
    eng.setPose(R, t)
    # markers: same fp32 expressions -> bit-exact
    pts_o, col_o = orc.markers(m, 300.0)
    pts_g, col_g = eng.getMarkers("model", 300.0)
    assert np.array_equal(pts_g, pts_o) and np.array_equal(col_g, col_o)
    # TUM line: host formatter, identical text
    assert eng.formatTumPose("1305031102.175304") == orc.format_tum_pose(R, t, "1305031102.175304")
    # deformation
    rs = np.random.RandomState(2)
    node_pos, node_rot, node_trans, w, nn = _graph(rs, 16, n, m.positions)
    want = orc.Surfels(n)
    for name, _, _ in orc.Surfels.FIELDS:
        getattr(want, name)[:] = getattr(m, name)
    orc.apply_deformation(want, node_pos, node_rot, node_trans, w, nn)
    eng.applyDeformation(node_pos, node_rot, node_trans, w, nn)
    got = eng.getModel(n)
    assert np.abs(got.positions - want.positions).max() < 1e-6
    assert np.abs(got.orientations - want.orientations).max() < 1e-6
    assert rel_err(got.shapes, want.shapes) < 1e-5
    assert np.array_equal(got.colors, want.colors) and np.array_equal(got.confidences, want.confidences)
    # exportModel: text format of supersurfel_fusion.cu:616-630, one 7-line record per stable supersurfel
    path = str(tmp_path / "model.txt")
    eng.exportModel(path)
    lines = open(path).read().split("\n")
    stable = int((got.confidences > TUM_PARAMS["conf_thresh"]).sum())
    assert len(lines) == 7 * stable + 1
    if stable:
        i = int(np.nonzero(got.confidences > TUM_PARAMS["conf_thresh"])[0][0])
        assert lines[0] == "%d %d %f" % (got.stamps[i, 0], got.stamps[i, 1], got.confidences[i])
        assert lines[1] == "%f %f %f" % tuple(got.positions[i])
    eng.close()
