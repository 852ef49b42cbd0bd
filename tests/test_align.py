"""Loop-closure registration DenseRegistration::align.
CPU: the oracle's own properties and (when the golden file exists) the oracle against the
reference's align run on a B200.  GPU: the one-launch CUDA loop against the oracle and against the
reference harness."""
import os

import numpy as np
import pytest

from conftest import TUM_PARAMS, rot_angle
from supersurfel_fusion_b200.synth import SyntheticSequence

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "align_golden_640x480.npz")


def _perturbed(Rrel, trel, scale=1.0):
    dR = np.array([[1, -0.004 * scale, 0.002 * scale], [0.004 * scale, 1, -0.003 * scale],
                   [-0.002 * scale, 0.003 * scale, 1]], np.float64)
    u, _, vt = np.linalg.svd(Rrel.astype(np.float64) @ dR)
    return (u @ vt).astype(np.float32), (trel + scale * np.array([0.006, -0.004, 0.005], np.float32)).astype(np.float32)


@pytest.fixture(scope="module")
def scene(orc):
    """keyframe (frame 2) and current frame (frame 5) of a VGA sequence, tracked by the oracle"""
    seq = SyntheticSequence(width=640, height=480, seed=77)
    cam = seq.cam_param()
    eng = orc.Engine(orc.default_config(cam=cam, **dict(TUM_PARAMS, nb_supersurfels_max=20000)))
    for k in range(3):
        eng.process_frame(*seq.frame(k))
    key = eng.frame()
    Rk, tk = eng.pose()
    for k in range(3, 6):
        eng.process_frame(*seq.frame(k))
    cur = eng.frame()
    Rc, tc = eng.pose()
    seg = eng.tps.get()
    return dict(seq=seq, cam=cam, key=key, cur=cur, labels=seg["labels"], slanted=seg["slanted"], Rrel=Rc.T @ Rk,
                trel=Rc.T @ (tk - tc))


def _orc_align(orc, s, Ri, ti, **kw):
    k, c = s["key"], s["cur"]
    return orc.align(orc.OrcCam(*s["cam"]), k.positions, k.colors, k.orientations, k.confidences, c.colors,
                     c.orientations, c.confidences, Ri, ti, s["labels"], s["slanted"], **kw)


def test_oracle_align_recovers_relative_pose(orc, scene):
    Ri, ti = _perturbed(scene["Rrel"], scene["trel"])
    ok, R, t, st = _orc_align(orc, scene, Ri, ti, nb_iter=10, cov_thresh=0.05)
    assert ok and st["iters"] == 10 and st["pairs"] > 400          # no early exit in align
    Rinc, tinc = R.T, -(R.T @ t)
    before = np.linalg.norm(ti - scene["trel"])
    after = np.linalg.norm(Rinc @ ti + tinc - scene["trel"])
    assert after < 0.35 * before and rot_angle(Rinc @ Ri, scene["Rrel"]) < 0.35 * rot_angle(Ri, scene["Rrel"])


def test_oracle_align_gates(orc, scene):
    Ri, ti = _perturbed(scene["Rrel"], scene["trel"])
    # far-off initial guess: fewer than 100 pairs -> invalid, identity returned (dense_registration.cu:141-146)
    ok, R, t, st = _orc_align(orc, scene, Ri, ti + np.float32(0.5), nb_iter=10, cov_thresh=0.05)
    assert not ok and st["iters"] == 1 and st["pairs"] < 100
    assert np.array_equal(R, np.eye(3, dtype=np.float32)) and not t.any()
    # covariance gate
    ok, _, _, _ = _orc_align(orc, scene, Ri, ti, nb_iter=10, cov_thresh=1e-9)
    assert not ok
    # no valid source supersurfel
    k = scene["key"]
    ok, _, _, st = orc.align(orc.OrcCam(*scene["cam"]), k.positions, k.colors, k.orientations, -np.ones_like(k.confidences),
                             scene["cur"].colors, scene["cur"].orientations, scene["cur"].confidences, Ri, ti,
                             scene["labels"], scene["slanted"])
    assert not ok and st["pairs"] == 0


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="golden vectors not generated yet")
def test_oracle_align_against_reference_golden(orc):
    g = np.load(GOLDEN)
    cam = tuple(g["cam"][:4]) + (int(g["cam"][4]), int(g["cam"][5]))
    ok, R, t, st = orc.align(orc.OrcCam(*cam), g["key_positions"], g["key_colors"], g["key_orientations"],
                             g["key_confidences"], g["cur_colors"], g["cur_orientations"], g["cur_confidences"],
                             g["R_init"], g["t_init"], g["labels"], g["slanted"], nb_iter=int(g["icp_iter"]),
                             cov_thresh=float(g["cov_thresh"]))
    assert ok == bool(g["valid"]) and ok
    # the reference runs with --use_fast_math (approximate division, powf, rsqrtf)
    assert np.linalg.norm(t - g["t"]) < 1e-4 and rot_angle(R, g["R"]) < 1e-4


def _gpu_engine_at(scene):
    from supersurfel_fusion_b200 import CamParam, SupersurfelFusion, Supersurfels
    eng = SupersurfelFusion().initialize(CamParam(*scene["cam"]), **dict(TUM_PARAMS, nb_supersurfels_max=20000))
    cur = Supersurfels.from_arrays(**scene["cur"].as_dict())
    eng.setSegmentation(labels=scene["labels"], slanted=scene["slanted"])
    eng.setFrame(cur)
    key = Supersurfels.from_arrays(**scene["key"].as_dict())
    return eng, key


@pytest.mark.gpu
def test_gpu_align_matches_oracle(orc, scene):
    eng, key = _gpu_engine_at(scene)
    for scale in (1.0, 0.3, 2.0):
        Ri, ti = _perturbed(scene["Rrel"], scene["trel"], scale)
        ok_o, R_o, t_o, st_o = _orc_align(orc, scene, Ri, ti, nb_iter=TUM_PARAMS["icp_iter"], cov_thresh=TUM_PARAMS["icp_cov_thresh"])
        ok_g, R_g, t_g, st_g = eng.align(key, Ri, ti)
        assert ok_g == ok_o and st_g["iters"] == st_o["iters"]
        assert abs(st_g["pairs"] - st_o["pairs"]) <= 1            # a borderline Lab gate (GPU cbrtf vs libm)
        assert np.linalg.norm(t_g - t_o) < 1e-5 and rot_angle(R_g, R_o) < 1e-5
        assert np.abs(st_g["system"] - st_o["system"]).max() <= 1e-4 * np.abs(st_o["system"]).max()
    # gates: starved and covariance
    ok_g, R_g, t_g, st_g = eng.align(key, Ri, ti + np.float32(0.5))
    assert not ok_g and st_g["iters"] == 1 and np.array_equal(R_g, np.eye(3, dtype=np.float32)) and not t_g.any()
    eng.close()


@pytest.mark.gpu
def test_gpu_align_matches_reference_harness(orc, scene):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libssf_ref.so not built")
    eng, key = _gpu_engine_at(scene)
    r = ref.RefEngine(scene["cam"], orc.Surfels, **dict(TUM_PARAMS, nb_supersurfels_max=20000))
    r.tps(*scene["seq"].frame(0))      # allocates the reference's images and textures
    r.set_segmentation(labels=scene["labels"], slanted=scene["slanted"])
    r.set_frame(scene["cur"])
    Ri, ti = _perturbed(scene["Rrel"], scene["trel"])
    ok_r, R_r, t_r = r.align(scene["key"], Ri, ti)
    ok_g, R_g, t_g, _ = eng.align(key, Ri, ti)
    assert ok_r and ok_g
    assert np.linalg.norm(t_g - t_r) < 1e-4 and rot_angle(R_g, R_r) < 1e-4     # north_star tolerance
    eng.close()
    r.close()
