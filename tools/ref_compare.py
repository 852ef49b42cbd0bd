#!/usr/bin/env python
"""Measure how the CPU oracle (and the CUDA product) deviate from the reference's own kernels
(oracle/_ref harness) on identical inputs.  Run on the GPU box; prints one JSON document.
TEST INFRASTRUCTURE -- not part of the product."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import TUM_PARAMS, rel_err, rot_angle  # noqa: E402
from oracle import orc, ref  # noqa: E402
from supersurfel_fusion_b200 import CamParam, SupersurfelFusion  # noqa: E402
from supersurfel_fusion_b200.synth import SyntheticSequence  # noqa: E402


def main():
    out = {}
    seq = SyntheticSequence(seed=1234)
    cam = seq.cam_param()
    p = dict(TUM_PARAMS)
    cfg = orc.default_config(cam=cam, **p)
    # ---- stage-level: one frame of segmentation by each side
    r = ref.RefEngine(cam, orc.Surfels, **p)
    o = orc.Engine(cfg)
    rgb, depth = seq.frame(0)
    rs = r.tps(rgb, depth)
    os_ = o.tps.compute(rgb, depth)
    out["tps_label_mismatch_frac_frame0"] = float((rs["labels"] != os_["labels"]).mean())
    out["tps_inlier_mismatch_frac_frame0"] = float(((rs["inliers"] > 0) != (os_["inliers"] > 0)).mean())
    sp_r, sp_o = rs["superpixels"], os_["superpixels"]
    out["tps_superpixel_mean_xy_maxabs"] = float(np.nanmax(np.abs(sp_r[:, :2] - sp_o[:, :2])))
    # ---- extraction on IDENTICAL segmentation (the oracle's), reference kernels vs oracle
    rgba = os_["rgba"]
    r.set_segmentation(os_["labels"], os_["bound"], os_["inliers"], os_["slanted"], rgba)
    fr = r.generate(0)
    fo = orc.generate_supersurfels(orc.cam_of(cfg), o.S, rgba, os_["slanted"], os_["labels"], os_["inliers"],
                                   os_["bound"], cfg.range_min, cfg.range_max, 0)
    both = (fr.confidences > 0) & (fo.confidences > 0)
    out["extract_valid_ref"] = int((fr.confidences > 0).sum())
    out["extract_valid_oracle"] = int((fo.confidences > 0).sum())
    out["extract_conf_equal_frac"] = float((fr.confidences == fo.confidences).mean())
    out["extract_pos_relerr"] = rel_err(fr.positions[both], fo.positions[both])
    out["extract_color_relerr"] = rel_err(fr.colors[both], fo.colors[both])
    out["extract_shape_relerr"] = rel_err(fr.shapes[both], fo.shapes[both])
    out["extract_dims_relerr"] = rel_err(fr.dims[both], fo.dims[both])
    nr, no = fr.orientations[both][:, 6:9], fo.orientations[both][:, 6:9]
    ang = np.arccos(np.clip(np.abs((nr * no).sum(1)), 0, 1))
    out["extract_normal_angle_rad_max"] = float(ang.max())
    out["extract_normal_angle_rad_median"] = float(np.median(ang))
    # ---- sequence: reference vs oracle vs CUDA, full path
    r2 = ref.RefEngine(cam, orc.Surfels, **p)
    o2 = orc.Engine(cfg)
    kw = dict(p); kw["seg_use_ransac"] = bool(kw["seg_use_ransac"])
    g2 = SupersurfelFusion().initialize(CamParam(*cam), **kw)
    traj = []
    for k in range(15):
        rgb, depth = seq.frame(k)
        sr = r2.process_frame(rgb, depth)
        so = o2.process_frame(rgb, depth)
        sg = g2.processFrame(rgb, depth)
        Rr, tr = r2.pose(); Ro, to = o2.pose(); Rg, tg = g2.getPose(); Rt, tt = seq.pose(k)
        lm = float((r2.segmentation()["labels"] != o2.tps.get()["labels"]).mean())
        traj.append(dict(k=k, ref_nb=sr["nb_supersurfels"], orc_nb=so["nb_supersurfels"], gpu_nb=sg["nb_supersurfels"],
                         ref_vis=sr["nb_visible"], orc_vis=so["nb_visible"], label_mismatch=lm,
                         dt_ref_orc=float(np.linalg.norm(tr - to)), dr_ref_orc=rot_angle(Rr, Ro),
                         dt_gpu_orc=float(np.linalg.norm(tg - to)), dt_ref_gt=float(np.linalg.norm(tr - tt)),
                         dt_orc_gt=float(np.linalg.norm(to - tt)), ref_ms=sr["ms_total"], ref_wall_ms=sr["wall_ms"],
                         gpu_ms=sg["gpu_ms"], ref_icp_ok=sr["icp_valid"], orc_icp_ok=so["icp_valid"]))
    out["sequence"] = traj
    # ---- ICP on identical state (the oracle's), reference vs oracle vs CUDA
    seg = o2.tps.get()
    frame, model = o2.frame(), o2.model()
    nb, nv = o2.last["nb_supersurfels"], o2.last["nb_visible"]
    r3 = ref.RefEngine(cam, orc.Surfels, **p)
    r3.tps(*seq.frame(0))     # allocate its images / textures
    r3.set_segmentation(seg["labels"], seg["bound"], seg["inliers"], seg["slanted"], seg["rgba"])
    r3.set_frame(frame)
    mcap = orc.Surfels(cfg.nb_supersurfels_max)
    for name, _, _ in orc.Surfels.FIELDS:
        getattr(mcap, name)[:nb] = getattr(model, name)[:nb]
    r3.set_model(mcap, nb, nv)
    R, t = o2.pose()
    Rv = R.T.copy(); tv = -(Rv @ t)
    s_ref = r3.icp_system(Rv, tv, nv)
    s_orc = orc.icp_system(orc.cam_of(cfg), model.positions[:nv], model.colors[:nv], model.orientations[:nv],
                           frame.colors, frame.orientations, frame.confidences, Rv, tv, seg["labels"], seg["slanted"])
    out["icp_system_inliers_ref_orc"] = [float(s_ref[28]), float(s_orc[28])]
    out["icp_system_relerr_ref_orc"] = rel_err(s_ref, s_orc)
    dR = np.array([[1, -0.006, 0.004], [0.006, 1, -0.005], [-0.004, 0.005, 1]], np.float64)
    u, _, vt = np.linalg.svd(R.astype(np.float64) @ dR)
    Rp = (u @ vt).astype(np.float32); tp = (t + np.array([0.01, -0.008, 0.006], np.float32)).astype(np.float32)
    Rv2 = Rp.T.copy(); tv2 = -(Rv2 @ tp)
    ok_r, Rr, tr = r3.icp(Rv2, tv2)
    ok_o, Ro, to, st = orc.icp(orc.cam_of(cfg), model.positions[:nv], model.colors[:nv], model.orientations[:nv],
                               frame.colors, frame.orientations, frame.confidences, Rv2, tv2, seg["labels"],
                               seg["slanted"], nb_iter=10, cov_thresh=cfg.icp_cov_thresh)
    out["icp_loop_valid_ref_orc"] = [ok_r, ok_o]
    out["icp_loop_dt_ref_orc"] = float(np.linalg.norm(tr - to))
    out["icp_loop_dr_ref_orc"] = rot_angle(Rr, Ro)
    out["icp_loop_orc_iters"] = st["iters"]
    # ---- fusion on identical state: reference vs oracle
    rgb, depth = seq.frame(15)
    tps = o2.tps
    seg15 = tps.compute(rgb, depth)
    f15 = orc.generate_supersurfels(orc.cam_of(cfg), o2.S, seg15["rgba"], seg15["slanted"], seg15["labels"],
                                    seg15["inliers"], seg15["bound"], cfg.range_min, cfg.range_max, 15)
    r3.set_segmentation(seg15["labels"], seg15["bound"], seg15["inliers"], seg15["slanted"], seg15["rgba"])
    r3.set_frame(f15)
    r3.set_model(mcap, nb, nv)
    r3.set_pose(R, t)
    c_ref = r3.fuse(15)
    m_o = orc.Surfels(cfg.nb_supersurfels_max)
    for name, _, _ in orc.Surfels.FIELDS:
        getattr(m_o, name)[:nb] = getattr(model, name)[:nb]
    c_orc = orc.fuse(orc.cam_of(cfg), f15, m_o, cfg.nb_supersurfels_max, R, t, seg15["labels"], seg15["slanted"],
                     cfg.range_min, cfg.range_max, 15, cfg.delta_t, cfg.conf_thresh, nb, nv)
    out["fuse_counts_ref"] = list(c_ref[:3])
    out["fuse_counts_orc"] = [c_orc["nb_supersurfels"], c_orc["nb_visible"], c_orc["nb_removed"]]
    if c_ref[0] == c_orc["nb_supersurfels"]:
        n = c_ref[0]
        mr = r3.model(n)
        def key(m):
            return np.lexsort((np.round(m.positions[:n, 2], 3), np.round(m.positions[:n, 1], 3),
                               np.round(m.positions[:n, 0], 3), m.stamps[:n, 0]))
        ir, io = key(mr), key(m_o)
        out["fuse_model_pos_relerr_sorted"] = rel_err(mr.positions[ir], m_o.positions[:n][io])
        out["fuse_model_conf_equal_frac"] = float((mr.confidences[ir] == m_o.confidences[:n][io]).mean())
        out["fuse_model_color_relerr_sorted"] = rel_err(mr.colors[ir], m_o.colors[:n][io])
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
