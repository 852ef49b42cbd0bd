#!/usr/bin/env python
"""Accuracy sanity on a real sequence (development container only: reads the TUM data the reference bundles
under /root/reference/rgbd_benchmark): runs the CPU oracle -- which the CUDA path matches to 5e-7 m per frame --
over freiburg1_xyz with the TUM launch parameters, no VO prior, no MOD, and reports the absolute trajectory
error against the ground truth next to the ATE of the authors' own bundled trajectory (full system:
ORB VO prior + MOD + YOLO), SURVEY.md section 6.

  python tools/tum_ate.py [n_frames] [--engine gpu]   ->  profiles/tum_fr1_xyz_ate.json
"""
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import TUM_PARAMS  # noqa: E402
from oracle import orc  # noqa: E402

SEQ = "/root/reference/rgbd_benchmark/rgbd_dataset_freiburg1_xyz"
CAM = (525.0, 525.0, 319.5, 239.5, 480, 640)


def horn_ate(est, gt):
    """RMSE of the translation error after the optimal rigid alignment (Horn), as the TUM tools compute it."""
    est, gt = np.asarray(est, np.float64), np.asarray(gt, np.float64)
    ce, cg = est.mean(0), gt.mean(0)
    H = (est - ce).T @ (gt - cg)
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1, 1, np.sign(np.linalg.det(Vt.T @ U.T))])
    R = Vt.T @ D @ U.T
    err = np.linalg.norm((R @ (est - ce).T).T + cg - gt, axis=1)
    return float(np.sqrt((err ** 2).mean())), float(err.mean()), float(err.max())


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 10 ** 9
    lines = [l.split() for l in open(os.path.join(SEQ, "associations_with_gt.txt"))][:n]
    eng = orc.Engine(orc.default_config(cam=CAM, **TUM_PARAMS))
    est, gt, valid = [], [], 0
    for k, w in enumerate(lines):
        rgb = np.ascontiguousarray(cv2.imread(os.path.join(SEQ, w[1]), cv2.IMREAD_COLOR)[:, :, ::-1])
        d16 = cv2.imread(os.path.join(SEQ, w[3]), cv2.IMREAD_UNCHANGED)
        depth = orc.bilateral_filter(orc.depth16_to_metres(d16, 0.0002))
        st = eng.process_frame(rgb, depth)
        valid += int(st["icp_valid"])
        est.append(eng.pose()[1].copy())
        gt.append([float(v) for v in w[5:8]])
        if k % 100 == 0:
            print("frame", k, "model", st["nb_supersurfels"], "icp_valid", st["icp_valid"], flush=True)
    rmse, mean, mx = horn_ate(est, gt)
    out = {"sequence": "freiburg1_xyz", "frames": len(est), "icp_valid_frames": valid, "ate_rmse_m": rmse, "ate_mean_m": mean,
           "ate_max_m": mx, "engine": "CPU oracle (the CUDA path matches it to 5e-7 m per frame)",
           "setup": "TUM launch parameters, in-library bilateral filter, pose prior = previous fused pose (no VO), no MOD"}
    ref_path = os.path.join(SEQ, "estimated.txt")
    if os.path.exists(ref_path):
        ref = {l.split()[0]: [float(v) for v in l.split()[1:4]] for l in open(ref_path) if l.strip() and l[0] != "#"}
        pairs = [(ref[w[0]], [float(v) for v in w[5:8]]) for w in lines if w[0] in ref]
        if len(pairs) > 10:
            r = horn_ate([p[0] for p in pairs], [p[1] for p in pairs])
            out["authors_bundled_trajectory"] = {"frames": len(pairs), "ate_rmse_m": r[0], "ate_mean_m": r[1], "ate_max_m": r[2],
                                                 "note": "full system: ORB VO prior + MOD + YOLO"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "tum_fr1_xyz_ate.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
