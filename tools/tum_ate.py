#!/usr/bin/env python
"""Accuracy on a real sequence: TUM freiburg1_xyz (the data the reference bundles under rgbd_benchmark/) through the hot
path alone -- TUM launch parameters, in-library bilateral filter, no VO prior, no MOD -- and the absolute trajectory
error against the ground truth, next to the ATE of the authors' own bundled trajectory (full system: ORB VO prior +
MOD + YOLO), SURVEY.md section 6.

  python tools/tum_ate.py --engine oracle --save-traj tests/golden/tum_fr1_xyz_oracle_traj.npz     (development container)
  python tools/tum_ate.py --engine gpu --seq gpurun_in/tum_fr1_xyz --compare-traj tests/golden/tum_fr1_xyz_oracle_traj.npz
                                                                                                  (GPU box, under gpurun)
--engine oracle runs the CPU oracle and can save its per-frame poses; --engine gpu runs the CUDA path (libssf through
the Python mirror: ssf_process_frame_depth16 with SSF_FLAG_BILATERAL) and reports, besides the ATE, the per-frame
deviation from a saved oracle trajectory.  The sequence directory needs rgb/, depth/ and associations_with_gt.txt.
"""
import argparse
import json
import os
import sys
import time

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import TUM_PARAMS, rot_angle  # noqa: E402

DEFAULT_SEQ = "/root/reference/rgbd_benchmark/rgbd_dataset_freiburg1_xyz"
CAM = (525.0, 525.0, 319.5, 239.5, 480, 640)
DEPTH_SCALE = 0.0002          # 1 / 5000: TUM depth PNGs


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
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="oracle", choices=["oracle", "gpu"])
    ap.add_argument("--seq", default=DEFAULT_SEQ)
    ap.add_argument("--frames", type=int, default=10 ** 9)
    ap.add_argument("--ingest", default="library", choices=["library", "oracle"],
                    help="gpu engine only: 'library' = 16-bit depth + the CUDA bilateral filter (what a node would call); "
                         "'oracle' = the oracle's CPU bilateral restatement feeds the CUDA path, i.e. both engines see "
                         "IDENTICAL inputs and the trajectory difference is the path's alone")
    ap.add_argument("--save-traj", default=None, help="write the per-frame poses (npz: R [n,3,3], t [n,3], valid [n])")
    ap.add_argument("--compare-traj", default=None, help="per-frame deviation from a trajectory saved with --save-traj")
    ap.add_argument("--out", default=None, help="JSON result file (default profiles/tum_fr1_xyz_ate[_gpu].json)")
    args = ap.parse_args()
    seq = args.seq
    lines = [l.split() for l in open(os.path.join(seq, "associations_with_gt.txt"))][:args.frames]
    params = dict(TUM_PARAMS)
    if args.engine == "oracle":
        from oracle import orc
        eng = orc.Engine(orc.default_config(cam=CAM, **params))
    else:
        from supersurfel_fusion_b200 import CamParam, SupersurfelFusion
        from supersurfel_fusion_b200.engine import SSF_FLAG_BILATERAL
        params["seg_use_ransac"] = True
        eng = SupersurfelFusion().initialize(CamParam(*CAM), **params)
    Rs, ts, gt, valid = [], [], [], []
    gpu_ms = 0.0
    t_start = time.time()
    for k, w in enumerate(lines):
        rgb = np.ascontiguousarray(cv2.imread(os.path.join(seq, w[1]), cv2.IMREAD_COLOR)[:, :, ::-1])
        d16 = cv2.imread(os.path.join(seq, w[3]), cv2.IMREAD_UNCHANGED)
        if args.engine == "oracle":
            depth = orc.bilateral_filter(orc.depth16_to_metres(d16, DEPTH_SCALE))
            st = eng.process_frame(rgb, depth)
            R, t = eng.pose()
        elif args.ingest == "oracle":
            from oracle import orc
            st = eng.processFrame(rgb, orc.bilateral_filter(orc.depth16_to_metres(d16, DEPTH_SCALE)))
            R, t = eng.getPose()
            gpu_ms += st["gpu_ms"]
        else:
            st = eng.processFrameDepth16(rgb, d16, DEPTH_SCALE, flags=SSF_FLAG_BILATERAL)
            R, t = eng.getPose()
            gpu_ms += st["gpu_ms"]
        valid.append(int(st["icp_valid"]))
        Rs.append(np.array(R, np.float32).copy())
        ts.append(np.array(t, np.float32).copy())
        gt.append([float(v) for v in w[5:8]])
        if k % 100 == 0:
            print("frame", k, "model", st["nb_supersurfels"], "icp_valid", st["icp_valid"], flush=True)
    rmse, mean, mx = horn_ate(ts, gt)
    out = {"sequence": "freiburg1_xyz", "frames": len(ts), "icp_valid_frames": int(sum(valid)), "ate_rmse_m": rmse,
           "ate_mean_m": mean, "ate_max_m": mx,
           "engine": "CPU oracle" if args.engine == "oracle" else
                     ("CUDA path (libssf, ssf_process_frame_depth16 + SSF_FLAG_BILATERAL)" if args.ingest == "library" else
                      "CUDA path (libssf, ssf_process_frame) on the oracle's filtered depth: identical inputs to the oracle run"),
           "setup": "TUM launch parameters, in-library bilateral filter, pose prior = previous fused pose (no VO), no MOD",
           "wall_s": time.time() - t_start}
    if args.engine == "gpu":
        out["gpu_ms_per_frame"] = gpu_ms / max(len(ts), 1)
    if args.save_traj:
        np.savez_compressed(args.save_traj, R=np.array(Rs), t=np.array(ts), valid=np.array(valid, np.int8),
                            stamps=np.array([w[0] for w in lines]))
    if args.compare_traj:
        ref = np.load(args.compare_traj)
        n = min(len(ts), len(ref["t"]))
        dt = np.linalg.norm(np.array(ts[:n]) - ref["t"][:n], axis=1)
        dr = np.array([rot_angle(Rs[k], ref["R"][k]) for k in range(n)])
        same_valid = int((np.array(valid[:n]) == ref["valid"][:n]).sum())
        out["vs_oracle_trajectory"] = {
            "frames": n, "max_dt_m": float(dt.max()), "median_dt_m": float(np.median(dt)), "p99_dt_m": float(np.percentile(dt, 99)),
            "max_dR_rad": float(dr.max()), "frames_with_same_icp_validity": same_valid,
            "first_frame_over_1e-4_m": int(np.argmax(dt > 1e-4)) if (dt > 1e-4).any() else None,
            "note": "per-frame pose of this engine vs the CPU oracle's saved trajectory on the same real frames; with --ingest "
                    "library the two ingest paths differ (oracle bilateral restatement vs the CUDA bilateral kernel, 2e-6 m "
                    "on the depth, enough to flip label decisions), with --ingest oracle the inputs are identical"}
    ref_path = os.path.join(seq, "estimated.txt")
    if os.path.exists(ref_path):
        ref = {l.split()[0]: [float(v) for v in l.split()[1:4]] for l in open(ref_path) if l.strip() and l[0] != "#"}
        pairs = [(ref[w[0]], [float(v) for v in w[5:8]]) for w in lines if w[0] in ref]
        if len(pairs) > 10:
            r = horn_ate([p[0] for p in pairs], [p[1] for p in pairs])
            out["authors_bundled_trajectory"] = {"frames": len(pairs), "ate_rmse_m": r[0], "ate_mean_m": r[1], "ate_max_m": r[2],
                                                 "note": "full system: ORB VO prior + MOD + YOLO"}
    path = args.out or os.path.join(ROOT, "profiles", "tum_fr1_xyz_ate%s.json" % ("" if args.engine == "oracle" else "_gpu"))
    out["ingest"] = args.ingest if args.engine == "gpu" else "oracle"
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
