#!/usr/bin/env python
"""Generates tests/golden/ref_golden_320x240.npz: inputs and the outputs of the REFERENCE'S OWN
kernels (oracle/_ref harness = reference sources compiled unmodified for sm_100a) for the
stages of the hot path, at 320x240.  Must run on a machine with a GPU:

    gpurun -- 'python tests/golden/make_ref_golden.py gpurun_out/ref_golden_320x240.npz'

then copy the file into tests/golden/.  tests/test_oracle_golden.py checks the CPU oracle
against these vectors without a GPU.  (The reference itself ships no golden vectors.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import TUM_PARAMS  # noqa: E402
from oracle import orc, ref  # noqa: E402
from supersurfel_fusion_b200.synth import SyntheticSequence  # noqa: E402


def main(path):
    seq = SyntheticSequence(width=320, height=240, seed=77)
    cam = seq.cam_param()
    p = dict(TUM_PARAMS, nb_supersurfels_max=4000, icp_cov_thresh=5.0)   # 300 superpixels: the 0.05 gate would reject every solve
    out = dict(cam=np.array(cam, np.float64), params_json=np.array(repr(sorted(p.items()))))
    r = ref.RefEngine(cam, orc.Surfels, **p)
    # five frames through the reference's own path; keep the state the 5th frame sees
    for k in range(4):
        r.process_frame(*seq.frame(k))
    nb, nv, _, stamp = r.counts()
    model = r.model(nb)
    R, t = r.pose()
    rgb, depth = seq.frame(4)
    seg = r.tps(rgb, depth)
    # rgba is not exposed by the harness getter: rebuild it (3 -> 4 channel copy)
    rgba = np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], -1)
    frame = r.generate(4)
    out.update(rgb=rgb, depth=depth, seg_labels=seg["labels"], seg_bound=seg["bound"], seg_inliers=seg["inliers"],
               seg_slanted=seg["slanted"], seg_superpixels=seg["superpixels"], rgba=rgba)
    for name, _, _ in orc.Surfels.FIELDS:
        out["frame_" + name] = getattr(frame, name)
        out["model_" + name] = getattr(model, name)
    out.update(nb=np.int32(nb), nv=np.int32(nv), pose_R=R, pose_t=t)
    Rv = R.T.copy()
    tv = -(Rv @ t)
    out.update(icp_Rv=Rv, icp_tv=tv, icp_system=r.icp_system(Rv, tv, nv))
    dR = np.array([[1, -0.005, 0.003], [0.005, 1, -0.004], [-0.003, 0.004, 1]], np.float64)
    u, _, vt = np.linalg.svd(R.astype(np.float64) @ dR)
    Rp = (u @ vt).astype(np.float32)
    tp = (t + np.array([0.008, -0.006, 0.005], np.float32)).astype(np.float32)
    Rv2 = Rp.T.copy()
    tv2 = -(Rv2 @ tp)
    ok, Rrel, trel = r.icp(Rv2, tv2)
    out.update(icp2_Rv=Rv2, icp2_tv=tv2, icp2_valid=np.int32(ok), icp2_Rrel=Rrel, icp2_trel=trel)
    # fusion of frame 4 at the pose above
    c = r.fuse(4)
    out.update(fuse_counts=np.array(c[:3], np.int32))
    m2 = r.model(c[0])
    for name, _, _ in orc.Surfels.FIELDS:
        out["fused_" + name] = getattr(m2, name)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1])
