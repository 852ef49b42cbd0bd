#!/usr/bin/env python
"""Generates tests/golden/align_golden_640x480.npz: inputs and the output of the REFERENCE'S OWN
DenseRegistration::align (core/src/dense_registration.cu:52-243, compiled unmodified into
oracle/_ref) for a keyframe -> current-frame registration.  Must run on a machine with a GPU:

    gpurun -- 'python tests/golden/make_align_golden.py gpurun_out/align_golden_640x480.npz'

then copy the file into tests/golden/.  tests/test_align.py checks the CPU oracle against it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import TUM_PARAMS  # noqa: E402
from oracle import orc, ref  # noqa: E402
from supersurfel_fusion_b200.synth import SyntheticSequence  # noqa: E402


def perturbed(Rrel, trel):
    dR = np.array([[1, -0.004, 0.002], [0.004, 1, -0.003], [-0.002, 0.003, 1]], np.float64)
    u, _, vt = np.linalg.svd(Rrel.astype(np.float64) @ dR)
    return (u @ vt).astype(np.float32), (trel + np.array([0.006, -0.004, 0.005], np.float32)).astype(np.float32)


def main(path):
    seq = SyntheticSequence(width=640, height=480, seed=77)
    cam = seq.cam_param()
    p = dict(TUM_PARAMS, nb_supersurfels_max=20000)
    r = ref.RefEngine(cam, orc.Surfels, **p)
    for k in range(3):
        r.process_frame(*seq.frame(k))
    key = r.frame()                      # keyframe supersurfels, camera frame of frame 2
    Rk, tk = r.pose()
    for k in range(3, 6):
        r.process_frame(*seq.frame(k))
    cur = r.frame()
    Rc, tc = r.pose()
    seg = r.segmentation()
    Rrel = Rc.T @ Rk
    trel = Rc.T @ (tk - tc)
    Ri, ti = perturbed(Rrel, trel)
    ok, R, t = r.align(key, Ri, ti)
    out = dict(cam=np.array(cam, np.float64), R_init=Ri, t_init=ti, R_true=Rrel, t_true=trel, valid=np.int32(ok), R=R, t=t,
               labels=seg["labels"], slanted=seg["slanted"], icp_iter=np.int32(p["icp_iter"]),
               cov_thresh=np.float64(p["icp_cov_thresh"]))
    for name in ("positions", "colors", "orientations", "confidences"):
        out["key_" + name] = getattr(key, name)
        out["cur_" + name] = getattr(cur, name)
    np.savez_compressed(path, **out)
    print("wrote", path, "valid", ok, "t", t)


if __name__ == "__main__":
    main(sys.argv[1])
