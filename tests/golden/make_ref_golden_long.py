#!/usr/bin/env python
"""Generates tests/golden/ref_golden_long_320x240.npz: the state of the REFERENCE'S OWN kernels (oracle/_ref harness =
reference sources compiled unmodified for sm_100a) after 27 frames, and what its model update
(findBestMatches / updateSupersurfels / insertSupersurfels / filterModel / sort_by_key,
core/src/supersurfel_fusion.cu:351-483) does to that state on frame 27 -- late enough for the age-based removal
`time_diff > delta_t && conf < conf_thresh && stamp > delta_t` (core/src/supersurfel_fusion_kernels.cu:429) with the
launch file's delta_t = 20 to fire.  Must run on a machine with a GPU:

    gpurun -- 'python tests/golden/make_ref_golden_long.py gpurun_out/ref_golden_long_320x240.npz'

then copy the file into tests/golden/.  tests/test_oracle_golden.py feeds the same inputs to the CPU oracle's fusion
and compares which supersurfels die and why."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import TUM_PARAMS  # noqa: E402
from oracle import orc, ref  # noqa: E402
from supersurfel_fusion_b200.synth import SyntheticSequence  # noqa: E402

N_BEFORE = 27       # frames 0..26 through the reference's own path; frame 27 is the recorded update


def main(path):
    seq = SyntheticSequence(width=320, height=240, seed=77)
    cam = seq.cam_param()
    p = dict(TUM_PARAMS, nb_supersurfels_max=8000, icp_cov_thresh=5.0)   # delta_t = 20 as in the TUM launch file
    out = dict(cam=np.array(cam, np.float64), params_json=np.array(repr(sorted(p.items()))), n_before=np.int32(N_BEFORE))
    r = ref.RefEngine(cam, orc.Surfels, **p)
    removed_per_frame = []
    for k in range(N_BEFORE):
        st = r.process_frame(*seq.frame(k))
        removed_per_frame.append(int(r.counts()[2]))
    nb, nv, _, stamp = r.counts()
    model = r.model(nb)
    R, t = r.pose()
    rgb, depth = seq.frame(N_BEFORE)
    seg = r.tps(rgb, depth)
    frame = r.generate(N_BEFORE)
    out.update(removed_per_frame=np.array(removed_per_frame, np.int32), stamp=np.int32(stamp), seg_labels=seg["labels"],
               seg_slanted=seg["slanted"], nb=np.int32(nb), nv=np.int32(nv), pose_R=R, pose_t=t)
    for name, _, _ in orc.Surfels.FIELDS:
        out["frame_" + name] = getattr(frame, name)
        out["model_" + name] = getattr(model, name)
    c = r.fuse(N_BEFORE)
    out.update(fuse_counts=np.array(c[:3], np.int32))
    m2 = r.model(c[0])
    for name, _, _ in orc.Surfels.FIELDS:
        out["fused_" + name] = getattr(m2, name)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; model", nb, "visible", nv, "-> counts", c[:3],
          "removed per frame", removed_per_frame)


if __name__ == "__main__":
    main(sys.argv[1])
