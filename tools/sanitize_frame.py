#!/usr/bin/env python
"""Workload for compute-sanitizer (run under gpurun, see tools/sanitize.sh): a few 640x480 frames through the
synchronous call, the 4-stage pipelined call, the stage entry points and the loop-closure registration, so that
every kernel of the frame path runs at least once under memcheck / racecheck."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import TUM_PARAMS  # noqa: E402
from supersurfel_fusion_b200 import CamParam, SupersurfelFusion, Supersurfels  # noqa: E402
from supersurfel_fusion_b200.engine import SSF_FLAG_BILATERAL  # noqa: E402
from supersurfel_fusion_b200.synth import SyntheticSequence  # noqa: E402

seq = SyntheticSequence(seed=1234)
cam = CamParam(*seq.cam_param())
params = dict(TUM_PARAMS, seg_use_ransac=True, nb_supersurfels_max=20000)
frames = [seq.frame(k) for k in range(6)]

eng = SupersurfelFusion().initialize(cam, **params)
for rgb, depth in frames[:3]:
    st = eng.processFrame(rgb, depth, flags=SSF_FLAG_BILATERAL)
print("synchronous:", st["nb_supersurfels"], st["icp_valid"])
kf = eng.getFrame()
ok, R, t, info = eng.align(kf, np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
print("align:", ok, info["pairs"])
eng.processFrameStaged(*frames[3], dynamic_mask=(np.arange(eng.nbSuperpixels) % 7 == 0).astype(np.uint8))
eng.extractLocalPointCloud()
eng.getMarkers()
eng.close()

os.environ["SSF_ICP_LOOP"] = "0"          # the multi-launch registration too
eng = SupersurfelFusion().initialize(cam, **params)
for rgb, depth in frames[:3]:
    st = eng.processFrame(rgb, depth)
print("multi-launch registration:", st["nb_supersurfels"], st["icp_valid"])
eng.close()
del os.environ["SSF_ICP_LOOP"]

pipe = SupersurfelFusion().initialize(cam, **params).prepare()
depth_p = pipe.pipelineDepth()
done = 0
for k, (rgb, depth) in enumerate(frames):
    if k >= depth_p:
        pipe.waitFrame(); done += 1
    pipe.submitFrame(rgb, depth)
while done < len(frames):
    st, R, t = pipe.waitFrame(); done += 1
print("pipelined (%d stages):" % depth_p, st["nb_supersurfels"], st["icp_valid"])
pipe.close()
print("sanitize workload done")
