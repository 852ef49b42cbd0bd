#!/usr/bin/env python
"""Throughput of the pipelined mode (ssf_submit_frame / ssf_wait_frame) against the synchronous one,
VGA synthetic sequence, inputs resident on the device and from pinned host memory."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from supersurfel_fusion_b200 import CamParam, SupersurfelFusion
seq, frames = bench.render_frames(1234)
cam = seq.cam_param()
d = [(torch.from_numpy(f[0]).cuda(), torch.from_numpy(f[1]).cuda()) for f in frames]
hp = [(torch.from_numpy(f[0]).pin_memory(), torch.from_numpy(f[1]).pin_memory()) for f in frames]
torch.cuda.synchronize()
def run(bufs, pipelined, steps=300):
    eng = SupersurfelFusion().initialize(CamParam(*cam), **bench.PARAMS)
    for s in range(10):
        eng.processFrame(*bufs[bench.frame_index(s)])
    t0 = time.perf_counter()
    if pipelined:
        depth = eng.pipelineDepth()
        for s in range(10, 10 + steps):
            if s - 10 >= depth:
                eng.waitFrame()
            eng.submitFrame(*bufs[bench.frame_index(s)])
        for _ in range(depth):
            eng.waitFrame()
    else:
        for s in range(10, 10 + steps):
            eng.processFrame(*bufs[bench.frame_index(s)])
    dt = time.perf_counter() - t0
    st = eng.getFrameStats()
    eng.close()
    return steps / dt, st["nb_supersurfels"]
print("stages", os.environ.get("SSF_PIPELINE_STAGES", "default"))
for name, bufs in (("device", d), ("pinned-host", hp)):
    print(name, "sync", run(bufs, False), "pipelined", run(bufs, True))
