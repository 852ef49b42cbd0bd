#!/usr/bin/env python
"""Profiling aid (run under gpurun): where the cycles of one fused relabelling pass go.  Thread 0 of every CTA of
tps_pass_tile_kernel stamps clock64() at its phase boundaries (SSF_TPS_TRACE); this prints the mean / max phase lengths
over the CTAs of the LAST pass of a 640x480 segmentation (a colour + disparity pass)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["SSF_TPS_TRACE"] = "/tmp/tps_pass_trace.bin"
from supersurfel_fusion_b200 import CamParam, SupersurfelFusion
from supersurfel_fusion_b200.synth import SyntheticSequence
from conftest import TUM_PARAMS
seq = SyntheticSequence(seed=1234)
eng = SupersurfelFusion().initialize(CamParam(*seq.cam_param()), **dict(TUM_PARAMS, seg_use_ransac=True))
rgb, d = seq.frame(0)
for _ in range(3):
    eng.tpsSegment(rgb, d)
t = np.fromfile("/tmp/tps_pass_trace.bin", dtype=np.int64)[:330 * 8].reshape(330, 8)
t = t[t[:, 0] != 0]
names = ["geometry + issue of all loads", "means / planes (waits for the sums)", "barrier + TMA wait + patch", "decide", "apply (atomics issued)", "buffer rotation"]
print("CTAs", len(t), "clock 1.965 GHz")
for k, n in enumerate(names):
    dcy = t[:, k + 1] - t[:, k]
    print("%-40s mean %7.0f cycles (%5.2f us)   max %7.0f" % (n, dcy.mean(), dcy.mean() / 1965.0, dcy.max()))
if t[:, 7].any():      # stamp between the CTA barrier and the wait for the label tile
    print("%-40s mean %7.0f cycles" % ("  of which: CTA barrier after the means", (t[:, 7] - t[:, 2]).mean()))
    print("%-40s mean %7.0f cycles" % ("  of which: wait for the label tile + patch", (t[:, 3] - t[:, 7]).mean()))
tot = t[:, 6] - t[:, 0]
print("%-40s mean %7.0f cycles (%5.2f us)   max %7.0f" % ("thread 0, entry to exit", tot.mean(), tot.mean() / 1965.0, tot.max()))

# the pass ends with its slowest CTA: where do the slow ones lose their time?
order = np.argsort(tot)
slow = order[-max(1, len(t) // 10):]
fast = order[:len(t) // 2]
print("lifetime percentiles (cycles): p50 %.0f  p90 %.0f  p99 %.0f  max %.0f" % tuple(np.percentile(tot, [50, 90, 99, 100])))
print("%-40s %10s %10s" % ("phase (mean cycles)", "fast half", "slowest 10%"))
for k, n in enumerate(names):
    dcy = t[:, k + 1] - t[:, k]
    print("%-40s %10.0f %10.0f" % (n, dcy[fast].mean(), dcy[slow].mean()))
