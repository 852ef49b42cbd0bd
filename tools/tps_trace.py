#!/usr/bin/env python
"""Profiling aid: per-CTA phase timeline of the persistent TPS kernel (SSF_TPS_TRACE)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["SSF_TPS_TRACE"] = "/tmp/tps_trace.bin"
from supersurfel_fusion_b200 import CamParam, SupersurfelFusion
from supersurfel_fusion_b200.synth import SyntheticSequence
from conftest import TUM_PARAMS
seq = SyntheticSequence(seed=1234)
p = dict(TUM_PARAMS); p["seg_use_ransac"] = True
eng = SupersurfelFusion().initialize(CamParam(*seq.cam_param()), **p)
rgb, d = seq.frame(0)
for _ in range(3):
    eng.tpsSegment(rgb, d)
t = np.fromfile("/tmp/tps_trace.bin", dtype=np.uint64).reshape(-1, 512).astype(np.int64)
active = t[:, 0] > 0
t = t[active]
n = int((t[0] > 0).sum())
t0 = t[:, 0].min()
t = t[:, :n] - t0
print("ctas", len(t), "stamps", n, "total us", t[:, n - 1].max() / 1e3)
def seg(name, a, b):
    print("%-28s mean %7.2f us   max %7.2f us" % (name, (t[:, b] - t[:, a]).mean() / 1e3, (t[:, b] - t[:, a]).max() / 1e3))
npass_rgb = 20
for k in (0, 1, 10, 19):
    o = 4 * k
    print("RGB pass", k, "start skew us", (t[:, o].max() - t[:, o].min()) / 1e3)
    seg("  cache fill", o, o + 1); seg("  items", o + 1, o + 2); seg("  barrier", o + 2, o + 3)
o = 4 * npass_rgb
print("after RGB (+merge)", t[:, o].mean() / 1e3, "RANSAC+init_disp+merge until", t[:, o + 1].mean() / 1e3)
for k in (0, 10, 19):
    q = o + 2 + 4 * k
    print("RGBD pass", k)
    seg("  cache fill", q, q + 1); seg("  items", q + 1, q + 2); seg("  barrier", q + 2, q + 3)
q = o + 2 + 80
print("RGBD end", t[:, q - 1].mean() / 1e3, "filter done", t[:, q].mean() / 1e3, "render done", t[:, q + 1].mean() / 1e3)
per_pass = (t[:, 4:4 * npass_rgb:4] - t[:, 0:4 * npass_rgb - 4:4]).mean()
print("mean RGB pass period us", per_pass / 1e3)
