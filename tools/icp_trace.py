"""Phase lengths of the ICP system kernel at the roofline sizing (run under gpurun with a trace build:
tools/build_variant.sh trace -DSSF_ICP_TRACE -DSSF_ICP_LD_PLAIN; SSF_LIB=$PWD/variants/libssf_trace.so).
The kernel sums clock64() stamps per phase over all threads (csrc/ssf_icp.cu, SSF_ICP_TRACE); this prints the
mean cycles per loop iteration between consecutive stamps."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from supersurfel_fusion_b200 import CamParam, SupersurfelFusion, Supersurfels  # noqa: E402
from supersurfel_fusion_b200.engine import SsfSurfels, _ptr, lib_path  # noqa: E402
from supersurfel_fusion_b200.synth import synthetic_icp_problem  # noqa: E402

n = int(os.environ.get("ICP_TRACE_N", 16 * 1024 * 1024))
prob = synthetic_icp_problem(n, width=2560, height=1920, seed=1234)
eng = SupersurfelFusion(0).initialize(CamParam(*prob["cam"]), nb_supersurfels_max=n)
frame = Supersurfels(prob["S"])
frame.colors[:] = prob["tgt_col"]; frame.orientations[:] = prob["tgt_ori"]; frame.confidences[:] = prob["tgt_conf"]
eng.setSegmentation(labels=prob["labels"], slanted=prob["depth"])
eng.setFrame(frame)
eng.setModelPointers(SsfSurfels(_ptr(prob["src_pos"]), _ptr(prob["src_col"]), None, _ptr(prob["src_ori"]), None, None, None), n, n)
R = np.eye(3, dtype=np.float32)
t = np.array([0.002, -0.001, 0.003], np.float32)
eng.icpSystem(R, t, n)
for _ in range(3):
    eng.icpSystemEnqueue(R, t, n, 1)
eng.synchronize()
lib = ctypes.CDLL(lib_path())
out = (ctypes.c_ulonglong * 8)()
assert lib.ssf_debug_icp_trace(out, 1) == 0
L = 10
eng.timerStart()
eng.icpSystemEnqueue(R, t, n, L)
ms = eng.timerStop() / L
assert lib.ssf_debug_icp_trace(out, 1) == 0
v = [int(x) for x in out]
iters = v[5]                      # thread-iterations
names = ["streams issued -> arrived", "-> texels of the second pair arrived", "-> frame records of the second pair arrived",
         "-> both pairs accumulated"]
res = {"us_per_launch": ms * 1e3, "thread_iterations_per_launch": iters / L, "cycles": {}}
for k, name in enumerate(names):
    res["cycles"][name] = ((v[k + 1] - v[k]) % 2**64) / iters   # the sums wrap; differences do not
res["cycles"]["whole iteration (top -> accumulated)"] = ((v[4] - v[0]) % 2**64) / iters
print(json.dumps(res, indent=1))
eng.close()
