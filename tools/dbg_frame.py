import sys, os
sys.path.insert(0, "tests")
from conftest import TUM_PARAMS
from supersurfel_fusion_b200 import CamParam, SupersurfelFusion
from supersurfel_fusion_b200.synth import SyntheticSequence
seq = SyntheticSequence(width=320, height=240, seed=77)
eng = SupersurfelFusion().initialize(CamParam(*seq.cam_param()), **dict(TUM_PARAMS, seg_use_ransac=True, nb_supersurfels_max=20000))
for k in range(2):
    print(k, eng.processFrame(*seq.frame(k))["nb_supersurfels"], flush=True)
print("ok")
