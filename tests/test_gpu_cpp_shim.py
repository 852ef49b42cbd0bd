"""The header-only C++ class of include/supersurfel_fusion.hpp (the reference's SupersurfelFusion surface over
the C-ABI) executed for real: a compiled program runs three frames through initialize / processFrame / getPose /
getModel / ... and must report what the ctypes path reports on the same frames."""
import os
import subprocess

import numpy as np
import pytest

from conftest import TUM_PARAMS
from supersurfel_fusion_b200.synth import SyntheticSequence

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shim_runs_frames_and_matches_ctypes(ssf_lib_path, tmp_path):
    from supersurfel_fusion_b200 import CamParam, SupersurfelFusion
    exe = str(tmp_path / "shim_frames")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++14", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "shim_frames.cpp"), "-o", exe, ssf_lib_path,
                           "-Wl,-rpath," + os.path.dirname(ssf_lib_path)])
    seq = SyntheticSequence(width=320, height=240, seed=77)
    cam = seq.cam_param()
    prefix = str(tmp_path / "f")
    frames = [seq.frame(k) for k in range(3)]
    for k, (rgb, depth) in enumerate(frames):
        rgb.tofile("%s_rgb_%d.bin" % (prefix, k))
        depth.tofile("%s_depth_%d.bin" % (prefix, k))
    export = str(tmp_path / "model_cpp.txt")
    out = subprocess.run([exe, str(cam[5]), str(cam[4]), repr(cam[0]), repr(cam[1]), repr(cam[2]), repr(cam[3]), "3", prefix, export],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stderr[-500:])
    rows = [l.split() for l in out.stdout.splitlines()]
    poses = {int(r[1]): np.array([float(v) for v in r[2:]], np.float32) for r in rows if r[0] == "pose"}
    stats = {int(r[1]): [int(v) for v in r[2:]] for r in rows if r[0] == "stats"}
    eng = SupersurfelFusion().initialize(CamParam(*cam), **dict(TUM_PARAMS, nb_supersurfels_max=20000, seg_use_ransac=True, conf_thresh=400.0))
    for k, (rgb, depth) in enumerate(frames):
        st = eng.processFrame(rgb, depth)
        R, t = eng.getPose()
        assert np.array_equal(poses[k], np.concatenate([R.reshape(9), t])), k       # same library, same bits
        assert stats[k] == [eng.getStamp(), st["nb_supersurfels"], st["nb_visible"], st["nb_removed"], st["icp_valid"],
                            st["icp_iters"]], k
    m = eng.getModel()
    model_row = next(r for r in rows if r[0] == "model")
    assert int(model_row[1]) == m.n > 0
    assert abs(float(model_row[2]) - float(m.positions.astype(np.float64).sum())) < 1e-3
    assert abs(float(model_row[3]) - float(m.confidences.astype(np.float64).sum())) < 1e-2
    assert int(next(r for r in rows if r[0] == "frame")[1]) == eng.nbSuperpixels
    assert int(next(r for r in rows if r[0] == "cloud")[1]) == len(eng.extractLocalPointCloud()[0])
    assert int(next(r for r in rows if r[0] == "markers")[1]) == m.n
    tum = next(l for l in out.stdout.splitlines() if l.startswith("tum "))
    assert tum[4:] + "\n" == eng.formatTumPose("1305031102.175304")
    mine = str(tmp_path / "model_py.txt")
    eng.exportModel(mine)
    assert open(export).read() == open(mine).read() and os.path.getsize(mine) > 0
