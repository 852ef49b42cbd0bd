"""CPU tests of the N>1 paths with world_size 2 on the gloo backend: shard ranges, the
rank-ordered 29-float reduction (bit-identical on every rank), the tile-parallel ICP loop, and
the max-over-ranks / aggregate-throughput arithmetic bench.py uses."""
import os
import socket
import subprocess
import sys

import numpy as np

from supersurfel_fusion_b200 import multi
from supersurfel_fusion_b200.synth import synthetic_icp_problem

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_ranges_tile_the_prefix():
    for n in (0, 1, 5, 1200, 4097, 100000, 16 * 1024 * 1024):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(world):
                b, c = multi.shard_range(n, r, world)
                assert b % 4 == 0 and c >= 0
                assert c == 0 or b == covered
                covered += c
            assert covered == n


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_tile_parallel_icp_world2_gloo(orc, tmp_path):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "multi_worker.py"), str(r), "2", str(port), str(tmp_path)])
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=240) == 0
    r0 = np.load(os.path.join(str(tmp_path), "rank0.npz"))
    r1 = np.load(os.path.join(str(tmp_path), "rank1.npz"))
    # both ranks hold the same bits after the rank-ordered reduction, in every iteration
    assert r0["systems"].shape == (2, 29)
    assert np.array_equal(r0["systems"].view(np.uint32), r1["systems"].view(np.uint32))
    assert np.array_equal(r0["Rr"], r1["Rr"])
    # shards tile the source range
    assert r0["shard"][0] == 0 and r1["shard"][0] == r0["shard"][1] and r0["shard"][1] + r1["shard"][1] == 20011
    # the reduced first-iteration system equals the single-process system
    prob = synthetic_icp_problem(20011, width=320, height=240, seed=3)
    whole = orc.icp_system(orc.OrcCam(*prob["cam"]), prob["src_pos"], prob["src_col"], prob["src_ori"],
                           prob["tgt_col"], prob["tgt_ori"], prob["tgt_conf"], np.eye(3, dtype=np.float32),
                           np.array([0.002, -0.001, 0.003], np.float32), prob["labels"], prob["depth"])
    assert r0["systems"][0][28] == whole[28] > 5000
    assert np.abs(r0["systems"][0] - whole).max() <= 1e-5 * np.abs(whole).max()
    # bench arithmetic: max over ranks, whole-job aggregate, per-rank seeds
    assert float(r0["max_ms"]) == float(r1["max_ms"]) == 11.0
    assert abs(float(r0["fps"]) - 200 / 0.011) < 1e-6
    assert (int(r0["seed"]), int(r1["seed"])) == (1234, 1235)
