#!/usr/bin/env python
"""BASELINE configs[4]: one 2560x1920 frame, frame-to-model ICP tiled across the ranks of a
torchrun job (one process per GPU), 29-float rank-ordered reduction per iteration.
Checks the result against the single-GPU loop on the same inputs and reports latencies.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tools/tile_icp_check.py [n_src]
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from supersurfel_fusion_b200 import CamParam, SupersurfelFusion, Supersurfels, multi  # noqa: E402
from supersurfel_fusion_b200.engine import SsfSurfels, _ptr  # noqa: E402
from supersurfel_fusion_b200.synth import synthetic_icp_problem  # noqa: E402


def note(msg):
    if os.environ.get("SSF_VERBOSE"):
        with open(os.path.join(ROOT, "gpurun_out", "tile_dbg_rank%s.log" % os.environ.get("RANK", "0")), "a") as f:
            f.write("%.3f %s\n" % (time.time(), msg))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2 * 1024 * 1024
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    prob = synthetic_icp_problem(n, width=2560, height=1920, seed=1234)
    eng = SupersurfelFusion(local).initialize(CamParam(*prob["cam"]), nb_supersurfels_max=n)
    frame = Supersurfels(prob["S"])
    frame.colors[:] = prob["tgt_col"]; frame.orientations[:] = prob["tgt_ori"]; frame.confidences[:] = prob["tgt_conf"]
    eng.setSegmentation(labels=prob["labels"], slanted=prob["depth"])
    eng.setFrame(frame)
    eng.setModelPointers(SsfSurfels(_ptr(prob["src_pos"]), _ptr(prob["src_col"]), None, _ptr(prob["src_ori"]), None,
                                    None, None), n, n)
    R = np.eye(3, dtype=np.float32)
    t = np.array([0.004, -0.003, 0.005], np.float32)
    # single GPU, whole loop on the device
    note("engine ready")
    ok1, R1, t1, st1 = eng.icp(R, t)
    torch.cuda.synchronize()
    note("single done iters=%d" % st1["iters"])
    t0 = time.perf_counter()
    for _ in range(5):
        eng.icp(R, t)
    one_ms = (time.perf_counter() - t0) / 5 * 1e3
    # tiled across the ranks
    okN, RN, tN, stN = multi.tile_parallel_icp(eng, dist if world > 1 else None, n, R, t, device=dev)
    note("nccl-tiled done iters=%d" % stN["iters"])
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        multi.tile_parallel_icp(eng, dist if world > 1 else None, n, R, t, device=dev)
    tiled_ms = (time.perf_counter() - t0) / 5 * 1e3
    tiled_ms = multi.allreduce_max(dist if world > 1 else None, tiled_ms, device=dev)
    fused = {}
    if world > 1:
        multi.connect_peers(eng, dist, device=dev)
        note("peers connected")
        okF, RF, tF, stF = multi.fused_tile_parallel_icp(eng, dist, n, R, t)
        note("fused done iters=%d" % stF["iters"])
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            multi.fused_tile_parallel_icp(eng, dist, n, R, t)
        fused_ms = multi.allreduce_max(dist, (time.perf_counter() - t0) / 5 * 1e3, device=dev)
        fused = dict(valid_fused=okF, iters_fused=stF["iters"], dt_fused=float(np.linalg.norm(t1 - tF)),
                     R_fused=[float(v) for v in RF.reshape(9)], t_fused=[float(v) for v in tF],
                     dR_fused=float(np.abs(R1 - RF).max()), fused_ms=fused_ms,
                     fused_equals_nccl_bits=bool(np.array_equal(tF, tN) and np.array_equal(RF, RN)))
    out = dict(world=world, n_src=n, valid_single=ok1, valid_tiled=okN, iters_single=st1["iters"],
               iters_tiled=stN["iters"], dt=float(np.linalg.norm(t1 - tN)), dR=float(np.abs(R1 - RN).max()),
               sys_rel=float(np.abs(st1["system"] - stN["system"]).max() / np.abs(st1["system"]).max()),
               inliers_single=float(st1["system"][28]), inliers_tiled=float(stN["system"][28]),
               single_gpu_ms=one_ms, tiled_ms=tiled_ms, shard=list(stN["shard"]),
               R_single=[float(v) for v in R1.reshape(9)], t_single=[float(v) for v in t1],
               R_tiled=[float(v) for v in RN.reshape(9)], t_tiled=[float(v) for v in tN],
               R_init=[float(v) for v in R.reshape(9)], t_init=[float(v) for v in t], **fused)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
