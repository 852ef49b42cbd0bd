#!/usr/bin/env python
"""bench.py -- frames/s of the supersurfel hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" is one RGB-D frame through the whole path (segmentation -> supersurfel
extraction -> frame-to-model ICP -> fusion/cull).  N=1 workload: BASELINE.json configs[1],
a TUM-fr1/desk-shaped synthetic 640x480 sequence, full track+fuse.  N>1: N independent
sequences (configs[3]), one process and one engine per GPU, no data-path collective
("weak" scaling); the only collectives are the timing barrier and the max-over-ranks.

Rank 0 prints ONE JSON line:
  value     frames/s with the frames already resident in HBM
  e2e       frames/s through the C-ABI with pinned HOST buffers: the H2D copy of the frame and
            the D2H read of the pose/stats are inside the timed region
            (both through ssf_submit_frame / ssf_wait_frame, one frame in flight per pipeline stage; the numbers
            of the synchronous ssf_process_frame, the reference's call shape, are reported
            under "synchronous"; --no-pipeline times only those)
  roofline  the ICP system kernel (the metric kernel) at HBM-bound sizing: 16 Mi source
            supersurfels vs a 2560x1920 frame, algorithmic 72 B per supersurfel
  cpu_baseline  the CPU oracle port of the same path on this box's host cores, bounded sample
  reference_gpu_kernels  the reference's own CUDA kernels (oracle/_ref harness) on this GPU
`--impl reference` times that CPU oracle port alone, one engine per host thread on all host
threads (the reference has no CPU implementation of this path and its full build needs
ROS/OpenCV-CUDA/g2o).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "configs[1]: TUM-fr1/desk-shaped synthetic 640x480 RGB-D sequence, full track+fuse"
PARAMS = dict(cell_size=16, lambda_pos=10.0, lambda_bound=1000.0, lambda_size=1000.0, lambda_disp=1e8,
              thresh_disp=1e-4, seg_iter=10, seg_use_ransac=True, nb_samples=16, filter_iter=3, filter_alpha=0.1,
              filter_beta=1.0, filter_threshold=0.05, range_min=0.2, range_max=5.0, delta_t=20, conf_thresh=2560.0,
              nb_supersurfels_max=100000, icp_iter=10, icp_cov_thresh=0.05)   # launch/supersurfel_fusion_rgbd_benchmark.launch
N_UNIQUE_FRAMES = 24
ICP_BYTES_PER_SRC = 72
ICP_ROOFLINE_N = 16 * 1024 * 1024


def frame_index(step, n=None):
    """Ping-pong over the n rendered frames so that inter-frame motion stays ~1 cm."""
    n = n or N_UNIQUE_FRAMES
    period = 2 * (n - 1)
    k = step % period
    return k if k < n else period - k


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of ONE GPU while the measurement runs (rank 0 only).  In-process
    NVML reads every 50 ms: no nvidia-smi fork per sample (on an 8-GPU node each fork takes driver
    locks for tens of ms and showed up in the round-1 scaling curve); if NVML is not importable, ONE
    looping nvidia-smi process (-lms 200, the recipe's form) is started instead."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self._proc = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu_index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    @staticmethod
    def _physical_index(logical):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if logical < len(ids) and ids[logical].isdigit():
                return int(ids[logical])
        return logical

    def run(self):
        if self._nvml is not None:
            nv = self._nvml
            while not self._halt.is_set():
                try:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                    for name, bit in self.REASONS:
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                self._halt.wait(0.05)
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self._proc = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index(self.gpu)), "--query-gpu=" + q,
                                           "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            for line in self._proc.stdout:
                f = [x.strip() for x in line.strip().split(",")]
                if len(f) < 6:
                    continue
                try:
                    self.samples.append(float(f[0]))
                    self.max_mhz = float(f[1])
                except ValueError:
                    continue
                for (name, _), v in zip(self.REASONS, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
                if self._halt.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._halt.set()
        if self._proc is not None:
            self._proc.terminate()        # the exact process started above
        self.join(timeout=5)
        # "under load": the idle samples between legs (engine creation, rendering) would pull the median down
        busy = [x for x in self.samples if self.max_mhz is None or x >= 0.5 * self.max_mhz] or self.samples
        med = float(np.median(busy)) if busy else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "how": "NVML in-process, 50 ms period, rank 0, across the whole measurement" if self._nvml is not None
                       else "one nvidia-smi -lms 200 process, rank 0"}


SIZE = [640, 480]   # --size WxH: 640x480 is the metric's configuration (configs[1]); 1280x960 is configs[2]


def render_frames(seed, size=None, count=None, threads=4):
    from concurrent.futures import ThreadPoolExecutor
    from supersurfel_fusion_b200.synth import SyntheticSequence
    w, h = size or SIZE
    seq = SyntheticSequence(width=w, height=h, seed=seed)
    with ThreadPoolExecutor(max_workers=threads) as pool:       # numpy releases the GIL in the heavy parts
        frames = list(pool.map(seq.frame, range(count or N_UNIQUE_FRAMES)))
    return seq, frames


def pin_rank_to_cores(local_rank, world):
    """One disjoint slice of the host cores per rank (all GPUs of a box report the same affinity mask, so
    without this N ranks' host threads migrate over each other's cores).  Returns the cores used."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // max(world, 1)
        if world > 1 and per >= 2:
            mine = cores[local_rank * per:(local_rank + 1) * per]
            os.sched_setaffinity(0, mine)
            return len(mine)
        return len(cores)
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def workload_config(world):
    """The `config` object of the JSON line, identical for both arms."""
    if world == 1:
        w = WORKLOAD if SIZE == [640, 480] else "configs[2]: %dx%d synthetic RGB-D stream, full pipeline" % tuple(SIZE)
    else:
        w = "configs[3]: %d independent %dx%d synthetic sequences, one per GPU, no NCCL on the data path" % (world, SIZE[0], SIZE[1])
    return {"workload": w, "params": "launch/supersurfel_fusion_rgbd_benchmark.launch",
            "l2_policy": "24 distinct frames cycle through the engine (ping-pong); the per-frame working set (~9 MB of "
                         "images + the model) is L2 resident by the nature of the workload, no flush between frames; "
                         "the roofline kernel streams 604 MB per launch (> 126 MB L2)"}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def time_cpu_oracle(frames, cam, n_frames, threads=1, warmup=0):
    """CPU oracle port on the host cores, bounded sample of the same workload.  `threads` host
    threads each own one engine and one copy of the sequence (the way N GPUs own N sequences):
    the path has no cross-frame parallelism, so frame-level replication is how it uses a
    multi-core host.  Returns (aggregate frames/s, seconds of wall clock)."""
    from oracle import orc
    p = dict(PARAMS)
    p["seg_use_ransac"] = int(p["seg_use_ransac"])
    cfg = orc.default_config(cam=cam, **p)
    orc.set_num_threads(1)
    engines = [orc.Engine(cfg) for _ in range(threads)]
    gate = threading.Barrier(threads + 1)

    def work(eng):
        eng.process_frame(*frames[0])         # bootstrap frame + warm-up outside the timed region
        for s in range(1, warmup + 1):
            eng.process_frame(*frames[frame_index(s)])
        gate.wait()
        for s in range(warmup + 1, warmup + n_frames + 1):
            eng.process_frame(*frames[frame_index(s)])
        gate.wait()

    pool = [threading.Thread(target=work, args=(e,), daemon=True) for e in engines]
    for t in pool:
        t.start()
    gate.wait()
    t0 = time.perf_counter()
    gate.wait()
    dt = time.perf_counter() - t0
    for t in pool:
        t.join()
    return threads * n_frames / dt, dt


def time_reference_gpu_kernels(n_frames):
    """Runs the reference-kernel harness in a CHILD process (its code prints to stdout and is not ours to
    trust with this process' life); returns its JSON object, or a note when it is unavailable."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--ref-gpu-only", "--steps", str(n_frames),
                              "--size", "%dx%d" % tuple(SIZE)], capture_output=True, text=True, timeout=300)
        lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not lines:
            return {"unavailable": "harness produced no result (exit %d)" % out.returncode}
        return json.loads(lines[-1])
    except Exception as exc:
        return {"unavailable": repr(exc)[:200]}


def _time_reference_gpu_kernels(frames, cam, n_frames):
    """The reference's OWN CUDA kernels (TPS_RGBD, DenseRegistration and the surfel kernels,
    compiled unmodified for sm_100a into oracle/_ref/libssf_ref.so) replaying processFrame on
    this GPU: context for the speed-up, next to the CPU arm the contract asks for."""
    try:
        from oracle import ref
        if not ref.available():
            return None
        from supersurfel_fusion_b200 import Supersurfels
        p = dict(PARAMS)
        p["seg_use_ransac"] = int(p["seg_use_ransac"])
        eng = ref.RefEngine(cam, Supersurfels, **p)
        for s in range(3):
            eng.process_frame(*frames[frame_index(s)])
        t0 = time.perf_counter()
        dev_ms = 0.0
        for s in range(3, 3 + n_frames):
            dev_ms += eng.process_frame(*frames[frame_index(s)])["ms_total"]
        wall = time.perf_counter() - t0
        eng.close()
        return {"value": n_frames / wall, "unit": "frames/s", "device_ms_per_frame": dev_ms / n_frames,
                "wall_ms_per_frame": wall / n_frames * 1e3, "frames": n_frames,
                "what": "reference kernels + launch sequence (oracle/_ref harness), pageable host inputs, same GPU"}
    except Exception as exc:   # the harness is optional context, never a reason to lose the bench line
        return {"unavailable": repr(exc)[:200]}


def run_reference(args, rank, world):
    if rank != 0:
        return
    seq, frames = render_frames(1234)
    steps = min(args.steps, 100)
    warmup = min(args.warmup, 5)
    threads = host_threads()
    # every host thread runs its own engine over W warm-up + K timed frames of the sequence
    fps, dt = time_cpu_oracle(frames, seq.cam_param(), steps, threads=threads, warmup=warmup)
    line = {
        "impl": "reference", "metric": "RGB-D frames/sec @640x480", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 / fps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world),
        "method": {"note": "the reference has no CPU implementation of this path and its full build needs "
                           "ROS/OpenCV-CUDA/g2o; this is the CPU oracle restatement of it (oracle/), one engine per "
                           "host thread, all host threads; rank 0 only"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d frames per thread x %d threads of the workload sequence (%.1f s)" % (steps, threads, dt)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        import torch
        if torch.cuda.is_available():
            line["reference_gpu_kernels"] = time_reference_gpu_kernels(60)
    except Exception:
        pass
    print(json.dumps(line))


def icp_roofline(device, peaks):
    """ICP system kernel at HBM-bound sizing; returns the roofline object."""
    import torch
    from supersurfel_fusion_b200 import CamParam, SupersurfelFusion, Supersurfels
    from supersurfel_fusion_b200.synth import synthetic_icp_problem
    n = ICP_ROOFLINE_N
    prob = synthetic_icp_problem(n, width=2560, height=1920, seed=1234)
    eng = SupersurfelFusion(device).initialize(CamParam(*prob["cam"]), nb_supersurfels_max=n)
    S = prob["S"]
    frame = Supersurfels(S)
    frame.colors[:] = prob["tgt_col"]; frame.orientations[:] = prob["tgt_ori"]; frame.confidences[:] = prob["tgt_conf"]
    eng.setSegmentation(labels=prob["labels"], slanted=prob["depth"])
    eng.setFrame(frame)
    # model arrays straight from numpy (only the members the kernel reads are uploaded)
    from supersurfel_fusion_b200.engine import SsfSurfels, _ptr
    view = SsfSurfels(_ptr(prob["src_pos"]), _ptr(prob["src_col"]), None, _ptr(prob["src_ori"]), None, None, None)
    eng.setModelPointers(view, n, n)
    R = np.eye(3, dtype=np.float32)
    t = np.array([0.002, -0.001, 0.003], np.float32)
    sys29 = eng.icpSystem(R, t, n)                  # warm-up + sanity
    for _ in range(3):
        eng.icpSystemEnqueue(R, t, n, 1)
    eng.synchronize()
    L = 20
    eng.timerStart()
    eng.icpSystemEnqueue(R, t, n, L)                # 604 MB streamed per launch >> 126 MB L2: no flush needed
    ms = eng.timerStop() / L
    achieved = n * ICP_BYTES_PER_SRC / (ms * 1e-3) / 1e9
    peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    traffic_note = "no ncu capture committed"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "icp_system_traffic.json")))
        traffic = tj["traffic_bytes"]
        traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full "
                        "capture of this command (%s); the same capture shows %.3g B moved from L2 to the SMs per launch"
                        % (tj["source"], tj["l2_to_sm_bytes"]))
    except Exception:
        tj = None
    out = {"bound": "hbm", "kernel": "icp_system_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
           "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
           "dram_gbs": (traffic / (ms * 1e-3) / 1e9) if traffic else None,
           "dram_frac": (traffic / (ms * 1e-3) / 1e9 / peak) if traffic else None,
           "l2_to_sm_gbs": (tj["l2_to_sm_bytes"] / (ms * 1e-3) / 1e9) if tj else None,
           "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
           "n_src": n, "us_per_launch": ms * 1e3, "algorithmic_bytes_per_src": ICP_BYTES_PER_SRC,
           "streamed_only_gbs": n * 36 / (ms * 1e-3) / 1e9, "inlier_frac": float(sys29[28]) / n}
    # the same sources binned by the 32x32-pixel tile they project to (what a model prefix kept in image order
    # looks like): the two gathers of neighbouring lanes then share sectors / L1 lines, and the kernel is back on
    # the HBM streams.  The scattered figure above stays the headline (SURVEY.md section 8d sizing).
    try:
        cam = prob["cam"]
        uu = np.clip(prob["src_pos"][:, 0] / prob["src_pos"][:, 2] * cam[0] + cam[2], 0, cam[5] - 1).astype(np.int32) >> 5
        vv = np.clip(prob["src_pos"][:, 1] / prob["src_pos"][:, 2] * cam[1] + cam[3], 0, cam[4] - 1).astype(np.int32) >> 5
        order = np.argsort(vv * ((cam[5] + 31) >> 5) + uu, kind="stable")
        bp, bc, bo = (np.ascontiguousarray(prob[k][order]) for k in ("src_pos", "src_col", "src_ori"))
        eng.setModelPointers(SsfSurfels(_ptr(bp), _ptr(bc), None, _ptr(bo), None, None, None), n, n)
        sys_b = eng.icpSystem(R, t, n)
        for _ in range(3):
            eng.icpSystemEnqueue(R, t, n, 1)
        eng.synchronize()
        eng.timerStart()
        eng.icpSystemEnqueue(R, t, n, L)
        ms_b = eng.timerStop() / L
        out["binned_by_tile"] = {"us_per_launch": ms_b * 1e3, "achieved": n * ICP_BYTES_PER_SRC / (ms_b * 1e-3) / 1e9,
                                 "frac": n * ICP_BYTES_PER_SRC / (ms_b * 1e-3) / 1e9 / peak,
                                 "streamed_only_gbs": n * 36 / (ms_b * 1e-3) / 1e9,
                                 "streamed_only_frac": n * 36 / (ms_b * 1e-3) / 1e9 / peak,
                                 "inliers_equal_scattered": bool(sys_b[28] == sys29[28]),
                                 "what": "same 16 Mi sources sorted by projected 32x32-pixel tile: gathers hit L1"}
        del bp, bc, bo
    except Exception as exc:
        out["binned_by_tile"] = {"unavailable": repr(exc)[:200]}
    # latency at realistic sizes (L2 resident, launch bound): microseconds per system build
    lat = {}
    for m in (1200, 5000, 50000, 100000):
        for _ in range(3):
            eng.icpSystemEnqueue(R, t, m, 1)
        eng.synchronize()
        eng.timerStart()
        eng.icpSystemEnqueue(R, t, m, 50)
        lat[str(m)] = eng.timerStop() / 50 * 1e3
    out["latency_us"] = lat
    eng.close()
    return out


def tile_parallel_icp_report(dist, dev, sizes=(100000, 16 * 1024 * 1024)):
    """configs[4]: one 2560x1920 frame, the frame-to-model registration with the ICP source range tiled
    across all ranks (supersurfel_fusion_b200/multi.py): single-GPU loop vs NCCL rank-ordered reduction vs
    the fused build + peer-memory exchange + solve kernel.  Collective: every rank calls it.  Milliseconds
    are whole registrations (all Gauss-Newton iterations), wall clock around 5 calls incl. the host
    synchronisation each call ends with, max over ranks."""
    import torch
    from supersurfel_fusion_b200 import CamParam, SupersurfelFusion, Supersurfels, multi
    from supersurfel_fusion_b200.engine import SsfSurfels, _ptr
    from supersurfel_fusion_b200.synth import synthetic_icp_problem
    world = dist.get_world_size()
    tdev = torch.device("cuda", dev)
    out = {"what": "configs[4]: one 2560x1920 frame (S = 19200 frame supersurfels), registration with the source "
                   "range tiled over %d GPUs; ms per whole registration, max over ranks" % world, "world": world}
    R = np.eye(3, dtype=np.float32)
    t = np.array([0.004, -0.003, 0.005], np.float32)
    for n in sizes:
        prob = synthetic_icp_problem(n, width=2560, height=1920, seed=1234)
        eng = SupersurfelFusion(dev).initialize(CamParam(*prob["cam"]), nb_supersurfels_max=n)
        frame = Supersurfels(prob["S"])
        frame.colors[:] = prob["tgt_col"]; frame.orientations[:] = prob["tgt_ori"]; frame.confidences[:] = prob["tgt_conf"]
        eng.setSegmentation(labels=prob["labels"], slanted=prob["depth"])
        eng.setFrame(frame)
        eng.setModelPointers(SsfSurfels(_ptr(prob["src_pos"]), _ptr(prob["src_col"]), None, _ptr(prob["src_ori"]), None,
                                        None, None), n, n)
        multi.connect_peers(eng, dist, device=tdev)

        def clock(fn, reps=5):
            fn()
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                res = fn()
            ms = (time.perf_counter() - t0) / reps * 1e3
            return res, multi.allreduce_max(dist, ms, device=tdev)

        (ok1, R1, t1, st1), one_ms = clock(lambda: eng.icp(R, t))
        (okN, RN, tN, stN), nccl_ms = clock(lambda: multi.tile_parallel_icp(eng, dist, n, R, t, device=tdev))
        (okF, RF, tF, stF), fused_ms = clock(lambda: multi.fused_tile_parallel_icp(eng, dist, n, R, t))
        out[str(n)] = {
            "n_src": n, "iters": st1["iters"], "valid": bool(ok1 and okN and okF),
            "single_gpu_ms": one_ms, "nccl_ordered_ms": nccl_ms, "fused_peer_memory_ms": fused_ms,
            "dt_m_nccl": float(np.linalg.norm(t1 - tN)), "dR_nccl": float(np.abs(R1 - RN).max()),
            "dt_m_fused": float(np.linalg.norm(t1 - tF)), "dR_fused": float(np.abs(R1 - RF).max()),
            "inliers_single": float(st1["system"][28]),
            "inlier_delta_nccl": float(stN["system"][28] - st1["system"][28]),
            "inlier_delta_fused": float(stF["system"][28] - st1["system"][28]),
            "fused_equals_nccl_bits": bool(np.array_equal(tF, tN) and np.array_equal(RF, RN)),
        }
        eng.close()
        del prob
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    from supersurfel_fusion_b200 import CamParam, SupersurfelFusion
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    host_cores = pin_rank_to_cores(local_rank, world)
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    torch.cuda.set_device(dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        tns = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    def timed(cam, bufs, device_resident, pipelined, steps, warmup, reps):
        """`reps` windows of EXACTLY `steps` frames each, after `warmup` frames THROUGH THE SAME PATH (so
        that every CUDA graph the path replays exists and is uploaded before the first timed frame;
        ssf_prepare builds them up front as well).  Each window is bracketed by barrier + synchronize,
        timed with CUDA events on the engine's stream (the pipeline is drained inside the window), max
        over ranks; the MEDIAN window is reported.  pipelined: ssf_submit_frame / ssf_wait_frame, one
        frame in flight per pipeline stage (results identical to the synchronous call,
        tests/test_gpu_engine.py)."""
        eng = SupersurfelFusion(dev).initialize(CamParam(*cam), **PARAMS)
        eng.prepare()
        pos = [0]

        def run_window(n):
            s0 = pos[0]
            if pipelined:
                depth = min(eng.pipelineDepth(), n)
                for i in range(n):
                    if i >= depth:
                        eng.waitFrame()               # `depth` frames in flight, one per stage
                    eng.submitFrame(*bufs[frame_index(s0 + i, len(bufs))])
                for _ in range(depth):
                    eng.waitFrame()
            elif device_resident:
                for i in range(n):
                    eng.processFrameDevice(*bufs[frame_index(s0 + i, len(bufs))])
            else:
                for i in range(n):
                    eng.processFrame(*bufs[frame_index(s0 + i, len(bufs))])
            pos[0] = s0 + n

        run_window(warmup)
        windows, walls, launches = [], [], 0
        for _ in range(reps):
            barrier()
            l0 = eng.launchCount()
            eng.timerStart()                      # CUDA event on the engine's stream
            w0 = time.perf_counter()
            run_window(steps)
            ms = eng.timerStop()                  # records + synchronises
            wall = (time.perf_counter() - w0) * 1e3
            barrier()
            launches = eng.launchCount() - l0
            windows.append(max_over_ranks(ms))
            walls.append(max_over_ranks(wall))
        stats = eng.getFrameStats()
        stats["pipeline_depth"] = eng.pipelineDepth()
        eng.close()
        return float(np.median(windows)), float(np.median(walls)), launches, stats, windows

    def stage_breakdown(cam, bufs, n=30):
        """Per-stage device milliseconds of the synchronous frame (SSF_FLAG_STAGE_TIMING: event nodes at the
        stage boundaries of the frame graph) -- what the reference prints per frame
        (supersurfel_fusion.cu:516-528); mean over n frames after warm-up."""
        from supersurfel_fusion_b200.engine import SSF_FLAG_STAGE_TIMING
        eng = SupersurfelFusion(dev).initialize(CamParam(*cam), **PARAMS)
        keys = ("ms_ingest", "ms_segmentation", "ms_extraction", "ms_registration", "ms_fusion", "gpu_ms")
        acc = dict.fromkeys(keys, 0.0)
        for s in range(5 + n):
            eng.processFrameDevice(*bufs[frame_index(s)], flags=SSF_FLAG_STAGE_TIMING)
            if s >= 5:
                st = eng.getFrameStats()
                for k in keys:
                    acc[k] += st[k] / n
        eng.close()
        return acc

    def multi_stream(cam, h_bufs, k, steps):
        """Throughput headroom of ONE GPU: k independent sequences, one engine (own stream, own CUDA
        graph) and one host thread each, end to end from pinned host buffers.  A VGA frame is a chain of
        small kernels that cannot fill 148 SMs; concurrent sequences can."""
        engines = [SupersurfelFusion(dev).initialize(CamParam(*cam), **PARAMS) for _ in range(k)]
        for eng in engines:
            for s in range(5):
                eng.processFrame(*h_bufs[frame_index(s)])
        gate = threading.Barrier(k + 1)

        def work(eng):
            gate.wait()
            for s in range(5, 5 + steps):
                eng.processFrame(*h_bufs[frame_index(s)])
            gate.wait()

        pool = [threading.Thread(target=work, args=(eng,), daemon=True) for eng in engines]
        for t in pool:
            t.start()
        torch.cuda.synchronize()
        gate.wait()
        t0 = time.perf_counter()
        gate.wait()
        dt = time.perf_counter() - t0
        for t in pool:
            t.join()
        for eng in engines:
            eng.close()
        return k * steps / dt

    def upload(frames):
        d = [(torch.from_numpy(f[0]).cuda(dev), torch.from_numpy(f[1]).cuda(dev)) for f in frames]
        h = [(torch.from_numpy(f[0]).pin_memory(), torch.from_numpy(f[1]).pin_memory()) for f in frames]
        torch.cuda.synchronize()
        return d, h

    render_threads = max(1, min(8, host_cores))
    seq, frames = render_frames(1234 + rank, threads=render_threads)   # configs[3]: independent sequences, seeds 1234..
    cam = seq.cam_param()
    dev_bufs, host_bufs = upload(frames)     # resident copies (value) and pinned host copies (e2e)

    sampler = None
    if rank == 0:
        sampler = ClockSampler(dev)
        sampler.start()
    pipelined = not args.no_pipeline
    reps = max(1, args.reps)
    ms, wall_ms, launches, stats, win = timed(cam, dev_bufs, True, pipelined, args.steps, args.warmup, reps)
    ms_e, wall_e, _, _, win_e = timed(cam, host_bufs, False, pipelined, args.steps, args.warmup, reps)
    pipe_depth = stats.pop("pipeline_depth")
    sync_ms = sync_ms_e = None
    if pipelined and not args.skip_extras:
        sync_ms = timed(cam, dev_bufs, True, False, args.steps, args.warmup, reps)[0]
        sync_ms_e = timed(cam, host_bufs, False, False, args.steps, args.warmup, reps)[0]
    clocks = sampler.stop() if sampler is not None else None

    # configs[4] (one large frame, ICP tiled over the ranks): collective, outside the timed frames region
    tile_icp = None
    if world > 1 and not args.skip_extras:
        try:
            tile_icp = tile_parallel_icp_report(dist, dev)
        except Exception as exc:      # context for the line, never a reason to lose it
            tile_icp = {"unavailable": repr(exc)[:300]}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roof = icp_roofline(dev, peaks) if (world == 1 and not args.skip_extras) else None
    cpu = ref_gpu = streams = stages = cfg2 = None
    if world == 1 and not args.skip_extras:
        stages = stage_breakdown(cam, dev_bufs)
        stages["what"] = ("device ms per stage of the synchronous frame, mean of 30 frames (SSF_FLAG_STAGE_TIMING; "
                          "reference prints the same breakdown, supersurfel_fusion.cu:516-528)")
        threads = host_threads()
        fps_cpu, dt = time_cpu_oracle(frames, cam, 60, threads=threads, warmup=2)
        cpu = {"value": fps_cpu, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "60 frames per thread x %d threads of the workload sequence (%.1f s), CPU oracle "
                         "restatement, one engine per host thread" % (threads, dt)}
        ref_gpu = time_reference_gpu_kernels(60)
        streams = {"unit": "frames/s", "what": "k independent sequences on this ONE GPU, one engine + host thread each, "
                   "end to end from pinned host buffers (wall clock); k = 1 is the synchronous e2e figure's setting",
                   "by_k": {str(k): multi_stream(cam, host_bufs, k, 150) for k in (1, 2, 4, 8)}}
        if SIZE == [640, 480]:
            # configs[2]: the 1280x960 stream, same engine, same legs (north_star names both resolutions)
            seq2, frames2 = render_frames(1234, size=(1280, 960), count=12, threads=render_threads)
            try:
                d2, h2 = upload(frames2)
                cam2 = seq2.cam_param()
                k2 = max(100, args.steps)
                v2 = timed(cam2, d2, True, True, k2, 10, 3)
                e2 = timed(cam2, h2, False, True, k2, 10, 3)
                s2 = timed(cam2, d2, True, False, k2, 10, 3)
                se2 = timed(cam2, h2, False, False, k2, 10, 3)
                cfg2 = {"workload": "configs[2]: 1280x960 synthetic RGB-D stream, full pipeline, S = 4800",
                        "unit": "frames/s", "frames_per_window": k2, "windows": 3,
                        "value": k2 / (v2[0] * 1e-3), "e2e": k2 / (e2[0] * 1e-3),
                        "synchronous": {"value": k2 / (s2[0] * 1e-3), "e2e": k2 / (se2[0] * 1e-3)},
                        "h2d_bytes_per_step": 1280 * 960 * 7, "gpu_launches_per_window": int(v2[2]),
                        "last_frame_stats": {k: v for k, v in v2[3].items() if k != "pipeline_depth"}}
                del d2, h2
            except Exception as exc:
                cfg2 = {"unavailable": repr(exc)[:300]}
    total_frames = args.steps * world
    value = total_frames / (ms * 1e-3)
    e2e = total_frames / (ms_e * 1e-3)
    line = {
        "metric": "RGB-D frames/sec @%dx%d" % tuple(SIZE), "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world),
        "method": {"frames_per_gpu": args.steps,
                   "pipeline": ("ssf_submit_frame / ssf_wait_frame: the frame's kernel chain cut into %d stages of equal cost on "
                                "separate streams, one frame in flight per stage; identical results to the synchronous "
                                "ssf_process_frame, whose numbers are under 'synchronous'" % pipe_depth) if pipelined
                               else "synchronous ssf_process_frame, one frame at a time",
                   "timing": "median of %d windows of exactly K frames each; every window bracketed by barrier + "
                             "synchronize, CUDA events on the engine stream (pipeline drained inside), max over ranks; "
                             "warm-up frames go through the timed path after ssf_prepare built all graphs" % reps,
                   "host_cores_per_rank": host_cores},
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": SIZE[0] * SIZE[1] * 7, "d2h_bytes_per_step": 112,
                "ms_per_step": ms_e / args.steps, "wall_ms_per_step": wall_e / args.steps, "window_ms": win_e},
        "wall_ms_per_step": wall_ms / args.steps,
        "window_ms": win,
        "synchronous": None if sync_ms is None else {
            "value": total_frames / (sync_ms * 1e-3), "e2e": total_frames / (sync_ms_e * 1e-3), "unit": "frames/s",
            "ms_per_frame": sync_ms / args.steps, "what": "one ssf_process_frame call per frame (the reference's call shape)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "stage_ms": stages,
        "configs2_1280x960": cfg2,
        "tile_parallel_icp": tile_icp,
        "reference_gpu_kernels": ref_gpu,
        "concurrent_sequences_one_gpu": streams,
        "last_frame_stats": stats,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--reps", type=int, default=5, help="timed windows of K frames each; the median window is reported")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="640x480", help="frame size WxH (default: the metric's 640x480; 1280x960 = configs[2])")
    ap.add_argument("--no-pipeline", action="store_true", help="time the synchronous ssf_process_frame instead of submit/wait")
    ap.add_argument("--skip-extras", action="store_true", help="frames only: no roofline / cpu_baseline legs (for ncu)")
    ap.add_argument("--roofline-only", action="store_true", help="only the ICP roofline leg (for ncu --set full)")
    ap.add_argument("--ref-gpu-only", action="store_true", help="(internal) only the reference-kernel harness leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3
    SIZE[0], SIZE[1] = (int(v) for v in args.size.lower().split("x"))
    if args.ref_gpu_only:
        seq, frames = render_frames(1234)
        print(json.dumps(_time_reference_gpu_kernels(frames, seq.cam_param(), args.steps)))
    elif args.roofline_only:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        print(json.dumps(icp_roofline(local_rank, peaks)))
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
