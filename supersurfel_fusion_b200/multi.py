"""Host-side logic of the two multi-GPU modes.

* independent sequences (BASELINE configs[3]): one process + one engine per GPU, no data-path
  collective; only the timing barrier and the max-over-ranks use torch.distributed;
* one large frame tiled across GPUs (configs[4]): the ICP system is a sum over source
  supersurfels, so every rank builds it over its slice of the visible model prefix, the
  29 floats are summed across ranks IN RANK ORDER (all-gather + ordered fp32 sum, identical
  bits on every rank, unlike a ring all-reduce whose order depends on the rank), and every
  rank applies the identical Gauss-Newton step.

Everything here works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import numpy as np


def sequence_seed(rank, base=1234):
    """configs[3]: seeds 1234 ... 1234 + world - 1."""
    return base + int(rank)


def shard_range(n, rank, world, align=4):
    """Contiguous slice [begin, begin + count) of n elements for `rank`; begins are multiples of
    `align` (float4 loads), the slices tile [0, n) exactly."""
    per = -(-n // world)                 # ceil
    per = -(-per // align) * align       # round up to the alignment
    begin = rank * per
    if begin >= n:
        return 0, 0                      # nothing left for this rank
    return begin, min(n, begin + per) - begin


def aggregate_throughput(units_per_rank, world, max_ms):
    """Whole-job throughput: units all ranks processed / the slowest rank's device time."""
    return units_per_rank * world / (max_ms * 1e-3)


def allreduce_max(dist, value, device=None):
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def ordered_sum(dist, vec29, device=None):
    """Sum of the per-rank 29-float systems, accumulated in fp32 in rank order 0,1,2,...; every
    rank returns the same bits."""
    import torch
    v = np.ascontiguousarray(vec29, np.float32).reshape(-1)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return v.copy()
    world = dist.get_world_size()
    mine = torch.from_numpy(v.copy())
    if device is not None:
        mine = mine.to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    total = np.zeros_like(v)
    for p in parts:                       # fixed order: rank 0 first
        total = (total + p.cpu().numpy()).astype(np.float32)
    return total


def tile_parallel_icp(engine, dist, n_visible, R_init=None, t_init=None, device=None, apply_to_pose=False):
    """featureConstrainedSymmetricICP (dense_registration.cu:245-424) with the system build
    sharded over the ranks of `dist`.  `engine` is this rank's SupersurfelFusion holding the full
    frame state and (at least its slice of) the model.  Returns (valid, R_rel, t_rel, info)."""
    rank = dist.get_rank() if (dist is not None and dist.is_initialized()) else 0
    world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
    begin, count = shard_range(n_visible, rank, world)
    engine.icpBegin(R_init, t_init)
    builds = 0
    last = None
    for _ in range(engine.cfg.icp_iter):
        local = engine.icpBuild(begin, count)
        last = ordered_sum(dist, local, device)
        builds += 1
        if engine.icpSolve(last):
            break
    valid, R, t, info = engine.icpFinish(apply_to_pose)
    info = dict(info, builds=builds, system=last, shard=(begin, count))
    return valid, R, t, info


def connect_peers(engine, dist, device=None):
    """Exchange the engines' peer-memory handles over `dist` and map every rank's exchange buffer
    (once, after creating the engine).  Needed for fused_tile_parallel_icp()."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.from_numpy(engine.peerHandle().copy())
    if device is not None:
        mine = mine.to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    handles = np.concatenate([p.cpu().numpy() for p in parts])
    engine.connectPeers(rank, world, handles)
    dist.barrier()      # every rank has mapped every buffer before anyone starts a fused loop


def fused_tile_parallel_icp(engine, dist, n_visible, R_init=None, t_init=None):
    """Same result as tile_parallel_icp(), but the slices' sums travel GPU-to-GPU by peer stores
    inside the system kernel and every rank solves on the device: no NCCL call and no host round
    trip per iteration (ssf_icp_tiled)."""
    begin, count = shard_range(n_visible, dist.get_rank(), dist.get_world_size())
    valid, R, t, info = engine.icpTiled(begin, count, R_init, t_init)
    return valid, R, t, dict(info, shard=(begin, count))
