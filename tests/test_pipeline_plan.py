"""The frame pipeline's stage planner (csrc/ssf_engine.cu plan_stages_impl, behind ssf_plan_pipeline): pure
host logic, checked without a GPU -- the cut is contiguous, covers every step once, is optimal for its
weights, and degrades gracefully."""
import ctypes as C
import itertools

import pytest

from supersurfel_fusion_b200 import load_library
from supersurfel_fusion_b200.engine import SsfConfig


def _weights(seg_iter, icp_iter, persistent):
    """The planner's own per-step cost estimates (ssf_plan_weights)."""
    lib = load_library()
    cfg = SsfConfig()
    lib.ssf_config_default(C.byref(cfg))
    cfg.seg_iter, cfg.icp_iter = seg_iter, icp_iter
    w = (C.c_int * 64)()
    lib.ssf_plan_weights.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    n = lib.ssf_plan_weights(C.byref(cfg), int(persistent), w, 64)
    assert n == (1 if persistent else seg_iter + 2) + 3
    assert all(v > 0 for v in w[:n])
    return list(w[:n])


def _plan(lib, seg_iter, icp_iter, stages, persistent=0):
    cfg = SsfConfig()
    lib.ssf_config_default(C.byref(cfg))
    cfg.seg_iter, cfg.icp_iter = seg_iter, icp_iter
    first = (C.c_int * 8)()
    steps = C.c_int(0)
    lib.ssf_plan_pipeline.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    used = lib.ssf_plan_pipeline(C.byref(cfg), stages, persistent, first, C.byref(steps))
    return used, list(first[:used + 1]), steps.value


@pytest.mark.parametrize("seg_iter,icp_iter", [(10, 10), (8, 6), (1, 10), (0, 3), (20, 10)])
def test_plan_is_a_contiguous_optimal_cut(seg_iter, icp_iter):
    lib = load_library()
    w = _weights(seg_iter, icp_iter, False)
    for stages in range(1, 7):
        used, first, steps = _plan(lib, seg_iter, icp_iter, stages)
        assert steps == len(w) == seg_iter + 5
        assert used == min(stages, steps) and first[0] == 0 and first[-1] == steps
        assert all(b > a for a, b in zip(first, first[1:]))                  # every stage holds at least one step
        heaviest = max(sum(w[a:b]) for a, b in zip(first, first[1:]))
        best = min(max(sum(w[a:b]) for a, b in zip((0,) + cuts, cuts + (steps,)))
                   for cuts in itertools.combinations(range(1, steps), used - 1))
        assert heaviest == best                                              # no cut has a lighter heaviest stage


def test_plan_defaults_and_degenerate_inputs():
    lib = load_library()
    used, first, steps = _plan(lib, 10, 10, 4)
    assert (used, steps) == (4, 15)
    assert first[1] > 1 and first[3] >= 12         # ingest is not alone; the last stage is registration + fusion (+ at most the tail before it)
    assert _plan(lib, 10, 10, 0)[0] == 1 and _plan(lib, 10, 10, 99)[0] == 6
    assert _plan(lib, 10, 10, 6, persistent=1) == (4, [0, 1, 2, 3, 4], 4)    # the one-kernel segmentation is one step
    assert _plan(lib, 1000, 10, 4)[0] == 1                                   # absurd iteration count: no pipelining
