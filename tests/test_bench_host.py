"""Host-side logic of bench.py that needs no GPU: the frame schedule, the config object shared by both arms, the
per-rank core partition, and the reference arm's JSON line (a short CPU run of the oracle port)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_frame_schedule_ping_pongs_without_jumps():
    n = bench.N_UNIQUE_FRAMES
    idx = [bench.frame_index(s) for s in range(5 * n)]
    assert min(idx) == 0 and max(idx) == n - 1
    assert all(abs(a - b) == 1 for a, b in zip(idx, idx[1:]))          # consecutive frames are neighbours: ~1 cm motion
    idx12 = [bench.frame_index(s, 12) for s in range(60)]
    assert max(idx12) == 11 and all(abs(a - b) == 1 for a, b in zip(idx12, idx12[1:]))


def test_config_object_is_shared_by_both_arms():
    for world in (1, 2, 8):
        a, b = bench.workload_config(world), bench.workload_config(world)
        assert a == b and set(a) == {"workload", "params", "l2_policy"}
    assert "configs[1]" in bench.workload_config(1)["workload"] and "configs[3]" in bench.workload_config(8)["workload"]


def test_rank_core_partition_is_disjoint():
    cores = sorted(os.sched_getaffinity(0))
    try:
        world = 2 if len(cores) >= 4 else 1
        seen = []
        for r in range(world):
            os.sched_setaffinity(0, cores)
            got = bench.pin_rank_to_cores(r, world)
            mine = sorted(os.sched_getaffinity(0))
            assert got == len(mine) >= 1
            seen.append(set(mine))
        if world == 2:
            assert not (seen[0] & seen[1]) and (seen[0] | seen[1]) <= set(cores)
    finally:
        os.sched_setaffinity(0, cores)


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"] == bench.workload_config(1)


def test_committed_driver_bench_line_honours_the_contract():
    """The bench line kept under profiles/ for the round (the driver's own command) carries every key of the
    bench contract and is self-consistent."""
    import json
    path = os.path.join(ROOT, "profiles", "bench_driver_r2_final.json")
    d = json.loads([l for l in open(path) if l.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] == 5 and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] * d["ms_per_step"] / 1000.0 - 1.0) < 1e-6            # frames/s x s/frame
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 640 * 480 * (3 + 4) and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]                                           # measured, not copied
    assert d["gpu_launches"] >= 57 * d["steps"]
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - 72 * r["n_src"] / (r["us_per_launch"] * 1e-6) / 1e9) < 1e-3 * r["achieved"]
    assert r["traffic"] and 0.4 < r["dram_frac"] < 1.0
    b = d["cpu_baseline"]
    assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
