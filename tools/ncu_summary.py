#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_rN.csv > profiles/launches_rN.txt
  python tools/ncu_summary.py kernel   gpurun_out/icp_rN.ncu-rep [profiles/icp_system_traffic.json] > profiles/icp_rN.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__t_sectors.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000.0 if unit in ("nsecond", "ns") else v * 1000.0 if unit in ("msecond", "ms") else v
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("# source: %s   launches: %d   total: %.1f us" % (path, n, tot))
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s n=%4d  total %9.1f us  avg %8.2f us  share %5.1f%%" % (k[:44], c, v, v / c, 100 * v / tot))


def _to_bytes(value, unit):
    v = float(value.replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * scale.get(unit, 1)


def kernel(path, json_out=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none --import-source on; source: %s" % path)
    if json_out:
        # per-launch DRAM traffic of the captured kernel, the `traffic` figure bench.py reports
        import json
        r = rows[2]
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        l2 = hdr.index("l1tex__m_xbar2l1tex_read_bytes.sum")
        info = {"kernel": r[hdr.index("Kernel Name")], "source": path,
                "dram_bytes_read": _to_bytes(r[rd], units[rd]), "dram_bytes_write": _to_bytes(r[wr], units[wr]),
                "l2_to_sm_bytes": _to_bytes(r[l2], units[l2]),
                "ncu_duration_us": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")),
                "grid": r[hdr.index("launch__grid_size")], "block": r[hdr.index("launch__block_size")]}
        info["traffic_bytes"] = info["dram_bytes_read"] + info["dram_bytes_write"]
        json.dump(info, open(json_out, "w"), indent=1)
    for r in rows[2:]:
        print("kernel: %s  (id %s)" % (r[hdr.index("Kernel Name")], r[0]))
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("  %-78s %s %s" % (m, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](*sys.argv[2:])
