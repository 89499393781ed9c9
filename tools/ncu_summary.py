#!/usr/bin/env python
"""Summaries of ncu outputs for profiles/ (run in the build container; no GPU needed).

  python tools/ncu_summary.py launches gpurun_out/x_launches.csv          > profiles/rNN_launches.txt
  python tools/ncu_summary.py full     gpurun_out/x.ncu-rep [regex]       > profiles/rNN_full.txt
  python tools/ncu_summary.py traffic  gpurun_out/x.ncu-rep               > profiles/ncu_traffic.json   (read by bench.py: roofline.traffic)
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value"); mu = hdr.index("Metric Unit")
    d = defaultdict(lambda: [0, 0.0])
    unit = None
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        unit = r[mu]
        name = re.sub(r"\(.*", "", r[kn])[:70]
        d[name][0] += 1; d[name][1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in d.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)  source: %s  unit: %s" % (path, unit))
    print("%-72s %6s %14s %12s %7s" % ("kernel", "n", "total", "avg", "share"))
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %6d %14.0f %12.0f %7.3f" % (k, v[0], v[1], v[1] / v[0], v[1] / tot))


def full(path, pat=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    print("# ncu --set full --clock-control none  source: %s" % path)
    for r in rows[2:]:
        if pat and not re.search(pat, r[kn]):
            continue
        print("== %s" % r[kn][:150])
        for k in KEYS:
            if k in hdr:
                print("   %-75s %18s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))


def traffic(path):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches) of the EAM tile kernels"""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name"); rd = hdr.index("dram__bytes_read.sum"); wr = hdr.index("dram__bytes_write.sum"); du = hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = defaultdict(list)
    for r in rows[2:]:
        key = "eam_force" if "EamForceTileOp" in r[kn] else "eam_rho_rewrite" if ("EamRhoTileOp" in r[kn] and ", 3, " in r[kn]) else "eam_rho" if "EamRhoTileOp" in r[kn] else None
        ms = float(r[du]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[du], 1.0)
        if key and ms > 0.05:              # the launch of a step whose mode is not current returns at once: not a measurement
            acc[key].append((float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]], float(r[du]), r[kn][:120]))
    res = {"source": path, "how": "ncu --set full --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum, mean per launch"}
    for k, v in acc.items():
        res[k] = {"dram_bytes_per_launch": sum(x[0] for x in v) / len(v), "launches": len(v), "kernel": v[0][2], "duration_under_ncu": "%g %s" % (sum(x[1] for x in v) / len(v), units[du])}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
