"""Summarise an `ncu --set full` report and an `ncu --metrics gpu__time_duration.sum` launch list into the
small tracked files under profiles/ (the .ncu-rep itself stays in gpurun_out/, which is scratch).

   python tools/ncu_summary.py <tag> [--rep gpurun_out/<tag>_bounce.ncu-rep] [--launches gpurun_out/<tag>_launches.csv]

Writes profiles/<tag>_bounce_ncu.csv (one row per captured launch, the metrics DESIGN.md quotes),
profiles/<tag>_bounce_traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic) and
profiles/<tag>_launches_summary.txt (per-kernel share of the GPU time of one bench.py run).
"""
import argparse
import collections
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
    "smsp__inst_executed_op_branch.sum", "sass__inst_executed_register_spilling",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "lts__t_sector_hit_rate.pct",
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--kernel", default="bounce")
    a = ap.parse_args()
    rep = a.rep or os.path.join(ROOT, "gpurun_out", a.tag + "_bounce.ncu-rep")
    launches = a.launches or os.path.join(ROOT, "gpurun_out", a.tag + "_launches.csv")
    out = os.path.join(ROOT, "profiles")
    os.makedirs(out, exist_ok=True)

    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
        iK = hdr.index("Kernel Name")
        with open(os.path.join(out, "%s_%s_ncu.csv" % (a.tag, a.kernel)), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel"] + ["%s [%s]" % (m, units[i]) for m, i in cols])
            for r in data:
                w.writerow([r[iK]] + [r[i] for _, i in cols])
        iR, iW, iT = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        per = [{"kernel": r[iK].strip(), "dram_read_bytes": to_bytes(r[iR], units[iR]), "dram_write_bytes": to_bytes(r[iW], units[iW]),
                "duration_us": float(r[iT].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[iT], 1)} for r in data]
        tot = sum(p["dram_read_bytes"] + p["dram_write_bytes"] for p in per)
        json.dump({"source": os.path.basename(rep), "launches_captured": len(per), "dram_bytes_per_launch": tot / max(1, len(per)),
                   "note": "dram__bytes_read.sum + dram__bytes_write.sum averaged over the captured launches (consecutive bounce kernels of bench.py's wavefront)",
                   "launches": per}, open(os.path.join(out, "%s_%s_traffic.json" % (a.tag, a.kernel)), "w"), indent=1)
        print("wrote ncu summary for", len(per), "launches;", tot / max(1, len(per)) / 1e6, "MB DRAM traffic per launch")

    if os.path.exists(launches):
        rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
        hdr = rows[0]
        iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        tot, cnt = collections.Counter(), collections.Counter()
        for r in rows[1:]:
            v = float(r[iV].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iU], 1.0)
            tot[r[iK]] += v
            cnt[r[iK]] += 1
        T = sum(tot.values())
        with open(os.path.join(out, "%s_launches_summary.txt" % a.tag), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, one bench.py run (cold-cache, serialised: shares, not absolutes)\n")
            f.write("# %d launches, %.1f us of GPU time in total\n" % (sum(cnt.values()), T))
            for k, v in tot.most_common():
                f.write("%6.2f%%  %10.1f us  %5d launches  avg %8.1f us  %s\n" % (100 * v / T, v, cnt[k], v / cnt[k], k))
        print(open(os.path.join(out, "%s_launches_summary.txt" % a.tag)).read())


if __name__ == "__main__":
    main()
