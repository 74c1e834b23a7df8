"""Brief per-launch table from an ncu report: python tools/ncu_brief.py <rep>"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, data = rows[0], rows[1], rows[2:]
cols = [("time", "gpu__time_duration.sum"), ("inst", "smsp__inst_executed.sum"), ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio"), ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"), ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"), ("barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        ("mathpipe", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
        ("notsel", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
        ("branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
        ("noinst", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
        ("dramR_MB", "dram__bytes_read.sum"), ("dramW_MB", "dram__bytes_write.sum")]
idx = [(n, hdr.index(m)) for n, m in cols if m in hdr]
print(" ".join("%9s" % n for n, _ in idx), "kernel")
iK = hdr.index("Kernel Name")
for r in data:
    out = []
    for n, i in idx:
        v = r[i].replace(",", "")
        try:
            f = float(v)
            out.append("%9.3g" % f if f < 1e6 else "%9.4g" % f)
        except ValueError:
            out.append("%9s" % v[:9])
    print(" ".join(out), r[iK][:40])
