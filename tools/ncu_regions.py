#!/usr/bin/env python
"""Summarise an ncu report: headline metrics + executed-instruction share per SASS region."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for h, u, v in zip(hdr, units, vals):
    if h in keep:
        print(f"{h:85s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
iS, iN, iE, iT = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
data = []
for r in rows[2:]:
    try:
        data.append((r[iS], int(r[iN] or 0), int(r[iE] or 0), float(r[iT] or 0)))
    except Exception:
        pass
tot = sum(d[2] for d in data); tots = sum(d[1] for d in data)
print(f"total warp-inst {tot}  samples {tots}  sass lines {len(data)}")
segs = []; prev = None; start = 0; acc = accs = 0
for i, d in enumerate(data):
    if prev is None or abs(d[2] - prev) > 0.03 * max(prev, 1):
        if prev is not None: segs.append((start, i - 1, prev, acc, accs))
        start, acc, accs = i, 0, 0
    prev = d[2]; acc += d[2]; accs += d[1]
segs.append((start, len(data) - 1, prev, acc, accs))
for s in segs:
    if s[3] > 0.01 * tot:
        print(f"sass {s[0]:5d}-{s[1]:5d} n={s[1]-s[0]+1:4d} exec={s[2]:>11d} inst%={100*s[3]/tot:5.1f} samp%={100*s[4]/max(tots,1):5.1f} thr={data[s[0]][3]:4.1f} | {data[s[0]][0][:70]}")
