#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU): key raw metrics + instruction/stall histogram per SASS block.
usage: python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [block]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; W = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio']
for w in want:
    if w in h:
        i = h.index(w); print(f"{w:95s} {rows[1][i]:12s} {[r[i][:50] for r in rows[2:]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
if hi:
    h = rows[hi[0]]; end = hi[1] - 1 if len(hi) > 1 else len(rows)
    body = rows[hi[0] + 1:end]
    ci, cs, sc = h.index('Instructions Executed'), h.index('# Samples'), h.index('Source')
    tot_i = sum(int(r[ci]) for r in body if r[ci].isdigit()); tot_s = sum(int(r[cs]) for r in body if r[cs].isdigit())
    print(f"SASS instrs {len(body)}  executed {tot_i/1e6:.1f}M  samples {tot_s}")
    for s in range(0, len(body), W):
        blk = body[s:s + W]
        ins = sum(int(r[ci]) for r in blk); smp = sum(int(r[cs]) for r in blk)
        if ins == 0 and smp == 0: continue
        ops = collections.Counter((r[sc].split()[1] if r[sc].strip().startswith('@') else r[sc].split()[0]).split('.')[0] for r in blk)
        print(f"{s:5d} inst={100*ins/tot_i:5.1f}% samples={100*smp/max(1,tot_s):5.1f}%  " + ' '.join(f"{k}:{v}" for k, v in ops.most_common(9)))
