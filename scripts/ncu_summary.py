"""Summarise an .ncu-rep (raw page + hottest source lines) into a text file for profiles/."""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
for vals in rows[2:]:
    print("=" * 100)
    for h, u, v in zip(hdr, units, vals):
        if h in want or any(h == w for w in want):
            print(f"{h:90s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = next(i for i, r in enumerate(rows[:10]) if "Source" in r)
hdr = rows[h]
ci = {n: k for k, n in enumerate(hdr)}
data = []
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        data.append((float(r[ci["# Samples"]] or 0), r))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1
print("=" * 100)
print(f"hottest SASS instructions by warp-stall samples (total {tot:.0f}); top two stall reasons each")
reasons = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: 0.0 for n in reasons}
for samp, r in data:
    for n in reasons:
        agg[n] += float(r[ci[n]] or 0)
print("kernel-wide stall samples: " + ", ".join(f"{n[6:]}={v / tot * 100:.1f}%" for n, v in sorted(agg.items(), key=lambda x: -x[1]) if v / tot > 0.005))
for samp, r in sorted(data, key=lambda x: -x[0])[:topn]:
    top = sorted(((float(r[ci[n]] or 0), n[6:]) for n in reasons), reverse=True)[:2]
    why = " ".join(f"{n}={v:.0f}" for v, n in top if v > 0)
    print(f"{samp / tot * 100:5.1f}%  execs={r[ci['Instructions Executed']]:>10}  {r[ci['Source']][:70]:70s} {why}")
