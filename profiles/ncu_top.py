"""Summarise an ncu report: `python profiles/ncu_top.py report.ncu-rep [n]` prints the key raw
metrics and the top-n SASS lines by stall samples (source page)."""
import csv, subprocess, sys, io

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + __import__("os").environ.get("NCU_FILTER", "").split(), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__throughput.avg.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread",
        "launch__occupancy_limit", "smsp__average_warps_issue_stalled_long_scoreboard_per",
        "smsp__average_warps_issue_stalled_barrier_per", "smsp__average_warps_issue_stalled_short_scoreboard_per",
        "smsp__average_warps_issue_stalled_wait_per", "smsp__average_warps_issue_stalled_not_selected_per",
        "smsp__average_warps_issue_stalled_math_pipe", "smsp__average_warps_issue_stalled_lg_throttle",
        "smsp__average_warps_issue_stalled_mio_throttle", "smsp__average_warps_issue_stalled_branch_resolving",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_bytes.sum ", "l1tex__data_bank_conflicts_pipe_lsu"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("== kernel:", name[:90])
    for h, u, v in zip(hdr, units, r):
        if any(h.startswith(w.strip()) for w in WANT) and "per_second" not in h and "pct_of_peak_sustained_elapsed" not in h.replace("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "").replace("sm__throughput.avg.pct_of_peak_sustained_elapsed", ""):
            print(f"  {h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + __import__("os").environ.get("NCU_FILTER", "").split(), capture_output=True, text=True).stdout
lines = src.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
data = []
for k, r in enumerate(rows[1:]):
    try:
        data.append((int(r[si]), int(r[ii]), int(r[ti]), k, r[1].strip()))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print(f"-- {len(data)} SASS lines, {tot} stall samples, {sum(d[1] for d in data)} warp instructions; top {n} by samples:")
for s, i, t, k, txt in sorted(data, reverse=True)[:n]:
    print(f"  {100.0 * s / max(tot, 1):5.1f}%  line {k:5d}  inst {i:9d}  thr/inst {t / max(i, 1):5.1f}  {txt[:100]}")
