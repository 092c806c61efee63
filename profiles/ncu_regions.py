"""`python profiles/ncu_regions.py report.ncu-rep` -- instruction counts of consecutive SASS regions
with the same execution count (cheap way to see which phase of a kernel issues how much)."""
import csv, subprocess, io, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + __import__("os").environ.get("NCU_FILTER", "").split(), capture_output=True, text=True).stdout
lines = src.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
data = []
for k, r in enumerate(rows[1:]):
    try:
        data.append((k, int(r[ii]), int(r[ti]), int(r[si]), r[1].strip()))
    except (ValueError, IndexError):
        pass
cur = None; out = []
for k, i, t, s, txt in data:
    if cur is None or abs(i - cur) > 0.02 * max(cur, 1):
        if cur is not None:
            out.append((first, k - 1, cur, n, acc, samp, thr_acc))
        cur, acc, n, samp, first, thr_acc = i, 0, 0, 0, k, 0
    acc += i; n += 1; samp += s; thr_acc += t
out.append((first, data[-1][0], cur, n, acc, samp, thr_acc))
tot = sum(o[4] for o in out)
for o in out:
    if o[4] > thr * tot:
        print(f"lines {o[0]:5d}-{o[1]:5d} count {o[2]:9d} x{o[3]:4d} = {o[4] / 1e6:7.2f}M inst ({100 * o[4] / tot:4.1f}%) "
              f"thr/inst {o[6] / max(o[4], 1):5.1f} samples {o[5]}")
print(f"total {tot / 1e6:.2f}M warp instructions")
