"""`python profiles/ncu_opmix.py report.ncu-rep [units]` -- dynamic SASS opcode mix of the first kernel in an
ncu report, per `units` (default 313290 = 32-triangle chunks of the bench scene)."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; units = float(sys.argv[2]) if len(sys.argv) > 2 else 313290.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + __import__("os").environ.get("NCU_FILTER", "").split(), capture_output=True, text=True).stdout
lines = src.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
ii = rows[0].index("Instructions Executed")
cnt = collections.Counter(); tot = 0
for r in rows[1:]:
    try: n = int(r[ii])
    except (ValueError, IndexError): continue
    toks = r[1].split()
    op = (toks[1] if toks and toks[0].startswith("@") else toks[0] if toks else "?").split(".")[0]
    cnt[op] += n; tot += n
for op, n in cnt.most_common(28):
    print(f"{op:10s} {n / units:7.1f}  {100 * n / tot:5.1f}%")
print(f"total {tot / units:.1f} warp instructions per unit")
