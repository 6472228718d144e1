"""Opcode histogram (+ stall samples) of one kernel from an ncu report's SASS source page.
    python scripts/ncu_opcodes.py report.ncu-rep kernel_regex [--hot N]"""
import csv, subprocess, sys
from collections import Counter
rep, rx = sys.argv[1], sys.argv[2]
hot = int(sys.argv[sys.argv.index("--hot") + 1]) if "--hot" in sys.argv else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ci, ce, cs = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
data = []
for r in rows[hi + 1:]:
    if len(r) <= cs or r[0] == "Kernel Name" or r[0] == "Address":
        if r and r[0] == "Kernel Name": break      # first launch only
        continue
    try:
        data.append((int(r[ce]), int(r[cs]), r[ci].strip()))
    except ValueError:
        pass
tot = sum(e for e, _, _ in data); ts = sum(s for _, s, _ in data)
c, st = Counter(), Counter()
for e, s, src in data:
    tk = src.split()
    op = (tk[1] if tk[0].startswith("@") else tk[0]).split(".")[0]
    c[op] += e; st[op] += s
print(f"{len(data)} SASS lines, {tot} warp instructions, {ts} stall samples")
for op, n in c.most_common(22):
    print(f"  {op:10s} {n:12d} {100 * n / tot:5.1f} %   stalls {100 * st[op] / max(ts, 1):5.1f} %")
if hot:
    print("hottest lines by stall samples:")
    for e, s, src in sorted(data, key=lambda x: -x[1])[:hot]:
        print(f"  {s:7d} {e:10d}  {src[:100]}")
