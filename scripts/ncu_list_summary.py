"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import collections, csv, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
    name = re.sub(r"^void ", "", row["Kernel Name"])
    name = re.sub(r"\(.*", "", name)[:100]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
print(f"{n} launches, {tot / 1e3:.3f} ms total (ncu times are cold-cache and serialised: compare SHARES)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{v[1] / tot * 100:5.1f}%  {v[0]:4d}x  avg {v[1] / v[0]:9.1f} us  {k}")
