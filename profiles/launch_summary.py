"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, mean and share per kernel (and grid).
usage: launch_summary.py file.csv [skip_first_fraction]"""
import collections, csv, re, sys
f = sys.argv[1]
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
lines = [l for l in open(f) if not l.startswith("==")]
rows = []
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v * 1e6 if u == "s" else v
        rows.append((r["Kernel Name"], v, r.get("Grid Size", "")))
tail = rows[int(len(rows) * skip):]
agg = collections.defaultdict(list)
for n, v, g in tail:
    agg[re.sub(r"\(.*", "", n)[:64] + " " + g].append(v)
tot = sum(v for _, v, _ in tail)
print(f"{f}: {len(rows)} launches, {len(tail)} summarised, {tot:.1f} us total")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"  {k:88s} n={len(v):4d} mean={sum(v) / len(v):8.2f} us  sum={sum(v):9.1f} ({100 * sum(v) / tot:.1f}%)")
