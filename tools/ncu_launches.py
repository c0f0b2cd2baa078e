"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (+grid) -> markdown table."""
import csv, re, sys, collections
path = sys.argv[1]; by_grid = len(sys.argv) > 2
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict(); tot = 0.0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
    tmpl = re.search(r"<[^>]*>", row["Kernel Name"])
    key = name + (" grid=" + row["Grid Size"] if by_grid else "")
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"total {tot/1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches\n")
print("| kernel | launches | us | share |\n|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| {k} | {c} | {t:.0f} | {t/tot*100:.1f}% |")
