"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = row["Kernel Name"].split("(")[0][-60:]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(row["Metric Value"].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"| launches | total ms | share | avg us | kernel |\n|---:|---:|---:|---:|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {a[0]} | {a[1]/1e6:.3f} | {100*a[1]/tot:.1f}% | {a[1]/a[0]/1e3:.1f} | `{k}` |")
