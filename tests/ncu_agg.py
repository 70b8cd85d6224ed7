"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) by kernel name."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    agg[row["Kernel Name"].split("(")[0]].append(v)
tot = sum(sum(v) for v in agg.values())
print(f"total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:64]:64s} n={len(v):4d} mean={sum(v)/len(v):9.1f}us total={sum(v):10.1f}us {100*sum(v)/tot:5.1f}%")
