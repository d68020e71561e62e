"""Sum an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: count, mean us, share."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) > mv:
        try:
            agg[r[kn]].append(float(r[mv].replace(",", "")))
        except ValueError:
            pass
tot = sum(sum(v) for k, v in agg.items() if "pbr::" in k) or 1.0
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    if "pbr::" in k:
        print(f"{len(v):5d} x {sum(v) / len(v) / 1000:10.2f} us  {100 * sum(v) / tot:5.1f}%  {k[:100]}")
