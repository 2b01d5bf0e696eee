"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...`):
    python tools/launch_summary.py x.csv > profiles/rNN_launches_summary.csv"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"^void ", "", r[ki].split("(")[0])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi])
tot = sum(v[1] for v in agg.values())
print("kernel,launches,total_ns,mean_ns,share")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{n},{t:.0f},{t / n:.0f},{t / tot:.4f}")
