"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  Usage: python tools/launch_summary.py launches.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("<unnamed>::", "").replace("void ", "")
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print("%-62s %6s %12s %10s %10s %6s" % ("kernel", "n", "total us", "mean us", "max us", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s %6d %12.1f %10.1f %10.1f %5.1f%%" % (k[:62], a[0], a[1], a[1] / a[0], a[2], 100 * a[1] / tot))
print("%-62s %6d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))
