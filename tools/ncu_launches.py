"""Kernel share of a command from an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...).
usage: python tools/ncu_launches.py <launches.csv> "<command>" """
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[mv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1e-6)
    a = agg.setdefault(r[kn], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"kernel share of `{sys.argv[2] if len(sys.argv) > 2 else '?'}` (ncu gpu__time_duration.sum, cold-cache serialised launches)")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:4d} launches {ms:10.3f} ms {ms / tot * 100:5.1f}%  {k[:100]}")
