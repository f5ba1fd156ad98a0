#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file <csv>): python scripts/launch_summary.py <csv>"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    k = r[ki].split("(")[0][-50:]; agg[k][0] += 1; agg[k][1] += v
tot = sum(t for _, t in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print("%-50s n %4d total %9.1f us (%4.1f %%) avg %8.1f us" % (k, n, t, 100 * t / tot, t / n))
