#!/usr/bin/env python
"""Bounce-loop kernels of the LAST event in an ncu launch list: per-launch durations in order, and the share of the event spent in
launches shorter than a threshold (the tail of tiny lists): python scripts/launch_tail.py <csv> [us]"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 250.0
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
seq = []
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    seq.append((r[ki].split("(")[0].replace("void ", "")[:24], v))
last = max(i for i, (k, v) in enumerate(seq) if k.startswith("k_wf_generate"))
ev = [(k, v) for k, v in seq[last:] if k.startswith("k_wf") or k.startswith("k_hit")]
tot = sum(v for _, v in ev)
pairs = [v for k, v in ev if k.startswith("k_wf_trace")]
tail = sum(v for k, v in ev[1:] if (k.startswith("k_wf_trace") or k.startswith("k_wf_propagate")) and v < thr)
print("event kernels %.0f us; trace launches: %s" % (tot, " ".join("%.0f" % v for v in pairs)))
print("launches under %.0f us: %.0f us = %.1f %% of the event" % (thr, tail, 100 * tail / tot))
