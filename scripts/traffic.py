#!/usr/bin/env python
"""DRAM bytes per photon / per ray of the two bounce-loop kernels from their ncu captures (bounce 10 of a 4 M-photon event) and the
event's live counts: python scripts/traffic.py <ncu_wf_propagate.ncu-rep> <ncu_wf_trace.ncu-rep> <live_counts.json> > profiles/traffic_rX.json"""
import csv, io, json, subprocess, sys

def metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    names, units, vals = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(names)}
    def get(n):
        v = float(vals[col[n]].replace(",", "")); u = units[col[n]]
        return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)
    return get("dram__bytes_read.sum"), get("dram__bytes_write.sum"), get("gpu__time_duration.sum")

prop, trace, lc = sys.argv[1:4]
lc = json.loads(open(lc).read().strip().splitlines()[-1])
b = 10
live, home = lc["rays_per_bounce"][b], lc["home_rays_per_bounce"][b]
pr, pw, pt = metrics(prop); tr, tw, tt = metrics(trace)
print(json.dumps({
    "source": "ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 10 -c 1, bench.py --no-cpu-baseline --steps 1 --warmup 1 --photons 4000000: "
              "the 11th launch of each kernel = bounce 10 of event id 1; live photons / home-settled rays of that bounce from scripts/live_counts.py 4000000 1",
    "bounce": b, "live_photons_in_captured_launch": live, "home_settled_rays_of_that_bounce": home, "pending_rays_in_captured_launch": live - home,
    "k_wf_propagate": {"dram_read_bytes": pr, "dram_write_bytes": pw, "dram_bytes_per_photon": (pr + pw) / live, "duration_us": pt},
    "k_wf_trace": {"dram_read_bytes": tr, "dram_write_bytes": tw, "dram_bytes_per_ray": (tr + tw) / (live - home), "duration_us": tt},
}, indent=1))
