"""Write what the C++ drivers read: python scripts/make_event_files.py <workload> <photons> <outdir>
-> <outdir>/geom (persisted CSGFoundry + SSim directory), <outdir>/gs.npy (quad6 gensteps), <outdir>/ip.npy (input photons, if any)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from eic_opticks_b200 import workloads, foundry as F

wl, n, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
os.makedirs(out, exist_ok=True)
w = workloads.WORKLOADS[wl](num_photon=n)
F.save_geometry(w["geom"], os.path.join(out, "geom"))
np.save(os.path.join(out, "gs.npy"), np.ascontiguousarray(w["gensteps"], dtype=np.float32))
if w["input_photons"] is not None:
    np.save(os.path.join(out, "ip.npy"), np.ascontiguousarray(w["input_photons"], dtype=np.float32))
print(wl, w["num_photon"], "photons,", len(w["gensteps"]), "gensteps, max_bounce", w["config"].get("max_bounce", 31))
