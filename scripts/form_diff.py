"""Which kernel pair disagrees, and where: for one workload, the photon / record arrays of persistent and wavefront form x production
and debug kernels; prints per pair the number of differing photons, the fields that differ and, from the debug records, the first step."""
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import eic_opticks_b200 as ph
from eic_opticks_b200 import workloads
name = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
w = workloads.WORKLOADS[name](num_photon=n); g = w["geom"]
out = {}
for mode in (ph.KERNEL_PERSISTENT, ph.KERNEL_WAVEFRONT):
    for em, extra in ((ph.MODE_HITPHOTON, {}), (ph.MODE_DEBUGLITE, dict(max_record=32))):
        kwc = dict(w["config"]); kwc.update(extra)
        sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"], g["icdf"], event_mode=em, kernel_mode=mode, **kwc)
        sim.simulate_np(w["gensteps"], 2, w["input_photons"])
        out[(mode, em)] = {"photon": sim.get_array("photon").copy()}
        if em == ph.MODE_DEBUGLITE: out[(mode, em)]["record"] = sim.get_array("record").copy()
        sim.close()
keys = list(out)
for i in range(len(keys)):
    for j in range(i + 1, len(keys)):
        a, b = out[keys[i]]["photon"].view(np.uint32), out[keys[j]]["photon"].view(np.uint32)
        d = (a != b).reshape(len(a), -1)
        bad = np.nonzero(d.any(axis=1))[0]
        print(keys[i], "vs", keys[j], ": differing photons", len(bad), "fields", np.nonzero(d.any(axis=0))[0].tolist())
        if len(bad) and "record" in out[keys[i]] and "record" in out[keys[j]]:
            ra, rb = out[keys[i]]["record"].view(np.uint32), out[keys[j]]["record"].view(np.uint32)
            for k in bad[:5]:
                dd = (ra[k] != rb[k]).reshape(ra.shape[1], -1)
                st = np.nonzero(dd.any(axis=1))[0]
                print("   photon", k, "first differing step", st[:1], "fields", np.nonzero(dd[st[0]])[0].tolist() if len(st) else None)
                if len(st):
                    s0 = st[0]
                    print("      A", out[keys[i]]["record"][k, s0].ravel()[:12], hex(ra[k, s0, 3, 0]), hex(ra[k, s0, 3, 3]))
                    print("      B", out[keys[j]]["record"][k, s0].ravel()[:12], hex(rb[k, s0, 3, 0]), hex(rb[k, s0, 3, 3]))
