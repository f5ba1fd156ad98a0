import sys, time
sys.path.insert(0, ".")
import numpy as np
import eic_opticks_b200 as ph
from eic_opticks_b200 import workloads
for wl in ("pmt_wall_torch", "boolean_zoo_torch", "scintillator_tank", "raindrop_cerenkov", "sipm8x8_scint"):
    for n in (100000, 250000, 500000, 1000000, 2000000):
        w = workloads.WORKLOADS[wl](num_photon=n)
        g = w["geom"]
        row = []
        for mode in (ph.KERNEL_PERSISTENT, ph.KERNEL_WAVEFRONT):
            sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"], g["icdf"], event_mode=ph.MODE_MINIMAL, kernel_mode=mode, **w["config"])
            ts, ks = [], []
            for k in range(8):
                t0 = time.perf_counter()
                sim.simulate_np(w["gensteps"], k, w["input_photons"])
                ts.append(time.perf_counter() - t0)
                ks.append(sim.stats()["simulate_kernel_seconds"])
            row.append((np.median(ts[2:]) * 1e3, np.median(ks[2:]) * 1e3))
            sim.close()
        print("%-18s n %8d  persistent wall %8.3f ms loop %8.3f ms | wavefront wall %8.3f ms loop %8.3f ms" % (wl, n, row[0][0], row[0][1], row[1][0], row[1][1]), flush=True)
