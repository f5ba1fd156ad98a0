"""small events of several workloads in both loop forms and several event modes, for compute-sanitizer (memcheck / racecheck)"""
import sys
sys.path.insert(0, ".")
import numpy as np
import eic_opticks_b200 as ph
from eic_opticks_b200 import workloads
for name, kw in (("sipm8x8_scint", dict(num_photon=3000, photons_per_genstep=100)), ("raindrop_cerenkov", dict(num_photon=2000)),
                 ("pmt_wall_torch", dict(num_photon=2000, nx=6, ny=6)), ("boolean_zoo_torch", dict(num_photon=2000)), ("box_maze_photons", dict(num_photon=2000))):
    w = workloads.WORKLOADS[name](**kw); g = w["geom"]
    for mode in (ph.KERNEL_WAVEFRONT, ph.KERNEL_PERSISTENT):
        for em, extra in ((ph.MODE_MINIMAL, {}), (ph.MODE_DEBUGLITE, dict(max_record=6))):
            cfg = dict(w["config"]); cfg.update(extra)
            sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"], g["icdf"], event_mode=em, kernel_mode=mode, **cfg)
            h = sim.simulate_np(w["gensteps"], 1, w["input_photons"])
            print(name, mode, em, len(h), sim.stats()["num_kernel"], flush=True)
            sim.close()
