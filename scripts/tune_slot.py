"""bounce-loop device time of the default workload as a function of photons per launch (L2-sized slices?)"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import eic_opticks_b200 as ph
from eic_opticks_b200 import workloads
w = workloads.sipm8x8_scint(num_photon=12_500_000)
g = w["geom"]
sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"], g["icdf"], event_mode=ph.MODE_MINIMAL, **w["config"])
for ms in (0, 125_000, 250_000, 500_000, 1_000_000, 2_000_000, 4_000_000):
    sim.set_config(max_slot=ms)
    for rep in range(3):
        t0 = time.perf_counter()
        h = sim.simulate_np(w["gensteps"], 1 + rep)
        dt = time.perf_counter() - t0
        st = sim.stats()
    print("max_slot %8d launches %3d loop_ms %7.2f compact_ms %6.2f wall_ms %7.2f hits %d" % (ms, st["num_launch"], st["simulate_kernel_seconds"] * 1e3, st["compact_kernel_seconds"] * 1e3, dt * 1e3, len(h)), flush=True)
