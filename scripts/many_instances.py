"""BVH build time and event rate with many instances (pmt_wall nx x ny)"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import eic_opticks_b200 as ph
from eic_opticks_b200 import workloads
for nx in (100, 300, 700):
    w = workloads.pmt_wall_torch(num_photon=2_000_000, nx=nx, ny=nx)
    g = w["geom"]
    t0 = time.perf_counter()
    sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"], g["icdf"], event_mode=ph.MODE_MINIMAL, **w["config"])
    t1 = time.perf_counter()
    ts = []
    for k in range(4):
        t = time.perf_counter(); h = sim.simulate_np(w["gensteps"], k, w["input_photons"]); ts.append(time.perf_counter() - t)
    st = sim.stats()
    # BVH vs brute on a sample of rays
    rng = np.random.default_rng(1)
    o = np.zeros((20000, 3), dtype=np.float32); o[:, :2] = rng.uniform(-nx * 120, nx * 120, (20000, 2)); o[:, 2] = 900
    d = rng.normal(size=(20000, 3)); d[:, 2] = -np.abs(d[:, 2]) - 0.5; d = (d / np.linalg.norm(d, axis=1)[:, None]).astype(np.float32)
    a = sim.intersect(o, d, 0.05, ph.ACCEL_BVH)
    same = None
    if nx <= 300:
        b = sim.intersect(o, d, 0.05, ph.ACCEL_BRUTE)
        same = float((a.view(np.uint32)[:, 1, 2:] == b.view(np.uint32)[:, 1, 2:]).all(axis=1).mean())
    print("instances %7d  create %.2f s  event %.1f ms (loop %.1f ms) hits %d  bvh==brute %s" % (nx * nx, t1 - t0, min(ts) * 1e3, st["simulate_kernel_seconds"] * 1e3, len(h), same), flush=True)
    sim.close()
