import sys, numpy as np
sys.path.insert(0,"."); sys.path.insert(0,"tests")
import eic_opticks_b200 as ph
from eic_opticks_b200 import workloads
w = workloads.sipm8x8_scint(num_photon=1000, photons_per_genstep=100)
g = w["geom"]; sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"], g["icdf"])
rng = np.random.default_rng(1); n = 8_000_000
cc = g["crystal_centers"]
o = (cc[rng.integers(0,64,n)] + rng.uniform(-0.99,0.99,(n,3))*np.array([1,1,3.99])).astype(np.float32)
d = rng.normal(size=(n,3)); d=(d/np.linalg.norm(d,axis=1)[:,None]).astype(np.float32)
for rep in range(3):
    out = sim.intersect(o, d, 0.05, ph.ACCEL_BVH)
    st = sim.stats()
    print("k_intersect: %.2f ms for %d rays -> %.2f Grays/s" % (st["simulate_kernel_seconds"]*1e3, n, n/st["simulate_kernel_seconds"]/1e9))
# coherent order: sort rays by crystal
