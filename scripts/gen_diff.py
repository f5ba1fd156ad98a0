import sys, os, subprocess, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
code = r'''
import sys; sys.path.insert(0, ".")
import numpy as np
import eic_opticks_b200 as ph
from eic_opticks_b200 import workloads
w = workloads.sipm8x8_scint(num_photon=30000, photons_per_genstep=100); g = w["geom"]
sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"], g["icdf"], event_mode=ph.MODE_DEBUGLITE, kernel_mode=ph.KERNEL_WAVEFRONT, max_record=2, **w["config"])
sim.simulate_np(w["gensteps"], 0)
np.save(sys.argv[1], sim.get_array("record")[:, 0].copy())
'''
for name, lib in (("inline", ""), ("outofline", "/root/repo/tune/oldgen.so")):
    env = dict(os.environ)
    if lib: env["PHOX_LIB"] = lib
    subprocess.run([sys.executable, "-c", code, "/tmp/gen_%s.npy" % name], env=env, check=True)
a = np.load("/tmp/gen_inline.npy").view(np.uint32).reshape(-1, 16); b = np.load("/tmp/gen_outofline.npy").view(np.uint32).reshape(-1, 16)
d = a != b
print("photons differing", d.any(axis=1).sum(), "of", len(a)); print("per word", d.sum(axis=0).tolist())
