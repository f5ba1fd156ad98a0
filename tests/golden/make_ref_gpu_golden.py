"""Generates tests/golden/ref_gpu_<workload>_<variant>.npz ON THE GPU BOX: photon (+ seq) arrays
produced by the REFERENCE's device headers (oracle/_ref/libphoxref_*.so = qsim.h, qbnd.h, qscint.h,
qcerenkov.h, storch.h, csg_intersect_*.h compiled unmodified, brute-force traversal) for small seeded
workloads.  They pin the CPU oracle without a GPU (tests/test_oracle.py).
    gpurun -- 'python tests/golden/make_ref_gpu_golden.py gpurun_out/golden'   then copy into tests/golden/
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from eic_opticks_b200 import workloads
from _ref import RefGPU

CASES = [("sipm8x8_scint", dict(num_photon=3000, photons_per_genstep=50)),
         ("raindrop_cerenkov", dict(num_photon=2000, photons_per_genstep=100)),
         ("sphere_leak_torch", dict(num_photon=1500)),
         ("pmt_wall_torch", dict(num_photon=3000)),
         ("boolean_zoo_torch", dict(num_photon=4000)),
         ("scintillator_tank", dict(num_photon=4000, photons_per_genstep=100))]

if __name__ == "__main__":
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for variant in ("debugtag", "production"):
        ref = RefGPU(variant)
        for name, kw in CASES:
            if variant == "production" and name not in ("sipm8x8_scint", "boolean_zoo_torch", "scintillator_tank"):
                continue
            w = workloads.WORKLOADS[name](**kw)
            r = ref.simulate(w["geom"], w["gensteps"], w["input_photons"], max_bounce=w["config"].get("max_bounce", 31))
            d = dict(workload=name, num_photon=kw["num_photon"], debug_tag=(variant == "debugtag"), photon=r["photon"])
            if "photons_per_genstep" in kw:
                d["photons_per_genstep"] = kw["photons_per_genstep"]
            if r["seq"] is not None:
                d["seq"] = r["seq"]
            np.savez_compressed(os.path.join(out_dir, "ref_gpu_%s_%s.npz" % (name, variant)), **d)
            print("wrote", name, variant, len(r["photon"]), "rays", r["nray"])
