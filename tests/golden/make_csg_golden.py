"""Generates tests/golden/csg_rays.npz: (t, normal) answers of the REFERENCE's CSG intersect headers
(CSG/csg_intersect_{leaf,node,tree}.h compiled for the host into oracle/_ref/libcsgref.so) for
seeded rays against every prim of the boolean zoo.  Run in the build container (needs /root/reference):
    python tests/golden/make_csg_golden.py
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import eic_opticks_b200 as ph
from _ref import RefCSGHost

def rays(center, n, seed):
    rng = np.random.default_rng(seed)
    o = (center + rng.uniform(-300, 300, (n, 3))).astype(np.float32)
    o[::3] = (center + rng.uniform(-60, 60, (len(o[::3]), 3))).astype(np.float32)       # a third start inside-ish
    t = (center + rng.uniform(-100, 100, (n, 3))).astype(np.float32)
    d = t - o
    d /= np.linalg.norm(d, axis=1)[:, None]
    d = d.astype(np.float32)
    d[::50] = np.array([0, 0, 1], dtype=np.float32)                                   # axis-parallel rays
    d[25::50] = np.array([1, 0, 0], dtype=np.float32)
    tmin = np.where(np.arange(n) % 2 == 0, 0.05, 0.0).astype(np.float32)
    return o, d, tmin

if __name__ == "__main__":
    g = ph.geometries.boolean_zoo(); fd = g["foundry"]
    ref = RefCSGHost()
    out = {}
    n = 400
    for pi in range(1, len(fd["prim"])):
        c = g["shape_centers"][pi - 2] if pi >= 2 else np.zeros(3, dtype=np.float32)
        o, d, tmin = rays(c, n, 100 + pi)
        isect, valid = ref.intersect_prim_batch(fd, pi, o, d, tmin)
        out["o_%d" % pi], out["d_%d" % pi], out["tmin_%d" % pi] = o, d, tmin
        out["isect_%d" % pi], out["valid_%d" % pi] = isect, valid
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "csg_rays.npz"), **out)
    print("wrote csg_rays.npz", sum(v.nbytes for v in out.values()))
