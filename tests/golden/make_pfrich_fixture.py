"""tests/golden/pfrich_min_geometry.npz : the pfRICH geometry the reference ships (tests/geom/pfrich_min_FINAL.gdml: aerogel, gas vessel,
inner / outer mirrors, 64 sensor pyramids + absorbing edges; tubes, cones, booleans, G4Trap) translated by eic-opticks_b200/gdml.py into
the CSGFoundry + bnd / optical arrays.  Run in the build container, where /root/reference exists:

    python tests/golden/make_pfrich_fixture.py

The GPU box has no /root/reference, so the translated arrays travel as this fixture (170 kB)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eic_opticks_b200 import gdml  # noqa: E402

g = gdml.translate("/root/reference/tests/geom/pfrich_min_FINAL.gdml")
fd = g["foundry"]
out = os.path.join(ROOT, "tests", "golden", "pfrich_min_geometry.npz")
np.savez_compressed(out, bnd=g["bnd"].astype(np.float32), optical=g["optical"], bnd_names=np.array(g["bnd_names"]), prim_names=np.array(g["prim_names"]),
                    sensitive_prims=np.array(g["sensitive_prims"], dtype=np.int32), **{k: v for k, v in fd.items() if hasattr(v, "shape")})
print(out, os.path.getsize(out), "bytes;", len(fd["prim"]), "prims,", len(fd["node"]), "nodes,", len(g["bnd_names"]), "boundaries")
