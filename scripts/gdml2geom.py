#!/usr/bin/env python
"""GDML file -> persisted geometry directory (<out>/CSGFoundry/{solid,prim,node,tran,itra,inst,plan}.npy + SSim/stree/standard/
{bnd,optical,icdf}.npy, the layout CSGFoundry::save_ writes, CSG/CSGFoundry.cc:2768-2802), which the C++ drivers take with -g:

    python scripts/gdml2geom.py tests/golden/mini_detector.gdml /tmp/mini [--freq-cut 500]
    eic-opticks_b200/apps/PhoxPhotonFileSource -g /tmp/mini -p photons.txt -o opticks_hits_output.txt

This is the step the reference apps do in process with Geant4 (G4GDMLParser + G4CXOpticks::SetGeometry, src/GPUPhotonFileSource.cpp).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("gdml")
    ap.add_argument("out")
    ap.add_argument("--freq-cut", type=int, default=None, help="stree::FREQ_CUT: subtrees repeated at least this often are instanced")
    a = ap.parse_args(argv)
    from eic_opticks_b200 import gdml, foundry
    g = gdml.translate(a.gdml, freq_cut=a.freq_cut)
    foundry.save_geometry(g, a.out)
    fd = g["foundry"]
    print("gdml2geom: %s -> %s : %d solids, %d prims, %d nodes, %d instances, %d boundaries, %d instanced factors%s"
          % (a.gdml, a.out, len(fd["solid"]), len(fd["prim"]), len(fd["node"]), len(fd["inst"]), len(g["bnd_names"]), g["num_factor"],
             ", scintillator %s" % g["scintillator"] if g.get("scintillator") else ""))
    return 0


if __name__ == "__main__":
    sys.exit(main())
