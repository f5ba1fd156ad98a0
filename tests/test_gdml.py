"""GDML translator (eic-opticks_b200/gdml.py): against the hand-built geometries that follow the reference's
GDML files volume by volume (needs /root/reference for those files; skipped on the GPU box), and on a small
GDML written for this repository (loop, boolean with displaced rhs, rotated placement, skin/border/implicit surfaces)."""
import os

import numpy as np
import pytest

import eic_opticks_b200 as ph
from eic_opticks_b200 import gdml, foundry as F, tables as T

REF_GEOM = "/root/reference/tests/geom"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same_geometry(a, b):
    for k in ("solid", "prim", "node", "itra", "tran", "inst"):
        assert a["foundry"][k].shape == b["foundry"][k].shape, k
    assert (a["foundry"]["node"].view(np.uint32)[:, :, :] == b["foundry"]["node"].view(np.uint32)).all()
    assert (a["foundry"]["prim"].view(np.uint32) == b["foundry"]["prim"].view(np.uint32)).all()
    assert np.allclose(a["foundry"]["itra"], b["foundry"]["itra"], atol=1e-6)
    assert a["bnd_names"] == b["bnd_names"]
    assert np.allclose(a["bnd"], b["bnd"], rtol=1e-6, atol=1e-6)
    assert (a["optical"][:, 1] == b["optical"][:, 1]).all()            # ems; .x is only a table index (the GDML may list unused surfaces)
    assert ((a["optical"][:, 0] > 0) == (b["optical"][:, 0] > 0)).all()


@pytest.mark.parametrize("fname,builder", [("opticks_raindrop.gdml", "raindrop"), ("sphere_leak.gdml", "sphere_leak"),
                                            ("8x8SiPM_w_CSI_optial_grease.gdml", "sipm8x8")])
def test_translation_of_reference_gdml_matches_hand_built(fname, builder):
    path = os.path.join(REF_GEOM, fname)
    if not os.path.exists(path):
        pytest.skip("reference GDML files are only present in the build container")
    t = gdml.translate(path)
    h = getattr(ph.geometries, builder)()
    _same_geometry(t, h)
    if builder == "sipm8x8":
        assert t["icdf"].shape == (3, 4096) and np.allclose(t["icdf"], h["icdf"], rtol=1e-5)
        assert t["scintillator"] == "Crystal" and abs(t["scintillation_time"] - 21.5) < 1e-9
        assert len(t["sensitive_prims"]) == 64


@pytest.mark.parametrize("fname", ["raindrop.gdml", "basic_detector.gdml", "opticks_raindrop_with_scintillation.gdml"])
def test_other_reference_gdml_files_translate(fname):
    path = os.path.join(REF_GEOM, fname)
    if not os.path.exists(path):
        pytest.skip("reference GDML files are only present in the build container")
    t = gdml.translate(path)
    assert len(t["foundry"]["prim"]) >= 4 and t["bnd"].shape[0] == len(t["bnd_names"])
    assert t["bnd_names"][0].split("/")[0] == t["bnd_names"][0].split("/")[3]           # world: omat == imat


def test_mini_detector_translation():
    t = gdml.translate(os.path.join(GOLD, "mini_detector.gdml"))
    fd = t["foundry"]
    names = t["bnd_names"]
    assert t["prim_names"] == ["WorldVol_PV", "plate_pv", "pipe_pv", "block_0", "block_1", "block_2", "block_3"]
    # steel has no RINDEX: water->steel gets an implicit perfect absorber on the outer side only
    assert "Water/Implicit_RINDEX_NoRINDEX_WorldVol_PV_plate_pv//Steel" in names
    # directional border surface: only for photons going world -> pipe (osur), nothing on the way out
    assert "Water/WaterToPipe//Glass" in names
    assert "Water/BlockSkin/BlockSkin/Glass" in names
    i = names.index("Water/WaterToPipe//Glass")
    assert np.allclose(t["bnd"][i, 1, 0, 0], [0.0, 0.1, 0.0, 0.9], atol=1e-6)              # ground finish -> diffuse
    j = names.index("Water/BlockSkin/BlockSkin/Glass")
    assert t["bnd"][j, 1, 0, :, 0].max() <= 0.3 + 1e-6 and t["bnd"][j, 1, 0, :, 0].min() >= 0.2 - 1e-6   # sensor: detect = EFFICIENCY
    # units: plate is 20 cm x 20 cm x 2 cm at z = -150 mm
    prim = fd["prim"].reshape(-1, 16)
    assert np.allclose(prim[1, 8:14], [-100, -100, -160, 100, 100, -140], atol=1e-4)
    # pipe rotated by 90 deg about x: its axis (local z) lies along world y
    assert np.allclose(prim[2, 8:14], [-10, 80 - 0.4, -10, 10, 120 + 0.4, 10], atol=0.5) or np.allclose(prim[2, 11:14] - prim[2, 8:11], [20, 40, 20], atol=1.0)
    # tube with rmin: difference positivised into intersection with a complemented, 1 % longer inner cylinder
    nu = fd["node"].view(np.uint32).reshape(-1, 16)
    no = fd["prim"].view(np.int32).reshape(-1, 16)[2, 1]
    assert nu[no, 14] == F.CSG_INTERSECTION and nu[no + 2, 15] >> 31 == 1
    assert np.allclose(fd["node"].reshape(-1, 16)[no + 2, 4:6], [-20.2, 20.2], atol=1e-4)
    # loop placement expression -half*3 + k*pitch
    xs = [0.5 * (prim[3 + k, 8] + prim[3 + k, 11]) for k in range(4)]
    assert np.allclose(xs[1] - xs[0], 30.0, atol=1e-3) and np.allclose(prim[3, 10] + 12, 0.0, atol=1e-3) and np.allclose(prim[6, 10] + 12, 30.0, atol=1e-3)
    # dispersion: GROUPVEL derived from RINDEX lies below c/n
    w = names.index("Water///Water")
    assert (t["bnd"][w, 0, 1, :, 0] <= 299.792458 / 1.33 + 1e-3).all() and (t["bnd"][w, 0, 1, :, 0] > 150).all()


@pytest.mark.gpu
def test_mini_detector_runs_and_bvh_matches_brute():
    from eic_opticks_b200 import gensteps as G
    t = gdml.translate(os.path.join(GOLD, "mini_detector.gdml"))
    sim = ph.Simulator.Create(t["foundry"], t["bnd"], t["optical"], event_mode=ph.MODE_HITPHOTONSEQ)
    tor = dict(pos=[0, 0, 0], time=0.0, mom=[0, 0, 1.0], pol=[1, 0, 0], wavelength=430.0, radius=-180.0, numphoton=20000, type="sphere")
    gs = G.torch_genstep(tor)
    a = sim.simulate_np(gs, 0).copy(); sa = sim.get_array("seq").copy()
    sim.set_config(accel=ph.ACCEL_BRUTE)
    b = sim.simulate_np(gs, 0); sb = sim.get_array("seq")
    assert (sa == sb).all(axis=(1, 2)).mean() > 0.999
    assert len(a) > 0 and abs(len(a) - len(b)) <= 5
    fl = set()
    for s in sa[:, 0, 0]:
        v = int(s)
        while v:
            fl.add(v & 0xf); v >>= 4
    assert {3, 7, 8}.issubset(fl)          # TO, SD (block skins), SA (implicit absorber on the steel plate)
    sim.close()
