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


@pytest.mark.parametrize("fname", ["raindrop.gdml", "basic_detector.gdml", "opticks_raindrop_with_scintillation.gdml", "pfrich_min_FINAL.gdml"])
def test_other_reference_gdml_files_translate(fname):
    path = os.path.join(REF_GEOM, fname)
    if not os.path.exists(path):
        pytest.skip("reference GDML files are only present in the build container")
    t = gdml.translate(path)
    assert len(t["foundry"]["prim"]) >= 4 and t["bnd"].shape[0] == len(t["bnd_names"])
    assert t["bnd_names"][0].split("/")[0] == t["bnd_names"][0].split("/")[3]           # world: omat == imat


def test_trap_becomes_a_convexpolyhedron_through_its_eight_vertices(tmp_path):
    """<trap> (G4Trap; tests/geom/pfrich_min_FINAL.gdml has eleven): six outward planes, each of the eight G4Trap vertices
    (G4Trap::MakePlanes layout from the halved GDML lengths) on exactly three of them and inside the rest"""
    import math
    z, th, phi, y1, x1, x2, a1, y2, x3, x4, a2 = 60.0, 0.2, 0.3, 40.0, 30.0, 36.0, 0.1, 28.0, 21.0, 25.2, 0.1      # top = 0.7 x bottom: planar side faces, as G4Trap demands
    gd = tmp_path / "trap.gdml"
    gd.write_text('''<?xml version="1.0"?><gdml><define><matrix name="RI" coldim="2" values="1.55*eV 1.5 6.2*eV 1.5"/></define><materials>
<material name="Vac"><D value="1e-25"/><property name="RINDEX" ref="RI"/></material><material name="Gl"><D value="2.2"/><property name="RINDEX" ref="RI"/></material></materials>
<solids><box name="w" x="500" y="500" z="500" lunit="mm"/>
<trap name="t" z="%g" theta="%g" phi="%g" y1="%g" x1="%g" x2="%g" alpha1="%g" y2="%g" x3="%g" x4="%g" alpha2="%g" lunit="mm" aunit="rad"/></solids>
<structure><volume name="tv"><materialref ref="Gl"/><solidref ref="t"/></volume>
<volume name="wv"><materialref ref="Vac"/><solidref ref="w"/><physvol name="tp"><volumeref ref="tv"/></physvol></volume></structure>
<setup name="Default" version="1.0"><world ref="wv"/></setup></gdml>''' % (z, th, phi, y1, x1, x2, a1, y2, x3, x4, a2))
    t = gdml.translate(str(gd))
    plan = t["foundry"]["plan"].reshape(-1, 4).astype(np.float64)
    assert len(plan) == 6
    dz, tc, ts = z / 2, math.tan(th) * math.cos(phi), math.tan(th) * math.sin(phi)
    verts = []
    for sz, dy, dxa, dxb, ta in ((-1, y1 / 2, x1 / 2, x2 / 2, math.tan(a1)), (1, y2 / 2, x3 / 2, x4 / 2, math.tan(a2))):
        for sy, dx in ((-1, dxa), (1, dxb)):
            for sx in (-1, 1):
                verts.append([sz * dz * tc + sy * dy * ta + sx * dx, sz * dz * ts + sy * dy, sz * dz])
    verts = np.array(verts)
    d = verts @ plan[:, :3].T - plan[:, 3]                 # signed distance of every vertex to every plane
    assert (d < 1e-4).all()                                # inside or on
    assert ((np.abs(d) < 1e-4).sum(axis=1) == 3).all()     # each vertex is the corner of three faces
    assert np.allclose(np.linalg.norm(plan[:, :3], axis=1), 1.0, atol=1e-6)
    prim = t["foundry"]["prim"].reshape(-1, 16)
    tp = prim[-1]
    assert np.allclose(tp[8:11], verts.min(axis=0), atol=1e-4) and np.allclose(tp[11:14], verts.max(axis=0), atol=1e-4)


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


def test_polycone_and_ellipsoid_follow_u4(tmp_path):
    """U4Polycone (u4/U4Polycone.h) and U4Solid::init_Ellipsoid restated: tree shape, nudges, and ray distances from the oracle"""
    from eic_opticks_b200 import gdml as GD, foundry as F
    from _ref import Oracle
    # outer: cylinder r50 z[-100,0], cone 50->20 z[0,60]; inner: r10 everywhere
    t = GD.polycone_tree([(10, 50, -100), (10, 50, 0), (10, 20, 60)])
    assert isinstance(t, F.Op) and t.typecode == F.CSG_DIFFERENCE
    outer, inner = t.left, t.right
    assert isinstance(outer, F.Op) and outer.typecode == F.CSG_UNION and isinstance(inner, F.Leaf) and inner.typecode == F.CSG_CYLINDER
    cyl, cone = outer.left, outer.right
    # joint at z = 0: both radii 50 -> "else" branch of ZNudgeOverlapJoint: the lower prim grows 1 mm upwards
    assert cyl.typecode == F.CSG_CYLINDER and tuple(cyl.param[3:6]) == (50.0, -100.0, 1.0)
    assert cone.typecode == F.CSG_CONE and tuple(cone.param[:4]) == (50.0, 0.0, 20.0, 60.0)
    assert tuple(inner.param[3:6]) == (10.0, -100.0, 60.0)            # single-radius inner: plain cylinder over the z range
    # varying inner radius: ends stick out by 1 mm, cone radii extrapolated along the slope
    t2 = GD.polycone_tree([(5, 50, 0), (15, 50, 100)])
    assert t2.left.typecode == F.CSG_CYLINDER and t2.right.typecode == F.CSG_CONE
    assert np.allclose(t2.right.param[:4], (4.9, -1.0, 15.1, 101.0))
    # descending z planes are reversed
    t3 = GD.polycone_tree([(0, 20, 60), (0, 50, 0), (0, 50, -100)])
    assert t3.typecode == F.CSG_UNION and t3.left.typecode == F.CSG_CYLINDER

    gd = tmp_path / "pc.gdml"
    gd.write_text("""<?xml version="1.0"?>
<gdml><define/><materials>
 <material name="Vac"><D value="1e-25"/><fraction n="1" ref="H"/></material><element name="H" formula="H" Z="1"><atom value="1"/></element>
</materials><solids>
 <box name="w" x="1000" y="1000" z="1000" lunit="mm"/>
 <polycone name="pc" startphi="0" deltaphi="360" aunit="deg" lunit="mm">
  <zplane rmin="10" rmax="50" z="-100"/><zplane rmin="10" rmax="50" z="0"/><zplane rmin="10" rmax="20" z="60"/></polycone>
 <ellipsoid name="el" ax="30" by="40" cz="50" zcut1="-20" zcut2="0" lunit="mm"/>
</solids><structure>
 <volume name="pcl"><materialref ref="Vac"/><solidref ref="pc"/></volume>
 <volume name="ell"><materialref ref="Vac"/><solidref ref="el"/></volume>
 <volume name="W"><materialref ref="Vac"/><solidref ref="w"/>
  <physvol name="a"><volumeref ref="pcl"/></physvol>
  <physvol name="b"><volumeref ref="ell"/><position x="200" y="0" z="0" unit="mm"/></physvol></volume>
</structure><setup name="Default" version="1.0"><world ref="W"/></setup></gdml>""")
    geom = GD.translate(str(gd))
    o = np.array([[200, 400, -10], [200, 0, -300]], dtype=np.float32)
    d = np.array([[0, -1, 0], [0, 0, 1]], dtype=np.float32)
    el = Oracle().intersect(geom, o, d)
    # ellipsoid centre (200,0,0), semi axes 30,40,50, kept between z=-20 and z=0; ray down -y at z=-10: y = 40*sqrt(1-(10/50)^2)
    assert abs(el[0, 0, 3] - (400 - 40 * np.sqrt(1 - 0.04))) < 1e-2
    assert abs(el[1, 0, 3] - (300 - 20)) < 1e-3                                # from below: the z = -20 cut plane
    pc = Oracle().intersect(geom, np.array([[0, 300, -50], [0, 300, 30], [30, 0, 300]], dtype=np.float32),
                            np.array([[0, -1, 0], [0, -1, 0], [0, 0, -1]], dtype=np.float32))
    assert abs(pc[0, 0, 3] - 250) < 1e-3                                       # cylinder part, r = 50
    assert abs(pc[1, 0, 3] - (300 - 35)) < 1e-3                                # cone part: r(30) = 50 - 30/60*30 = 35
    assert abs(pc[2, 0, 3] - (300 - 40)) < 1e-3                                # from above at r = 30: cone surface z = 60*(50-30)/30 = 40


def test_sphere_theta_cone_and_multiunion_follow_u4(tmp_path):
    """U4Solid::init_Sphere_ theta slices (zsphere layers), init_Cons (no inner nudge), init_MultiUnion (contiguous list node)"""
    from eic_opticks_b200 import gdml as GD, foundry as F
    from _ref import Oracle
    gd = tmp_path / "s.gdml"
    gd.write_text("""<?xml version="1.0"?>
<gdml><define/><materials>
 <material name="Vac"><D value="1e-25"/><fraction n="1" ref="H"/></material><element name="H" formula="H" Z="1"><atom value="1"/></element>
</materials><solids>
 <box name="w" x="2000" y="2000" z="2000" lunit="mm"/>
 <sphere name="cap" rmin="40" rmax="50" starttheta="0" deltatheta="60" startphi="0" deltaphi="360" aunit="deg" lunit="mm"/>
 <cone name="co" rmin1="10" rmax1="30" rmin2="20" rmax2="60" z="100" startphi="0" deltaphi="360" aunit="deg" lunit="mm"/>
 <orb name="o1" r="30" lunit="mm"/><box name="b1" x="40" y="40" z="100" lunit="mm"/>
 <multiUnion name="mu">
  <multiUnionNode name="n1"><solid ref="o1"/></multiUnionNode>
  <multiUnionNode name="n2"><solid ref="b1"/><position x="0" y="0" z="60" unit="mm"/></multiUnionNode>
 </multiUnion>
</solids><structure>
 <volume name="capl"><materialref ref="Vac"/><solidref ref="cap"/></volume>
 <volume name="col"><materialref ref="Vac"/><solidref ref="co"/></volume>
 <volume name="mul"><materialref ref="Vac"/><solidref ref="mu"/></volume>
 <volume name="W"><materialref ref="Vac"/><solidref ref="w"/>
  <physvol name="a"><volumeref ref="capl"/></physvol>
  <physvol name="b"><volumeref ref="col"/><position x="300" y="0" z="0" unit="mm"/></physvol>
  <physvol name="c"><volumeref ref="mul"/><position x="-300" y="0" z="0" unit="mm"/></physvol></volume>
</structure><setup name="Default" version="1.0"><world ref="W"/></setup></gdml>""")
    g = GD.GDML(str(gd))
    cap = g.solid_tree("cap")
    assert isinstance(cap, F.Op) and cap.typecode == F.CSG_DIFFERENCE
    assert cap.left.typecode == F.CSG_ZSPHERE and cap.right.typecode == F.CSG_ZSPHERE
    assert np.allclose(cap.left.param[3:6], (50.0, 25.0, 50.0)) and np.allclose(cap.right.param[3:6], (40.0, 20.0, 40.0))   # r, zmin = r cos 60, zmax = r
    co = g.solid_tree("co")
    assert co.typecode == F.CSG_DIFFERENCE and tuple(co.left.param[:4]) == (30.0, -50.0, 60.0, 50.0) and tuple(co.right.param[:4]) == (10.0, -50.0, 20.0, 50.0)
    mu = g.solid_tree("mu")
    assert isinstance(mu, F.ListNode) and mu.typecode == F.CSG_CONTIGUOUS and [s.typecode for s in mu.subs] == [F.CSG_SPHERE, F.CSG_BOX3]

    geom = GD.translate(str(gd))
    o = np.array([[0, 0, 500], [10, 0, -500], [300, 500, 0], [-300, 0, 500], [-300, 500, 0]], dtype=np.float32)
    d = np.array([[0, 0, -1], [0, 0, 1], [0, -1, 0], [0, 0, -1], [0, -1, 0]], dtype=np.float32)
    r = Oracle().intersect(geom, o, d)
    assert abs(r[0, 0, 3] - 450) < 1e-3                                        # top of the cap, r = 50
    assert abs(r[1, 0, 3] - (500 + np.sqrt(1600 - 100))) < 1e-3                # from below at x = 10: the outer layer's cut plane (z = 25) lies inside the subtracted inner layer -> first surface is the r = 40 sphere
    assert abs(r[2, 0, 3] - (500 - 45)) < 1e-3                                 # cone centred at x = 300: outer radius at z = 0 is 45
    assert abs(r[3, 0, 3] - (500 - 110)) < 1e-3                                # multi-union: box top at z = 60 + 50
    assert abs(r[4, 0, 3] - (500 - 30)) < 1e-3                                 # ... and the orb's equator


def test_assembly_daughters_are_imprinted_into_the_mother(tmp_path):
    """<assembly> (G4AssemblyVolume): its daughters land in the mother volume with the composed placement; same arrays as placing them directly"""
    from eic_opticks_b200 import gdml as GD
    head = """<?xml version="1.0"?>
<gdml><define/><materials>
 <material name="Vac"><D value="1e-25"/><fraction n="1" ref="H"/></material><element name="H" formula="H" Z="1"><atom value="1"/></element>
</materials><solids>
 <box name="w" x="2000" y="2000" z="2000" lunit="mm"/><box name="b" x="10" y="20" z="30" lunit="mm"/><orb name="o" r="5" lunit="mm"/>
</solids><structure>
 <volume name="bl"><materialref ref="Vac"/><solidref ref="b"/></volume>
 <volume name="ol"><materialref ref="Vac"/><solidref ref="o"/></volume>
"""
    tail = """</structure><setup name="Default" version="1.0"><world ref="W"/></setup></gdml>"""
    a = tmp_path / "a.gdml"
    a.write_text(head + """
 <assembly name="asm">
  <physvol name="p1"><volumeref ref="bl"/><position x="100" y="0" z="0" unit="mm"/></physvol>
  <physvol name="p2"><volumeref ref="ol"/><position x="0" y="50" z="0" unit="mm"/><rotation x="0" y="0" z="90" unit="deg"/></physvol>
 </assembly>
 <volume name="W"><materialref ref="Vac"/><solidref ref="w"/>
  <physvol name="A1"><volumeref ref="asm"/><position x="0" y="0" z="200" unit="mm"/><rotation x="0" y="0" z="90" unit="deg"/></physvol>
  <physvol name="A2"><volumeref ref="asm"/><position x="0" y="0" z="-200" unit="mm"/></physvol>
 </volume>""" + tail)
    ga = GD.translate(str(a))
    assert ga["prim_names"] == ["W_PV", "A1_p1", "A1_p2", "A2_p1", "A2_p2"]
    fa = ga["foundry"]
    assert fa["prim"].shape[0] == 5
    # second imprint is a pure translation: box centre at (100, 0, -200), orb at (0, 50, -200)
    bb = fa["prim"].reshape(-1, 16)[:, 8:14]
    assert np.allclose(bb[3], (95, -10, -215, 105, 10, -185), atol=1e-3) and np.allclose(bb[4], (-5, 45, -205, 5, 55, -195), atol=1e-3)
    # first imprint is rotated about z by 90 degrees (frame rotation): the box's x offset turns into a y offset, extents swap
    c = 0.5 * (bb[1, :3] + bb[1, 3:]); e = bb[1, 3:] - bb[1, :3]
    assert np.allclose(np.abs(c), (0, 100, 200), atol=1e-3) and np.allclose(e, (20, 10, 30), atol=1e-3)


def test_repeated_subtrees_become_instanced_solids(tmp_path):
    """stree::factorize restated (sysrap/stree.h:5263-5545) + stree::add_inst + the default sensor identifier: 600 placements of a
    two-volume PMT become one compound solid and 600 instance transforms; the nested repeat is not a factor of its own;
    the rays see the same surfaces as in the flat translation."""
    from eic_opticks_b200 import gdml as GD
    from _ref import Oracle
    pvs = []
    for i in range(24):
        for j in range(25):
            k = i * 25 + j
            pvs.append('<physvol name="PMT_%d" copynumber="%d"><volumeref ref="glassl"/><position x="%g" y="%g" z="0" unit="mm"/></physvol>'
                       % (k, 1000 + k, (i - 11.5) * 250.0, (j - 12.0) * 250.0))
    gd = tmp_path / "wall.gdml"
    gd.write_text("""<?xml version="1.0"?>
<gdml><define>
 <matrix name="RI" coldim="2" values="1.55*eV 1.4 6.2*eV 1.4"/><matrix name="RW" coldim="2" values="1.55*eV 1.33 6.2*eV 1.33"/>
 <matrix name="RV" coldim="2" values="1.55*eV 1.0 6.2*eV 1.0"/>
</define><materials>
 <element name="H" formula="H" Z="1"><atom value="1"/></element>
 <material name="Water"><property name="RINDEX" ref="RW"/><D value="1"/><fraction n="1" ref="H"/></material>
 <material name="Glass"><property name="RINDEX" ref="RI"/><D value="2"/><fraction n="1" ref="H"/></material>
 <material name="Vac"><property name="RINDEX" ref="RV"/><D value="1e-25"/><fraction n="1" ref="H"/></material>
</materials><solids>
 <box name="w" x="8000" y="8000" z="2000" lunit="mm"/><orb name="glass" r="100" lunit="mm"/><orb name="vac" r="95" lunit="mm"/>
</solids><structure>
 <volume name="vacl"><materialref ref="Vac"/><solidref ref="vac"/><auxiliary auxtype="SensDet" auxvalue="PhotonDetector"/></volume>
 <volume name="glassl"><materialref ref="Glass"/><solidref ref="glass"/><physvol name="inner"><volumeref ref="vacl"/></physvol></volume>
 <volume name="W"><materialref ref="Water"/><solidref ref="w"/>
""" + "\n".join(pvs) + """
 </volume>
</structure><setup name="Default" version="1.0"><world ref="W"/></setup></gdml>""")
    t = GD.translate(str(gd))
    flat = GD.translate(str(gd), freq_cut=10 ** 9)
    assert t["num_factor"] == 1 and flat["num_factor"] == 0
    fd, ff = t["foundry"], flat["foundry"]
    assert fd["solid"].shape[0] == 2 and fd["prim"].shape[0] == 3 and ff["prim"].shape[0] == 1201
    si = fd["solid"][:, 1]
    assert tuple(si[0, :2]) == (1, 0) and tuple(si[1, :2]) == (2, 1)                     # numPrim, primOffset
    assert bytes(fd["solid"][1].reshape(-1).view(np.uint8)[:2]) == b"f1"
    inst = fd["inst"]
    ii = inst.view(np.int32)
    assert inst.shape[0] == 601 and tuple(ii[0, :, 3]) == (0, 0, 0, -1) and (ii[1:, 1, 3] == 1).all() and (ii[:, 0, 3] == np.arange(601)).all()
    assert (ii[1:, 2, 3] == 1000 + np.arange(600) + 1).all() and (ii[1:, 3, 3] == np.arange(600)).all()      # copy number + 1 ; sensor index in preorder
    assert np.allclose(inst[1, 3, :3], (-11.5 * 250, -12 * 250, 0)) and np.allclose(inst[600, 3, :3], (11.5 * 250, 12 * 250, 0))
    # prims of the compound solid are in the frame of its outer volume
    bb = fd["prim"].reshape(-1, 16)[:, 8:14]
    assert np.allclose(bb[1], (-100, -100, -100, 100, 100, 100)) and np.allclose(bb[2], (-95, -95, -95, 95, 95, 95))
    assert t["bnd_names"] == flat["bnd_names"]
    # same surfaces along random rays (distance and normal), instanced vs flat
    rng = np.random.default_rng(5)
    o = np.stack([rng.uniform(-3000, 3000, 400), rng.uniform(-3000, 3000, 400), np.full(400, 900.0)], axis=1).astype(np.float32)
    d = np.stack([rng.normal(0, 0.3, 400), rng.normal(0, 0.3, 400), -np.ones(400)], axis=1)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    a, b = Oracle().intersect(t, o, d), Oracle().intersect(flat, o, d)
    assert np.allclose(a[:, 0, 3], b[:, 0, 3], rtol=1e-5, atol=1e-3) and np.allclose(a[:, 0, :3], b[:, 0, :3], atol=1e-4)
    assert ((a[:, 0, 3] < 1500).sum() > 100)                                              # a good share of the rays do hit PMTs (the rest reach the world box)


def test_gdml2geom_cli_writes_a_loadable_geometry_directory(tmp_path):
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "geom"
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "gdml2geom.py"), os.path.join(GOLD, "mini_detector.gdml"), str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0 and "7 prims" in r.stdout, r.stdout + r.stderr
    back = F.load_geometry(str(out))
    t = gdml.translate(os.path.join(GOLD, "mini_detector.gdml"))
    for k in ("solid", "prim", "node", "itra", "inst"):
        assert (back["foundry"][k].view(np.uint8) == t["foundry"][k].view(np.uint8)).all(), k
    assert back["bnd_names"] == t["bnd_names"] and (back["bnd"] == t["bnd"].astype(np.float32)).all()
