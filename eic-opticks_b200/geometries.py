"""Hand-built CSGFoundry + boundary tables for the geometries BASELINE.json's configs name.

These follow the reference's GDML files volume for volume (tests/geom/*.gdml) using the
translation rules of SURVEY section 9 (boundary = omat/osur/isur/imat, surface lookup order of
u4/U4Surface.h:422-443, flattening of non-repeated volumes into solid 0, surface payload rule of
u4/U4SurfaceArray.h:159-240).  They exist so that tests and the bench do not depend on the GDML
translator; `gdml.py` produces the same arrays from the files themselves.

Each builder returns a dict: foundry (arrays), bnd, optical, icdf (or None), names, and helper
indices (material lines for gensteps).
"""
import numpy as np

from . import foundry as F
from . import tables as T

EV = 1.0


def _finish(fd, bt, icdf=None, extra=None):
    bnd, optical = bt.arrays()
    out = dict(foundry=fd.arrays(), bnd=bnd, optical=optical, icdf=icdf, bnd_names=bt.names(), table=bt)
    if extra:
        out.update(extra)
    return out


def raindrop():
    """tests/geom/opticks_raindrop.gdml: Vacuum 240 > Pb 220 > Air 200 > Water 100 (mm boxes).
    The only detecting surface is the directional border surface drop_pv -> medium_pv
    (water -> air, EFFICIENCY 1, REFLECTIVITY 0)."""
    bt = T.BoundaryTable()
    e2 = ([1.55, 15.5], None)
    bt.add_material(T.Material("G4_WATER", RINDEX=([1.55, 15.5], [1.333, 1.333]), GROUPVEL=([1.55, 15.5], [224.901, 224.901])))
    bt.add_material(T.Material("G4_AIR", RINDEX=([1.512, 5.512], [1.1, 1.1]), GROUPVEL=([1.55, 15.5], [224.901, 224.901])))
    bt.add_material(T.Material("G4_Pb", RINDEX=([1.512, 5.512], [1.0, 1.1]), GROUPVEL=([1.55, 15.5], [224.901, 224.901])))
    bt.add_material(T.Material("VACUUM", RINDEX=([1.512, 5.512], [1.0, 1.1]), GROUPVEL=([1.55, 15.5], [224.901, 224.901])))
    bt.add_surface(T.Surface("medium_container_bs", REFLECTIVITY=0.0, EFFICIENCY=1.0))
    fd = F.Foundry()
    fd.begin_solid("r0")
    fd.add_prim(F.box3(240, 240, 240), bt.boundary("VACUUM", "", "", "VACUUM"), name="VACUUM_solid")
    fd.add_prim(F.box3(220, 220, 220), bt.boundary("VACUUM", "", "", "G4_Pb"), name="G4_Pb_solid")
    fd.add_prim(F.box3(200, 200, 200), bt.boundary("G4_Pb", "", "", "G4_AIR"), name="G4_AIR_solid")
    fd.add_prim(F.box3(100, 100, 100), bt.boundary("G4_AIR", "", "medium_container_bs", "G4_WATER"), name="G4_WATER_solid")
    fd.end_solid()
    return _finish(fd, bt, extra=dict(water_line=bt.material_line("G4_WATER"), n_water=1.333))


def sphere_leak():
    """tests/geom/sphere_leak.gdml: world sphere r50 (water); Mirror = sphere r30 - sphere r25
    (n 1.5, sensitive skin EFFICIENCY 1); Glass sphere r15 (water) wrapped in a perfect mirror skin."""
    bt = T.BoundaryTable()
    water = dict(RINDEX=([1.55, 15.5], [1.333, 1.333]), ABSLENGTH=([1.55, 15.5], [100000.0, 100000.0]),
                 RAYLEIGH=([1.55, 15.5], [1000.0, 1000.0]))
    bt.add_material(T.Material("WaterMaterial", **water))
    bt.add_material(T.Material("MirrorMaterial", RINDEX=([1.55, 15.5], [1.5, 1.5])))
    bt.add_surface(T.Surface("MirrorSkinSurface", EFFICIENCY=1.0))
    bt.add_surface(T.Surface("GlassSkinSurface", REFLECTIVITY=1.0, polished=True))
    fd = F.Foundry()
    fd.begin_solid("r0")
    fd.add_prim(F.sphere(50), bt.boundary("WaterMaterial", "", "", "WaterMaterial"), name="WorldBox")
    fd.add_prim(F.difference(F.sphere(30), F.sphere(25)),
                bt.boundary("WaterMaterial", "MirrorSkinSurface", "MirrorSkinSurface", "MirrorMaterial"), name="MirrorSphere")
    fd.add_prim(F.sphere(15), bt.boundary("WaterMaterial", "GlassSkinSurface", "GlassSkinSurface", "WaterMaterial"), name="GlassSphere")
    fd.end_solid()
    return _finish(fd, bt)


def sipm8x8():
    """tests/geom/8x8SiPM_w_CSI_optial_grease.gdml: 8x8 CsI crystals (2x2x8 mm, pitch 2.2) with a
    0.98 polished skin, grease and window layers, 8x8 SiPM pixels with a detecting skin and the
    dead strips between them; 194 sibling volumes flattened into solid 0."""
    bt = T.BoundaryTable()
    en = [1.0, 4.0]
    bt.add_material(T.Material("Air", RINDEX=(en, [1.0, 1.0]), ABSLENGTH=(en, [10000.0, 10000.0])))
    bt.add_material(T.Material("Crystal", RINDEX=(en, [1.82, 1.82]), ABSLENGTH=(en, [400.0, 400.0]), REEMISSIONPROB=(en, [0.0, 0.0])))
    bt.add_material(T.Material("OpticalGrease", RINDEX=(en, [1.47, 1.47]), ABSLENGTH=(en, [1000.0, 1000.0])))
    bt.add_material(T.Material("EntranceWindow", RINDEX=(en, [1.55, 1.55]), ABSLENGTH=(en, [1000.0, 1000.0])))
    bt.add_surface(T.Surface("SiPMActiveSkin", REFLECTIVITY=(en, [0.0, 0.0]), EFFICIENCY=(en, [1.0, 1.0])))
    bt.add_surface(T.Surface("CrystalSkin", REFLECTIVITY=(en, [0.98, 0.98]), EFFICIENCY=(en, [0.0, 0.0])))
    bt.add_surface(T.Surface("DeadSkinH", REFLECTIVITY=(en, [0.3, 0.3]), EFFICIENCY=(en, [0.0, 0.0])))
    bt.add_surface(T.Surface("DeadSkinV", REFLECTIVITY=(en, [0.3, 0.3]), EFFICIENCY=(en, [0.0, 0.0])))
    b_world = bt.boundary("Air", "", "", "Air")
    b_grease = bt.boundary("Air", "", "", "OpticalGrease")
    b_window = bt.boundary("Air", "", "", "EntranceWindow")
    b_xtal = bt.boundary("Air", "CrystalSkin", "CrystalSkin", "Crystal")
    b_sipm = bt.boundary("Air", "SiPMActiveSkin", "SiPMActiveSkin", "EntranceWindow")
    b_deadv = bt.boundary("Air", "DeadSkinV", "DeadSkinV", "EntranceWindow")
    b_deadh = bt.boundary("Air", "DeadSkinH", "DeadSkinH", "EntranceWindow")
    fd = F.Foundry()
    fd.begin_solid("r0")
    fd.add_prim(F.box3(100, 100, 100), b_world, name="WorldBox")
    fd.add_prim(F.box3(17.4, 17.4, 0.1), b_grease, F.translate(0, 0, 8.05), name="GreaseLayer")
    fd.add_prim(F.box3(17.4, 17.4, 0.1), b_window, F.translate(0, 0, 8.15), name="WindowLayer")
    centers = []
    for i in range(8):
        for j in range(8):
            c = (-8.7 + i * 2.2 + 1.0 - 1.0, -8.7 + j * 2.2, 4.0)
            fd.add_prim(F.box3(2.0, 2.0, 8.0), b_xtal, F.translate(-8.7 + i * 2.2, -8.7 + j * 2.2, 4.0), name="CrystalPixel")
            centers.append(c)
    for i in range(8):
        for j in range(8):
            fd.add_prim(F.box3(2.0, 2.0, 0.2), b_sipm, F.translate(-8.7 + i * 2.2, -8.7 + j * 2.2, 8.3), name="SiPMActive")
    for i in range(8):
        for j in range(7):
            fd.add_prim(F.box3(0.2, 2.0, 0.2), b_deadv, F.translate(-8.7 + i * 2.2 + 1.1, -8.7 + j * 2.2, 8.3), name="SiPMDeadV")
    for i in range(7):
        fd.add_prim(F.box3(17.4, 0.2, 0.2), b_deadh, F.translate(0, -8.7 + i * 2.2 + 1.1, 8.3), name="SiPMDeadH")
    fd.end_solid()
    # scintillation spectrum SCINT_SPECTRUM: 1.5 eV 0, 2.896 eV 1, 4.0 eV 0
    icdf = T.make_icdf([1.5, 2.896, 4.0], [0.0, 1.0, 0.0])      # the three points of the GDML matrix, integrated like Geant4 does
    return _finish(fd, bt, icdf, extra=dict(crystal_centers=np.array(centers, dtype=np.float32),
                                           crystal_line=bt.material_line("Crystal"), n_crystal=1.82,
                                           scintillation_time=21.5))


def pmt_wall(nx=100, ny=100, pitch=250.0, sensor_a=False):
    """Synthetic instanced PMT wall (BASELINE config 4): solid 0 = water world box, solid 1 = one
    PMT (glass bulb = sphere union neck cylinder, inner vacuum sphere behind a photocathode
    surface), instanced nx*ny times on a grid with sensor identifiers 0..n-1."""
    bt = T.BoundaryTable()
    en = [1.55, 6.2]
    bt.add_material(T.Material("Water", RINDEX=(en, [1.333, 1.333]), ABSLENGTH=(en, [30000.0, 30000.0]), RAYLEIGH=(en, [50000.0, 50000.0])))
    bt.add_material(T.Material("Pyrex", RINDEX=(en, [1.47, 1.47]), ABSLENGTH=(en, [1000.0, 1000.0])))
    bt.add_material(T.Material("Vacuum", RINDEX=(en, [1.0, 1.0])))
    # sensor_a: the optical surface is named '#...' so that its optical row carries ems 3 (smatsur_Surface_zplus_sensor_A,
    # sysrap/smatsur.h:8-16, sstandard.h:311-441): hits on the upper hemisphere of the bulb (lposcost >= 0) detect
    # unconditionally (qsim::propagate_at_surface_Detect), the lower hemisphere falls back to the ordinary surface model
    bt.add_surface(T.Surface("Photocathode", EFFICIENCY=(en, [0.25, 0.25]), optical_surface_name="#Photocathode" if sensor_a else None))
    b_world = bt.boundary("Water", "", "", "Water")
    b_glass = bt.boundary("Water", "", "", "Pyrex")
    b_vac = bt.boundary("Pyrex", "Photocathode", "Photocathode", "Vacuum")
    half_x, half_y = nx * pitch / 2 + 500.0, ny * pitch / 2 + 500.0
    fd = F.Foundry()
    fd.begin_solid("r0")
    fd.add_prim(F.box3(2 * half_x, 2 * half_y, 4000.0), b_world, name="WorldBox")
    fd.end_solid()
    fd.begin_solid("r1")
    bulb = F.union(F.sphere(100.0), F.cylinder(40.0, -170.0, -80.0))
    fd.add_prim(bulb, b_glass, name="PMT_glass")
    fd.add_prim(F.sphere(95.0), b_vac, name="PMT_vacuum")
    fd.end_solid()
    fd.add_instance(np.eye(4), 0, -1, -1)
    k = 0
    for i in range(nx):
        for j in range(ny):
            fd.add_instance(F.translate((i + 0.5) * pitch - nx * pitch / 2, (j + 0.5) * pitch - ny * pitch / 2, 0.0), 1, k, k)
            k += 1
    return _finish(fd, bt, extra=dict(half=(half_x, half_y), pitch=pitch, n_pmt=nx * ny))


def boolean_zoo():
    """CSG boolean-heavy solids (BASELINE config 5): difference / intersection / union trees of
    depth 1-3, zsphere with both caps, cone, tubs with the 1 % inner nudge (u4/U4Solid.h:813-821),
    polycone-like unions cut by a phi wedge, a convex polyhedron, hyperboloid and list nodes, all
    glass in scattering / absorbing water inside an absorbing rock box."""
    bt = T.BoundaryTable()
    en = [1.55, 6.2]
    bt.add_material(T.Material("Rock"))
    bt.add_material(T.Material("Water", RINDEX=(en, [1.333, 1.35]), ABSLENGTH=(en, [2000.0, 1500.0]), RAYLEIGH=(en, [800.0, 400.0])))
    bt.add_material(T.Material("Glass", RINDEX=(en, [1.48, 1.52]), ABSLENGTH=(en, [500.0, 300.0])))
    bt.add_surface(T.implicit_surface("Implicit_RINDEX_NoRINDEX_water_rock"))
    bt.add_surface(T.Surface("DiffuseSkin", REFLECTIVITY=(en, [0.8, 0.8]), polished=False))
    bt.add_surface(T.Surface("SensorSkin", EFFICIENCY=(en, [0.5, 0.5])))
    b_world = bt.boundary("Rock", "", "", "Rock")
    b_water = bt.boundary("Rock", "Implicit_RINDEX_NoRINDEX_water_rock", "Implicit_RINDEX_NoRINDEX_water_rock", "Water")
    b_glass = bt.boundary("Water", "", "", "Glass")
    b_diff = bt.boundary("Water", "DiffuseSkin", "DiffuseSkin", "Glass")
    b_sens = bt.boundary("Water", "SensorSkin", "SensorSkin", "Glass")
    fd = F.Foundry()
    fd.begin_solid("r0")
    fd.add_prim(F.box3(2200, 2200, 2200), b_world, name="Rock")
    fd.add_prim(F.box3(2000, 2000, 2000), b_water, name="Water")
    shapes = []
    # depth-1 trees
    shapes.append(("box_minus_sphere", F.difference(F.box3(200, 200, 200), F.sphere(120)), b_glass))
    shapes.append(("sphere_and_box", F.intersection(F.sphere(130), F.box3(200, 200, 200)), b_glass))
    shapes.append(("cyl_union_cone", F.union(F.cylinder(60, -100, 0), F.cone(60, 0, 10, 120)), b_diff))
    # tubs: outer cylinder minus inner cylinder lengthened by 1 % of hz each end
    hz = 100.0
    shapes.append(("tubs", F.difference(F.cylinder(100, -hz, hz), F.cylinder(70, -hz * 1.01, hz * 1.01)), b_glass))
    shapes.append(("zsphere", F.zsphere(110, -60, 80), b_sens))
    shapes.append(("cone", F.cone(100, -80, 30, 80), b_glass))
    # depth-2 / depth-3 trees
    shapes.append(("box_minus_2", F.difference(F.difference(F.box3(220, 220, 220), F.sphere(100).placed(F.translate(80, 0, 0))),
                                              F.cylinder(40, -150, 150).placed(F.translate(-60, 0, 0))), b_glass))
    shapes.append(("depth3", F.intersection(F.union(F.sphere(100), F.box3(120, 120, 240)),
                                            F.difference(F.cylinder(110, -130, 130), F.union(F.sphere(50), F.cone(40, 40, 5, 130)))), b_glass))
    # polycone-like union of cylinder + cone + cylinder cut by a phi wedge
    poly = F.union(F.union(F.cylinder(50, -120, -40), F.cone(50, -40, 90, 40)), F.cylinder(90, 40, 120))
    shapes.append(("polycone_phicut", F.intersection(poly, F.phicut(20.0, 250.0)), b_glass))
    shapes.append(("zsphere_rot", F.zsphere(100, -100 * 0.999, 40).placed(F.rotate_x(35.0)), b_glass))
    shapes.append(("hyperboloid", F.hyperboloid(50, 80, -100, 100), b_glass))
    # convex polyhedron: a truncated pyramid (trapezoid)
    pl = []
    for sx, sy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
        n = np.array([sx, sy, 0.3]); n = n / np.linalg.norm(n)
        pl.append([n[0], n[1], n[2], 90.0 * np.linalg.norm([sx, sy]) / np.linalg.norm([sx, sy, 0.3])])
    pl.append([0, 0, 1, 100.0]); pl.append([0, 0, -1, 100.0])
    shapes.append(("trapezoid", F.convexpolyhedron(pl, [-130, -130, -100, 130, 130, 100]), b_glass))
    # list nodes
    shapes.append(("discontiguous", F.ListNode(F.CSG_DISCONTIGUOUS, [F.sphere(40).placed(F.translate(-70, 0, 0)), F.box3(60, 60, 60).placed(F.translate(70, 0, 0)),
                                                                   F.cylinder(30, -40, 40).placed(F.translate(0, 80, 0))]), b_glass))
    shapes.append(("contiguous", F.ListNode(F.CSG_CONTIGUOUS, [F.sphere(60).placed(F.translate(-40, 0, 0)), F.sphere(60).placed(F.translate(40, 0, 0)),
                                                             F.box3(100, 50, 50)]), b_glass))
    shapes.append(("overlap", F.ListNode(F.CSG_OVERLAP, [F.sphere(90), F.box3(140, 140, 140), F.cylinder(80, -100, 100)]), b_glass))
    grid = 4
    centers = []
    for k, (name, shape, b) in enumerate(shapes):
        ix, iy = k % grid, k // grid
        c = (-600.0 + ix * 400.0, -600.0 + iy * 400.0, 0.0)
        centers.append(c)
        frame = F.rotate_z(10.0 * k) @ F.translate(*c)
        fd.add_prim(shape, b, frame, mesh_idx=k + 2, name=name)
    fd.end_solid()
    return _finish(fd, bt, extra=dict(shape_centers=np.array(centers, dtype=np.float32), shape_names=[s[0] for s in shapes]))


def scintillator_tank():
    """Parity stress geometry (not a reference file): a sphere of liquid scintillator with strongly wavelength-dependent
    RINDEX / ABSLENGTH / RAYLEIGH / REEMISSIONPROB inside an acrylic shell inside water, enclosed by a detecting skin.
    Exercises bulk re-emission (ICDF lookups), Rayleigh scattering and texture interpolation at fractional wavelengths."""
    bt = T.BoundaryTable()
    en = [1.55, 2.07, 2.48, 3.10, 4.13, 6.2]
    bt.add_material(T.Material("LS", RINDEX=(en, [1.47, 1.48, 1.49, 1.51, 1.55, 1.60]), ABSLENGTH=(en, [8000.0, 6000.0, 3000.0, 600.0, 60.0, 5.0]),
                               RAYLEIGH=(en, [9000.0, 4000.0, 2000.0, 800.0, 300.0, 100.0]), REEMISSIONPROB=(en, [0.0, 0.1, 0.4, 0.8, 0.8, 0.6])))
    bt.add_material(T.Material("Acrylic", RINDEX=(en, [1.48, 1.49, 1.50, 1.51, 1.53, 1.56]), ABSLENGTH=(en, [4000.0, 4000.0, 3000.0, 1000.0, 100.0, 10.0])))
    bt.add_material(T.Material("Water", RINDEX=(en, [1.33, 1.333, 1.337, 1.343, 1.36, 1.40]), ABSLENGTH=(en, [20000.0, 30000.0, 25000.0, 9000.0, 900.0, 90.0]),
                               RAYLEIGH=(en, [100000.0, 60000.0, 30000.0, 12000.0, 4000.0, 1000.0])))
    bt.add_surface(T.Surface("PMTSkin", EFFICIENCY=(en, [0.02, 0.1, 0.25, 0.3, 0.15, 0.02])))
    fd = F.Foundry()
    fd.begin_solid("r0")
    fd.add_prim(F.sphere(2200.0), bt.boundary("Water", "", "", "Water"), name="World")
    fd.add_prim(F.difference(F.sphere(2000.0), F.sphere(1990.0)), bt.boundary("Water", "PMTSkin", "PMTSkin", "Water"), name="PMTShell")
    fd.add_prim(F.sphere(1020.0), bt.boundary("Water", "", "", "Acrylic"), name="Acrylic")
    fd.add_prim(F.sphere(1000.0), bt.boundary("Acrylic", "", "", "LS"), name="LS")
    fd.end_solid()
    # emission spectrum peaked at ~2.9 eV (430 nm)
    e = np.linspace(2.0, 4.0, 41)
    icdf = T.make_icdf(e, np.exp(-0.5 * ((e - 2.9) / 0.25) ** 2))
    return _finish(fd, bt, icdf, extra=dict(ls_line=bt.material_line("LS"), scintillation_time=4.5))


def far_wall(distance=12000.0):
    """Large scene for PropagateRefine (CSGOptiX/CSGOptiX7.cu:146-185): intersect distances well beyond the default
    PropagateRefineDistance of 5000 mm.  A 30 m water box with a glass sphere, a tubs and a sensor slab ~2 x distance
    away from the source plane."""
    bt = T.BoundaryTable()
    en = [1.55, 6.2]
    bt.add_material(T.Material("Rock"))
    bt.add_material(T.Material("Water", RINDEX=(en, [1.333, 1.35]), ABSLENGTH=(en, [60000.0, 40000.0]), RAYLEIGH=(en, [80000.0, 40000.0])))
    bt.add_material(T.Material("Glass", RINDEX=(en, [1.48, 1.52]), ABSLENGTH=(en, [5000.0, 3000.0])))
    bt.add_surface(T.implicit_surface("Implicit_RINDEX_NoRINDEX_water_rock"))
    bt.add_surface(T.Surface("SensorSkin", EFFICIENCY=(en, [0.5, 0.5])))
    fd = F.Foundry()
    fd.begin_solid("r0")
    fd.add_prim(F.box3(32000, 32000, 32000), bt.boundary("Rock", "", "", "Rock"), name="Rock")
    fd.add_prim(F.box3(30000, 30000, 30000), bt.boundary("Rock", "Implicit_RINDEX_NoRINDEX_water_rock", "Implicit_RINDEX_NoRINDEX_water_rock", "Water"), name="Water")
    fd.add_prim(F.sphere(1500.0), bt.boundary("Water", "", "", "Glass"), F.translate(0, 0, -distance), name="FarSphere")
    fd.add_prim(F.difference(F.cylinder(1200.0, -400.0, 400.0), F.cylinder(900.0, -404.0, 404.0)), bt.boundary("Water", "", "", "Glass"),
                F.rotate_x(20.0) @ F.translate(2500.0, 0, -distance + 2000.0), name="FarTubs")
    fd.add_prim(F.box3(6000, 6000, 100), bt.boundary("Water", "SensorSkin", "SensorSkin", "Glass"), F.translate(-3000.0, 1000.0, -distance - 2000.0), name="FarSensor")
    fd.end_solid()
    return _finish(fd, bt)


def halfspace_zoo():
    """Solids cut by CSG_HALFSPACE leaves (CSG/csg_intersect_leaf_halfspace.h:156-193): the unbounded leaf, its
    'exit at infinity' signalling (isect.y = -0.f) inside intersections, a complemented halfspace (difference) and a
    transformed one, all glass in water inside an absorbing rock box."""
    bt = T.BoundaryTable()
    en = [1.55, 6.2]
    bt.add_material(T.Material("Rock"))
    bt.add_material(T.Material("Water", RINDEX=(en, [1.333, 1.35]), ABSLENGTH=(en, [2000.0, 1500.0]), RAYLEIGH=(en, [800.0, 400.0])))
    bt.add_material(T.Material("Glass", RINDEX=(en, [1.48, 1.52]), ABSLENGTH=(en, [500.0, 300.0])))
    bt.add_surface(T.implicit_surface("Implicit_RINDEX_NoRINDEX_water_rock"))
    bt.add_surface(T.Surface("SensorSkin", EFFICIENCY=(en, [0.5, 0.5])))
    b_water = bt.boundary("Rock", "Implicit_RINDEX_NoRINDEX_water_rock", "Implicit_RINDEX_NoRINDEX_water_rock", "Water")
    b_glass = bt.boundary("Water", "", "", "Glass")
    b_sens = bt.boundary("Water", "SensorSkin", "SensorSkin", "Glass")
    fd = F.Foundry()
    fd.begin_solid("r0")
    fd.add_prim(F.box3(1400, 1400, 1400), bt.boundary("Rock", "", "", "Rock"), name="Rock")
    fd.add_prim(F.box3(1200, 1200, 1200), b_water, name="Water")
    s3 = 1.0 / np.sqrt(3.0)
    shapes = [
        ("sphere_cut_z", F.intersection(F.sphere(120), F.halfspace(0, 0, 1, 40.0)), b_glass),
        ("box_minus_halfspace", F.difference(F.box3(200, 200, 200), F.halfspace(s3, s3, s3, 20.0)), b_sens),
        ("cyl_two_cuts", F.intersection(F.intersection(F.cylinder(90, -120, 120), F.halfspace(1, 0, 0, 30.0)), F.halfspace(0, -1, 0, 50.0)), b_glass),
        ("sphere_cut_rotated", F.intersection(F.sphere(110), F.halfspace(0, 0, 1, 0.0).placed(F.rotate_x(40.0) @ F.translate(0, 0, 25.0))), b_glass),
    ]
    centers = []
    for k, (name, shape, b) in enumerate(shapes):
        c = (-250.0 + (k % 2) * 500.0, -250.0 + (k // 2) * 500.0, 0.0)
        centers.append(c)
        fd.add_prim(shape, b, F.rotate_z(15.0 * k) @ F.translate(*c), mesh_idx=k + 2, name=name)
    fd.end_solid()
    return _finish(fd, bt, extra=dict(shape_centers=np.array(centers, dtype=np.float32), shape_names=[s[0] for s in shapes]))


def box_maze():
    """Stress geometry for the home cells and the exact-box paths (not a reference file): blocks of boxes that TOUCH (coincident
    faces, edges and corners: every face hit is a tie between two or more prims), boxes nested three deep, a thin layer across
    a whole block, pitches that are exact in float32 and pitches that are not; glass of two indices, a mirror skin, a
    half-reflecting skin and a detecting skin, all in water inside an absorbing rock box."""
    bt = T.BoundaryTable()
    en = [1.55, 6.2]
    bt.add_material(T.Material("Rock"))
    bt.add_material(T.Material("Water", RINDEX=(en, [1.333, 1.35]), ABSLENGTH=(en, [5000.0, 4000.0]), RAYLEIGH=(en, [3000.0, 1500.0])))
    bt.add_material(T.Material("GlassA", RINDEX=(en, [1.48, 1.52]), ABSLENGTH=(en, [800.0, 600.0])))
    bt.add_material(T.Material("GlassB", RINDEX=(en, [1.70, 1.80]), ABSLENGTH=(en, [300.0, 200.0]), RAYLEIGH=(en, [500.0, 300.0])))
    bt.add_surface(T.implicit_surface("Implicit_RINDEX_NoRINDEX_water_rock"))
    bt.add_surface(T.Surface("MirrorSkin", REFLECTIVITY=(en, [0.97, 0.97]), polished=True))
    bt.add_surface(T.Surface("RoughSkin", REFLECTIVITY=(en, [0.6, 0.6]), polished=False))
    bt.add_surface(T.Surface("SensorSkin", EFFICIENCY=(en, [0.7, 0.7])))
    b_water = bt.boundary("Rock", "Implicit_RINDEX_NoRINDEX_water_rock", "Implicit_RINDEX_NoRINDEX_water_rock", "Water")
    kinds = [bt.boundary("Water", "", "", "GlassA"), bt.boundary("Water", "", "", "GlassB"), bt.boundary("Water", "MirrorSkin", "MirrorSkin", "GlassA"),
             bt.boundary("Water", "RoughSkin", "RoughSkin", "GlassB"), bt.boundary("Water", "SensorSkin", "SensorSkin", "GlassA")]
    fd = F.Foundry()
    fd.begin_solid("r0")
    fd.add_prim(F.box3(640, 640, 640), bt.boundary("Rock", "", "", "Rock"), name="Rock")
    fd.add_prim(F.box3(600, 600, 600), b_water, name="Water")
    centers = []
    k = 0
    # block 1: 4 x 4 x 2 touching boxes, pitch 25 (exact in float32)
    for i in range(4):
        for j in range(4):
            for l in range(2):
                c = (-150.0 + 25.0 * i, -150.0 + 25.0 * j, -40.0 + 25.0 * l)
                fd.add_prim(F.box3(25.0, 25.0, 25.0), kinds[k % len(kinds)], F.translate(*c), name="touch%d" % k)
                centers.append(c); k += 1
    # a thin layer lying on block 1 (its bottom face is the top face of 16 boxes)
    fd.add_prim(F.box3(100.0, 100.0, 0.7), kinds[1], F.translate(-112.5, -112.5, -2.15), name="layer")
    # block 2: 5 x 3 touching boxes, pitch 17.3 (not exact in float32: faces coincide only up to rounding)
    for i in range(5):
        for j in range(3):
            c = (60.0 + 17.3 * i, -120.1 + 17.3 * j, 33.3)
            fd.add_prim(F.box3(17.3, 17.3, 40.0), kinds[(k + 2) % len(kinds)], F.translate(*c), name="pitch%d" % k)
            centers.append(c); k += 1
    # nested three deep, sharing one face plane (x = 50) and one corner
    fd.add_prim(F.box3(100.0, 100.0, 100.0), kinds[0], F.translate(0.0, 150.0, 100.0), name="outer")
    fd.add_prim(F.box3(60.0, 60.0, 60.0), bt.boundary("GlassA", "", "", "GlassB"), F.translate(20.0, 150.0, 100.0), name="middle")
    fd.add_prim(F.box3(20.0, 20.0, 20.0), bt.boundary("GlassB", "SensorSkin", "SensorSkin", "GlassA"), F.translate(40.0, 170.0, 120.0), name="inner")
    centers += [(0.0, 150.0, 100.0), (20.0, 150.0, 100.0), (40.0, 170.0, 120.0)]
    fd.end_solid()
    return _finish(fd, bt, extra=dict(box_centers=np.array(centers, dtype=np.float32)))


def pfrich(path=None):
    """The pfRICH detector the reference ships as tests/geom/pfrich_min_FINAL.gdml (aerogel radiator, nitrogen vessel, inner and outer
    mirrors, 64 sensor pyramids, absorbing edges; 103 prims, 39 boundaries), as translated by gdml.py and kept in
    tests/golden/pfrich_min_geometry.npz (tests/golden/make_pfrich_fixture.py) - the GPU box has no copy of the reference."""
    import os
    if path is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pfrich_min_geometry.npz")
    z = np.load(path)
    fd = {k: z[k] for k in ("solid", "prim", "node", "tran", "itra", "plan", "inst")}
    return dict(foundry=fd, bnd=z["bnd"], optical=z["optical"], icdf=None, bnd_names=[str(n) for n in z["bnd_names"]],
                prim_names=[str(n) for n in z["prim_names"]], sensitive_prims=z["sensitive_prims"])
