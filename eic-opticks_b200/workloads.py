"""Synthetic workloads for the BASELINE.json configs: geometry + tables + gensteps.

Gensteps normally come from Geant4 (absent here), so each workload fabricates them with the field
semantics of the reference collectors (u4/U4.cc:82-143, 196-252) from a fixed numpy seed.
Every function returns a dict: geom, gensteps (n,6,4) float32, input_photons or None, config
overrides (e.g. max_bounce), name, and the expected total photon count.
"""
import numpy as np

from . import geometries as GEO
from . import gensteps as G

SEED = 20261017


def _poisson_split(rng, total, n):
    """n positive counts summing exactly to total, Poisson-like spread"""
    mean = total / n
    c = rng.poisson(mean, n).astype(np.int64)
    c = np.maximum(c, 1)
    diff = total - int(c.sum())
    # spread the remainder deterministically
    step = 1 if diff > 0 else -1
    k = 0
    while diff != 0:
        if c[k % n] + step >= 1:
            c[k % n] += step
            diff -= step
        k += 1
    return c


def sipm8x8_scint(num_photon=12_500_000, photons_per_genstep=1000, cerenkov_fraction=0.1, seed=SEED):
    """BASELINE config 3: 8x8 CsI + SiPM, scintillation (+ Cerenkov) gensteps inside random crystals,
    OPTICKS_MAX_BOUNCE=32 as in tests/test_GPUPhotonSource_8x8SiPM.sh:5."""
    geom = GEO.sipm8x8()
    rng = np.random.default_rng(seed)
    ngs = max(1, num_photon // photons_per_genstep)
    counts = _poisson_split(rng, num_photon, ngs)
    cc = geom["crystal_centers"]
    pos = cc[rng.integers(0, len(cc), ngs)] + rng.uniform(-0.95, 0.95, (ngs, 3)) * np.array([1.0, 1.0, 3.95])
    dirs = rng.normal(size=(ngs, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    step = rng.uniform(0.02, 0.05, ngs)               # short steps keep the whole step inside the crystal
    is_ck = rng.uniform(size=ngs) < cerenkov_fraction
    gs = G.scint_gensteps(pos, dirs, 1.0, counts, geom["crystal_line"], geom["scintillation_time"])
    gs[:, 2, :3] = dirs * step[:, None]
    gs[:, 2, 3] = step
    n = geom["n_crystal"]
    ck = G.cerenkov_gensteps(pos, dirs, 1.0, counts, geom["crystal_line"], 1.0, 300.0, 700.0, n,
                             pre_velocity=299.0, post_velocity=298.5, mean_photons=(30.0, 29.0))
    ck[:, 2, :3] = dirs * step[:, None]
    ck[:, 2, 3] = step
    gs[is_ck] = ck[is_ck]
    return dict(name="sipm8x8_scint", geom=geom, gensteps=np.ascontiguousarray(gs), input_photons=None,
                config=dict(max_bounce=32), num_photon=int(counts.sum()))


def raindrop_cerenkov(num_photon=10_000_000, photons_per_genstep=1000, seed=SEED):
    """BASELINE config 2: muon-like track (dir (0,0.2,0.8)) through the water box, Cerenkov gensteps
    along it, ~1000 photons each."""
    geom = GEO.raindrop()
    rng = np.random.default_rng(seed)
    ngs = max(1, num_photon // photons_per_genstep)
    counts = _poisson_split(rng, num_photon, ngs)
    d = np.array([0.0, 0.2, 0.8])
    d /= np.linalg.norm(d)
    s = np.linspace(-48.0, 47.0, ngs)
    pos = s[:, None] * d[None, :] + rng.normal(0, 0.05, (ngs, 3))
    n = geom["n_water"]
    gs = G.cerenkov_gensteps(pos, d, 1.0, counts, geom["water_line"], 1.0, 80.0, 800.0, n, pre_velocity=299.79, post_velocity=299.78,
                             mean_photons=(2.0, 1.9))
    gs[:, 1, 3] = (s + 48.0) / 299.79
    return dict(name="raindrop_cerenkov", geom=geom, gensteps=np.ascontiguousarray(gs), input_photons=None, config=dict(), num_photon=int(counts.sum()))


def sphere_leak_torch(num_photon=1_000_000, seed=0):
    """BASELINE config 1: config/sphere_leak.json torch (disc r 0.1 at the origin, mom (0,0.3,1),
    420 nm) scaled to 1e6 photons, generated on the host and fed as input photons like
    GPUPhotonSource does (src/GPUPhotonSourceMinimal.h:57-88)."""
    geom = GEO.sphere_leak()
    t = dict(pos=[0.0, 0.0, 0.0], time=0.0, mom=G._normalize_f32([0.0, 0.3, 1.0]), pol=[1.0, 0.0, 0.0], wavelength=420.0, radius=0.1,
             numphoton=num_photon, type="disc")
    ph = G.torch_photons(t, num_photon, seed=seed)
    return dict(name="sphere_leak_torch", geom=geom, gensteps=G.input_photon_genstep(num_photon), input_photons=ph, config=dict(), num_photon=num_photon)


def pmt_wall_torch(num_photon=10_000_000, nx=100, ny=100, seed=0):
    """BASELINE config 4 (scaled per launch): instanced PMT wall, photons from a wide disc above it,
    supplied as an input-photon array like GPUPhotonFileSource (src/GPUPhotonFileSource.h:51-87)."""
    geom = GEO.pmt_wall(nx, ny)
    hx, hy = geom["half"]
    t = dict(pos=[0.0, 0.0, 1500.0], time=0.0, mom=[0.0, 0.0, -1.0], pol=[1.0, 0.0, 0.0], wavelength=420.0, radius=float(min(hx, hy) - 600.0),
             numphoton=num_photon, type="disc")
    ph = G.torch_photons(t, num_photon, seed=seed)
    pu = ph.view(np.uint32)
    pu[:, 3, :] = 0                                   # file-source photons carry zero flags
    return dict(name="pmt_wall_torch", geom=geom, gensteps=G.input_photon_genstep(num_photon), input_photons=ph, config=dict(), num_photon=num_photon)


def boolean_zoo_torch(num_photon=1_000_000):
    """BASELINE config 5: boolean-heavy solids lit from an inward-facing sphere source (torch genstep)."""
    geom = GEO.boolean_zoo()
    t = dict(pos=[0.0, 0.0, 0.0], time=0.0, mom=[0.0, 0.0, 1.0], pol=[1.0, 0.0, 0.0], wavelength=440.0, radius=-950.0, numphoton=num_photon,
             type="sphere", zenith=[0.0, 1.0], azimuth=[0.0, 1.0])
    return dict(name="boolean_zoo_torch", geom=geom, gensteps=G.torch_genstep(t, num_photon), input_photons=None, config=dict(), num_photon=num_photon)


WORKLOADS = dict(sipm8x8_scint=sipm8x8_scint, raindrop_cerenkov=raindrop_cerenkov, sphere_leak_torch=sphere_leak_torch,
                 pmt_wall_torch=pmt_wall_torch, boolean_zoo_torch=boolean_zoo_torch)


def scintillator_tank(num_photon=1_000_000, photons_per_genstep=500, seed=SEED):
    """parity stress workload: scintillation + Cerenkov gensteps inside the LS sphere of geometries.scintillator_tank"""
    geom = GEO.scintillator_tank()
    rng = np.random.default_rng(seed)
    ngs = max(1, num_photon // photons_per_genstep)
    counts = _poisson_split(rng, num_photon, ngs)
    r = 900.0 * rng.uniform(size=ngs) ** (1 / 3.0)
    u = rng.normal(size=(ngs, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    pos = u * r[:, None]
    dirs = rng.normal(size=(ngs, 3)); dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    gs = G.scint_gensteps(pos, dirs, 1.0, counts, geom["ls_line"], geom["scintillation_time"], mean_velocity=200.0)
    gs[:, 2, :3] = dirs * 2.0
    gs[:, 2, 3] = 2.0
    ck = G.cerenkov_gensteps(pos, dirs, 2.0, counts, geom["ls_line"], 1.05, 200.0, 700.0, 1.60, pre_velocity=290.0, post_velocity=288.0,
                             mean_photons=(40.0, 38.0))
    is_ck = rng.uniform(size=ngs) < 0.3
    gs[is_ck] = ck[is_ck]
    return dict(name="scintillator_tank", geom=geom, gensteps=np.ascontiguousarray(gs), input_photons=None, config=dict(), num_photon=int(counts.sum()))


WORKLOADS["scintillator_tank"] = scintillator_tank


# ---- workloads for the arms of the path the BASELINE configs do not reach -----------------------------
def far_wall_torch(num_photon=100_000):
    """PropagateRefine scene: torch disc 24 m from the solids it lights, so 0.99 t exceeds the default 5000 mm refine distance"""
    geom = GEO.far_wall()
    t = dict(pos=[0.0, 0.0, 12000.0], time=0.0, mom=[0.0, 0.0, -1.0], pol=[1.0, 0.0, 0.0], wavelength=430.0, radius=4000.0, numphoton=num_photon,
             type="disc")
    return dict(name="far_wall_torch", geom=geom, gensteps=G.torch_genstep(t, num_photon), input_photons=None,
                config=dict(propagate_refine=1, refine_distance=5000.0), num_photon=num_photon)


# rectangle first: its side index is photon_id / (numphoton / 4) with the ABSOLUTE photon id, so only a genstep that
# starts the event (ids 0 .. numphoton-1) lands on sides 0..3 (sysrap/storch.h:470-510)
TORCH_SHAPES = ("rectangle", "disc", "sphere", "sphere_marsaglia", "line", "point", "circle")


def torch_shapes(num_photon=20_000, shapes=TORCH_SHAPES):
    """one on-device storch genstep per source type (sysrap/storch.h:189-516) inside the boolean zoo; the index-driven types
    (line, circle, rectangle) take their fraction from the ABSOLUTE photon id over the genstep's numphoton like the reference"""
    geom = GEO.boolean_zoo()
    per = max(4, num_photon // len(shapes)) // 4 * 4
    gss = []
    for k, ty in enumerate(shapes):
        t = dict(pos=[30.0 * k, -20.0 * k, 250.0], time=0.1 * k, mom=G._normalize_f32([0.2, -0.1, -1.0]), pol=[1.0, 0.0, 0.0], wavelength=400.0 + 20.0 * k,
                 radius=150.0, numphoton=per, type=ty, zenith=[0.1, 0.9], azimuth=[0.05, 0.95], distance=0.3)
        if ty == "rectangle":
            t["zenith"], t["azimuth"] = [-300.0, 300.0], [-400.0, 400.0]           # z and x extents of the frame
        if ty in ("sphere", "sphere_marsaglia"):
            t["radius"] = -900.0 if ty == "sphere" else 60.0
            t["pos"] = [0.0, 0.0, 0.0]
        gss.append(G.torch_genstep(t, per))
    gs = np.ascontiguousarray(np.concatenate(gss))
    return dict(name="torch_shapes", geom=geom, gensteps=gs, input_photons=None, config=dict(), num_photon=per * len(shapes))


def carrier_photons(num_photon=150):
    """scarrier gensteps (sysrap/scarrier.h:47-58): the photon is carried in the genstep itself, y displaced by 10 mm per photon id"""
    geom = GEO.boolean_zoo()
    gs = G.empty_gensteps(2)
    u = gs.view(np.uint32)
    for k, (pos, mom, pol, wl) in enumerate((((-900.0, -800.0, 10.0), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), 500.0),
                                             ((-900.0, -1500.0, -40.0), G._normalize_f32([1.0, 0.05, 0.02]), (0.0, 0.0, 1.0), 430.0))):
        u[k, 0] = (G.GS_CARRIER, 0, 0, num_photon // 2)
        gs[k, 2] = (pos[0], pos[1], pos[2], 0.5 * k)
        gs[k, 3, :3] = mom; u[k, 3, 3] = 0
        gs[k, 4] = (pol[0], pol[1], pol[2], wl)
    return dict(name="carrier_photons", geom=geom, gensteps=gs, input_photons=None, config=dict(), num_photon=2 * (num_photon // 2))


def pmt_wall_sensor_a(num_photon=40_000, nx=12, ny=12):
    """ems 3 (zplus_sensor_A) photocathode: photons from above hit the upper hemisphere (lposcost >= 0: unconditional detect,
    qsim.h:2184-2327, 1749-1755), photons from below reach the lower hemisphere through the glass (lposcost < 0: ordinary
    surface model with the same optical row)"""
    geom = GEO.pmt_wall(nx, ny, sensor_a=True)
    hx, hy = geom["half"]
    half = num_photon // 2
    r = float(min(hx, hy) - 600.0)
    top = G.torch_photons(dict(pos=[0.0, 0.0, 1500.0], time=0.0, mom=[0.0, 0.0, -1.0], pol=[1.0, 0.0, 0.0], wavelength=420.0, radius=r, numphoton=half, type="disc"), half, seed=0)
    bot = G.torch_photons(dict(pos=[0.0, 0.0, -1500.0], time=0.0, mom=G._normalize_f32([0.05, 0.02, 1.0]), pol=[1.0, 0.0, 0.0], wavelength=450.0, radius=r, numphoton=half, type="disc"),
                          half, seed=1)
    ph = np.ascontiguousarray(np.concatenate([top, bot]))
    ph.view(np.uint32)[:, 3, :] = 0
    return dict(name="pmt_wall_sensor_a", geom=geom, gensteps=G.input_photon_genstep(len(ph)), input_photons=ph, config=dict(), num_photon=len(ph))


def halfspace_zoo_torch(num_photon=30_000):
    """halfspace-cut solids lit from an inward-facing sphere source"""
    geom = GEO.halfspace_zoo()
    t = dict(pos=[0.0, 0.0, 0.0], time=0.0, mom=[0.0, 0.0, 1.0], pol=[1.0, 0.0, 0.0], wavelength=440.0, radius=-580.0, numphoton=num_photon,
             type="sphere", zenith=[0.0, 1.0], azimuth=[0.0, 1.0])
    return dict(name="halfspace_zoo_torch", geom=geom, gensteps=G.torch_genstep(t, num_photon), input_photons=None, config=dict(), num_photon=num_photon)


def box_maze_photons(num_photon=40_000, seed=11):
    """input photons aimed at what makes box geometry awkward: starts exactly ON faces, edges and corners of touching boxes,
    axis-parallel directions (1 / d = inf), directions lying IN a face plane, plus isotropic photons from inside the boxes"""
    geom = GEO.box_maze()
    rng = np.random.default_rng(seed)
    cen = geom["box_centers"]
    n = num_photon // 4 * 4
    q = n // 4
    pos = np.zeros((n, 3), dtype=np.float32); mom = np.zeros((n, 3), dtype=np.float32)
    pick = cen[rng.integers(0, len(cen), size=n)]
    # 1: isotropic, from random points around box centres (inside boxes and in the water between blocks)
    pos[:q] = pick[:q] + rng.uniform(-14.0, 14.0, size=(q, 3))
    v = rng.normal(size=(q, 3)); mom[:q] = v / np.linalg.norm(v, axis=1)[:, None]
    # 2: axis-parallel directions from random points
    pos[q:2 * q] = pick[q:2 * q] + rng.uniform(-14.0, 14.0, size=(q, 3))
    ax = rng.integers(0, 3, size=q); sg = rng.choice([-1.0, 1.0], size=q)
    mom[q + np.arange(q), ax] = sg
    # 3: starts exactly on a face / edge / corner of a 25 mm box of block 1 (coordinates that are face values), random directions
    c1 = cen[rng.integers(0, 32, size=q)]
    off = rng.uniform(-12.5, 12.5, size=(q, 3))
    snap = rng.integers(1, 8, size=q)                       # bit k: axis k sits on a face
    for a in range(3):
        on = (snap >> a) & 1 == 1
        off[on, a] = rng.choice([-12.5, 12.5], size=int(on.sum()))
    pos[2 * q:3 * q] = c1 + off
    v = rng.normal(size=(q, 3)); mom[2 * q:3 * q] = v / np.linalg.norm(v, axis=1)[:, None]
    # 4: starts on a face, direction IN that face plane (grazing along coincident faces), half of them axis-parallel
    c2 = cen[rng.integers(0, 32, size=q)]
    off = rng.uniform(-12.5, 12.5, size=(q, 3))
    fa = rng.integers(0, 3, size=q)
    off[np.arange(q), fa] = rng.choice([-12.5, 12.5], size=q)
    pos[3 * q:] = c2 + off
    v = rng.normal(size=(q, 3)); v[np.arange(q), fa] = 0.0
    axp = rng.random(q) < 0.5
    oth = (fa + 1 + rng.integers(0, 2, size=q)) % 3
    v[axp] = 0.0; v[np.nonzero(axp)[0], oth[axp]] = rng.choice([-1.0, 1.0], size=int(axp.sum()))
    mom[3 * q:] = v / np.linalg.norm(v, axis=1)[:, None]
    ph = np.zeros((n, 4, 4), dtype=np.float32)
    ph[:, 0, :3] = pos; ph[:, 0, 3] = 0.0
    ph[:, 1, :3] = mom
    pol = np.cross(mom, np.array([0.3, -0.5, 0.81], dtype=np.float32)); pol /= np.linalg.norm(pol, axis=1)[:, None]
    ph[:, 2, :3] = pol; ph[:, 2, 3] = rng.uniform(300.0, 600.0, size=n)
    return dict(name="box_maze_photons", geom=geom, gensteps=G.input_photon_genstep(n), input_photons=np.ascontiguousarray(ph), config=dict(), num_photon=n)


def pfrich_photons(num_photon=100_000, seed=3):
    """Cherenkov-like photons in the reference's own pfRICH geometry (geometries.pfrich): born in the aerogel annulus, heading down
    the vessel within 0.1 - 0.4 rad of +z, 300 - 600 nm; they cross the aerogel face, fly through the nitrogen, bounce off the
    inner / outer mirrors and end on the sensor pyramids or the absorbing edges"""
    geom = GEO.pfrich()
    rng = np.random.default_rng(seed)
    n = num_photon
    r = np.sqrt(rng.uniform(135.0 ** 2, 585.0 ** 2, size=n)); a = rng.uniform(0, 2 * np.pi, size=n)
    ph = np.zeros((n, 4, 4), dtype=np.float32)
    ph[:, 0, 0] = r * np.cos(a); ph[:, 0, 1] = r * np.sin(a); ph[:, 0, 2] = rng.uniform(-236.0, -216.0, size=n)
    th = rng.uniform(0.1, 0.4, size=n); az = rng.uniform(0, 2 * np.pi, size=n)
    mom = np.stack([np.sin(th) * np.cos(az), np.sin(th) * np.sin(az), np.cos(th)], axis=1)
    ph[:, 1, :3] = mom
    pol = np.cross(mom, np.array([0.0, 0.0, 1.0])); pol /= np.linalg.norm(pol, axis=1)[:, None]
    ph[:, 2, :3] = pol; ph[:, 2, 3] = rng.uniform(300.0, 600.0, size=n)
    return dict(name="pfrich_photons", geom=geom, gensteps=G.input_photon_genstep(n), input_photons=np.ascontiguousarray(ph), config=dict(), num_photon=n)


ARM_WORKLOADS = dict(box_maze_photons=box_maze_photons, pfrich_photons=pfrich_photons, far_wall_torch=far_wall_torch, torch_shapes=torch_shapes, carrier_photons=carrier_photons, pmt_wall_sensor_a=pmt_wall_sensor_a,
                     halfspace_zoo_torch=halfspace_zoo_torch)
WORKLOADS.update(ARM_WORKLOADS)
