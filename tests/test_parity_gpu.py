"""GPU parity tests: libphox.so (through the C ABI) against
   (1) the reference's own device headers compiled unmodified (oracle/_ref/libphoxref_*.so),
   (2) the CPU oracle,
   (3) itself across launch slicing / rank sharding / BVH vs brute force / host vs device-resident paths.

Tolerances (north_star): history flags, boundaries, identities, indices bit-exact; positions, times,
wavelengths within 1e-4 relative.

* With FMA contraction off on both sides (csrc/libphox_nofma.so against oracle/_ref/libphoxref_*_nofma.so, `make parity`)
  EVERY photon of every workload is identical - integer data AND every float bit
  (test_nofma_builds_are_bit_identical_to_reference_headers): the two code bases evaluate the same expressions in the
  same order.
* In the default build (the reference's own flags: no -fmad option, CSGOptiX/CMakeLists.txt:47-57) nvcc fuses a*b+c
  differently in the two code bases - it does so even between two inlining contexts of ONE code base - so a photon whose
  branch variable lies within the last bit of its threshold takes the other branch.  Measured (profiles/parity_r2.json,
  with the bounce, the decision and the distance of every such photon): 0 to 8 photons of 30 000; the tests bound the
  fraction at 5e-4 and check the floats of ALL matching photons (max, not a quantile) against a per-workload bound.
"""
import os

import numpy as np
import pytest

import eic_opticks_b200 as ph
from eic_opticks_b200 import gensteps as G, workloads, parallel
from _ref import RefGPU, Oracle, ORACLE

pytestmark = pytest.mark.gpu

CASES = [("sipm8x8_scint", dict(num_photon=30000, photons_per_genstep=100)),
         ("raindrop_cerenkov", dict(num_photon=20000, photons_per_genstep=100)),
         ("sphere_leak_torch", dict(num_photon=10000)),
         ("pmt_wall_torch", dict(num_photon=30000, nx=20, ny=20)),
         ("boolean_zoo_torch", dict(num_photon=40000)),
         ("scintillator_tank", dict(num_photon=30000, photons_per_genstep=100)),      # re-emission, Rayleigh, dispersive tables
         ("box_maze_photons", dict(num_photon=40000))]      # touching / nested boxes, starts on faces and edges, axis-parallel and in-plane rays


def make_sim(w, **cfg):
    g = w["geom"]
    kw = dict(w["config"]); kw.update(cfg)
    return ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"], g["icdf"], **kw)


def rel_err(a, b):
    scale = np.maximum(1.0, np.abs(b[:, :3, :]))
    return np.abs(a[:, :3, :] - b[:, :3, :]) / scale


# Largest float error over ALL photons with matching histories, default build (profiles/parity_r2.json).  Beyond 1e-4 the
# gap opens at a grazing intersect with a curved surface (sqrt of a tiny discriminant: one ulp in, 1e-3 of t out) and, in
# the sphere_leak billiard, is then amplified by 31 specular bounces; flat-faced geometry stays at 1e-5.  The bound that
# holds everywhere is the one of the nofma pair: 0.
MAX_FLOAT_ERR = {"sipm8x8_scint": 1e-4, "raindrop_cerenkov": 1e-6, "sphere_leak_torch": 3e-2, "pmt_wall_torch": 3e-2,
                 "boolean_zoo_torch": 5e-2, "scintillator_tank": 1e-2,
                 # the maze is made of ties: two photons can part at a shared face and still collect the same flags and boundaries (neighbouring
                 # boxes of one kind), so "matching history" does not mean "same path" there: quantile bound only (the nofma pair is bit-identical)
                 "box_maze_photons": None}
MIN_SAME = 0.9995            # measured: >= 0.99973 on every workload, both RNG modes
# ... except the box maze, which is MADE of ties: every face of its touching boxes is shared by two to eight prims, a third of its
# photons start exactly on faces / edges / corners or travel inside a face plane, so which prim answers hangs on the last bit of t
# far more often (measured 0.99898; with FMA contraction off on both sides it is 1.0 like everywhere else, see the nofma test)
MIN_SAME_BY = {"box_maze_photons": 0.998}


def check_against(name, p, seq, ref_p, ref_seq, min_same=MIN_SAME, float_q=0.9995, max_err=None, q_tol=1e-4):
    pu, ru = p.view(np.uint32), ref_p.view(np.uint32)
    same = (pu[:, 3, :] == ru[:, 3, :]).all(axis=1) & (pu[:, 1, 3] == ru[:, 1, 3])          # q3 flags/identity/index + hitcount_iindex
    if seq is not None and ref_seq is not None:
        same &= (seq == ref_seq).all(axis=(1, 2))                                            # seqhis AND seqbnd
    assert same.mean() >= min_same, "%s: identical integer data for only %.5f of photons" % (name, same.mean())
    r = rel_err(p[same], ref_p[same])
    assert np.quantile(r, float_q) < q_tol, "%s: float q%.4f rel err %.3g" % (name, float_q, np.quantile(r, float_q))
    if max_err is not None:
        assert r.max() <= max_err, "%s: max float rel err over all matching photons %.3g" % (name, r.max())
    return same.mean()


@pytest.mark.parametrize("name,kw", CASES)
@pytest.mark.parametrize("variant", ["debugtag", "production"])
def test_photon_by_photon_vs_reference_headers(name, kw, variant):
    if not os.path.exists(os.path.join(ORACLE, "_ref", "libphoxref_%s.so" % variant)):
        pytest.fail("oracle/_ref/libphoxref_%s.so missing: run __graft_entry__.build() where /root/reference exists" % variant)
    w = workloads.WORKLOADS[name](**kw)
    ref = RefGPU(variant).simulate(w["geom"], w["gensteps"], w["input_photons"], max_bounce=w["config"].get("max_bounce", 31))
    sim = make_sim(w, event_mode=ph.MODE_DEBUGHEAVY, rng_mode=(ph.RNG_DEBUG_TAG if variant == "debugtag" else ph.RNG_PRODUCTION))
    for accel in (ph.ACCEL_BRUTE, ph.ACCEL_BVH):
        sim.set_config(accel=accel)
        hits = sim.simulate_np(w["gensteps"], 0, w["input_photons"])
        p, seq = sim.get_array("photon"), sim.get_array("seq")
        # the production build of the reference keeps no seq array: photons that part and end on the same final flags cannot be
        # told from matching ones there, so the max over "matching" photons is only asserted where histories are compared
        frac = check_against("%s/%s/accel%d" % (name, variant, accel), p, seq, ref["photon"], ref["seq"], min_same=MIN_SAME_BY.get(name, MIN_SAME),
                             max_err=MAX_FLOAT_ERR[name] if variant == "debugtag" else None,
                             q_tol=5e-4 if name.startswith("sphere_leak") else 1e-4)
        # hits = stable compaction of the photon array
        fm = p.view(np.uint32)[:, 3, 3]
        sel = (fm & 0x40) == 0x40
        assert len(hits) == sel.sum() and (hits.view(np.uint32) == p[sel].view(np.uint32)).all()
        if variant == "debugtag":
            rec, prd = sim.get_array("record"), sim.get_array("prd")
            same = (seq == ref["seq"]).all(axis=(1, 2))
            ns = 4 if name.startswith("sphere_leak") else rec.shape[1]          # step records: first bounces of the billiard
            rr = np.abs(rec[same][:, :ns, :3, :] - ref["record"][same][:, :ns, :3, :]) / np.maximum(1.0, np.abs(ref["record"][same][:, :ns, :3, :]))
            assert np.quantile(rr, 0.9995) < 1e-4
            assert MAX_FLOAT_ERR[name] is None or rr.max() <= MAX_FLOAT_ERR[name]
            assert (prd.view(np.uint32)[same][:, :, 1, 2:] == ref["prd"].view(np.uint32)[same][:, :, 1, 2:]).mean() > 0.9999   # identity, prim|boundary
        print(name, variant, accel, "identical fraction %.5f" % frac, "hits", len(hits), "rays", sim.stats()["num_ray"], ref["nray"])
    sim.close()


def test_nofma_builds_are_bit_identical_to_reference_headers(tmp_path):
    """with -fmad=false on both sides the engine and the reference's device headers agree on EVERY photon
    of all six workloads x both RNG-consumption modes x brute force / BVH - history flags, boundaries, identities, seqhis,
    seqbnd, and every bit of every float.  Runs tests/_parity.py in a process of its own (the library is chosen at import)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "eic-opticks_b200", "csrc", "libphox_nofma.so")
    assert os.path.exists(lib), "run __graft_entry__.build() (make -C eic-opticks_b200/csrc parity)"
    for v in ("debugtag", "production"):
        assert os.path.exists(os.path.join(ORACLE, "_ref", "libphoxref_%s_nofma.so" % v)), "oracle/_ref/*_nofma.so missing: build where /root/reference exists"
    out = tmp_path / "parity_nofma.json"
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "_parity.py"), "--build", "nofma", "--out", str(out)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rep = json.load(open(out))
    assert len(rep["entries"]) == (6 + 7) * 2 * 2 - 2        # the six workloads and the seven arm workloads of tests/_parity.py (carrier: production only)
    for e in rep["entries"]:
        tag = (e["workload"], e["rng_mode"], e["accel"])
        assert e["identical_integer_data"] == e["photons"], tag
        assert e["float_bits_identical_fraction_of_matching"] == 1.0 and e["max_ulp_matching"] == 0, tag
        assert e["rays_phox"] == e["rays_ref"], tag
        if e["rng_mode"] == "debugtag":
            assert e["max_rel_err_step_records"] == 0.0, tag
    assert rep["all_identical"]


def _arm(name, kw, variant="debugtag", **extra):
    """run one arm workload on the reference headers and on libphox (brute force and BVH); returns (w, ref, [(p, seq, rec, prd, hits)])"""
    w = workloads.WORKLOADS[name](**kw)
    cfg = w["config"]
    ref = RefGPU(variant).simulate(w["geom"], w["gensteps"], w["input_photons"], max_bounce=cfg.get("max_bounce", 31),
                                   refine=cfg.get("propagate_refine", 0), refine_distance=cfg.get("refine_distance", 5000.0))
    sim = make_sim(w, event_mode=ph.MODE_DEBUGHEAVY, rng_mode=(ph.RNG_DEBUG_TAG if variant == "debugtag" else ph.RNG_PRODUCTION), **extra)
    got = []
    for accel in (ph.ACCEL_BRUTE, ph.ACCEL_BVH):
        sim.set_config(accel=accel)
        hits = sim.simulate_np(w["gensteps"], 0, w["input_photons"]).copy()
        got.append(tuple(sim.get_array(k).copy() for k in ("photon", "seq", "record", "prd")) + (hits, sim.stats()["num_ray"]))
    sim.close()
    return w, ref, got


def test_propagate_refine_beyond_refine_distance():
    """row a2: trace<true> (CSGOptiX7.cu:146-185) - a second trace from 0.99 t whenever 0.99 t exceeds PropagateRefineDistance;
    scene with 24 m flights (default 5000 mm applies) and the boolean zoo with a 50 mm distance (most bounces refine)"""
    for name, kw, extra in (("far_wall_torch", dict(num_photon=20000), {}), ("boolean_zoo_torch", dict(num_photon=20000), dict(propagate_refine=1, refine_distance=50.0))):
        w = workloads.WORKLOADS[name](**kw)
        w["config"].update(extra)
        cfg = w["config"]
        ref = RefGPU("debugtag").simulate(w["geom"], w["gensteps"], w["input_photons"], refine=1, refine_distance=cfg["refine_distance"])
        plain = RefGPU("debugtag").simulate(w["geom"], w["gensteps"], w["input_photons"], refine=0)
        assert ref["nray"] > 1.3 * plain["nray"]                                  # the re-trace really runs
        for mode in (ph.KERNEL_PERSISTENT, ph.KERNEL_WAVEFRONT):
            sim = make_sim(w, event_mode=ph.MODE_DEBUGHEAVY, kernel_mode=mode)
            sim.simulate_np(w["gensteps"], 0, w["input_photons"])
            p, seq, prd = sim.get_array("photon"), sim.get_array("seq"), sim.get_array("prd")
            assert abs(int(sim.stats()["num_ray"]) - int(ref["nray"])) <= 1e-3 * ref["nray"], (name, mode)     # equal in the nofma pair
            check_against(name + "/refine", p, seq, ref["photon"], ref["seq"], max_err=5e-2)
            same = (seq == ref["seq"]).all(axis=(1, 2))
            # the refined distance (t_approx + second t) agrees with the reference's and differs from the unrefined one in its last bits
            t, tr, tp = prd[same][:, 0, 0, 3], ref["prd"][same][:, 0, 0, 3], plain["prd"][same][:, 0, 0, 3]
            assert np.abs(t - tr).max() <= 1e-4 * np.abs(tr).max()
            if name == "far_wall_torch":
                assert (t > 5000.0).mean() > 0.9 and (tr != tp).mean() > 0.2
            sim.close()


def test_every_torch_source_type_on_device():
    """row a12: storch::generate (sysrap/storch.h:189-516) for DISC, SPHERE, SPHERE_MARSAGLIA, LINE, POINT, CIRCLE, RECTANGLE gensteps
    generated ON the device, against the reference's own storch.h running on the B200: generation step bit-exact"""
    w, ref, got = _arm("torch_shapes", dict(num_photon=21000))
    per = w["num_photon"] // len(workloads.TORCH_SHAPES)
    for p, seq, rec, prd, hits, nray in got:
        # slot 0 of the step record = the generated photon: every field of every source type
        g, r = rec[:, 0].view(np.uint32), ref["record"][:, 0].view(np.uint32)
        for k, ty in enumerate(workloads.TORCH_SHAPES):
            sl = slice(k * per, (k + 1) * per)
            assert (g[sl, 3] == r[sl, 3]).all(), ty
            err = (np.abs(rec[sl, 0, :3] - ref["record"][sl, 0, :3]) / np.maximum(1.0, np.abs(ref["record"][sl, 0, :3]))).max()
            assert err <= 1e-4, (ty, err)                                         # sinf / cosf arguments one fma apart (bit-equal in the nofma pair)
            assert len(np.unique(rec[sl, 0, 0, :3], axis=0)) > (1 if ty != "point" else 0)
        check_against("torch_shapes", p, seq, ref["photon"], ref["seq"], max_err=5e-2)
        assert nray == ref["nray"] or abs(nray - ref["nray"]) < 1e-3 * nray


def test_carrier_gensteps():
    """row a11: scarrier::generate (sysrap/scarrier.h:47-58) - photon carried by the genstep, y + 10 mm per absolute photon id.
    Against the reference's PRODUCTION build only: in the as-built DEBUG_TAG layout offsetof(sctx, p) = 36 and the float4 stores
    of scarrier::generate through (quad4&)sphoton are misaligned - the reference kernel itself faults on a carrier genstep."""
    w = workloads.carrier_photons(150)
    refp = RefGPU("production")
    gen = refp.simulate(w["geom"], w["gensteps"], None, max_bounce=0)["photon"]              # no bounces: the generated photons themselves
    ref = refp.simulate(w["geom"], w["gensteps"], None)
    for mode in (ph.KERNEL_PERSISTENT, ph.KERNEL_WAVEFRONT):
        sim = make_sim(w, event_mode=ph.MODE_DEBUGLITE, rng_mode=ph.RNG_PRODUCTION, kernel_mode=mode)
        sim.simulate_np(w["gensteps"], 0)
        p, rec = sim.get_array("photon"), sim.get_array("record")
        assert (rec[:, 0].view(np.uint32) == gen.view(np.uint32)).all()                      # generation is pure data movement: bit-exact
        y = rec[:, 0, 0, 1]
        assert np.allclose(np.diff(y[:75]), 10.0) and np.allclose(np.diff(y[75:]), 10.0)
        check_against("carrier", p, None, ref["photon"], None, min_same=0.99)
        sim.close()


def test_sensor_a_surface_arm_and_lower_hemisphere_fallback():
    """row a20: optical ems 3 (smatsur_Surface_zplus_sensor_A): lposcost >= 0 -> propagate_at_surface_Detect (one draw, always SD),
    lposcost < 0 -> the ordinary surface model (qsim.h:2296-2312, 1749-1755)"""
    w, ref, got = _arm("pmt_wall_sensor_a", dict(num_photon=30000))
    opt = np.asarray(w["geom"]["optical"]).reshape(-1, 4)
    assert (opt[:, 1] == 3).any()
    for p, seq, rec, prd, hits, nray in got:
        check_against("sensor_a", p, seq, ref["photon"], ref["seq"], max_err=5e-2)
        # photons whose LAST hit was on a photocathode row: split by the sign of lposcost
        pu = p.view(np.uint32)
        flag = pu[:, 3, 0] & 0xffff
        nstep = (rec.view(np.uint32)[:, :, 3, 3] != 0).sum(axis=1)                 # steps recorded (flagmask non-zero)
        last_prd = prd[np.arange(len(prd)), np.maximum(nstep - 2, 0)]
        bnd = last_prd.view(np.uint32)[:, 1, 3] & 0xffff
        orient = pu[:, 3, 0] >> 31
        su_line = 4 * bnd + np.where(orient == 1, 1, 2)                            # cosTheta < 0 -> osur, else isur (qbnd.h:184-214)
        on_a = (opt[np.minimum(su_line, len(opt) - 1), 1] == 3) & np.isin(flag, (64, 128, 256, 512))
        up, down = on_a & (last_prd[:, 1, 0] >= 0), on_a & (last_prd[:, 1, 0] < 0)
        assert up.sum() > 500 and down.sum() > 200, (up.sum(), down.sum())
        assert (flag[up] == 64).all()                                              # upper hemisphere: always detected
        assert (flag[down] == 128).mean() > 0.5 and (flag[down] == 64).mean() > 0.1  # lower: absorb 0.75 / detect 0.25 of the surface row


def test_halfspace_cut_solids():
    """row a10: CSG_HALFSPACE leaves (csg_intersect_leaf_halfspace.h:156-193) inside intersections / differences, incl. a transformed
    and a complemented one: geometry queries against the reference headers, then full histories"""
    w = workloads.halfspace_zoo_torch(num_photon=1000)
    g = w["geom"]
    sim = make_sim(w)
    rng = np.random.default_rng(11)
    n = 200000
    o = rng.uniform(-580, 580, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1)[:, None]).astype(np.float32)
    a = sim.intersect(o, d, 0.05, ph.ACCEL_BVH)
    b = sim.intersect(o, d, 0.05, ph.ACCEL_BRUTE)
    assert a.tobytes() == b.tobytes()                                              # the boxes only cull
    r = RefGPU("debugtag").intersect(g, o, d, 0.05)
    bu, ru = b.view(np.uint32), r.view(np.uint32)
    same = (bu[:, 1, 2:] == ru[:, 1, 2:]).all(axis=1)
    assert same.mean() > 0.9995, same.mean()
    prim = bu[:, 1, 3] >> 16
    assert (np.bincount(prim[bu[:, 1, 3] != 0xffffffff], minlength=6)[2:6] > 1000).all()   # every cut solid is hit
    # flat cut faces: the normal of a halfspace hit is the plane normal itself
    err = np.abs(b[same, 0, :] - r[same, 0, :]) / np.maximum(1, np.abs(r[same, 0, :]))
    assert np.quantile(err, 0.999) < 1e-4
    sim.close()
    w, ref, got = _arm("halfspace_zoo_torch", dict(num_photon=30000))
    for p, seq, rec, prd, hits, nray in got:
        # default build: 0.9975.  A photon that sits ON a cut face after a transmission re-enters intersect_leaf_halfspace with
        # on_w = dot(o, n) - w a rounding residue of either sign; `inside = on_w < -1e-9f` (csg_intersect_leaf_halfspace.h:168-169) then
        # decides between "no hit" and "exit at infinity" - a knife edge of the reference's algorithm, settled by whichever way the
        # compiler contracted the dot product.  The nofma pair agrees on all 30 000 photons (test_nofma_builds_...).
        check_against("halfspace_zoo", p, seq, ref["photon"], ref["seq"], min_same=0.995, max_err=5e-2)
        assert len(hits) > 100


def test_pfrich_the_detector_geometry_the_reference_ships():
    """tests/geom/pfrich_min_FINAL.gdml (aerogel, nitrogen vessel, inner / outer mirrors, 64 sensor pyramids, absorbing edges: tubes,
    cones, booleans, G4Trap as convexpolyhedron; 103 prims, 39 boundaries) translated by gdml.py and carried as
    tests/golden/pfrich_min_geometry.npz: geometry queries BVH = brute force = the reference's intersect headers, then Cherenkov-like
    photons from the aerogel against the reference's device headers, photon by photon"""
    w = workloads.pfrich_photons(num_photon=1000)
    g = w["geom"]
    sim = make_sim(w)
    rng = np.random.default_rng(12)
    n = 200000
    o = (rng.uniform(-1, 1, (n, 3)) * np.array([640.0, 640.0, 240.0])).astype(np.float32)
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1)[:, None]).astype(np.float32)
    a = sim.intersect(o, d, 0.05, ph.ACCEL_BVH)
    b = sim.intersect(o, d, 0.05, ph.ACCEL_BRUTE)
    assert a.tobytes() == b.tobytes()
    r = RefGPU("debugtag").intersect(g, o, d, 0.05)
    bu, ru = b.view(np.uint32), r.view(np.uint32)
    same = (bu[:, 1, 2:] == ru[:, 1, 2:]).all(axis=1)
    assert same.mean() > 0.9995, same.mean()
    prim = bu[:, 1, 3] >> 16
    assert len(np.unique(prim[bu[:, 1, 3] != 0xffffffff])) > 80                     # rays reach most of the 103 prims
    err = np.abs(b[same, 0, 3] - r[same, 0, 3]) / np.maximum(1, np.abs(r[same, 0, 3]))
    assert np.quantile(err, 0.999) < 1e-4
    sim.close()
    w, ref, got = _arm("pfrich_photons", dict(num_photon=40000))
    for p, seq, rec, prd, hits, nray in got:
        frac = check_against("pfrich", p, seq, ref["photon"], ref["seq"], min_same=0.999, max_err=None)
        assert len(hits) > 4000 and nray == ref["nray"] or frac < 1.0
        print("pfrich identical fraction %.5f hits %d rays %d (reference %d)" % (frac, len(hits), nray, ref["nray"]))


@pytest.mark.parametrize("name,kw", CASES)
def test_photon_by_photon_vs_cpu_oracle(name, kw):
    kw = dict(kw); kw["num_photon"] = min(kw["num_photon"], 10000)
    w = workloads.WORKLOADS[name](**kw)
    orc = Oracle().simulate(w["geom"], w["gensteps"], w["input_photons"], max_bounce=w["config"].get("max_bounce", 31))
    sim = make_sim(w, event_mode=ph.MODE_HITPHOTONSEQ)
    hits = sim.simulate_np(w["gensteps"], 0, w["input_photons"])
    p, seq = sim.get_array("photon"), sim.get_array("seq")
    # dispersive tables: the oracle's emulation of the texture filter matches the hardware bit for bit on 97.4 % of
    # fetches (see test_oracle_texture_emulation_vs_hardware); the others move lengths by up to 1/256 of a table step
    fq = 0.98 if name == "scintillator_tank" else 0.9995
    check_against(name + "/oracle", p, seq, orc["photon"], orc["seq"], min_same=0.99 if name in ("scintillator_tank", "box_maze_photons") else 0.995, float_q=fq,
                  q_tol=5e-3 if name.startswith("sphere_leak") else 1e-4)      # host libm (sinf, logf ...) differs from CUDA's by ulps: the billiard amplifies them
    assert abs(len(hits) - orc["nhit"]) <= max(5, 0.005 * len(p))
    sim.close()


def test_known_answer_file_source_ten_photons_ten_hits():
    g = ph.geometries.raindrop()
    ip = G.photons_from_text(os.path.join(os.path.dirname(__file__), "golden", "photons_file_source.txt"))
    sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"])
    sim.event.set_input_photon(ip)
    dt = sim.simulate(0, False)
    assert dt >= 0 and sim.event.get_num_hit() == 10
    hits = sim.event.hits
    assert sorted(hits[:, 2, 3].tolist()) == [420.0] * 3 + [450.0] * 2 + [500.0] * 5
    assert (hits.view(np.uint32)[:, 3, 2] == np.arange(10)).all()
    sim.reset(0)
    assert sim.num_hit() == 0 and sim.simulate(1) == -1.0                     # nothing collected -> -1. like QSim::simulate
    sim.close()


@pytest.mark.parametrize("name,kw", [CASES[0], CASES[3], CASES[4]])
def test_intersect_bvh_equals_brute_and_reference(name, kw):
    w = workloads.WORKLOADS[name](**dict(kw, num_photon=1000))
    g = w["geom"]
    sim = make_sim(w)
    rng = np.random.default_rng(5)
    fd = g["foundry"]
    lo = fd["prim"].reshape(-1, 16)[:, 8:11].min(0); hi = fd["prim"].reshape(-1, 16)[:, 11:14].max(0)
    if name == "pmt_wall_torch":
        lo, hi = np.array([-2500, -2500, -400.0]), np.array([2500, 2500, 400.0])
    if name == "sipm8x8_scint":
        lo, hi = np.array([-10, -10, -1.0]), np.array([10, 10, 9.0])
    n = 200000
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1)[:, None]).astype(np.float32)
    a = sim.intersect(o, d, 0.05, ph.ACCEL_BVH)
    b = sim.intersect(o, d, 0.05, ph.ACCEL_BRUTE)
    au, bu = a.view(np.uint32), b.view(np.uint32)
    same_int = (au[:, 1, 2:] == bu[:, 1, 2:]).all(axis=1)                       # identity, prim|boundary
    assert same_int.mean() > 0.9995, same_int.mean()
    assert np.abs(a[same_int, 0, 3] - b[same_int, 0, 3]).max() <= 1e-4 * np.abs(b[same_int, 0, 3]).max()
    r = RefGPU("debugtag").intersect(g, o, d, 0.05)
    ru = r.view(np.uint32)
    same_ref = (bu[:, 1, 2:] == ru[:, 1, 2:]).all(axis=1)
    assert same_ref.mean() > 0.9995, same_ref.mean()
    err = np.abs(b[same_ref, 0, :] - r[same_ref, 0, :]) / np.maximum(1, np.abs(r[same_ref, 0, :]))
    assert np.quantile(err, 0.999) < 1e-4                   # grazing hits on curved surfaces are ill-conditioned: sqrt of a tiny discriminant
    assert np.quantile(err, 0.9999) < 5e-3
    sim.close()


def test_multi_launch_slicing_equals_single_launch():
    w = workloads.sipm8x8_scint(num_photon=50000, photons_per_genstep=100)
    sim = make_sim(w, event_mode=ph.MODE_HITPHOTONSEQ)
    h1 = sim.simulate_np(w["gensteps"], 3).copy()
    p1, s1 = sim.get_array("photon").copy(), sim.get_array("seq").copy()
    assert sim.stats()["num_launch"] == 1
    sim.set_config(max_slot=7000)
    h2 = sim.simulate_np(w["gensteps"], 3)
    p2, s2 = sim.get_array("photon"), sim.get_array("seq")
    assert sim.stats()["num_launch"] >= 7
    assert (p1.view(np.uint32) == p2.view(np.uint32)).all() and (s1 == s2).all() and (h1.view(np.uint32) == h2.view(np.uint32)).all()
    idx = h2.view(np.uint32)[:, 3, 2]
    assert (np.diff(idx.astype(np.int64)) > 0).all()                            # ascending absolute photon index
    sim.close()


@pytest.mark.parametrize("name,kw", CASES)
def test_wavefront_form_is_bit_identical_to_persistent_form(name, kw):
    """The two forms of the bounce loop (include/phox.h PHOX_KERNEL_*), and the production and debug instantiations of each,
    compile the same physics / generation / intersect bodies into different kernels.  nvcc fuses a*b+c per compilation
    context, so their agreement is checked, not assumed: every output byte must be the same - hits of the production
    kernels (Minimal), photon arrays of the production kernels (HitPhoton) against those of the debug kernels (DebugLite,
    DebugHeavy), records / seq / prd between the forms.  (scripts/form_consistency.py is the same check at 1 M photons.)"""
    w = workloads.WORKLOADS[name](**kw)
    out = {}
    modes = ((ph.MODE_MINIMAL, {}), (ph.MODE_HITPHOTON, {}), (ph.MODE_DEBUGLITE, dict(max_record=8)), (ph.MODE_DEBUGHEAVY, dict(max_record=12)))
    for mode in (ph.KERNEL_PERSISTENT, ph.KERNEL_WAVEFRONT):
        for em, extra in modes:
            sim = make_sim(w, event_mode=em, kernel_mode=mode, **extra)
            h = sim.simulate_np(w["gensteps"], 2, w["input_photons"]).copy()
            arrs = {"hit": h}
            if em != ph.MODE_MINIMAL:
                names = {ph.MODE_HITPHOTON: ("photon",), ph.MODE_DEBUGLITE: ("photon", "seq", "record"), ph.MODE_DEBUGHEAVY: ("photon", "seq", "record", "prd")}[em]
                for k in names:
                    arrs[k] = sim.get_array(k).copy()
            arrs["num_ray"] = np.array([sim.stats()["num_ray"]])
            out[(mode, em)] = arrs
            sim.close()
    for em, _ in modes:
        a, b = out[(ph.KERNEL_PERSISTENT, em)], out[(ph.KERNEL_WAVEFRONT, em)]
        assert a.keys() == b.keys()
        for k in a:
            assert a[k].shape == b[k].shape, (name, k, a[k].shape, b[k].shape)
            assert a[k].tobytes() == b[k].tobytes(), (name, em, k)
    for mode in (ph.KERNEL_PERSISTENT, ph.KERNEL_WAVEFRONT):           # production kernels against debug kernels
        prod, lite, heavy = out[(mode, ph.MODE_HITPHOTON)], out[(mode, ph.MODE_DEBUGLITE)], out[(mode, ph.MODE_DEBUGHEAVY)]
        assert prod["photon"].tobytes() == lite["photon"].tobytes(), (name, mode, "photon")
        assert prod["hit"].tobytes() == lite["hit"].tobytes() == out[(mode, ph.MODE_MINIMAL)]["hit"].tobytes(), (name, mode, "hit")
        # DebugHeavy runs the tag-recording physics body (propagate_t<true>: every tagged draw is also written to the tag / flat
        # arrays), a second compiled body whose multiply-adds are fused differently: same integer data and histories, floats equal
        # to a few last bits
        pu, hu = prod["photon"].view(np.uint32), heavy["photon"].view(np.uint32)
        assert (pu[:, 3, :] == hu[:, 3, :]).all() and (pu[:, 1, 3] == hu[:, 1, 3]).all(), (name, mode, "DebugHeavy integer data")
        assert (lite["seq"] == heavy["seq"]).all(), (name, mode, "DebugHeavy seq")
        assert rel_err(heavy["photon"], prod["photon"]).max() < 2e-5, (name, mode, rel_err(heavy["photon"], prod["photon"]).max())
    assert len(out[(ph.KERNEL_WAVEFRONT, ph.MODE_MINIMAL)]["hit"]) > 0


HOME_CASES = CASES + [("far_wall_torch", dict(num_photon=20000)), ("halfspace_zoo_torch", dict(num_photon=20000)),
                      ("pmt_wall_sensor_a", dict(num_photon=20000)), ("pfrich_photons", dict(num_photon=30000))]


@pytest.mark.parametrize("name,kw", HOME_CASES)
def test_home_cells_only_cull(name, kw):
    """The home-cell shortcut of the traversal (phox_kernels.cuh traverse_bvh: candidate list of the prim the photon sits in,
    accepted when the nearest candidate hit ends inside that prim's box) must give the bytes of the plain BVH and of the
    brute-force loop (the persistent form of the loop has no home pass and serves as one more reference); and it has to be
    exercised, not just compiled."""
    w = workloads.WORKLOADS[name](**kw)
    out, home_rays = {}, {}
    for mode in (ph.KERNEL_WAVEFRONT, ph.KERNEL_PERSISTENT):
        for accel in (ph.ACCEL_BVH, ph.ACCEL_BVH_NOHOME, ph.ACCEL_BRUTE):
            sim = make_sim(w, event_mode=ph.MODE_DEBUGLITE, kernel_mode=mode, accel=accel, max_record=8)
            h = sim.simulate_np(w["gensteps"], 1, w["input_photons"]).copy()
            out[(mode, accel)] = {"hit": h, "photon": sim.get_array("photon").copy(), "seq": sim.get_array("seq").copy(),
                                  "record": sim.get_array("record").copy()}
            st = sim.stats()
            home_rays[(mode, accel)] = (st["num_home_ray"], st["num_ray"])
            sim.close()
    for mode in (ph.KERNEL_WAVEFRONT, ph.KERNEL_PERSISTENT):
        ref = out[(mode, ph.ACCEL_BRUTE)]
        for accel in (ph.ACCEL_BVH, ph.ACCEL_BVH_NOHOME):
            for k in ref:
                assert out[(mode, accel)][k].tobytes() == ref[k].tobytes(), (name, "mode", mode, "accel", accel, k)
    for k in out[(ph.KERNEL_WAVEFRONT, ph.ACCEL_BRUTE)]:
        assert out[(ph.KERNEL_WAVEFRONT, ph.ACCEL_BRUTE)][k].tobytes() == out[(ph.KERNEL_PERSISTENT, ph.ACCEL_BRUTE)][k].tobytes(), (name, "wavefront vs persistent", k)
    for mode in (ph.KERNEL_WAVEFRONT, ph.KERNEL_PERSISTENT):
        assert home_rays[(mode, ph.ACCEL_BVH_NOHOME)][0] == 0 and home_rays[(mode, ph.ACCEL_BRUTE)][0] == 0
        assert home_rays[(mode, ph.ACCEL_BVH)][1] == home_rays[(mode, ph.ACCEL_BRUTE)][1]          # same rays traced
    print(name, "rays settled by their home cell: %d of %d" % home_rays[(ph.KERNEL_WAVEFRONT, ph.ACCEL_BVH)])
    if name == "sipm8x8_scint":
        assert home_rays[(ph.KERNEL_WAVEFRONT, ph.ACCEL_BVH)][0] > 0.3 * home_rays[(ph.KERNEL_WAVEFRONT, ph.ACCEL_BVH)][1]
    if name == "box_maze_photons":
        assert home_rays[(ph.KERNEL_WAVEFRONT, ph.ACCEL_BVH)][0] > 0.05 * home_rays[(ph.KERNEL_WAVEFRONT, ph.ACCEL_BVH)][1]


def test_wavefront_form_with_launch_slicing_and_time_cut():
    w = workloads.sipm8x8_scint(num_photon=50000, photons_per_genstep=100)
    res = []
    for mode in (ph.KERNEL_PERSISTENT, ph.KERNEL_WAVEFRONT):
        # distinct epsilons: tmin0 applies after the flags of epsilon0_mask (here SI|CK|SC|RE only: not after TORCH... the sources
        # of this workload), the wavefront form carries that decision as a bit of the list entry
        sim = make_sim(w, event_mode=ph.MODE_HITPHOTONSEQ, kernel_mode=mode, max_slot=7000, max_time=3.0, max_bounce=9,
                       propagate_epsilon=0.07, propagate_epsilon0=0.001)
        h = sim.simulate_np(w["gensteps"], 1).copy()
        res.append((h, sim.get_array("photon").copy(), sim.get_array("seq").copy()))
        sim.close()
    for a, b in zip(*res):
        assert a.shape == b.shape and a.tobytes() == b.tobytes()


def max_bounce_cfg(w):
    return w["config"].get("max_bounce", 32)


def test_bounce_loop_stops_once_the_list_is_empty():
    """The wavefront form stops launching bounce kernels once no photon is alive (the BVH kernel posts the list length to a pinned
    host word every 4 bounces, the host reads one batch behind): a one-bounce workload runs 8 to 12 of its 32 bounce pairs, a
    long-history workload all of them, and neither changes a byte of output against the persistent form."""
    for name, kw, short in (("raindrop_cerenkov", dict(num_photon=300000), True), ("sipm8x8_scint", dict(num_photon=60000, photons_per_genstep=100), False)):
        w = workloads.WORKLOADS[name](**kw)
        out = {}
        for mode in (ph.KERNEL_PERSISTENT, ph.KERNEL_WAVEFRONT):
            sim = make_sim(w, event_mode=ph.MODE_HITPHOTON, kernel_mode=mode, max_bounce=max_bounce_cfg(w))
            h = sim.simulate_np(w["gensteps"], 1, w["input_photons"]).copy()
            out[mode] = (h, sim.get_array("photon").copy(), sim.stats())
            sim.close()
        a, b = out[ph.KERNEL_PERSISTENT], out[ph.KERNEL_WAVEFRONT]
        assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes(), name
        assert a[2]["num_ray"] == b[2]["num_ray"], name
        max_bounce = max_bounce_cfg(w)
        full = 2 * max_bounce + 5                                  # generate + a kernel pair per bounce + hit selection (+ genstep home)
        if short:
            assert b[2]["num_kernel"] <= 2 * 12 + 6, (name, b[2]["num_kernel"])
        else:
            assert full - 2 <= b[2]["num_kernel"] <= full + 2, (name, b[2]["num_kernel"], full)


def test_auto_mode_hands_the_tail_of_an_event_to_the_persistent_kernel():
    """PHOX_KERNEL_AUTO: once the posted list length is at or below the tail threshold, the rest of the histories runs in the persistent
    kernel (resume mode: photon records and parked draw counts taken over from the live list).  Same bytes as the pure wavefront and
    the pure persistent form, fewer kernels than the pure wavefront form."""
    for name, kw in (("scintillator_tank", dict(num_photon=400000, photons_per_genstep=100)), ("pmt_wall_torch", dict(num_photon=300000, nx=20, ny=20))):
        w = workloads.WORKLOADS[name](**kw)
        out = {}
        for mode in (ph.KERNEL_PERSISTENT, ph.KERNEL_WAVEFRONT, ph.KERNEL_AUTO):
            sim = make_sim(w, event_mode=ph.MODE_HITPHOTON, kernel_mode=mode)
            h = sim.simulate_np(w["gensteps"], 1, w["input_photons"]).copy()
            out[mode] = (h, sim.get_array("photon").copy(), sim.stats())
            sim.close()
        ref = out[ph.KERNEL_PERSISTENT]
        for mode in (ph.KERNEL_WAVEFRONT, ph.KERNEL_AUTO):
            assert out[mode][0].tobytes() == ref[0].tobytes() and out[mode][1].tobytes() == ref[1].tobytes(), (name, mode)
            assert out[mode][2]["num_ray"] == ref[2]["num_ray"], (name, mode)
        assert 10 < out[ph.KERNEL_AUTO][2]["num_kernel"] < out[ph.KERNEL_WAVEFRONT][2]["num_kernel"], (name, out[ph.KERNEL_AUTO][2]["num_kernel"], out[ph.KERNEL_WAVEFRONT][2]["num_kernel"])


@pytest.mark.parametrize("name,kw", [CASES[0], CASES[3]])
def test_rank_sharding_concatenates_to_single_gpu_result(name, kw):
    w = workloads.WORKLOADS[name](**dict(kw, num_photon=24000))
    sim = make_sim(w)
    whole = sim.simulate_np(w["gensteps"], 0, w["input_photons"]).copy()
    parts = []
    for r in range(4):
        gs_r, ip_r, off, cnt = parallel.shard_event(w["gensteps"], r, 4, w["input_photons"])
        parts.append(sim.simulate_np(gs_r, 0, ip_r, off).copy())
    cat = np.concatenate(parts, axis=0)
    assert cat.shape == whole.shape and (cat.view(np.uint32) == whole.view(np.uint32)).all()
    sim.close()


def test_device_resident_path_equals_host_path():
    import torch
    w = workloads.sipm8x8_scint(num_photon=40000, photons_per_genstep=100)
    sim = make_sim(w)
    h_host = sim.simulate_np(w["gensteps"], 1).copy()
    d_gs = torch.from_numpy(w["gensteps"]).cuda()
    sim.set_stream(torch.cuda.current_stream().cuda_stream)
    sim.simulate_device(d_gs.data_ptr(), len(w["gensteps"]), 0, 0, 1, 0)
    n = sim.num_hit()
    out = torch.empty((n, 4, 4), dtype=torch.float32, device="cuda")
    sim.get_hits_device(out.data_ptr())
    torch.cuda.synchronize()
    assert n == len(h_host) and (out.cpu().numpy().view(np.uint32) == h_host.view(np.uint32)).all()
    st = sim.stats()
    assert st["num_kernel"] >= 3 and st["simulate_kernel_seconds"] > 0
    sim.set_stream(0)
    sim.close()


def test_simtrace_matches_oracle_and_intersect():
    """SSimulator::simtrace (CSGOptiX7.cu:536-577): FRAME gensteps and caller-supplied rays"""
    for name, ce, cegs in (("sipm8x8_scint", (0.0, 0.0, 4.0, 12.0), [8, 0, 8, 200]), ("pmt_wall_torch", (0.0, 0.0, 0.0, 2000.0), [6, 6, 0, 100]),
                           ("boolean_zoo_torch", (0.0, 0.0, 0.0, 400.0), [5, 5, 5, 30])):
        w = workloads.WORKLOADS[name](num_photon=1000)
        sim = make_sim(w)
        gs = G.frame_gensteps(ce, cegs, gridscale=0.1)
        st = sim.simtrace(gs)
        ref = Oracle().simtrace(w["geom"], gs)
        assert st.shape == ref.shape and len(st) == int(gs.view(np.uint32)[:, 0, 3].sum())
        # origins are exact, directions differ by the ulps of sincosf (device) vs sinf/cosf (host)
        assert (st[:, 2, :3] == ref[:, 2, :3]).all() and np.abs(st[:, 3, :3] - ref[:, 3, :3]).max() < 1e-6
        su, ru = st.view(np.uint32), ref.view(np.uint32)
        same = (su[:, 2, 3] == ru[:, 2, 3]) & (su[:, 3, 3] == ru[:, 3, 3])
        assert same.mean() > 0.995, (name, same.mean())
        hit = same & (su[:, 2, 3] != 0xffffffff)
        assert hit.sum() > 0.5 * len(st)
        assert np.quantile(np.abs(st[hit, 0, 3] - ref[hit, 0, 3]) / np.maximum(1.0, ref[hit, 0, 3]), 0.999) < 1e-4
        assert np.allclose(st[hit, 1, :3], st[hit, 2, :3] + st[hit, 0, 3][:, None] * st[hit, 3, :3], rtol=0, atol=1e-3 * max(1.0, ce[3] / 100))
        miss = su[:, 2, 3] == 0xffffffff
        if miss.any():                                                          # miss program: background colour, t = 1
            assert (st[miss, 0] == np.array([0.6, 0.6, 0.6, 1.0], dtype=np.float32)).all() and (su[miss, 3, 3] == 0xffffffff).all()
        # against the reference's own generate_photon_simtrace_frame + add_simtrace compiled for the B200 (oracle/_ref)
        rg = RefGPU("debugtag").simtrace(w["geom"], gs)
        assert (st[:, 2, :3] == rg[:, 2, :3]).all()                              # origins: same transform arithmetic
        assert np.abs(st[:, 3, :3] - rg[:, 3, :3]).max() < 2e-7                  # directions: same sincosf, at most an fma contraction apart
        gu = rg.view(np.uint32)
        same_g = (su[:, 2, 3] == gu[:, 2, 3]) & (su[:, 3, 3] == gu[:, 3, 3])
        assert same_g.mean() > 0.998, (name, same_g.mean())
        hg = same_g & (su[:, 2, 3] != 0xffffffff)
        assert np.quantile(np.abs(st[hg, 0, 3] - rg[hg, 0, 3]) / np.maximum(1.0, rg[hg, 0, 3]), 0.999) < 1e-4
        assert (st[:, 1, 3] == rg[:, 1, 3]).all()                                # tmin column
        # the same rays fed back as INPUT_PHOTON_SIMTRACE: bit-identical records, and t / identities equal to phox_intersect
        rays = np.zeros_like(st); rays[:, 0, :3] = st[:, 2, :3]; rays[:, 1, :3] = st[:, 3, :3]
        st2 = sim.simtrace(G.input_simtrace_genstep(len(rays)), rays)
        assert st2.tobytes() == st.tobytes()
        prd = sim.intersect(rays[:, 0, :3], rays[:, 1, :3], tmin=0.05)
        pu = prd.view(np.uint32)
        h2 = pu[:, 1, 3] != 0xffffffff
        assert (h2 == ~miss).all() and (pu[h2, 1, 3] == su[h2, 2, 3]).all() and (pu[h2, 1, 2] == su[h2, 3, 3]).all() and (prd[h2, 0, 3] == st[h2, 0, 3]).all()
        sim.close()


def test_hit_merging_equals_oracle_bit_for_bit():
    """SPM::merge_partial_select semantics (sysrap/SPM.cu:153-290, sphoton.h:277-304): integer/byte work, so bit-exact"""
    w = workloads.pmt_wall_torch(num_photon=300000, nx=20, ny=20)
    sim = make_sim(w)
    hits = sim.simulate_np(w["gensteps"], 0, w["input_photons"]).copy()
    assert len(hits) > 1000
    orc = Oracle()
    for tw in (0.5, 5.0, 1000.0):
        a = sim.merge_hits(tw)
        b = orc.merge(hits, tw)
        assert a.shape == b.shape and a.tobytes() == b.tobytes(), tw
        key = (a.view(np.uint32)[:, 3, 1].astype(np.uint64) << np.uint64(48)) | (a[:, 0, 3] / np.float32(tw)).astype(np.uint32).astype(np.uint64)
        assert (np.diff(key.astype(np.int64)) > 0).all()                        # one record per key, ascending
        assert (a.view(np.uint32)[:, 1, 3] >> 16).sum() == len(hits)            # hitcounts add up to the unmerged hits
    assert sim.merge_hits(0.0).tobytes() == hits.tobytes()                      # window 0 = no merging
    # FinalMerge of two halves merged separately == merge of the whole (rank / launch concatenation), and any-bit selection
    tw = 5.0
    half = len(hits) // 2
    parts = np.concatenate([sim.merge(hits[:half], tw), sim.merge(hits[half:], tw)])
    assert sim.merge(parts, tw).view(np.uint32)[:, 3, 1:].tobytes() == sim.merge(hits, tw).view(np.uint32)[:, 3, 1:].tobytes()
    mixed = hits.copy(); mixed.view(np.uint32)[::3, 3, 3] = 0x8                  # every third record loses the SD bit
    assert sim.merge(mixed, tw, select_mask=0x40).tobytes() == orc.merge(mixed, tw, select_mask=0x40).tobytes()
    big = np.tile(hits, (max(1, 300000 // len(hits)), 1, 1))                    # several sort tiles per digit, long groups
    assert sim.merge(big, 2.0).tobytes() == orc.merge(big, 2.0).tobytes()
    sim.close()


def test_lite_hits_and_lite_merging():
    """sphotonlite hits (sysrap/sphotonlite.h; raygen CSGOptiX7.cu:455-463) and their merge; both forms of the loop"""
    w = workloads.pmt_wall_torch(num_photon=100000, nx=20, ny=20)
    orc = Oracle()
    ref = orc.simulate(w["geom"], w["gensteps"], w["input_photons"], max_bounce=w["config"].get("max_bounce", 31), lite=True)
    res = {}
    for mode in (ph.KERNEL_PERSISTENT, ph.KERNEL_WAVEFRONT):
        sim = make_sim(w, event_mode=ph.MODE_HITPHOTON, mode_lite=1, kernel_mode=mode)
        hits = sim.simulate_np(w["gensteps"], 0, w["input_photons"]).copy()
        lite = sim.get_hits_lite()
        assert lite.shape == (len(hits), 4) and len(hits) > 1000
        hu = hits.view(np.uint32)
        assert (lite[:, 0] == ((1 << 16) | (hu[:, 3, 1] & 0xffff))).all() and (lite[:, 1] == hu[:, 0, 3]).all() and (lite[:, 3] == hu[:, 3, 3]).all()
        # packed local positions against the CPU oracle on the photons whose history agrees (float ulps move the u16 by <= 1)
        p = sim.get_array("photon")
        idx = hu[:, 3, 2]
        same = (p.view(np.uint32)[idx, 3, :] == ref["photon"].view(np.uint32)[idx, 3, :]).all(axis=1)
        a, b = lite[same, 2], ref["lite"][idx[same], 2]
        assert same.mean() > 0.99
        assert (np.abs((a >> 16).astype(int) - (b >> 16).astype(int)) <= 1).mean() > 0.999
        dphi = np.abs((a & 0xffff).astype(int) - (b & 0xffff).astype(int)); dphi = np.minimum(dphi, 65535 - dphi)
        assert (dphi <= 2).mean() > 0.995
        assert ((a >> 16) > 0).mean() > 0.5                                    # hits on the front hemisphere of the bulb
        for tw in (1.0, 50.0):
            m = sim.merge_hits_lite(tw)
            assert m.tobytes() == orc.merge_lite(lite, tw).tobytes()
            assert (m[:, 0] >> 16).sum() == len(hits)
        assert sim.merge_hits_lite(0.0).tobytes() == lite.tobytes()
        res[mode] = (hits, lite)
        sim.close()
    assert res[ph.KERNEL_PERSISTENT][1].tobytes() == res[ph.KERNEL_WAVEFRONT][1].tobytes()
    sim = make_sim(w)
    sim.simulate_np(w["gensteps"], 0, w["input_photons"])
    with pytest.raises(ph.lib.PhoxError):
        sim.get_hits_lite()                                                     # mode_lite off: loud, not empty
    sim.close()


def test_edge_cases_empty_ragged_offsets_and_limits():
    """empty and ragged gensteps, a zero-bounce event, absolute photon indices beyond 2^32 (sphoton::set_index puts bits
    32..39 into the top byte of identity, sysrap/sphoton.h:224-232), a genstep larger than max_slot"""
    w = workloads.sipm8x8_scint(num_photon=20000, photons_per_genstep=100)
    gs = w["gensteps"]
    for mode in (ph.KERNEL_PERSISTENT, ph.KERNEL_WAVEFRONT):
        sim = make_sim(w, event_mode=ph.MODE_HITPHOTONSEQ, kernel_mode=mode)
        base = sim.simulate_np(gs, 0).copy()
        p0 = sim.get_array("photon").copy()
        # ragged: zero-photon gensteps interleaved do not move any photon index
        empty = gs[:3].copy(); empty.view(np.uint32)[:, 0, 3] = 0
        ragged = np.concatenate([empty[:1], gs[:50], empty[1:2], gs[50:], empty[2:]])
        assert sim.simulate_np(ragged, 0).tobytes() == base.tobytes() and sim.get_array("photon").tobytes() == p0.tobytes()
        # only empty gensteps: an event with no photons and no hits
        h = sim.simulate_np(empty, 0)
        assert h.shape == (0, 4, 4) and sim.num_hit() == 0 and len(sim.get_array("photon")) == 0
        # one photon
        one = gs[:1].copy(); one.view(np.uint32)[0, 0, 3] = 1
        sim.simulate_np(one, 0)
        assert sim.get_array("photon").shape == (1, 4, 4) and (sim.get_array("photon").view(np.uint32)[0, 3, 2] == 0)
        # zero bounces: the generated photons themselves
        sim.set_config(max_bounce=0)
        assert len(sim.simulate_np(gs, 0)) == 0
        g0 = sim.get_array("photon")
        assert (g0.view(np.uint32)[:, 3, 0] & 0xffff).tolist().count(2) + (g0.view(np.uint32)[:, 3, 0] & 0xffff).tolist().count(1) == len(g0)   # SI or CK
        assert (sim.get_array("seq")[:, 0, 0] < 16).all()
        sim.set_config(max_bounce=w["config"].get("max_bounce", 31))
        # absolute index beyond 2^32
        off = (5 << 32) + 123
        hb = sim.simulate_np(gs[:20], 0, None, off)
        pb = sim.get_array("photon")
        assert (pb.view(np.uint32)[:, 3, 2] == (np.arange(len(pb), dtype=np.uint64) + np.uint64(123)).astype(np.uint32)).all()
        assert ((pb.view(np.uint32)[:, 3, 1] >> 24) == 5).all() and ((hb.view(np.uint32)[:, 3, 1] >> 24) == 5).all()
        ref = Oracle().simulate(w["geom"], gs[:20], None, photon_offset=off, max_bounce=w["config"].get("max_bounce", 31))
        same = (pb.view(np.uint32)[:, 3, :] == ref["photon"].view(np.uint32)[:, 3, :]).all(axis=1)
        assert same.mean() > 0.99
        # a genstep that does not fit max_slot is an error, not a truncation
        sim.set_config(max_slot=50)
        with pytest.raises(ph.lib.PhoxError):
            sim.simulate_np(gs, 0)
        sim.close()


@pytest.mark.parametrize("name,kw", [CASES[0], CASES[5], CASES[3]])
def test_tag_and_flat_arrays_equal_the_reference_tagr(name, kw):
    """DebugHeavy tag / flat arrays (sysrap/stag.h, written by sctx::end) against the reference's own stagr running inside its
    qsim.h on the B200 (as-built DEBUG_TAG build), for both forms of the loop: integer work, so bit-exact on every photon whose
    history agrees"""
    w = workloads.WORKLOADS[name](**dict(kw, num_photon=10000))
    ref = RefGPU("debugtag").simulate(w["geom"], w["gensteps"], w["input_photons"], max_bounce=w["config"].get("max_bounce", 31), tags=True)
    orc = Oracle().simulate(w["geom"], w["gensteps"], w["input_photons"], max_bounce=w["config"].get("max_bounce", 31), tags=True)
    got = {}
    for mode in (ph.KERNEL_PERSISTENT, ph.KERNEL_WAVEFRONT):
        sim = make_sim(w, event_mode=ph.MODE_DEBUGHEAVY, kernel_mode=mode)
        sim.simulate_np(w["gensteps"], 0, w["input_photons"])
        tag, flat, seq = sim.get_array("tag").copy(), sim.get_array("flat").copy(), sim.get_array("seq").copy()
        assert tag.shape == (10000, 4) and flat.shape == (10000, 64)
        same = (seq == ref["seq"]).all(axis=(1, 2))
        assert same.mean() > 0.995
        assert (tag[same] == ref["tag"][same]).all(), name
        assert (flat[same].view(np.uint32) == ref["flat"][same].view(np.uint32)).all(), name
        same_o = (seq == orc["seq"]).all(axis=(1, 2))
        assert (tag[same_o] == orc["tag"][same_o]).all() and (flat[same_o].view(np.uint32) == orc["flat"][same_o].view(np.uint32)).all()
        got[mode] = (tag, flat)
        sim.close()
    assert got[ph.KERNEL_PERSISTENT][0].tobytes() == got[ph.KERNEL_WAVEFRONT][0].tobytes()
    assert got[ph.KERNEL_PERSISTENT][1].tobytes() == got[ph.KERNEL_WAVEFRONT][1].tobytes()
    sim = make_sim(w, event_mode=ph.MODE_DEBUGLITE)
    sim.simulate_np(w["gensteps"], 0, w["input_photons"])
    assert len(sim.get_array("tag")) == 0 and len(sim.get_array("flat")) == 0                  # only DebugHeavy keeps them
    sim.close()


def test_event_index_skipahead_and_rng_sequence():
    w = workloads.sipm8x8_scint(num_photon=1000, photons_per_genstep=100)
    sim = make_sim(w)
    u = sim.rng_sequence(64, 16, id0=5, event_id=2)
    want = G.curand_uniform_matrix(0, 5, 64, 200000, 16)
    assert (u.view(np.uint32) == want.view(np.uint32)).all()
    a = sim.simulate_np(w["gensteps"], 0).copy()
    b = sim.simulate_np(w["gensteps"], 1).copy()
    c = sim.simulate_np(w["gensteps"], 0)
    assert (a.view(np.uint32) == c.view(np.uint32)).all()
    assert a.shape != b.shape or not (a.view(np.uint32) == b.view(np.uint32)).all()
    sim.close()


def test_event_modes_keep_the_documented_arrays():
    w = workloads.boolean_zoo_torch(num_photon=5000)
    sim = make_sim(w, event_mode=ph.MODE_MINIMAL)
    h0 = sim.simulate_np(w["gensteps"], 0).copy()
    assert len(sim.get_array("photon")) == 0 and len(sim.get_array("seq")) == 0
    sim.set_config(event_mode=ph.MODE_DEBUGLITE, max_record=10)
    h1 = sim.simulate_np(w["gensteps"], 0)
    assert (h0.view(np.uint32) == h1.view(np.uint32)).all()
    rec, seq, p = sim.get_array("record"), sim.get_array("seq"), sim.get_array("photon")
    assert rec.shape == (5000, 10, 4, 4) and seq.shape == (5000, 2, 2) and p.shape == (5000, 4, 4)
    assert (rec[:, 0].view(np.uint32)[:, 3, 0] & 0xffff == 4).all()             # slot 0 holds the TORCH generation flag
    nib0 = (seq[:, 0, 0] & np.uint64(0xf)).astype(int)
    assert (nib0 == 3).all()                                                    # FFS(TORCH)
    sim.close()


def test_bad_arguments_give_error_codes_not_crashes():
    g = ph.geometries.raindrop()
    sim = ph.Simulator()
    with pytest.raises(ph.PhoxError):
        sim.simulate_np(G.input_photon_genstep(4), 0, np.zeros((4, 4, 4), np.float32))   # no geometry yet
    sim.set_geometry(g["foundry"])
    with pytest.raises(ph.PhoxError):
        sim.simulate_np(G.input_photon_genstep(4), 0, np.zeros((4, 4, 4), np.float32))   # no tables yet
    sim.set_tables(g["bnd"], g["optical"])
    with pytest.raises(ph.PhoxError):
        sim.simulate_np(G.input_photon_genstep(5), 0, np.zeros((4, 4, 4), np.float32))   # count mismatch
    with pytest.raises(ph.PhoxError):
        sim.simulate_np(G.scint_gensteps([[0, 0, 0]], [0, 0, 1], 1.0, [10], 3, 1.0), 0)  # scintillation without icdf
    bad = {k: v.copy() for k, v in g["foundry"].items() if hasattr(v, "copy")}
    bad["prim"].view(np.int32)[0, 0, 1] = 10 ** 6
    with pytest.raises(ph.PhoxError):
        sim.set_geometry(bad)
    with pytest.raises(ph.PhoxError):
        sim.set_config(max_record=99)
    sim.close()


def test_cxx_file_source_driver_known_answer(tmp_path):
    """the C++ host driver (apps/PhoxPhotonFileSource.cpp on include/PhoxSimulator.h) reproduces the
    GPUPhotonFileSource contract: 10 file photons -> 'Opticks: NumHits:  10', hit text file with the
    wavelengths preserved (tests/test_GPUPhotonFileSource.sh:23-118)."""
    import subprocess
    from eic_opticks_b200 import foundry as F
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "eic-opticks_b200", "apps", "PhoxPhotonFileSource")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    F.save_geometry(ph.geometries.raindrop(), str(tmp_path / "geom"))
    out = tmp_path / "opticks_hits_output.txt"
    r = subprocess.run([exe, "-g", str(tmp_path / "geom"), "-p", os.path.join(root, "tests", "golden", "photons_file_source.txt"), "-o", str(out)],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "Loaded 10 photons" in r.stdout and "Opticks: NumHits:  10" in r.stdout
    lines = open(out).read().strip().split("\n")
    assert len(lines) == 10
    assert sorted(float(l.split()[1]) for l in lines) == [420.0] * 3 + [450.0] * 2 + [500.0] * 5
    r2 = subprocess.run([exe, "-g", str(tmp_path / "geom")], capture_output=True, text=True, timeout=60)      # missing -p must fail
    assert r2.returncode != 0


def test_cxx_torch_driver_matches_python_path(tmp_path):
    """apps/PhoxPhotonSourceMinimal.cpp (GPUPhotonSourceMinimal contract): config/dev.json torch on the raindrop"""
    import subprocess
    from eic_opticks_b200 import foundry as F
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "eic-opticks_b200", "apps", "PhoxPhotonSourceMinimal")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    g = ph.geometries.raindrop()
    F.save_geometry(g, str(tmp_path / "geom"))
    cfg = os.path.join(root, "tests", "golden", "config_dev.json")
    t, _ = G.torch_config(cfg)
    sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"])
    want_host = len(sim.simulate_np(G.input_photon_genstep(t["numphoton"]), 0, G.torch_photons(t, seed=0)))
    want_gs = len(sim.simulate_np(G.torch_genstep(t), 0))
    sim.close()
    for extra, want in (([], want_host), (["--genstep"], want_gs)):
        out = tmp_path / "hits.txt"
        r = subprocess.run([exe, "-g", str(tmp_path / "geom"), "-c", cfg, "-o", str(out)] + extra, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        assert "Opticks: NumHits:  %d" % want in r.stdout, (r.stdout, want)
        assert want > 50 and len(open(out).read().strip().split("\n")) == want


@pytest.mark.parametrize("name,kw", [("sipm8x8_scint", dict(num_photon=200000, photons_per_genstep=500)), ("pmt_wall_torch", dict(num_photon=60000, nx=20, ny=20))])
def test_cxx_multi_gpu_host_equals_single_gpu(tmp_path, name, kw):
    """apps/PhoxMultiGPU.cpp on include/PhoxMultiGPU.h: one thread + context per rank, genstep ranges with absolute photon
    offsets (input photons: photon ranges), hits into one pinned buffer at the prefix offsets of the ranks' counts while the
    next event runs.  The bytes must be those of one context simulating the whole event.  On a 1-GPU box the ranks share
    device 0 (--devices 0,0,0), which exercises the same host logic."""
    import subprocess
    import torch
    from eic_opticks_b200 import foundry as F
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "eic-opticks_b200", "apps", "PhoxMultiGPU")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    w = workloads.WORKLOADS[name](**kw)
    g = w["geom"]
    F.save_geometry(g, str(tmp_path / "geom"))
    np.save(tmp_path / "gs.npy", np.ascontiguousarray(w["gensteps"], dtype=np.float32))
    args = [exe, "-g", str(tmp_path / "geom"), "-G", str(tmp_path / "gs.npy"), "--events", "3", "--max-bounce", str(w["config"].get("max_bounce", 31))]
    if w["input_photons"] is not None:
        np.save(tmp_path / "ip.npy", np.ascontiguousarray(w["input_photons"], dtype=np.float32))
        args += ["-I", str(tmp_path / "ip.npy")]
    sim = make_sim(w)
    want = sim.simulate_np(w["gensteps"], 2, w["input_photons"]).copy()          # the app writes the hits of its last event, id 2
    sim.close()
    assert len(want) > 100
    ndev = torch.cuda.device_count()
    layouts = ["0", "0,0,0"] + ([",".join(str(d) for d in range(ndev))] if ndev > 1 else [])
    for devs in layouts:
        out = tmp_path / ("hits_%s.npy" % devs.replace(",", "_"))
        r = subprocess.run(args + ["--devices", devs, "-o", str(out)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (r.stdout, r.stderr)
        got = np.load(out)
        assert got.shape == want.shape, (devs, got.shape, want.shape)
        assert got.tobytes() == want.tobytes(), devs
        assert "Opticks: NumHits:  %d" % len(want) in r.stdout


def test_oracle_texture_emulation_vs_hardware():
    """the CPU oracle's restatement of CUDA linear texture filtering (8-bit fraction, lerp form) against the
    B200 texture unit on a dispersive boundary table at random fractional wavelengths"""
    g = ph.geometries.scintillator_tank()
    sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"], g["icdf"])
    rng = np.random.default_rng(0)
    n = 100000
    nm = rng.uniform(60, 820, n).astype(np.float32)
    nm[:1000] = np.round(nm[:1000])
    line = rng.integers(0, 4 * len(g["bnd_names"]), n).astype(np.uint32)
    k = rng.integers(0, 2, n).astype(np.uint32)
    hw = sim.boundary_lookup(nm, line, k)
    tex = np.ascontiguousarray(g["bnd"].reshape(-1, 761, 4), dtype=np.float32)
    x = ((nm - np.float32(60.0)) / np.float32(1.0) + np.float32(0.5)) / np.float32(761)
    y = ((2 * line + k).astype(np.float32) + np.float32(0.5)) / np.float32(tex.shape[0])
    emu = Oracle().tex2d4(tex, np.stack([x, y], axis=1))               # tex is (ny, nx, 4)
    same = (emu.view(np.uint32) == hw.view(np.uint32)).all(axis=1)
    rel = np.abs(emu - hw) / np.maximum(np.abs(hw), 1e-6)
    assert same.mean() > 0.95, same.mean()
    assert np.quantile(rel, 0.999) < 1e-4 and rel.max() < 0.03
    assert same[:1000].all()                        # integer wavelengths hit table samples exactly
    sim.close()


def test_reference_side_ssimulator_and_scompprovider(tmp_path):
    """include/PhoxSimulator.h compiled inside the reference's own header tree (SSimulator.h, SComp.h, NP.hh,
    sslice.h read in place) and driven only through SSimulator* / SCompProvider* with the slice loop of QSim::simulate
    (oracle/ref_ssimulator_test.cc -> oracle/_ref/phox_ssimulator_test; docs/reference_side.patch shows the QSim / G4CXOpticks hunks).
    The hit array SEvt would receive equals the Python path's, byte for byte."""
    import subprocess
    from eic_opticks_b200 import foundry as F
    exe = os.path.join(ORACLE, "_ref", "phox_ssimulator_test")
    if not os.path.exists(exe):
        pytest.fail("oracle/_ref/phox_ssimulator_test missing: run __graft_entry__.build() where /root/reference exists")
    w = workloads.sipm8x8_scint(num_photon=60000, photons_per_genstep=100)
    F.save_geometry(w["geom"], str(tmp_path / "geom"))
    np.save(tmp_path / "igs.npy", w["gensteps"])
    out = tmp_path / "hit.npy"
    r = subprocess.run([exe, str(tmp_path / "geom"), str(tmp_path / "igs.npy"), "7000", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("PASS"), r.stdout + r.stderr
    hit = np.load(out)
    sim = make_sim(w)
    sim.set_config(max_bounce=31)                                  # the adaptor runs phox_default_config (SEventConfig's max_bounce 31)
    want = sim.simulate_np(w["gensteps"], 3)
    assert hit.shape == want.shape and hit.tobytes() == want.tobytes()
    sim.close()


def test_empty_shards_and_device_path_validation():
    """(i) a rank's share may hold no gensteps - an empty event, not an error; (ii) the device-resident path checks the
    gensteps like the host path does (no icdf for scintillation, input-photon genstep without / with the wrong photons);
    (iii) malformed foundries are refused before they reach the BVH builder"""
    import torch
    w = workloads.sipm8x8_scint(num_photon=3000, photons_per_genstep=1000)           # 3 gensteps
    sim = make_sim(w)
    whole = sim.simulate_np(w["gensteps"], 0).copy()
    parts = []
    for r in range(8):                                                            # more ranks than gensteps
        gs_r, ip_r, off, cnt = parallel.shard_event(w["gensteps"], r, 8)
        h = sim.simulate_np(gs_r, 0, ip_r, off)
        assert sim.num_photon() == cnt and (cnt > 0 or (len(gs_r) == 0 and len(h) == 0))
        parts.append(h.copy())
    assert sum(len(p) == 0 for p in parts) >= 5
    assert np.concatenate(parts).tobytes() == whole.tobytes()
    sim.simulate_device(0, 0)                                                      # empty device-resident event
    assert sim.num_hit() == 0 and sim.num_photon() == 0
    sim.close()
    # (ii)
    g = ph.geometries.raindrop()
    sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"])                # no icdf
    d_sc = torch.from_numpy(G.scint_gensteps([[0, 0, 0]], [0, 0, 1], 1.0, [10], 3, 1.0)).cuda()
    with pytest.raises(ph.PhoxError):
        sim.simulate_device(d_sc.data_ptr(), 1)
    d_ip_gs = torch.from_numpy(G.input_photon_genstep(5)).cuda()
    d_ph = torch.zeros((4, 4, 4), dtype=torch.float32, device="cuda")
    with pytest.raises(ph.PhoxError):
        sim.simulate_device(d_ip_gs.data_ptr(), 1, 0, 0)                           # no photons
    with pytest.raises(ph.PhoxError):
        sim.simulate_device(d_ip_gs.data_ptr(), 1, d_ph.data_ptr(), 4)             # 5 announced, 4 given
    ip = G.photons_from_text(os.path.join(os.path.dirname(__file__), "golden", "photons_file_source.txt"))
    d_ok_gs, d_ok = torch.from_numpy(G.input_photon_genstep(len(ip))).cuda(), torch.from_numpy(ip).cuda()
    sim.simulate_device(d_ok_gs.data_ptr(), 1, d_ok.data_ptr(), len(ip))           # the context survived the refusals
    assert sim.num_hit() == 10
    # (iii)
    fd = g["foundry"]
    for mutate in ("inf_box", "inverted_box", "empty_solid_instance", "list_range", "tree_subnum"):
        bad = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in fd.items()}
        if mutate == "inf_box":
            bad["prim"][1, 2, 0] = np.inf
        elif mutate == "inverted_box":
            bad["prim"][1, 2, 0], bad["prim"][1, 2, 3] = 50.0, -50.0
        elif mutate == "empty_solid_instance":
            bad["solid"] = np.concatenate([bad["solid"], np.zeros((1, 3, 4), np.int32)])
            inst = np.concatenate([bad["inst"], bad["inst"][:1]]); inst.view(np.int32)[1, 1, 3] = 1
            bad["inst"] = inst
        elif mutate == "list_range":
            z = ph.geometries.boolean_zoo()["foundry"]
            bad = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in z.items()}
            tc = bad["node"].view(np.uint32)[:, 3, 2]
            k = int(np.flatnonzero(tc == 12)[0])                                   # a discontiguous list node
            bad["node"].view(np.uint32)[k, 0, 0] = 1000
        else:
            z = ph.geometries.boolean_zoo()["foundry"]
            bad = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in z.items()}
            tc = bad["node"].view(np.uint32)[:, 3, 2]
            k = int(np.flatnonzero((tc >= 1) & (tc <= 3))[0])                      # first tree root
            bad["node"].view(np.uint32)[k, 0, 0] = 6
        with pytest.raises(ph.PhoxError):
            sim.set_geometry(bad)
    sim.close()
