"""Photon-by-photon comparison of libphox.so with the reference's device headers (oracle/_ref), with a
first-divergence analysis of every photon whose history differs.  TEST INFRASTRUCTURE.

    python tests/_parity.py --build default --out profiles/parity_r2.json      # the shipped build vs oracle/_ref/libphoxref_*.so
    python tests/_parity.py --build nofma   --out profiles/parity_r2_nofma.json  # both sides built with -fmad=false

`--build nofma` selects eic-opticks_b200/csrc/libphox_nofma.so and oracle/_ref/libphoxref_*_nofma.so: the same sources
compiled with FMA contraction off.  Every float operation then rounds on its own, so the two code bases must agree bit
for bit if (and only if) they evaluate the same expressions in the same order; tests/test_parity_gpu.py asserts exactly
that.  In the default build nvcc fuses a*b+c differently in the two code bases (it does so even between two inlining
contexts of ONE code base), which moves the last bit of some intermediate results; a photon whose branch variable lands
within that bit of its threshold takes the other branch.  The report lists each such photon with the bounce where the
histories part, the decision that flipped and how far apart (in ulps) the two sides were just before it.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [("sipm8x8_scint", dict(num_photon=30000, photons_per_genstep=100)),
         ("raindrop_cerenkov", dict(num_photon=20000, photons_per_genstep=100)),
         ("sphere_leak_torch", dict(num_photon=10000)),
         ("pmt_wall_torch", dict(num_photon=30000, nx=20, ny=20)),
         ("boolean_zoo_torch", dict(num_photon=40000)),
         ("scintillator_tank", dict(num_photon=30000, photons_per_genstep=100))]
# arms of the path the six workloads above do not reach: PropagateRefine beyond 5000 mm, every torch
# source type on the device, carrier gensteps, the zplus_sensor_A surface arm incl. lposcost < 0, halfspace-cut solids
ARM_CASES = [("far_wall_torch", dict(num_photon=20000)),
             ("torch_shapes", dict(num_photon=21000)),
             # carrier: PRODUCTION build only.  scarrier::generate stores float4s through (quad4&)sphoton (sysrap/scarrier.h:49-56); in the
             # as-built DEBUG_TAG layout offsetof(sctx, p) is 36, so those stores are misaligned and the reference kernel itself faults
             ("carrier_photons", dict(num_photon=150, variants=("production",))),
             ("pmt_wall_sensor_a", dict(num_photon=30000)),
             ("halfspace_zoo_torch", dict(num_photon=30000)),
             # touching / nested boxes, photons starting on faces, edges and corners, axis-parallel and in-plane directions: ties everywhere
             ("box_maze_photons", dict(num_photon=40000)),
             # the pfRICH geometry the reference ships (tests/geom/pfrich_min_FINAL.gdml through gdml.py, tests/golden/pfrich_min_geometry.npz)
             ("pfrich_photons", dict(num_photon=40000))]

FLAG_NAMES = {1: "CK", 2: "SI", 4: "TO", 8: "AB", 16: "RE", 32: "SC", 64: "SD", 128: "SA", 256: "DR", 512: "SR", 1024: "BR", 2048: "BT", 0: "--"}


def ulp_distance(a, b):
    """distance between float32 arrays in units of the last place (0 = same bits, +0 and -0 count as equal)"""
    ia = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    ib = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7fffffff), ia)
    ib = np.where(ib < 0, -(ib & 0x7fffffff), ib)
    return np.abs(ia - ib)


def integer_identity(p, seq, ref_p, ref_seq):
    """per photon: q3 (orient|boundary|flag, identity, index, flagmask), hitcount|iindex, seqhis and seqbnd all equal"""
    pu, ru = p.view(np.uint32), ref_p.view(np.uint32)
    same = (pu[:, 3, :] == ru[:, 3, :]).all(axis=1) & (pu[:, 1, 3] == ru[:, 1, 3])
    if seq is not None and ref_seq is not None:
        same &= (seq == ref_seq).all(axis=(1, 2))
    return same


def decision_name(fa, fb, same_prim, same_face):
    if not same_prim:
        return "nearest surface: two candidate intersects of different prims closer together than their rounding"
    if not same_face:
        return "CSG solid: another face / root of the same prim (grazing or coincident constituent surfaces)"
    s = {fa, fb}
    if s == {1024, 2048}:
        return "Fresnel: u_reflect against TransCoeff"
    if s & {8, 16, 32}:
        return "bulk: absorption / scattering distance against the distance to the boundary (or u against reemission_prob)"
    if s <= {64, 128, 256, 512}:
        return "surface: u_surface against the absorb / detect / diffuse thresholds"
    return "other"


def scaled_err(a, b):
    """max |a - b| / max(1, |b|) : the north_star float metric (components near zero do not blow it up like ulps do)"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max())


def first_divergence(i, a, b):
    """a, b: dicts with record (n, R, 4, 4), prd (n, R, 2, 4); returns a small dict describing where photon i parts"""
    ra, rb = a["record"][i], b["record"][i]
    ua, ub = ra.view(np.uint32), rb.view(np.uint32)
    R = ra.shape[0]
    differs = [(ua[k, 3, 0] != ub[k, 3, 0]) or (ua[k, 3, 1] != ub[k, 3, 1]) or (ua[k, 3, 3] != ub[k, 3, 3]) or (ua[k, 1, 3] != ub[k, 1, 3]) for k in range(R)]
    if not any(differs):
        return {"photon": int(i), "bounce": None, "decision": "beyond the recorded steps"}
    k = differs.index(True)
    out = {"photon": int(i), "bounce": int(k - 1), "flag_phox": FLAG_NAMES.get(int(ua[k, 3, 0] & 0xffff), hex(int(ua[k, 3, 0] & 0xffff))),
           "flag_ref": FLAG_NAMES.get(int(ub[k, 3, 0] & 0xffff), hex(int(ub[k, 3, 0] & 0xffff)))}
    if k == 0:
        out["decision"] = "generation"
        return out
    pa, pb = a["prd"][i, k - 1], b["prd"][i, k - 1]
    same_prim = bool((pa.view(np.uint32)[1, 2:] == pb.view(np.uint32)[1, 2:]).all())
    gap_t = float(abs(float(pa[0, 3]) - float(pb[0, 3])) / max(1.0, abs(float(pb[0, 3]))))
    same_face = same_prim and gap_t < 1e-4 and scaled_err(pa[0, :3], pb[0, :3]) < 1e-2
    out["decision"] = decision_name(int(ua[k, 3, 0] & 0xffff), int(ub[k, 3, 0] & 0xffff), same_prim, same_face)
    out["same_prim"], out["same_face"] = same_prim, bool(same_face)
    # how far apart were the two sides going INTO the deciding bounce: photon state after the previous step and the hit distance
    out["err_state_before"] = scaled_err(ra[k - 1, :3, :], rb[k - 1, :3, :])
    out["ulp_time_before"] = int(ulp_distance(ra[k - 1, 0, 3:4], rb[k - 1, 0, 3:4])[0])
    out["t_phox"], out["t_ref"] = float(pa[0, 3]), float(pb[0, 3])
    out["ulp_hit_t"] = int(ulp_distance(pa[0, 3:4], pb[0, 3:4])[0])
    out["gap_hit_t"] = gap_t
    return out


def compare_case(name, kw, build, accels=(1, 0), max_listed=40):
    """one workload, both RNG-consumption modes; returns the report entries"""
    import eic_opticks_b200 as ph
    from eic_opticks_b200 import workloads
    from _ref import RefGPU
    sfx = "_nofma" if build == "nofma" else ""
    kw = dict(kw)
    variants = kw.pop("variants", ("debugtag", "production"))
    w = workloads.WORKLOADS[name](**kw)
    g = w["geom"]
    entries = []
    for variant in variants:
        ref = RefGPU(variant + sfx).simulate(g, w["gensteps"], w["input_photons"], max_bounce=w["config"].get("max_bounce", 31),
                                             refine=w["config"].get("propagate_refine", 0), refine_distance=w["config"].get("refine_distance", 5000.0))
        kwc = dict(w["config"])
        kwc.update(event_mode=ph.MODE_DEBUGHEAVY, rng_mode=(ph.RNG_DEBUG_TAG if variant == "debugtag" else ph.RNG_PRODUCTION))
        sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"], g["icdf"], **kwc)
        for accel in accels:
            sim.set_config(accel=accel)
            sim.simulate_np(w["gensteps"], 0, w["input_photons"])
            got = {k: sim.get_array(k).copy() for k in ("photon", "seq", "record", "prd")}
            same = integer_identity(got["photon"], got["seq"], ref["photon"], ref["seq"])
            n = len(same)
            e = {"workload": name, "photons": int(n), "rng_mode": variant, "accel": "bvh" if accel == 0 else "brute",
                 "rays_phox": int(sim.stats()["num_ray"]), "rays_ref": int(ref["nray"]),
                 "identical_integer_data": int(same.sum()), "identical_fraction": float(same.mean())}
            fa, fb = got["photon"][same][:, :3, :], ref["photon"][same][:, :3, :]
            ulp = ulp_distance(fa, fb)
            rel = np.abs(fa - fb) / np.maximum(1.0, np.abs(fb))
            e["float_bits_identical_fraction_of_matching"] = float((ulp.reshape(len(ulp), -1).max(axis=1) == 0).mean()) if len(ulp) else 1.0
            e["max_rel_err_matching"] = float(rel.max()) if rel.size else 0.0
            e["q9995_rel_err_matching"] = float(np.quantile(rel, 0.9995)) if rel.size else 0.0
            e["max_ulp_matching"] = int(ulp.max()) if ulp.size else 0
            if variant == "debugtag":
                sm = (got["seq"] == ref["seq"]).all(axis=(1, 2))
                rr = np.abs(got["record"][sm][:, :, :3, :] - ref["record"][sm][:, :, :3, :]) / np.maximum(1.0, np.abs(ref["record"][sm][:, :, :3, :]))
                e["max_rel_err_step_records"] = float(rr.max()) if rr.size else 0.0
                # matching histories whose floats are further apart than the 1e-4 of north_star: where the gap opens and how
                # ill-conditioned that intersect was (|cos| of the incidence angle: a grazing hit on a curved surface takes the
                # square root of a tiny discriminant, which turns one ulp of its input into a large relative error of t)
                per_photon = rel.reshape(len(rel), -1).max(axis=1) if len(rel) else np.zeros(0)
                idx_same = np.flatnonzero(same)
                outl = []
                for j in np.flatnonzero(per_photon > 1e-4)[:max_listed]:
                    i = idx_same[j]
                    ra, rb = got["record"][i], ref["record"][i]
                    step_err = [scaled_err(ra[k, :3, :], rb[k, :3, :]) for k in range(ra.shape[0])]
                    ks = [k for k, v in enumerate(step_err) if v > 1e-5]
                    o = {"photon": int(i), "final_err": float(per_photon[j]), "first_step_over_1e-5": int(ks[0]) if ks else None}
                    if ks and ks[0] > 0:
                        k = ks[0]
                        nrm, mom = ref["prd"][i, k - 1, 0, :3].astype(np.float64), rb[k - 1, 1, :3].astype(np.float64)
                        o["abs_cos_incidence"] = float(abs(np.dot(nrm, mom)))
                        o["gap_hit_t"] = float(abs(float(got["prd"][i, k - 1, 0, 3]) - float(ref["prd"][i, k - 1, 0, 3])) / max(1.0, abs(float(ref["prd"][i, k - 1, 0, 3]))))
                        o["steps_after"] = int(np.count_nonzero(rb.view(np.uint32)[k:, 3, 3]))
                    outl.append(o)
                e["matching_photons_over_1e-4"] = int((per_photon > 1e-4).sum())
                e["float_outliers_listed"] = outl
                bad = np.flatnonzero(~same)
                div = [first_divergence(i, got, ref) for i in bad[:max_listed]]
                e["divergent_photons_listed"] = div
                allb = [first_divergence(i, got, ref) for i in bad]
                tally = {}
                for d in allb:
                    tally[d["decision"]] = tally.get(d["decision"], 0) + 1
                e["divergence_tally"] = tally
                errs = [d["err_state_before"] for d in allb if d.get("err_state_before") is not None]
                e["max_err_state_before_divergence"] = float(max(errs)) if errs else 0.0
                e["median_err_state_before_divergence"] = float(np.median(errs)) if errs else 0.0
            entries.append(e)
        sim.close()
    return entries


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", default="default", choices=["default", "nofma"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--cases", default="")
    args = ap.parse_args()
    if args.build == "nofma":
        os.environ["PHOX_LIB"] = os.path.join(ROOT, "eic-opticks_b200", "csrc", "libphox_nofma.so")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    want = [c for c in args.cases.split(",") if c]
    entries = []
    for name, kw in CASES + ARM_CASES:
        if want and name not in want:
            continue
        entries += compare_case(name, kw, args.build)
        for e in entries[-4:]:
            print("%-20s %-10s %-5s identical %.6f (%d of %d) float-bits-identical %.4f max_rel %.3g" % (
                e["workload"], e["rng_mode"], e["accel"], e["identical_fraction"], e["identical_integer_data"], e["photons"],
                e["float_bits_identical_fraction_of_matching"], e["max_rel_err_matching"]), flush=True)
    rep = {"build": args.build, "library": os.environ.get("PHOX_LIB", "eic-opticks_b200/csrc/libphox.so"),
           "reference": "oracle/_ref/libphoxref_{debugtag,production}%s.so (reference device headers, brute-force closest hit)" % ("_nofma" if args.build == "nofma" else ""),
           "all_identical": all(e["identical_fraction"] == 1.0 for e in entries), "entries": entries}
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(rep, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
