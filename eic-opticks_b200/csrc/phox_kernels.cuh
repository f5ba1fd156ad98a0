// phox_kernels.cuh : the event kernels.
//
//   trace_inline() / trace()   nearest CSGPrim along a ray: instance BVH -> per-solid BVH -> prim test -> hit_finish.
//                  Replaces optixTrace + __intersection__is + __closesthit__ch + __miss__ms
//                  (CSGOptiX/CSGOptiX7.cu:110-216, 655-682, 749-847, 869-940).
//   wavefront form of the simulate raygen (CSGOptiX7.cu:405-503), the default:
//     k_wf_generate   seed lookup (binary search in the prefix sum of genstep.numphoton: no seed array, replaces
//                     qudarap/QEvt.cu:181-237 + sysrap/iexpand.h:86-142), RNG init, generate, photon + list entry
//     k_wf_trace      one ray per live photon -> 32 B hit record
//     k_wf_propagate  hit record -> qsim::propagate -> photon; survivors appended, in order, to the next list
//   k_simulate     persistent form: the whole loop per photon in one kernel, warps refill finished lanes.
//   k_hit_count / k_hit_offsets / k_hit_compact   stable stream compaction of hit photons (and their sphotonlite
//                  records), ascending photon index.  Replaces thrust count_if + copy_if (sysrap/SU.cu:48-56, 91-92, 157-159).
//   k_simtrace, k_intersect, k_boundary_lookup, k_rng_sequence   simtrace mode and query kernels.
#pragma once
#include "phox_types.h"
#include "phox_math.cuh"
#include "phox_philox.cuh"
#include "phox_csg.cuh"
#ifndef PHOX_SIM_MIN_BLOCKS
#define PHOX_SIM_MIN_BLOCKS 8      // 64 registers, 32 warps/SM: the loop is latency bound, occupancy beats spill-free code (profiles/)
#endif
#ifndef PHOX_HOT_LEAF
#define PHOX_HOT_LEAF 0            // 0: one out-of-line copy of the CSG leaf code serves every site (instruction-fetch bound kernel)
#endif
#ifndef PHOX_WF_STREAM
#define PHOX_WF_STREAM 1           // wavefront kernels read/write the per-photon streams with evict-first hints (ld/st.global.cs)
#endif
#ifndef PHOX_EXACT_BOX
#define PHOX_EXACT_BOX 1           // exit-distance bound for prims that are exactly their box (see traverse_bvh)
#endif
#ifndef PHOX_PROP_SSA
#define PHOX_PROP_SSA 1            // the physics body works on private copies of photon / stream / hit (see propagate_body)
#endif
#ifndef PHOX_PROP_INLINE_ALL
#define PHOX_PROP_INLINE_ALL 1     // propagate() is compiled into its (single) call site of every kernel; 0: one out-of-line body
#endif
#ifndef PHOX_WF_PROP_INLINE
#define PHOX_WF_PROP_INLINE 1      // same for k_wf_propagate alone (what PHOX_PROP_INLINE_ALL = 0 builds compare against)
#endif
#ifndef PHOX_LEAF_DIRECT
#define PHOX_LEAF_DIRECT 0          // single-leaf prims: 1 = call the leaf dispatcher directly, 2 = compile it in place (measured: mixed, profiles/r1_summary.md)
#endif
#ifndef PHOX_HITFIN_INLINE
#define PHOX_HITFIN_INLINE 0
#endif
#ifndef PHOX_TRAV_SPLIT
#define PHOX_TRAV_SPLIT 1          // exact-box leaves inline, node-loop state parked by hand around the out-of-line prim test
#endif
#include "phox_bvh.cuh"
#include "phox_physics.cuh"

namespace phox {

struct Scene {
    Geo geo;
    const float4* prim;             // Prim[nprim] viewed as 4 x float4
    const float4* exact;            // 2 x float4 per CSGPrim that is exactly a box (sizes ; translation), see intersect_exact_box
    const BvhNode* nodes;           // pool: instance tree first, then one tree per solid
    const InstanceRec* inst;
    int ninst;
    int tlas_root;                  // node index of the instance tree
    int accel;                      // PHOX_ACCEL_*
    const float4* home;             // 2 x float4 per CSGPrim (HomeRec, see traverse_bvh); null = home cells off
    const float4* cand;             // candidate lists of the home cells, two float4 per candidate: half sizes | prim, translation | instance
    const unsigned* prim_pb;        // per CSGPrim: the prd word (global prim index & 0xffff) << 16 | (boundary of its root node & 0xffff)
};

struct SimParams {
    Scene scene;
    Tables tables;
    // event
    const Genstep* genstep;
    const unsigned long long* gs_prefix;    // [num_genstep+1] exclusive prefix of numphoton within this launch
    int num_genstep;
    const Photon* input_photon;
    unsigned long long input_base;          // absolute photon index of input_photon[0]
    unsigned long long photon_offset;       // absolute index of slot 0 of this launch
    unsigned num_photon;                    // slots in this launch
    int event_index;
    // outputs (null = not kept)
    Photon* photon;
    Seq* seq;
    Photon* record;
    Prd* prd;
    unsigned long long* tag;                // stag[N]: 4 u64 per slot (DebugHeavy), zeroed before the launch
    float* flat;                            // sflat[N]: 64 floats per slot
    unsigned* tagslot;                      // per slot: tagged draws so far (wavefront form: the recorder is parked between kernels)
    unsigned* lpos;                         // per slot, lite mode: packed local position of the photon's last intersect (sphotonlite::set_lpos)
    int max_record;
    unsigned* work_counter;                 // next unclaimed photon slot of this launch
    unsigned long long* counters;           // [0] = rays traced
    // config
    int max_bounce;
    float tmin, tmin0, tmax, max_time, refine_distance;
    unsigned eps0_mask, hit_mask, refine;
    unsigned long long seed, rng_offset, skipahead;
    int burn;
    // persistent kernel as the FINISHER of a wavefront event (PHOX_KERNEL_AUTO, phox_engine.cu): instead of generating photons it takes
    // the photons of a live list over - record in `photon`, draw count in its index word - and runs each to the end of its history
    const unsigned* resume_list;            // null: generate
    const unsigned* resume_count;           // device-side length of resume_list
    int resume_bounce;                      // bounces every photon of the list has done
};

constexpr unsigned kHitFphi = 1u;          // fill lposfphi (only the prd debug array and simtrace read it); implies kHitCost
constexpr unsigned kHitCost = 4u;          // fill lposcost: the physics reads it only behind a surface row whose ems is neither NoSurface nor Surface (sensor_A ...), see hit_flags_of
// which optional parts of the hit record this launch needs: lposfphi for the prd debug array and the lite hits, lposcost also when
// the optical table has a row whose ems sends the physics to the lposcost test (qsim.h:2296-2312); a square root and a division per ray otherwise saved
PHOX_D unsigned hit_flags_of(const SimParams& P, bool debug) {
    return (((debug && P.prd != nullptr) || P.lpos != nullptr) ? kHitFphi : 0u) | (P.tables.need_lposcost ? kHitCost : 0u);
}

struct Nearest {
    float t;
    float3 n;                       // object-frame normal
    int prim;                       // global CSGPrim index
    int inst;
};

PHOX_D void keep_nearest(Nearest& best, const float4& is, int prim_idx, int inst_idx, float tmin) {
    float t = is.w;
    if (!(t > tmin)) return;                 // OptiX rejects reports outside (tmin, tmax]
    // ties go to the lower (instance, prim) pair so the answer does not depend on traversal order
    bool closer = t < best.t ||
                  (t == best.t && (best.prim < 0 || inst_idx < best.inst || (inst_idx == best.inst && prim_idx < best.prim)));
    if (closer) {
        best.t = t;
        best.n = f3(is.x, is.y, is.z);
        best.prim = prim_idx;
        best.inst = inst_idx;
    }
}

// A CSGPrim that is one un-complemented box3 leaf without rotation (most world, mother and crystal volumes): the same
// arithmetic as intersect_leaf's transform + leaf_box3 + normal back-transform, which for a pure translation reduce to
// o + t, d, n exactly (products with the 0 / 1 matrix entries are exact), without the dependent loads of prim -> node ->
// transform.  Out of line so that every kernel shares one compiled body, like intersect_prim_cold.
__device__ __noinline__ bool intersect_exact_box(float4& is, const float4* rec, float tmin, const float3& ro, const float3& rd, const float3& idir) {
    float4 q0 = __ldg(rec), tr = __ldg(rec + 1);
    float3 o = f3(ro.x + tr.x, ro.y + tr.y, ro.z + tr.z);
    return leaf_box3_idir(is, q0, tmin, o, rd, idir);       // idir = 1 / rd as the traversal computed it: the direction is not transformed
}
constexpr int kLeafExactBox = 0x40000000;        // leaf item flag: the prim qualifies for intersect_exact_box
constexpr int kLeafSingle = 0x20000000;          // leaf item flag: the prim is one leaf node (no tree, no list): the leaf dispatcher is called directly
constexpr int kLeafItemMask = 0x1fffffff;

constexpr int kBvhStack = 64;
constexpr int kTravReturn = (int)0x80000000;     // stack marker: leave the current solid, back to the instance tree
constexpr int kTravDone = (int)0x80000001;

// Traversal of both BVH levels in "while-while" form (Aila & Laine): every lane first walks internal
// nodes until it holds a leaf, then the lanes that hold one run the CSG prim test together - so the
// expensive, divergent part (intersect_prim) executes with as many lanes as possible.  `cur` is a node
// index (>= 0, relative to the current tree root), a leaf (~item: an instance in the top tree, a
// CSGPrim inside a solid) or one of the two markers.  Children are visited near-first; the far one is
// parked on the stack with its entry distance so it is dropped once a nearer hit is known.
PHOX_D void traverse_bvh(Nearest& best, const Scene& sc, float tmin, const float3& o_w, const float3& d_w) {
#if PHOX_TRAV_SPLIT
    volatile float park[10];
#endif
    int2 stack[kBvhStack];                                   // (item, entry distance bits): one 8 B local store / load per push / pop
    int sp = 0;
    auto push = [&](int item, float t) {
        if (sp < kBvhStack) { stack[sp] = make_int2(item, __float_as_int(t)); sp++; }
    };
    auto pop = [&]() -> int {
        while (sp > 0) {
            sp--;
            int2 e = stack[sp];
            if (__int_as_float(e.y) <= best.t) return e.x;
        }
        return kTravDone;
    };
    float3 o = o_w, d = d_w;
    float3 idir = f3(1.f / d.x, 1.f / d.y, 1.f / d.z);
    int root = sc.tlas_root;
    const BvhNode* tree = sc.nodes + root;                   // root of the tree being walked
    bool in_solid = false;
    int inst_idx = 0;
    int cur = sc.ninst == 1 ? ~0 : 0;
    while (true) {
        while (cur >= 0) {                                     // internal nodes
            const float4* np = reinterpret_cast<const float4*>(tree + cur);
            float4 a = __ldg(np), b = __ldg(np + 1), c = __ldg(np + 2);
            int4 ch = __ldg(reinterpret_cast<const int4*>(np + 3));
            float t0, t1 = 0.f, e0, e1 = 0.f;
            bool h0 = box_hit(a.x, a.y, a.z, a.w, b.x, b.y, o, idir, tmin, best.t, t0, e0);
            bool h1 = ch.y != kBvhNoChild && box_hit(b.z, b.w, c.x, c.y, c.z, c.w, o, idir, tmin, best.t, t1, e1);
#if PHOX_EXACT_BOX
            // c0/c1: no hit of the child can be nearer than this.  Normally the entry distance; for a prim that
            // is exactly its box and holds the ray origin, (a lower bound of) the exit distance - which lets the
            // enclosing volumes (the world box of every Geant4 geometry first of all) go untested once the
            // photon's own volume has answered.
            float c0 = t0, c1 = t1;
            if (ch.z | ch.w) {
                if (ch.z != 0 && h0) { c0 = box_exit_bound(a.x, a.y, a.z, a.w, b.x, b.y, o, idir, __int_as_float(ch.z), t0); h0 = c0 <= best.t; }
                if (ch.w != 0 && h1) { c1 = box_exit_bound(b.z, b.w, c.x, c.y, c.z, c.w, o, idir, __int_as_float(ch.w), t1); h1 = c1 <= best.t; }
            }
#else
            const float c0 = t0, c1 = t1;
#endif
            if (h0 && h1) {
                int nearc = ch.x, farc = ch.y; float tfar = c1;
                // nearer entry first; when the ray starts inside both boxes (equal entries) the box it
                // leaves sooner is the likelier home of the nearest surface
                if (t1 < t0 || (t1 == t0 && e1 < e0)) { nearc = ch.y; farc = ch.x; tfar = c0; }
                push(farc, tfar);
                cur = nearc;
            } else if (h0) cur = ch.x;
            else if (h1) cur = ch.y;
            else cur = pop();
        }
        if (cur == kTravDone) break;
        if (cur == kTravReturn) {                              // back to the instance tree
            if (in_solid && !(o.x == o_w.x && o.y == o_w.y && o.z == o_w.z && d.x == d_w.x && d.y == d_w.y && d.z == d_w.z)) {
                o = o_w; d = d_w;
                idir = f3(1.f / d.x, 1.f / d.y, 1.f / d.z);
            }
            in_solid = false;
            root = sc.tlas_root;
            tree = sc.nodes + root;
            cur = pop();
            continue;
        }
        if (!in_solid) {                                       // leaf of the instance tree: enter the solid
            inst_idx = ~cur;
            const InstanceRec* ir = sc.inst + inst_idx;
            int4 meta = __ldg(reinterpret_cast<const int4*>(&ir->solid));   // solid, identity, is_identity, bvh_root
            if (!meta.z) {
                float4 r0 = __ldg(&ir->inv[0]), r1 = __ldg(&ir->inv[1]), r2 = __ldg(&ir->inv[2]), r3 = __ldg(&ir->inv[3]);
                o = xform(r0, r1, r2, r3, o_w, 1.f);
                d = xform(r0, r1, r2, r3, d_w, 0.f);
                idir = f3(1.f / d.x, 1.f / d.y, 1.f / d.z);
            }
            root = meta.w;
            tree = sc.nodes + root;
            in_solid = true;
            if (sc.ninst > 1) push(kTravReturn, -CUDART_INF_F);
            cur = 0;
            continue;
        }
        {                                                      // a CSGPrim: the one intersect site
            int item = ~cur;
            int prim_idx = item & kLeafItemMask;
            float4 is = make_float4(0.f, 0.f, 0.f, 0.f);
            bool ok;
#if PHOX_EXACT_BOX && PHOX_TRAV_SPLIT
            if (item & kLeafExactBox) {                        // inlined: no call, nothing to preserve
                const float4* rec = sc.exact + 2 * prim_idx;
                float4 q0 = __ldg(rec), tr = __ldg(rec + 1);
                ok = leaf_box3_idir(is, q0, tmin, f3(o.x + tr.x, o.y + tr.y, o.z + tr.z), d, idir);
            } else
#elif PHOX_EXACT_BOX
            if (item & kLeafExactBox) {
                ok = intersect_exact_box(is, sc.exact + 2 * prim_idx, tmin, o, d, idir);
            } else
#endif
            {
                float4 p0 = __ldg(sc.prim + 4 * prim_idx);
                const float4* nroot = sc.geo.node + 4 * __float_as_int(p0.y);
#if PHOX_TRAV_SPLIT && PHOX_LEAF_DIRECT == 2
                if (item & kLeafSingle) {                      // leaf dispatcher compiled in place, on value copies
                    const float3 o_c = o, d_c = d;
                    float4 is_c;
                    ok = intersect_leaf(is_c, nroot, sc.geo, tmin, o_c, d_c);
                    is = is_c;
                } else {
#endif
#if PHOX_TRAV_SPLIT
                // The values the node loop reads every visit are parked by hand around the one out-of-line call and come
                // back as NEW values: their live ranges end at the call, so the register allocator has no reason to keep
                // them in spill slots - which it otherwise reloads at the head of every node visit (9 local loads per
                // visit at 64 registers, profiles/r1_summary.md).
                park[0] = o.x; park[1] = o.y; park[2] = o.z; park[3] = idir.x; park[4] = idir.y; park[5] = idir.z;
                park[6] = tmin; park[7] = best.t; park[8] = __int_as_float(root); park[9] = __int_as_float(sp);
                float4 is_c = make_float4(0.f, 0.f, 0.f, 0.f);
                const float3 o_c = o, d_c = d;
#if PHOX_LEAF_DIRECT
                if (item & kLeafSingle) ok = intersect_leaf_cold(is_c, nroot, sc.geo, tmin, o_c, d_c);    // what intersect_prim_cold would call
                else
#endif
                ok = intersect_prim_cold(is_c, nroot, sc.geo, tmin, o_c, d_c);
                is = is_c;
                o = f3(park[0], park[1], park[2]); idir = f3(park[3], park[4], park[5]);
                tmin = park[6]; best.t = park[7]; root = __float_as_int(park[8]); sp = __float_as_int(park[9]);
                tree = sc.nodes + root;
#elif PHOX_HOT_LEAF
                ok = intersect_prim(is, nroot, sc.geo, tmin, o, d);
#else
                ok = intersect_prim_cold(is, nroot, sc.geo, tmin, o, d);
#endif
#if PHOX_TRAV_SPLIT && PHOX_LEAF_DIRECT == 2
                }
#endif
            }
            if (ok) keep_nearest(best, is, prim_idx, inst_idx, tmin);
        }
        cur = pop();
    }
}

//
// Home cells.  A photon in a detector spends most of its bounces inside one small volume (a crystal, a fibre, a
// light guide), and from inside a box the nearest surface can only belong to a prim whose box reaches into it.
// `home` names a CSGPrim (of an identity-transform, single-instance solid: in practice the flattened remainder
// solid 0) whose HomeRec holds its box HB, grown by a pad, and the list of EVERY prim whose padded box overlaps
// HB grown once more (built in phox_set_geometry; prims with more than kHomeMaxCand neighbours, with a neighbour that
// is not an exact box, or whose box is touched by a transformed instance, have no list).  When the ray origin lies in HB, the candidates are tested
// first (tail of k_wf_propagate<.., HOME>).  If the nearest answer ends before the ray leaves HB
// (best.t * 1.000001 < exit distance of HB, the slab expressions of box_hit) the search is over: for any other
// prim Q some axis k has Q.lo_k > HB.hi_k + pad (or the mirror image), float subtraction and multiplication by
// idir_k are monotonic, so Q's slab entry is >= HB's slab exit > best.t * 1.0000004 and box_hit(Q) - the test
// the BVH culls with - is false as well.  Otherwise the ray goes to the full traversal (k_wf_trace over the pending list).  The home
// only culls; results are those of the brute-force loop, bit for bit (tests: BVH == brute, home on == off).
// On the way out the home becomes the prim that was hit if that prim has a list (the photon now sits on its
// surface, i.e. inside its padded box), else it is kept.
constexpr unsigned kNoHome = 0xffffffffu;
constexpr int kHomeMaxCand = 16;

// the home a photon takes along from a hit: the prim it hit if that prim has a candidate list, else the one it had
PHOX_D void home_update(unsigned& home, const Scene& sc, int hit_prim) {
    if (sc.home != nullptr && hit_prim >= 0 && (unsigned)hit_prim != home) {
        if (__float_as_int(__ldg(sc.home + 2 * hit_prim + 1).z) > 0) home = (unsigned)hit_prim;
    }
}

// Candidates of `home` against the ray; true when they settle it (best is then the answer of the whole geometry).
// Every candidate is an exact box (phox_set_geometry gives no list to a prim with any other neighbour), so the loop is
// call-free and runs in registers.  Only distances are worked out in the loop; the normal is computed once, for the
// winner (the same expressions as leaf_box3_idir, which is these two halves back to back).
PHOX_D bool home_search(Nearest& best, const Scene& sc, float tmin, const float3& o, const float3& d, unsigned home) {
    if (home == kNoHome) return false;
    const float4* hr = sc.home + 2 * home;
    const float4 ha = __ldg(hr), hb = __ldg(hr + 1);           // lo.xyz hi.x | hi.y hi.z count offset
    const int n = __float_as_int(hb.z);
    if (!(n > 0 && o.x >= ha.x && o.x <= ha.w && o.y >= ha.y && o.y <= hb.x && o.z >= ha.z && o.z <= hb.y)) return false;
    const float3 idir = f3(1.f / d.x, 1.f / d.y, 1.f / d.z);
    const float tx0 = (ha.x - o.x) * idir.x, tx1 = (ha.w - o.x) * idir.x;
    const float ty0 = (ha.y - o.y) * idir.y, ty1 = (hb.x - o.y) * idir.y;
    const float tz0 = (ha.z - o.z) * idir.z, tz1 = (hb.y - o.z) * idir.z;
    const float t_home = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fmaxf(tz0, tz1));
    const float4* cand = sc.cand + __float_as_int(hb.w);      // records of this home's candidates, in ascending (instance, prim) order
    float bt = best.t;
    int bi = -1;
    for (int i = 0; i < n; i++) {
        const float4 hs = __ldg(cand + 2 * i), tr = __ldg(cand + 2 * i + 1);
        float t;
        if (box3_t_half(t, f3(hs.x, hs.y, hs.z), tmin, f3(o.x + tr.x, o.y + tr.y, o.z + tr.z), d, idir)) {          // t > tmin
            // keep_nearest, whose ties go to the lower (instance, prim) pair: in list order that is the earlier candidate
            if (t < bt || (t == bt && bi < 0)) { bt = t; bi = i; }
        }
    }
    float3 nb = f3(0.f, 0.f, 0.f);
    int bprim = -1, binst = 0;
    if (bi >= 0) {
        const float4 hs = __ldg(cand + 2 * bi), tr = __ldg(cand + 2 * bi + 1);
        bprim = __float_as_int(hs.w); binst = __float_as_int(tr.w);
        nb = box3_normal_half(f3(hs.x, hs.y, hs.z), f3(o.x + tr.x, o.y + tr.y, o.z + tr.z), d, bt);
    }
    best.t = bt; best.n = nb; best.prim = bprim; best.inst = binst;
    return bt * 1.000001f < t_home;
}

// validation path: every prim of every instance, no boxes involved
__device__ __noinline__ void traverse_brute(Nearest& best, const Scene& sc, float tmin, const float3& o_w, const float3& d_w) {
    for (int i = 0; i < sc.ninst; i++) {
        const InstanceRec* ir = sc.inst + i;
        int4 meta = __ldg(reinterpret_cast<const int4*>(&ir->solid));
        float3 o = o_w, d = d_w;
        if (!meta.z) {
            float4 r0 = __ldg(&ir->inv[0]), r1 = __ldg(&ir->inv[1]), r2 = __ldg(&ir->inv[2]), r3 = __ldg(&ir->inv[3]);
            o = xform(r0, r1, r2, r3, o_w, 1.f);
            d = xform(r0, r1, r2, r3, d_w, 0.f);
        }
        int2 pr = __ldg(reinterpret_cast<const int2*>(&ir->prim_offset));
        for (int k = 0; k < pr.y; k++) {
            int prim_idx = pr.x + k;
            float4 p0 = __ldg(sc.prim + 4 * prim_idx);
            const float4* nroot = sc.geo.node + 4 * __float_as_int(p0.y);
            float4 is = make_float4(0.f, 0.f, 0.f, 0.f);
            if (intersect_prim_cold(is, nroot, sc.geo, tmin, o, d)) keep_nearest(best, is, prim_idx, i, tmin);
        }
    }
}

// Turns the winner of a traversal into the prd-equivalent (the closest-hit program's job, CSGOptiX7.cu:749-847, and
// the IS program's local-position terms, :919-934).  Out of line: every float that reaches the physics is computed by
// ONE compiled body (this function and intersect_prim_cold), whatever kernel ran the traversal around it - that is
// what keeps the persistent and the wavefront form, and the debug and production kernels, bit-identical.
constexpr unsigned kHitRawNormal = 2u;     // leave the normal as the closest-hit program delivers it (simtrace); the simulate raygen normalises
PHOX_D bool hit_finish_core(HitInfo& h, const Scene& sc, const Nearest& best, const float3& o, const float3& d, unsigned flags) {
    if (best.prim < 0) {
        h.normal = f3(0.f, 0.f, 0.f); h.t = 1.f; h.lposcost = 0.f; h.lposfphi = 0.f;
        h.iindex_identity = 0xffffffffu; h.prim_boundary = 0xffffffffu;
        return false;
    }
    const InstanceRec* ir = sc.inst + best.inst;
    int4 meta = __ldg(reinterpret_cast<const int4*>(&ir->solid));
    float3 oo = o, dd = d, n = best.n;
    if (!meta.z) {
        float4 r0 = __ldg(&ir->inv[0]), r1 = __ldg(&ir->inv[1]), r2 = __ldg(&ir->inv[2]), r3 = __ldg(&ir->inv[3]);
        oo = xform(r0, r1, r2, r3, o, 1.f);
        dd = xform(r0, r1, r2, r3, d, 0.f);
        n = xform_normal(r0, r1, r2, best.n);       // object -> world uses the inverse-transpose
    }
    float3 lpos = oo + best.t * dd;
    // the simulate raygen normalises every normal (CSGOptiX7.cu:470-471): n * (1 / sqrt(n.n)).  A box face normal has n.n == 1
    // exactly, for which that expression returns n bit for bit (1 / sqrt(1) = 1, x * 1 = x): skip the square root and the division
    const float nn = dot(n, n);
    h.normal = ((flags & kHitRawNormal) || nn == 1.f) ? n : n * (1.0f / sqrtf(nn));
    h.t = best.t;
    h.lposcost = (flags & (kHitFphi | kHitCost)) ? lpos.z / sqrtf(dot(lpos, lpos)) : 0.f;
    h.lposfphi = (flags & kHitFphi) ? (atan2f(lpos.y, lpos.x) + kPi) / (2.0f * kPi) : 0.f;
    h.iindex_identity = (((unsigned)best.inst & 0xffffu) << 16) | ((unsigned)meta.y & 0xffffu);
    // (global prim index << 16 | boundary of the prim's root node, put together per prim by phox_set_geometry: one load instead of prim -> node -> prim)
    h.prim_boundary = __ldg(sc.prim_pb + best.prim);
    return true;
}

// value copies in, value copy out: the same expression graph wherever it is compiled (see propagate_body)
PHOX_D bool hit_finish_body(HitInfo& h_out, const Scene& sc, const Nearest& best_in, const float3& o_in, const float3& d_in, unsigned flags) {
    const Nearest best = best_in;
    const float3 o = o_in, d = d_in;
    HitInfo h;
    const bool ok = hit_finish_core(h, sc, best, o, d, flags);
    h_out = h;
    return ok;
}
__device__ __noinline__ bool hit_finish(HitInfo& h, const Scene& sc, const Nearest& best, const float3& o, const float3& d, unsigned flags) {
    return hit_finish_body(h, sc, best, o, d, flags);
}

// nearest intersect in (tmin, tmax]; fills the prd-equivalent.  Returns false on a miss (the reference's miss
// program sets boundary 0xffff).  Inline form: the traversal is compiled into the calling kernel, where the scene
// pointers are kernel parameters (constant bank) instead of loads through a reference.  The boxes only cull; hit
// distances and normals come from the out-of-line prim evaluators and hit_finish.
template <bool HITFIN_INLINE = (PHOX_HITFIN_INLINE != 0)>
PHOX_D bool trace_inline(HitInfo& h, const Scene& sc, const float3& o, const float3& d, float tmin, float tmax, unsigned flags, unsigned& home) {
    Nearest best;
    best.t = tmax; best.prim = -1; best.inst = 0; best.n = f3(0.f, 0.f, 0.f);
    if (sc.accel == 0) traverse_bvh(best, sc, tmin, o, d);
    else traverse_brute(best, sc, tmin, o, d);
    home_update(home, sc, best.prim);
    // hit_finish compiled in place or called: a value-copy body either way, the same bits.  In place the BVH kernel is 3 - 5 % quicker on
    // geometries of boxes and single-leaf prims and 5 % slower on the boolean-tree ones (zoo, pfRICH: the tree evaluator's register
    // pressure); the engine takes the in-place instance for geometries without a boolean tree beyond one operator (profiles/r2_summary.md)
    if (HITFIN_INLINE) return hit_finish_body(h, sc, best, o, d, flags);
    return hit_finish(h, sc, best, o, d, flags);
}

// out-of-line form for kernels with several trace sites (persistent kernel, geometry queries)
__device__ __noinline__ bool trace(HitInfo& h, const Scene& sc, const float3& o, const float3& d, float tmin, float tmax, unsigned flags) {
    unsigned home = kNoHome;
    return trace_inline(h, sc, o, d, tmin, tmax, flags, home);
}

PHOX_D void seq_add(Seq& s, unsigned slot, unsigned flag, unsigned boundary) {      // sseq::add_nibble
    unsigned iseq = slot / 16u;
    unsigned shift = 4u * (slot - iseq * 16u);
    if (iseq < 2u) {
        s.seqhis[iseq] |= ((unsigned long long)(__ffs(flag) & 0xf)) << shift;
        s.seqbnd[iseq] |= ((unsigned long long)(boundary & 0xfu)) << shift;
    }
}

// sphotonlite::set_lpos (sysrap/sphotonlite.h:234-245): two u16 fractions; the float -> u16 conversion saturates like
// the cvt.rzi.u16.f32 nvcc emits for the reference's (uint16_t) cast
PHOX_D unsigned pack_lpos(float lposcost, float lposfphi) {
    unsigned a = __float2uint_rz(fminf(fmaxf(lposcost * 65535.f + 0.5f, 0.f), 65535.f));
    unsigned b = __float2uint_rz(fminf(fmaxf(lposfphi * 65535.f + 0.5f, 0.f), 65535.f));
    return (a << 16) | b;
}

// Persistent bounce-loop kernel.  The grid is sized to the machine (SMs x resident blocks), not to
// the event: each warp pulls photon slots from a global counter and REFILLS lanes whose photon has
// finished, so lanes do not idle while the longest history of the warp runs out (bounce counts are
// long-tailed: 1 .. max_bounce).  A photon's result depends only on its absolute index (RNG
// subsequence), never on which lane ran it, so this reordering cannot change any output.
constexpr unsigned kListEps0 = 0x80000000u;     // list entry bit: the photon's last flag is in PropagateEpsilon0Mask (-> tmin0), so that
constexpr unsigned kListSlotMask = 0x7fffffffu; // the trace kernel does not have to read the flag word of the photon record
constexpr unsigned kWaveNoHit = 0xffffffffu;    // prim_boundary of a list entry whose photon is final (miss or time over)
constexpr int kSimThreads = 128;
constexpr int kRefillMin = 8;        // refill when at least this many lanes of the warp are idle

template <bool DEBUG>
__global__ void __launch_bounds__(kSimThreads, PHOX_SIM_MIN_BLOCKS) k_simulate(const __grid_constant__ SimParams P) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    bool active = false, exhausted = false;
    unsigned idx = 0, nray = 0;
    int bounce = 0;
    PhotonState p;
    Philox rng;
    Seq seq;
    unsigned tag_slot = 0u;              // DebugHeavy: tagged draws of this photon so far
    unsigned last_lpos = 0u;             // lite mode: packed lposcost/lposfphi of the last trace (0 after a miss, like the miss program)
    const unsigned hit_flags = hit_flags_of(P, DEBUG);
    const unsigned work_n = P.resume_list ? *P.resume_count : P.num_photon;

    while (true) {
        unsigned need = __ballot_sync(0xffffffffu, !active);
        if (!exhausted && (__popc(need) >= kRefillMin)) {
            unsigned base = 0;
            int leader = __ffs(need) - 1;
            if ((int)lane == leader) base = atomicAdd(P.work_counter, (unsigned)__popc(need));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + (unsigned)__popc(need) >= work_n) exhausted = true;
            unsigned mine = base + (unsigned)__popc(need & lt_mask);
            if (!DEBUG && P.resume_list != nullptr) {
                if (!active && mine < work_n) {              // take a live photon of the wavefront loop over (production modes only)
                    idx = P.resume_list[mine] & kListSlotMask;
                    p.load(P.photon + idx);
                    const unsigned nd = p.index;             // draw count parked in the index word
                    p.index = (unsigned)(P.photon_offset + idx);
                    rng.init(P.seed, P.photon_offset + idx, P.rng_offset + P.skipahead * (unsigned long long)P.event_index + nd);
                    bounce = P.resume_bounce;
                    active = true;
                    last_lpos = 0u;
                    tag_slot = 0u;
                }
            } else
            if (!active && mine < P.num_photon) {
                idx = mine;
                // seed : which genstep owns slot idx (binary search in the numphoton prefix sum)
                int lo = 0, hi = P.num_genstep;
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (__ldg(P.gs_prefix + mid) <= (unsigned long long)idx) lo = mid; else hi = mid;
                }
                const Genstep& gs = P.genstep[lo];          // read where it lies (see k_wf_generate)
                unsigned long long photon_idx = P.photon_offset + idx;
                rng.init(P.seed, photon_idx, P.rng_offset + P.skipahead * (unsigned long long)P.event_index);
                generate_photon(p, rng, gs, P.tables, P.input_photon, P.input_base, photon_idx);
                bounce = 0;
                active = true;
                last_lpos = 0u;
                tag_slot = 0u;
                if (DEBUG) {
                    seq.seqhis[0] = seq.seqhis[1] = seq.seqbnd[0] = seq.seqbnd[1] = 0ull;
                    if (P.record && 0 < P.max_record) p.store(P.record + (size_t)P.max_record * idx);
                    if (P.seq) seq_add(seq, 0u, p.flag(), p.boundary());
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) {
            if (exhausted) break;
            continue;
        }
        if (active) {
            bool finished = !(bounce < P.max_bounce && p.time < P.max_time);
            if (!finished) {
                float tmin = (p.obf & P.eps0_mask) ? P.tmin0 : P.tmin;
                HitInfo h;
                bool ok = trace(h, P.scene, p.pos, p.mom, tmin, P.tmax, hit_flags);
                nray++;
                if (P.refine) {                                 // trace<true>: decided by 0.99 x prd distance, which is 1 after a miss (CSGOptiX7.cu:165-184)
                    float t_approx = 0.99f * h.t;
                    if (t_approx > P.refine_distance) {
                        float3 closer = p.pos + t_approx * p.mom;
                        ok = trace(h, P.scene, closer, p.mom, tmin, P.tmax, hit_flags);
                        nray++;
                        h.t += t_approx;
                    }
                }
                if (P.lpos) last_lpos = ok ? pack_lpos(h.lposcost, h.lposfphi) : 0u;
                if (!ok) finished = true;                       // photon left the world
                else {
                    // (the normal was normalised at the end of trace(), CSGOptiX7.cu:470-471)
                    if (DEBUG) {
                        if (P.prd && bounce < P.max_record) {
                            Prd r;
                            r.nx = h.normal.x; r.ny = h.normal.y; r.nz = h.normal.z; r.t = h.t;
                            r.lposcost = h.lposcost; r.lposfphi = h.lposfphi;
                            r.iindex_identity = h.iindex_identity; r.prim_boundary = h.prim_boundary;
                            P.prd[(size_t)P.max_record * idx + bounce] = r;
                        }
                    }
                    int command;
                    if (DEBUG && P.tag) {
                        Tagr tg;
                        tg.tag = P.tag + 4 * (size_t)idx; tg.flat = P.flat + 64 * (size_t)idx; tg.slot = tag_slot;
                        command = propagate_t<true>(p, rng, h, P.tables, P.burn != 0, &tg);
                        tag_slot = tg.slot;
                    } else {
                        command = propagate(p, rng, h, P.tables, P.burn != 0);
                    }
                    bounce++;
                    if (DEBUG) {
                        if (P.record && bounce < P.max_record) p.store(P.record + (size_t)P.max_record * idx + bounce);
                        if (P.seq) seq_add(seq, (unsigned)bounce, p.flag(), p.boundary());
                    }
                    if (command == FLOW_BREAK || !(bounce < P.max_bounce && p.time < P.max_time)) finished = true;
                }
            }
            if (finished) {
                if (DEBUG) { if (P.seq) P.seq[idx] = seq; }
                if (P.photon) p.store(P.photon + idx);
                if (P.lpos) P.lpos[idx] = last_lpos;
                active = false;
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) nray += __shfl_down_sync(0xffffffffu, nray, off);
    if (lane == 0 && nray) atomicAdd(P.counters, (unsigned long long)nray);
}

// ---- wavefront form of the same loop ----------------------------------------------------------------
// One bounce = two small kernels over the list of photons still alive: k_wf_trace (ray -> hit record)
// and k_wf_propagate (hit record -> physics, survivors appended to the next list).  Photon state lives in
// the output photon array itself (64 B per slot, final as soon as the photon stops), the random stream
// as a 4-byte draw count (Philox is counter based), the hit as a 32 B quad2 indexed by list position.
// Compared with the persistent kernel each phase has a small instruction footprint and few live
// registers, every lane works on a live photon, and because survivors are appended block by block in
// list order, the photons of one genstep - which mostly sit in the same volume - stay together in a
// warp bounce after bounce (coherent traversal).  trace() and propagate() are the same out-of-line
// bodies the persistent kernel calls, so both forms give bit-identical results.
struct WaveParams {
    SimParams sim;
    unsigned* active_in;            // photon slots alive at this bounce (bit 31: kListEps0)
    unsigned* active_out;           // survivors (next bounce)
    const unsigned* count_in;       // device-side length of active_in
    unsigned* count_out;            // device-side length of active_out (zeroed beforehand)
    unsigned* ndraw;                // (unused since the draw count travels in the index word of the live photon records)
    unsigned* home;                 // per position of active_in: home cell of that photon (CSGPrim index or kNoHome); null = no home pass
    unsigned* home_next;            // ... of active_out (the home travels with the list entry: coalesced, no per-slot array)
    const unsigned* gs_home;        // per genstep: home cell its photons start with (k_genstep_home), or null
    uint2* pending;                 // (list position, list entry) of the rays their home cell did not settle (null: k_wf_trace takes the whole list); the entry
                                    // rides along so that the BVH kernel goes from the pending entry straight to the photon record, one scattered load less
    unsigned* pending_count;        // device-side length of pending (zeroed beforehand)
    Seq* seq_state;                 // per slot history being built (debug modes)
    Prd* hits;                      // per list position: hit of this bounce
    Prd* hits_next;                 // HOME: hit records of the next bounce, per position in active_out (the physics kernel reads `hits` while it fills these)
    int bounce;                     // bounces done so far by every photon of active_in
    unsigned* live_report;          // pinned host word that k_wf_trace posts the length of active_in to (null: not this bounce)
};

#ifndef PHOX_WF_TRACE_MIN_BLOCKS
#define PHOX_WF_TRACE_MIN_BLOCKS 4      // resident 256-thread blocks per SM the trace kernel is compiled for (register cap 65536/(256*N))
#endif
constexpr int kWaveThreads = 256;
#ifndef PHOX_WF_TRACE_THREADS
#define PHOX_WF_TRACE_THREADS 256       // block of the trace kernel; with PHOX_WF_TRACE_MIN_BLOCKS it sets the register budget
#endif
constexpr int kTraceThreads = PHOX_WF_TRACE_THREADS;
#ifndef PHOX_WF_PROP_THREADS
#define PHOX_WF_PROP_THREADS 256        // block of the physics kernel = run length of the ordered survivor append
#endif
constexpr int kPropThreads = PHOX_WF_PROP_THREADS;
#ifndef PHOX_APPEND_SORT
#define PHOX_APPEND_SORT 0              // experiment: survivors grouped by the boundary they meet next (measured and rejected, profiles/r2_summary.md)
#endif
#ifndef PHOX_PROP_PREFETCH
#define PHOX_PROP_PREFETCH 1            // physics kernel: L2 prefetch of the next chunk's lines (measured, see profiles/r2_summary.md)
#endif
#ifndef PHOX_PROP_STAGE
#define PHOX_PROP_STAGE 0               // physics kernel: the next chunk's hit record / photon / draw count / home come in by cp.async while this chunk computes
#endif
#ifndef PHOX_WF_PROP_MIN_BLOCKS
#define PHOX_WF_PROP_MIN_BLOCKS 4       // 64 registers: the inlined physics body fits without spills (5 blocks = 48 registers: 0.509 vs 0.490 ms per launch)
#endif

// hit record of list position a (streaming store: the physics kernel reads it once)
PHOX_D void wave_store_hit(Prd* hits, unsigned a, const Prd& r) {
#if PHOX_WF_STREAM
    stcs256(hits + a, make_float4(r.nx, r.ny, r.nz, r.t),
            make_float4(r.lposcost, r.lposfphi, __uint_as_float(r.iindex_identity), __uint_as_float(r.prim_boundary)));
#else
    hits[a] = r;
#endif
}
PHOX_D void wave_no_hit(Prd& r) {
    r.nx = r.ny = r.nz = 0.f; r.t = -1.f; r.lposcost = r.lposfphi = 0.f; r.iindex_identity = 0xffffffffu; r.prim_boundary = kWaveNoHit;
}
PHOX_D void wave_hit_record(Prd& r, const HitInfo& h) {
    r.nx = h.normal.x; r.ny = h.normal.y; r.nz = h.normal.z; r.t = h.t;
    r.lposcost = h.lposcost; r.lposfphi = h.lposfphi;
    r.iindex_identity = h.iindex_identity; r.prim_boundary = h.prim_boundary;
}

// Home cell a genstep's photons start with: the smallest home box that holds the genstep's position (mid-point of the
// step for Cerenkov / scintillation gensteps, whose photons are spread along it).  Only a hint - every ray checks for
// itself that its origin lies in the box of the home it carries - so a genstep near a wall, or a wide torch source, just
// sends some photons of the first bounce to the BVH.
__global__ void k_genstep_home(const Genstep* __restrict__ gs, int ngs, const float4* __restrict__ home, int nprim, unsigned* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ngs) return;
    const Genstep& g = gs[i];
    const int code = g.gencode();
    float3 p = f3(g.f[4], g.f[5], g.f[6]);
    if (code == GS_CERENKOV || code == GS_SCINTILLATION || code == GS_DsG4Scintillation_r4695)
        p = f3(g.f[4] + 0.5f * g.f[8], g.f[5] + 0.5f * g.f[9], g.f[6] + 0.5f * g.f[10]);
    unsigned best = kNoHome;
    float best_vol = CUDART_INF_F;
    if (code != GS_INPUT_PHOTON) {
        for (int k = 0; k < nprim; k++) {
            const float4 ha = __ldg(home + 2 * k), hb = __ldg(home + 2 * k + 1);
            if (__float_as_int(hb.z) <= 0) continue;
            if (p.x >= ha.x && p.x <= ha.w && p.y >= ha.y && p.y <= hb.x && p.z >= ha.z && p.z <= hb.y) {
                const float vol = (ha.w - ha.x) * (hb.x - ha.y) * (hb.y - ha.z);
                if (vol < best_vol) { best_vol = vol; best = (unsigned)k; }
            }
        }
    }
    out[i] = best;
}

// HOME: photons take their genstep's home cell along (k_genstep_home) and try its candidate list at once, exactly like
// the survivors of a physics pass do (k_wf_propagate): settled rays get the hit record of bounce 0 from here, the rest
// goes to the pending list of bounce 0.
template <bool DEBUG, bool HOME>
__global__ void __launch_bounds__(kWaveThreads) k_wf_generate(const __grid_constant__ WaveParams W) {
    __shared__ unsigned s_warp[kWaveThreads / 32];
    __shared__ unsigned s_pbase;
    const SimParams& P = W.sim;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned nhome = 0;
    for (unsigned base_idx = blockIdx.x * blockDim.x; base_idx < P.num_photon; base_idx += gridDim.x * blockDim.x) {
        const unsigned idx = base_idx + threadIdx.x;
        bool pend = false;
        unsigned entry0 = 0u;                  // this slot's entry of the first list
        if (idx < P.num_photon) {
            int lo = 0, hi = P.num_genstep;
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (__ldg(P.gs_prefix + mid) <= (unsigned long long)idx) lo = mid; else hi = mid;
            }
            const Genstep& gs = P.genstep[lo];          // read where it lies (the out-of-line generators take it by reference: a local copy would sit in the frame)
            unsigned long long photon_idx = P.photon_offset + idx;
            unsigned long long base = P.rng_offset + P.skipahead * (unsigned long long)P.event_index;
            Philox rng;
            rng.init(P.seed, photon_idx, base);
            PhotonState p;
            generate_photon(p, rng, gs, P.tables, P.input_photon, P.input_base, photon_idx);
            // A live photon's record carries its draw count in the `index` word (the index itself is photon_offset + slot, put
            // back by whichever physics pass writes the record for the last time): one scattered 4 B load and store less per bounce.
            {
                const unsigned true_index = p.index;
                if (P.max_bounce > 0) p.index = rng.consumed(base);
                entry0 = idx | ((p.obf & P.eps0_mask) ? kListEps0 : 0u);
#if PHOX_WF_STREAM
                p.store_cs(P.photon + idx);
                __stcs(W.active_out + idx, entry0);
#else
                p.store(P.photon + idx);
                W.active_out[idx] = entry0;
#endif
                p.index = true_index;
            }
            if (P.lpos) P.lpos[idx] = 0u;
            if (DEBUG) {
                Seq seq;
                seq.seqhis[0] = seq.seqhis[1] = seq.seqbnd[0] = seq.seqbnd[1] = 0ull;
                if (P.record && 0 < P.max_record) p.store(P.record + (size_t)P.max_record * idx);
                if (P.seq) { seq_add(seq, 0u, p.flag(), p.boundary()); P.seq[idx] = seq; }
            }
            if (W.home_next) {
                unsigned home = W.gs_home ? __ldg(W.gs_home + lo) : kNoHome;
                if (HOME) {
                    pend = true;
                    if (0 < P.max_bounce && p.time < P.max_time) {         // else the first trace kernel writes the no-hit record
                        const float tmin = (p.obf & P.eps0_mask) ? P.tmin0 : P.tmin;
                        const float3 o = p.pos, d = p.mom;
                        Nearest best;
                        best.t = P.tmax; best.prim = -1; best.inst = 0; best.n = f3(0.f, 0.f, 0.f);
                        if (home_search(best, P.scene, tmin, o, d, home)) {
                            home_update(home, P.scene, best.prim);
                            const Nearest best_c = best;
                            const float3 o_c = o, d_c = d;
                            HitInfo h_c;
                            hit_finish_body(h_c, P.scene, best_c, o_c, d_c, hit_flags_of(P, DEBUG));
                            Prd r;
                            wave_hit_record(r, h_c);
                            if (DEBUG) { if (P.prd && 0 < P.max_record) P.prd[(size_t)P.max_record * idx] = r; }
                            wave_store_hit(W.hits_next, idx, r);
                            pend = false;
                            nhome++;
                        }
                    }
                }
                __stcs(W.home_next + idx, home);
            }
        }
        if (HOME) {        // pending list of bounce 0, in slot order within the chunk
            const unsigned pballot = __ballot_sync(0xffffffffu, pend);
            if (lane == 0) s_warp[warp] = __popc(pballot);
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned ptot = 0;
                for (int w = 0; w < kWaveThreads / 32; w++) { const unsigned c = s_warp[w]; s_warp[w] = ptot; ptot += c; }
                s_pbase = ptot ? atomicAdd(W.pending_count, ptot) : 0u;
            }
            __syncthreads();
            if (pend) W.pending[s_pbase + s_warp[warp] + __popc(pballot & ((1u << lane) - 1u))] = make_uint2(idx, entry0);
            __syncthreads();
        }
    }
    if (HOME) {
        for (int off = 16; off > 0; off >>= 1) nhome += __shfl_down_sync(0xffffffffu, nhome, off);
        if (lane == 0 && nhome) { atomicAdd(P.counters, (unsigned long long)nhome); atomicAdd(P.counters + 2, (unsigned long long)nhome); }
    }
}

// One ray per live photon (W.pending == null) or per entry of the pending list the physics kernel of the previous bounce
// left behind (the rays their home cell could not settle).
template <bool DEBUG, bool HITFIN_INLINE = false>
__global__ void __launch_bounds__(kTraceThreads, PHOX_WF_TRACE_MIN_BLOCKS) k_wf_trace(const __grid_constant__ WaveParams W) {
    const SimParams& P = W.sim;
    const unsigned count = W.pending ? *W.pending_count : *W.count_in;
    if (W.live_report != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {         // lets the host stop launching once the list is empty (phox_engine.cu)
        *reinterpret_cast<volatile unsigned*>(W.live_report) = *W.count_in;
        __threadfence_system();
    }
    unsigned nray = 0;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        unsigned a = i, entry;
        if (W.pending) { const uint2 pe = __ldcs(W.pending + i); a = pe.x; entry = pe.y; }
        else entry = __ldcs(W.active_in + a);
        const unsigned idx = entry & kListSlotMask;
        const float4* ph = reinterpret_cast<const float4*>(P.photon + idx);
        float4 q0, q1;
        ldcs256(ph, q0, q1);
        Prd r;
        wave_no_hit(r);
        if (q0.w < P.max_time) {                                // else the while-condition of the raygen loop fails: photon is final
            const unsigned home_in = W.home ? __ldcs(W.home + a) : kNoHome;
            unsigned home = home_in;
            float tmin = (entry & kListEps0) ? P.tmin0 : P.tmin;   // the list entry carries (last flag & PropagateEpsilon0Mask) != 0
            float3 o = f3(q0.x, q0.y, q0.z), d = f3(q1.x, q1.y, q1.z);
            HitInfo h;
            bool ok;
            float3 from = o;
            float t_add = 0.f;
            for (int pass = 0;; pass++) {                       // one inlined trace site; pass 1 = PropagateRefine re-trace from 0.99 t
                ok = trace_inline<HITFIN_INLINE || (PHOX_HITFIN_INLINE != 0)>(h, P.scene, from, d, tmin, P.tmax, hit_flags_of(P, DEBUG), home);
                nray++;
                if (pass == 1) { h.t += t_add; break; }
                if (!P.refine) break;
                t_add = 0.99f * h.t;                            // h.t is 1 after a miss, like the miss program's prd (CSGOptiX7.cu:165-184, 667)
                if (!(t_add > P.refine_distance)) break;
                from = o + t_add * d;
            }
            if (W.home && home != home_in) W.home[a] = home;
            if (ok) {
                wave_hit_record(r, h);
                if (DEBUG) { if (P.prd && W.bounce < P.max_record) P.prd[(size_t)P.max_record * idx + W.bounce] = r; }
            }
        }
        wave_store_hit(W.hits, a, r);
    }
    for (int off = 16; off > 0; off >>= 1) nray += __shfl_down_sync(0xffffffffu, nray, off);
    if ((threadIdx.x & 31u) == 0 && nray) atomicAdd(P.counters, (unsigned long long)nray);
}

// HOME: the geometry has home cells (traverse/home_search above).  Each survivor then tries the candidate list of its home
// right here, with its new position and direction still in registers: a settled ray gets its hit record for the NEXT
// bounce written from this kernel (W.hits_next, indexed by its position in the next list), the others go to the pending
// list that k_wf_trace walks before the next physics pass.  The candidate pass costs a few box tests; as a kernel of its
// own it had to re-read list entry, photon and home (~90 B of DRAM traffic per ray behind a chain of dependent loads).
template <bool DEBUG, bool HOME>
__global__ void __launch_bounds__(kPropThreads, PHOX_WF_PROP_MIN_BLOCKS) k_wf_propagate(const __grid_constant__ WaveParams W) {
    __shared__ unsigned s_warp[2][kPropThreads / 32];      // double-buffered by chunk parity: two barriers per chunk instead of three
    __shared__ unsigned s_base[2], s_pbase[2];
#if PHOX_APPEND_SORT
    __shared__ unsigned s_cls[2][8][kPropThreads / 32];
#endif
    unsigned par = 0;
    unsigned nhome = 0;
    const SimParams& P = W.sim;
    const unsigned count = *W.count_in;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#if PHOX_PROP_STAGE
    // Staging: what a thread needs at the head of its NEXT chunk (hit record, photon, draw count, home: 104 B behind the
    // dependent load of the list entry) is copied into shared memory by cp.async while the current chunk computes, each
    // thread into its own column - so the head of a chunk costs a shared-memory read instead of two DRAM round trips, no
    // registers are held across the physics, and nobody but the copying thread reads a column (no barrier: wait_group).
    __shared__ float4 s_ph[4][kPropThreads];
    __shared__ float4 s_hit[2][kPropThreads];
    __shared__ unsigned s_hm[kPropThreads];
    auto stage = [&](unsigned a_s, unsigned entry_s) {
        const unsigned idx_s = entry_s & kListSlotMask;
        const float4* hp = reinterpret_cast<const float4*>(W.hits + a_s);
        const float4* pp = reinterpret_cast<const float4*>(P.photon + idx_s);
#pragma unroll
        for (int k = 0; k < 2; k++)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&s_hit[k][threadIdx.x])), "l"(hp + k) : "memory");
#pragma unroll
        for (int k = 0; k < 4; k++)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&s_ph[k][threadIdx.x])), "l"(pp + k) : "memory");
        if (HOME) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(&s_hm[threadIdx.x])), "l"(W.home + a_s) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    unsigned entry_cur = 0xffffffffu;
    {
        const unsigned a0 = blockIdx.x * blockDim.x + threadIdx.x;
        if (a0 < count) { entry_cur = __ldcs(W.active_in + a0); stage(a0, entry_cur); }
    }
#endif
    for (unsigned base_a = blockIdx.x * blockDim.x; base_a < count; base_a += gridDim.x * blockDim.x) {
        unsigned a = base_a + threadIdx.x;
#if PHOX_PROP_PREFETCH || PHOX_PROP_STAGE
        // list entry of this thread's photon in the NEXT chunk: it arrives while this chunk's physics runs (then the data it
        // points to are staged; without staging: asked into L2 before the threads meet at the barrier)
        const unsigned a_next = a + gridDim.x * blockDim.x;
        unsigned entry_next = 0xffffffffu;
        if (a_next < count) entry_next = __ldcs(W.active_in + a_next);
#endif
        bool survive = false, settled = false, have = false;
        unsigned idx = 0, entry_out = 0, home = kNoHome;
        int bounce = W.bounce + 1;
        Prd r2;                                             // HOME: hit of the next bounce, when the home cell settles it
        PhotonState p;
        if (a < count) {
#if PHOX_PROP_STAGE
            idx = entry_cur & kListSlotMask;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            Prd r;
            {
                const float4 h0 = s_hit[0][threadIdx.x], h1 = s_hit[1][threadIdx.x];
                r.nx = h0.x; r.ny = h0.y; r.nz = h0.z; r.t = h0.w; r.lposcost = h1.x; r.lposfphi = h1.y;
                r.iindex_identity = __float_as_uint(h1.z); r.prim_boundary = __float_as_uint(h1.w);
            }
            if (HOME) home = s_hm[threadIdx.x];
#elif PHOX_WF_STREAM
            idx = __ldcs(W.active_in + a) & kListSlotMask;
            Prd r;
            {
                float4 h0, h1;
                ldcs256(W.hits + a, h0, h1);
                r.nx = h0.x; r.ny = h0.y; r.nz = h0.z; r.t = h0.w; r.lposcost = h1.x; r.lposfphi = h1.y;
                r.iindex_identity = __float_as_uint(h1.z); r.prim_boundary = __float_as_uint(h1.w);
            }
#else
            idx = W.active_in[a] & kListSlotMask;
            Prd r = W.hits[a];
#endif
#if !PHOX_PROP_STAGE
            if (HOME) home = __ldcs(W.home + a);            // coalesced, asked for with the hit record: back long before the candidate pass wants it
#endif
            if (r.prim_boundary != kWaveNoHit) {            // a miss (or time over) leaves the photon as it is: final
                have = true;
#if PHOX_PROP_STAGE
                {
                    const float4 qa = s_ph[0][threadIdx.x], qb = s_ph[1][threadIdx.x], qc = s_ph[2][threadIdx.x], qd = s_ph[3][threadIdx.x];
                    p.pos = f3(qa.x, qa.y, qa.z); p.time = qa.w;
                    p.mom = f3(qb.x, qb.y, qb.z); p.hitcount_iindex = __float_as_uint(qb.w);
                    p.pol = f3(qc.x, qc.y, qc.z); p.wavelength = qc.w;
                    p.obf = __float_as_uint(qd.x); p.identity = __float_as_uint(qd.y); p.index = __float_as_uint(qd.z); p.flagmask = __float_as_uint(qd.w);
                }
#elif PHOX_WF_STREAM
                p.load_cs(P.photon + idx);
#else
                p.load_rw(P.photon + idx);
#endif
                const unsigned nd = p.index;                 // draw count parked in the index word (k_wf_generate)
                p.index = (unsigned)(P.photon_offset + idx);
                unsigned long long base = P.rng_offset + P.skipahead * (unsigned long long)P.event_index;
                Philox rng;
                rng.init(P.seed, P.photon_offset + idx, base + nd);
                HitInfo h;
                h.normal = f3(r.nx, r.ny, r.nz); h.t = r.t; h.lposcost = r.lposcost; h.lposfphi = r.lposfphi;
                h.iindex_identity = r.iindex_identity; h.prim_boundary = r.prim_boundary;
                int command;
                if (DEBUG && P.tag) {
                    Tagr tg;
                    tg.tag = P.tag + 4 * (size_t)idx; tg.flat = P.flat + 64 * (size_t)idx; tg.slot = P.tagslot[idx];
                    command = propagate_t<true>(p, rng, h, P.tables, P.burn != 0, &tg);
                    P.tagslot[idx] = tg.slot;
                } else {
#if PHOX_WF_PROP_INLINE
                    command = propagate_body<false>(p, rng, h, P.tables, P.burn != 0, nullptr);
#else
                    command = propagate(p, rng, h, P.tables, P.burn != 0);
#endif
                }
                if (DEBUG) {
                    if (P.record && bounce < P.max_record) p.store(P.record + (size_t)P.max_record * idx + bounce);
                    if (P.seq) { Seq seq = P.seq[idx]; seq_add(seq, (unsigned)bounce, p.flag(), p.boundary()); P.seq[idx] = seq; }
                }
                survive = !(command == FLOW_BREAK) && bounce < P.max_bounce && p.time < P.max_time;
                entry_out = idx | ((p.obf & P.eps0_mask) ? kListEps0 : 0u);
                if (survive) p.index = rng.consumed(base);   // still alive: the index word parks the draw count again
#if PHOX_WF_STREAM
                p.store_cs(P.photon + idx);
#else
                p.store(P.photon + idx);
#endif
            } else {
                P.photon[idx].index = (unsigned)(P.photon_offset + idx);    // final as it is, but for the draw count parked in its index word
            }
        }
#if PHOX_PROP_STAGE
        // this chunk's columns are in registers: the next chunk's copies may start (they land during the candidate pass and the append)
        entry_cur = entry_next;
        if (entry_next != 0xffffffffu) stage(a_next, entry_next);
#endif
        {
            {
                if (HOME && survive) {
                    const float tmin = (p.obf & P.eps0_mask) ? P.tmin0 : P.tmin;
                    const float3 o = p.pos, d = p.mom;
                    Nearest best;
                    best.t = P.tmax; best.prim = -1; best.inst = 0; best.n = f3(0.f, 0.f, 0.f);
                    if (home_search(best, P.scene, tmin, o, d, home)) {
                        home_update(home, P.scene, best.prim);
                        const Nearest best_c = best;
                        const float3 o_c = o, d_c = d;
                        HitInfo h_c;
                        hit_finish_body(h_c, P.scene, best_c, o_c, d_c, hit_flags_of(P, DEBUG));     // a candidate answered: never a miss
                        wave_hit_record(r2, h_c);
                        if (DEBUG) { if (P.prd && bounce < P.max_record) P.prd[(size_t)P.max_record * idx + bounce] = r2; }
                        settled = true;
                        nhome++;
                    }
                }
            }
        }
        if (P.lpos && a < count && !survive) {              // lite mode: local position of the photon's last intersect, re-read from
            const Prd* hp = W.hits + a;                     // the hit record so that nothing extra stays live across propagate();
            unsigned pb = hp->prim_boundary;                // the miss program clears it
            P.lpos[idx] = pb == kWaveNoHit ? 0u : pack_lpos(hp->lposcost, hp->lposfphi);
        }
#if PHOX_PROP_PREFETCH && !PHOX_PROP_STAGE
        if (entry_next != 0xffffffffu) {
            const unsigned idx_next = entry_next & kListSlotMask;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(W.hits + a_next));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.photon + idx_next));
        }
#endif
#if PHOX_APPEND_SORT
        // EXPERIMENT (profiles/r2_summary.md, "re-sorting by boundary"): survivors of the chunk are appended grouped by the
        // boundary they will meet next (known for the rays their home cell settled; the pending ones form the last group),
        // so that the warps of the next physics pass see fewer different surface types.  Stable within a group.
        const unsigned ballot = __ballot_sync(0xffffffffu, survive);
        const unsigned pballot = HOME ? __ballot_sync(0xffffffffu, survive && !settled) : 0u;
        const unsigned key = (HOME && settled) ? min(r2.prim_boundary & 0xffffu, 6u) : 7u;
        unsigned my_class_ballot = 0u;
#pragma unroll
        for (unsigned c = 0; c < 8u; c++) {
            const unsigned bc = __ballot_sync(0xffffffffu, survive && key == c);
            if (key == c) my_class_ballot = bc;
            if (lane == 0) s_cls[par][c][warp] = __popc(bc);
        }
        if (lane == 0) s_warp[par][warp] = __popc(pballot) << 16;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned tot = 0, ptot = 0;
            for (int c = 0; c < 8; c++)
                for (int w = 0; w < kPropThreads / 32; w++) { const unsigned n = s_cls[par][c][w]; s_cls[par][c][w] = tot; tot += n; }
            for (int w = 0; w < kPropThreads / 32; w++) { const unsigned n = s_warp[par][w] >> 16; s_warp[par][w] = ptot << 16; ptot += n; }
            s_base[par] = tot ? atomicAdd(W.count_out, tot) : 0u;
            if (HOME) s_pbase[par] = ptot ? atomicAdd(W.pending_count, ptot) : 0u;
        }
        __syncthreads();
        if (survive) {
            const unsigned wo = s_warp[par][warp];
            const unsigned pos = s_base[par] + s_cls[par][key][warp] + __popc(my_class_ballot & ((1u << lane) - 1u));
            __stcs(W.active_out + pos, entry_out);
            if (HOME) {
                __stcs(W.home_next + pos, home);
                if (settled) wave_store_hit(W.hits_next, pos, r2);
                else W.pending[s_pbase[par] + (wo >> 16) + __popc(pballot & ((1u << lane) - 1u))] = make_uint2(pos, entry_out);
            }
        }
#else
        const unsigned ballot = __ballot_sync(0xffffffffu, survive);
        const unsigned pballot = HOME ? __ballot_sync(0xffffffffu, survive && !settled) : 0u;
        // append the survivors of this chunk to the next list, in order within the chunk (HOME: and the unsettled ones
        // among them to the pending list; the two counts share one word per warp, 16 bits each: a chunk has 256 entries)
        if (lane == 0) s_warp[par][warp] = __popc(ballot) | (__popc(pballot) << 16);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned tot = 0, ptot = 0;
            for (int w = 0; w < kPropThreads / 32; w++) {
                const unsigned c = s_warp[par][w];
                s_warp[par][w] = tot | (ptot << 16);
                tot += c & 0xffffu; ptot += c >> 16;
            }
            s_base[par] = tot ? atomicAdd(W.count_out, tot) : 0u;
            if (HOME) s_pbase[par] = ptot ? atomicAdd(W.pending_count, ptot) : 0u;
        }
        __syncthreads();
        if (survive) {
            const unsigned wo = s_warp[par][warp];
            const unsigned pos = s_base[par] + (wo & 0xffffu) + __popc(ballot & ((1u << lane) - 1u));
#if PHOX_WF_STREAM
            __stcs(W.active_out + pos, entry_out);
#else
            W.active_out[pos] = entry_out;
#endif
            if (HOME) {
                __stcs(W.home_next + pos, home);
                if (settled) wave_store_hit(W.hits_next, pos, r2);
                else W.pending[s_pbase[par] + (wo >> 16) + __popc(pballot & ((1u << lane) - 1u))] = make_uint2(pos, entry_out);
            }
        }
#endif
        par ^= 1u;
    }
    if (HOME) {
        for (int off = 16; off > 0; off >>= 1) nhome += __shfl_down_sync(0xffffffffu, nhome, off);
        if (lane == 0 && nhome) { atomicAdd(P.counters, (unsigned long long)nhome); atomicAdd(P.counters + 2, (unsigned long long)nhome); }
    }
}

// after a solid's BVH is built: leaf children whose CSGPrim is exactly its box get that prim's slack in d.z / d.w
// slack < 0 marks a prim that is a single leaf node but not an exact box (kLeafSingle)
__global__ void k_mark_exact_boxes(BvhNode* nodes, int nnode, const float* __restrict__ slack) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnode) return;
    int4 d = nodes[i].d;
    float s0 = (d.x < 0) ? slack[~d.x] : 0.f;
    float s1 = (d.y < 0 && d.y != kBvhNoChild) ? slack[~d.y] : 0.f;
    d.z = s0 > 0.f ? __float_as_int(s0) : 0;
    d.w = s1 > 0.f ? __float_as_int(s1) : 0;
    if (s0 > 0.f) d.x = ~((~d.x) | kLeafExactBox); else if (s0 < 0.f) d.x = ~((~d.x) | kLeafSingle);
    if (s1 > 0.f) d.y = ~((~d.y) | kLeafExactBox); else if (s1 < 0.f) d.y = ~((~d.y) | kLeafSingle);
    nodes[i].d = d;
}

// hits per tile of kHitTile photons (reads only the flagmask word of each photon)
constexpr int kHitTile = 128;
__global__ void __launch_bounds__(kHitTile) k_hit_count(const Photon* __restrict__ photon, unsigned num_photon, unsigned hit_mask,
                                                       unsigned* __restrict__ block_hits) {
    unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    bool is_hit = false;
    if (idx < num_photon) {
        unsigned fm = __ldg(&photon[idx].flagmask);
        is_hit = (fm & hit_mask) == hit_mask;
    }
    int n = __syncthreads_count(is_hit);
    if (threadIdx.x == 0) block_hits[blockIdx.x] = (unsigned)n;
}

// exclusive scan of block_hits[n] -> block_off[n], total -> total_out[0] (single block of 1024 threads, any n): each of the 32 warps
// owns a contiguous segment, sums it with coalesced 32-wide loads, the 32 segment sums are scanned once, then each warp walks its
// segment again with a shuffle scan per 32 entries and a running carry - two block barriers in all.
// (The former version scanned 1024 entries at a time with 20 barriers each: 170 us for the 97 k tiles of a 12.5 M-photon event.)
__global__ void __launch_bounds__(1024) k_hit_offsets(const unsigned* __restrict__ block_hits, int n, unsigned long long* __restrict__ block_off,
                                                       unsigned long long* __restrict__ total_out) {
    __shared__ unsigned long long s_warp[32];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int per = ((n + 31) / 32 + 31) / 32 * 32;               // segment length, a multiple of 32
    const int lo = min(n, (int)warp * per), hi = min(n, lo + per);
    unsigned long long sum = 0ull;
    for (int i = lo + (int)lane; i < hi; i += 32) sum += block_hits[i];
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (lane == 0) s_warp[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long w = s_warp[lane];
        unsigned long long winc = w;
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, off);
            if ((int)lane >= off) winc += t;
        }
        s_warp[lane] = winc - w;                                   // exclusive offset of each warp's segment
        if (lane == 31u) total_out[0] = winc;
    }
    __syncthreads();
    unsigned long long carry = s_warp[warp];
    for (int base = lo; base < hi; base += 32) {
        const int i = base + (int)lane;
        const unsigned v = i < hi ? block_hits[i] : 0u;
        unsigned inc = v;
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, off);
            if ((int)lane >= off) inc += t;
        }
        if (i < hi) block_off[i] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// block b re-reads the flagmasks of its photons and copies the hits, in order, to
// hit[hit_base + block_off[b] + rank]
__global__ void __launch_bounds__(kHitTile) k_hit_compact(const Photon* __restrict__ photon, unsigned num_photon, unsigned hit_mask,
                                                      const unsigned long long* __restrict__ block_off, Photon* __restrict__ hit,
                                                      const unsigned* __restrict__ lpos, PhotonLite* __restrict__ hitlite) {
    __shared__ unsigned warp_count[4];
    unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    bool is_hit = false;
    if (idx < num_photon) {
        unsigned fm = __ldg(&photon[idx].flagmask);
        is_hit = (fm & hit_mask) == hit_mask;
    }
    unsigned ballot = __ballot_sync(0xffffffffu, is_hit);
    unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (lane == 0) warp_count[warp] = __popc(ballot);
    __syncthreads();
    unsigned before = 0;
    for (unsigned w = 0; w < warp; w++) before += warp_count[w];
    if (is_hit) {
        unsigned rank = before + __popc(ballot & ((1u << lane) - 1u));
        const float4* src = reinterpret_cast<const float4*>(photon + idx);
        float4* dst = reinterpret_cast<float4*>(hit + block_off[blockIdx.x] + rank);
        float4 q0 = __ldg(src), q3 = __ldg(src + 3);
        dst[0] = q0; dst[1] = __ldg(src + 1); dst[2] = __ldg(src + 2); dst[3] = q3;
        if (hitlite) {                                       // sphotonlite::init + set_lpos (CSGOptiX7.cu:455-463)
            PhotonLite l;
            l.hitcount_identity = (1u << 16) | (__float_as_uint(q3.y) & 0xffffu);
            l.time = q0.w;
            l.lposcost_lposfphi = lpos[idx];
            l.flagmask = __float_as_uint(q3.w);
            hitlite[block_off[blockIdx.x] + rank] = l;
        }
    }
}

// numphoton prefix of device-resident gensteps (one block)
// info (zeroed by the caller): [0] bit 0 = a scintillation genstep is present, bit 1 = an input-photon genstep is present,
// [1] = number of input-photon gensteps, [2] = numphoton of (the last) one - the device-path form of check_gensteps
__global__ void k_genstep_prefix(const Genstep* __restrict__ gs, int n, unsigned long long* __restrict__ prefix, unsigned* __restrict__ info) {
    __shared__ unsigned long long s[1024];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) { carry = 0ull; prefix[0] = 0ull; }
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        int i = base + threadIdx.x;
        unsigned long long v = i < n ? (unsigned long long)gs[i].u[3] : 0ull;
        if (i < n && info) {
            int code = gs[i].gencode();
            if (code == GS_SCINTILLATION || code == GS_DsG4Scintillation_r4695) atomicOr(info, 1u);
            if (code == GS_INPUT_PHOTON) { atomicOr(info, 2u); atomicAdd(info + 1, 1u); info[2] = gs[i].u[3]; }
        }
        s[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            unsigned long long t = threadIdx.x >= off ? s[threadIdx.x - off] : 0ull;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n) prefix[i + 1] = carry + s[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
}

// geometry-only queries (the role of simtrace / CSGScan)
__global__ void k_intersect(Scene sc, const float4* __restrict__ ray_o_tmin, const float4* __restrict__ ray_d, unsigned n, float tmax,
                            Prd* __restrict__ out) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 o = ray_o_tmin[i], d = ray_d[i];
    HitInfo h;
    trace(h, sc, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), o.w, tmax, kHitFphi);
    Prd r;
    r.nx = h.normal.x; r.ny = h.normal.y; r.nz = h.normal.z; r.t = h.t;
    r.lposcost = h.lposcost; r.lposfphi = h.lposfphi;
    r.iindex_identity = h.iindex_identity; r.prim_boundary = h.prim_boundary;
    out[i] = r;
}

// simtrace (CSGOptiX7.cu:536-577): one ray per slot from FRAME gensteps (qsim.h:2459-2511: local position gs.q1,
// direction from 2 uniforms in the plane named by gridaxes, both through the genstep's own transform gs.q2..q5) or
// from caller-supplied rays (INPUT_PHOTON_SIMTRACE, qsim.h:2455), one trace, record in sevent::add_simtrace layout
// (sevent.h:670-697): q0 normal+t, q1 intersect position+tmin, q2 origin+prim/boundary, q3 direction+iindex/identity.
struct SimtraceParams {
    Scene scene;
    const Genstep* genstep;
    const unsigned long long* gs_prefix;
    int num_genstep;
    const float4* input;                 // quad4 per slot (q0 position, q1 direction) for INPUT_PHOTON_SIMTRACE gensteps
    unsigned long long input_base;
    unsigned long long photon_offset;
    unsigned num;
    float tmin, tmax, refine_distance;
    unsigned refine;
    unsigned long long seed, rng_offset;
    float4* out;                         // quad4 per slot
};
enum : int { GS_FRAME = 17, GS_INPUT_PHOTON_SIMTRACE = 20 };        // OpticksGenstep.h:38,41
enum : int { AX_XYZ = 0, AX_YZ = 1, AX_XZ = 2, AX_XY = 3 };          // sxyz.h:3

__global__ void k_simtrace(const __grid_constant__ SimtraceParams S) {
    unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S.num) return;
    int lo = 0, hi = S.num_genstep;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(S.gs_prefix + mid) <= (unsigned long long)idx) lo = mid; else hi = mid;
    }
    const Genstep& gs = S.genstep[lo];
    unsigned long long photon_idx = S.photon_offset + idx;
    float3 pos = f3(0.f, 0.f, 0.f), mom = f3(0.f, 0.f, 1.f);
    if (gs.gencode() == GS_INPUT_PHOTON_SIMTRACE) {
        float4 a = __ldg(S.input + 4 * (photon_idx - S.input_base)), b = __ldg(S.input + 4 * (photon_idx - S.input_base) + 1);
        pos = f3(a.x, a.y, a.z); mom = f3(b.x, b.y, b.z);
    } else if (gs.gencode() == GS_FRAME) {
        Philox rng;
        rng.init(S.seed, photon_idx, S.rng_offset);                  // sim->rng->init(rng, 0, photon_idx): event index 0
        float u0 = rng.uniform();
        float sinPhi, cosPhi;
        sincosf(2.f * kPi * u0, &sinPhi, &cosPhi);
        float u1 = rng.uniform();
        float cosTheta = 2.f * u1 - 1.f;
        float sinTheta = sqrtf(1.f - cosTheta * cosTheta);
        float3 l = f3(gs.f[4], gs.f[5], gs.f[6]), m;
        switch (gs.i[1]) {
            case AX_YZ: m = f3(0.f, cosPhi, sinPhi); break;
            case AX_XZ: m = f3(cosPhi, 0.f, sinPhi); break;
            case AX_XY: m = f3(cosPhi, sinPhi, 0.f); break;
            default:    m = f3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta); break;
        }
        const float* q = gs.f + 8;                                   // rows q2..q5 of the genstep = qat4, row-vector convention
        pos = f3(q[0] * l.x + q[4] * l.y + q[8] * l.z + q[12] * 1.f, q[1] * l.x + q[5] * l.y + q[9] * l.z + q[13] * 1.f,
                 q[2] * l.x + q[6] * l.y + q[10] * l.z + q[14] * 1.f);
        mom = f3(q[0] * m.x + q[4] * m.y + q[8] * m.z + q[12] * 0.f, q[1] * m.x + q[5] * m.y + q[9] * m.z + q[13] * 0.f,
                 q[2] * m.x + q[6] * m.y + q[10] * m.z + q[14] * 0.f);
    }
    HitInfo h;
    bool ok = trace(h, S.scene, pos, mom, S.tmin, S.tmax, kHitFphi | kHitRawNormal);
    if (S.refine) {
        float t_approx = 0.99f * h.t;
        if (t_approx > S.refine_distance) {
            float3 closer = pos + t_approx * mom;
            ok = trace(h, S.scene, closer, mom, S.tmin, S.tmax, kHitFphi | kHitRawNormal);
            h.t += t_approx;
        }
    }
    if (!ok) { h.normal = f3(0.6f, 0.6f, 0.6f); h.t = 1.f; }          // miss program: background colour in q0.xyz, t = 1 (CSGOptiX7.cu:655-682, SBT.cc:181-193)
    float4* o = S.out + 4 * (size_t)idx;
    o[0] = make_float4(h.normal.x, h.normal.y, h.normal.z, h.t);
    o[1] = make_float4(pos.x + h.t * mom.x, pos.y + h.t * mom.y, pos.z + h.t * mom.z, S.tmin);
    o[2] = make_float4(pos.x, pos.y, pos.z, __uint_as_float(h.prim_boundary));
    o[3] = make_float4(mom.x, mom.y, mom.z, __uint_as_float(h.iindex_identity));
}

// boundary texture readback (QSim.cu boundary_lookup_line role)
__global__ void k_boundary_lookup(Tables tb, const float* __restrict__ nm, const unsigned* __restrict__ line, const unsigned* __restrict__ k,
                                  unsigned n, float4* __restrict__ out) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = bnd_lookup(tb, nm[i], line[i], k[i]);
}

// precooked random streams (qudarap/QSim.cu:43-68)
__global__ void k_rng_sequence(float* __restrict__ out, unsigned ni, unsigned nv, unsigned long long id0,
                               unsigned long long seed, unsigned long long element_offset) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ni) return;
    Philox rng;
    rng.init(seed, id0 + i, element_offset);
    for (unsigned k = 0; k < nv; k++) out[(size_t)i * nv + k] = rng.uniform();
}

}  // namespace phox
