// phox_csg.cuh : ray / CSG-solid intersection for one CSGPrim.
//
// What is computed follows the reference's device headers so that (t, normal) agree to the last
// bit on the same inputs; how it is organised does not: one flat `Hit` register struct, a single
// leaf dispatcher, the boolean-tree evaluator with its stacks kept in a compact local-memory
// struct, and no debug plumbing.
//
//   leaf dispatch + transform/complement handling    CSG/csg_intersect_leaf.h:174-324
//   sphere  csg_intersect_leaf_sphere.h:15-55        zsphere  ..._zsphere.h:44-151
//   box3    ..._box3.h:64-155                        cylinder ..._cylinder.h:34-83
//   cone    ..._newcone.h:50-121                     convexpolyhedron ..._convexpolyhedron.h:20-102
//   hyperboloid ..._hyperboloid.h                    halfspace ..._halfspace.h:156-193
//   phicut  ..._phicut.h:580-635 (the "simple" variant the dispatcher uses)
//   robust quadratic roots                           csg_robust_quadratic_roots.h:95-135
//   list nodes contiguous/discontiguous/overlap      csg_intersect_node.h:392-640, 654-686, 819-905
//   boolean tree (Kensler single-hit state machine)  csg_intersect_tree.h:276-672
//   action table                                     csg_classify.h:55-101
//   postorder bit-twiddling                          csg_postorder.h:56-71
//   prim dispatch                                    csg_intersect_tree.h:683-719
//
// Conventions kept: a hit is valid iff t > tmin strictly; isect = (normal.xyz, t); a complemented
// leaf that misses reports x = -0.f, a leaf whose far side is at infinity reports y = -0.f
// (sign bits are data: csg_intersect_leaf.h:264-274, csg_intersect_tree.h:455-494).
#pragma once
#include "phox_types.h"
#include "phox_math.cuh"

namespace phox {

struct Geo {                        // device pointers to the CSGFoundry arrays
    const float4* node;             // Node[nnode] viewed as 4 x float4
    const float4* plan;             // float4[nplan]
    const float4* itra;             // Qat4[nitra] viewed as 4 x float4
};

PHOX_D unsigned node_u(const float4& q, int k) { return __float_as_uint(k == 0 ? q.x : k == 1 ? q.y : k == 2 ? q.z : q.w); }

// Solve d t^2 + 2 b t + c = 0 without catastrophic cancellation (csg_robust_quadratic_roots.h:95-111)
PHOX_D void quad_roots(float& t1, float& t2, float& disc, float& sdisc, float d, float b, float c) {
    disc = b * b - d * c;
    sdisc = disc > 0.f ? sqrtf(disc) : 0.f;
    float q = b > 0.f ? -(b + sdisc) : -(b - sdisc);
    float r1 = q / d, r2 = c / q;
    t1 = fminf(r1, r2);
    t2 = fmaxf(r1, r2);
}
// variant that parks both roots at `park` when there is no real root (:122-135)
PHOX_D void quad_roots_park(float park, float& t1, float& t2, float& disc, float& sdisc, float d, float b, float c) {
    disc = b * b - d * c;
    sdisc = disc > 0.f ? sqrtf(disc) : 0.f;
    float q = b > 0.f ? -(b + sdisc) : -(b - sdisc);
    float r1 = sdisc > 0.f ? q / d : park;
    float r2 = sdisc > 0.f ? c / q : park;
    t1 = fminf(r1, r2);
    t2 = fmaxf(r1, r2);
}

PHOX_D bool leaf_sphere(float4& is, const float4& q0, float tmin, const float3& ro, const float3& rd) {
    float3 O = ro - f3(q0.x, q0.y, q0.z);
    float radius = q0.w;
    float b = dot(O, rd);
    float c = dot(O, O) - radius * radius;
    float d = dot(rd, rd);
    float r1, r2, disc, sdisc;
    quad_roots(r1, r2, disc, sdisc, d, b, c);
    float t = sdisc > 0.f ? (r1 > tmin ? r1 : r2) : tmin;
    bool ok = t > tmin;
    if (ok) {
        is.x = (O.x + t * rd.x) / radius;
        is.y = (O.y + t * rd.y) / radius;
        is.z = (O.z + t * rd.z) / radius;
        is.w = t;
    }
    return ok;
}

PHOX_D bool leaf_zsphere(float4& is, const float4& q0, const float4& q1, float tmin, const float3& ro, const float3& rd) {
    float3 center = f3(q0.x, q0.y, q0.z);
    float3 O = ro - center;
    float radius = q0.w;
    float b = dot(O, rd);
    float c = dot(O, O) - radius * radius;
    if (c > 0.f && b > 0.f) return false;          // outside and heading away
    float zmax = center.z + q1.y;
    float zmin = center.z + q1.x;
    float d = dot(rd, rd);
    float t1s, t2s, disc, sdisc;
    quad_roots(t1s, t2s, disc, sdisc, d, b, c);
    float z1s = ro.z + t1s * rd.z;
    float z2s = ro.z + t2s * rd.z;
    float idz = 1.f / rd.z;
    float tQ = (zmax - ro.z) * idz;                // upper cap plane
    float tP = (zmin - ro.z) * idz;                // lower cap plane
    float t1c = fminf(tQ, tP);
    float t2c = fmaxf(tQ, tP);
    if (t1c < t1s || t1c > t2s) t1c = tmin;        // cap hits outside the sphere are void
    if (t2c < t1s || t2c > t2s) t2c = tmin;
    float t = tmin;
    if (sdisc > 0.f) {
        if (t1s > tmin && z1s > zmin && z1s <= zmax) t = t1s;
        else if (t1c > tmin) t = t1c;
        else if (t2c > tmin) t = t2c;
        else if (t2s > tmin && z2s > zmin && z2s <= zmax) t = t2s;
    }
    bool ok = t > tmin;
    if (ok) {
        is.w = t;
        if (t == t1s || t == t2s) {
            is.x = (O.x + t * rd.x) / radius;
            is.y = (O.y + t * rd.y) / radius;
            is.z = (O.z + t * rd.z) / radius;
        } else {
            is.x = 0.f; is.y = 0.f;
            is.z = t == tP ? -1.f : 1.f;
        }
    }
    return ok;
}

// idir = 1 / rd, supplied by the caller (the traversal already holds it for its slab tests; IEEE division gives the same bits)
PHOX_D bool leaf_box3_idir(float4& is, const float4& q0, float tmin, const float3& ro, const float3& rd, const float3& idir);

PHOX_D bool leaf_box3(float4& is, const float4& q0, float tmin, const float3& ro, const float3& rd) {
    return leaf_box3_idir(is, q0, tmin, ro, rd, f3(1.f / rd.x, 1.f / rd.y, 1.f / rd.z));
}

// The box leaf in two halves, so that a caller with several boxes to compare (the home-cell pass) can settle the
// nearest distance first and work out one normal, for the winner.  box3_t: does the ray meet the box beyond tmin, and where.
// (half = the three full sizes / 2, an exact operation, so a caller may hold the halves ready made)
PHOX_D bool box3_t_half(float& t_out, const float3& half, float tmin, const float3& ro, const float3& rd, const float3& idir) {
    float3 bmin = f3(-half.x, -half.y, -half.z);
    float3 bmax = half;
    float3 t0 = f3((bmin.x - ro.x) * idir.x, (bmin.y - ro.y) * idir.y, (bmin.z - ro.z) * idir.z);
    float3 t1 = f3((bmax.x - ro.x) * idir.x, (bmax.y - ro.y) * idir.y, (bmax.z - ro.z) * idir.z);
    float3 nr = f3(fminf(t0.x, t1.x), fminf(t0.y, t1.y), fminf(t0.z, t1.z));
    float3 fr = f3(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y), fmaxf(t0.z, t1.z));
    float t_near = fmaxf(fmaxf(nr.x, nr.y), nr.z);
    float t_far = fminf(fminf(fr.x, fr.y), fr.z);

    // axis-parallel rays (two zero components) are decided by the origin's position in the other two slabs; everything else by
    // the slab distances.  The rare case is kept off the common path (same decisions as csg_intersect_leaf_box3.h:100-125).
    const int nzero = (rd.x == 0.f) + (rd.y == 0.f) + (rd.z == 0.f);
    bool has;
    if (nzero == 2) {
        bool in_x = ro.x > bmin.x && ro.x < bmax.x;
        bool in_y = ro.y > bmin.y && ro.y < bmax.y;
        bool in_z = ro.z > bmin.z && ro.z < bmax.z;
        has = rd.x != 0.f ? (in_y && in_z) : (rd.y != 0.f ? (in_x && in_z) : (in_x && in_y));
    } else has = (t_far > t_near && t_far > 0.f);
    if (!has) return false;
    float t = tmin < t_near ? t_near : (tmin < t_far ? t_far : tmin);
    t_out = t;
    return t > tmin;
}
PHOX_D bool box3_t(float& t_out, const float4& q0, float tmin, const float3& ro, const float3& rd, const float3& idir) {
    return box3_t_half(t_out, f3(q0.x / 2.f, q0.y / 2.f, q0.z / 2.f), tmin, ro, rd, idir);
}
// ... and the face normal at distance t along the ray
PHOX_D float3 box3_normal_half(const float3& half, const float3& ro, const float3& rd, float t) {
    float3 bmin = f3(-half.x, -half.y, -half.z);
    float3 bmax = half;
    float3 p = f3(ro.x + t * rd.x - 0.f, ro.y + t * rd.y - 0.f, ro.z + t * rd.z - 0.f);
    float3 pa = f3(fabsf(p.x) / (bmax.x - bmin.x), fabsf(p.y) / (bmax.y - bmin.y), fabsf(p.z) / (bmax.z - bmin.z));
    float3 n = f3(0.f, 0.f, 0.f);
    if (pa.x >= pa.y && pa.x >= pa.z) n.x = copysignf(1.f, p.x);
    else if (pa.y >= pa.x && pa.y >= pa.z) n.y = copysignf(1.f, p.y);
    else if (pa.z >= pa.x && pa.z >= pa.y) n.z = copysignf(1.f, p.z);
    return n;
}
PHOX_D float3 box3_normal(const float4& q0, const float3& ro, const float3& rd, float t) {
    return box3_normal_half(f3(q0.x / 2.f, q0.y / 2.f, q0.z / 2.f), ro, rd, t);
}

PHOX_D bool leaf_box3_idir(float4& is, const float4& q0, float tmin, const float3& ro, const float3& rd, const float3& idir) {
    float t;
    if (!box3_t(t, q0, tmin, ro, rd, idir)) return false;
    float3 n = box3_normal(q0, ro, rd, t);
    is.x = n.x; is.y = n.y; is.z = n.z; is.w = t;
    return true;
}

PHOX_D bool leaf_cylinder(float4& is, const float4& q0, const float4& q1, float tmin, const float3& ro, const float3& rd) {
    float r = q0.w, z1 = q1.x, z2 = q1.y;
    float ox = ro.x, oy = ro.y, oz = ro.z, vx = rd.x, vy = rd.y, vz = rd.z;
    float r2 = r * r;
    float a = vx * vx + vy * vy;
    float b = ox * vx + oy * vy;
    float c = ox * ox + oy * oy - r2;
    float t_near, t_far, disc, sdisc;
    quad_roots_park(tmin, t_near, t_far, disc, sdisc, a, b, c);
    float z_near = oz + t_near * vz;
    float z_far = oz + t_far * vz;
    float t_c1 = (z1 - oz) / vz;
    float rr1 = (ox + t_c1 * vx) * (ox + t_c1 * vx) + (oy + t_c1 * vy) * (oy + t_c1 * vy);
    float t_c2 = (z2 - oz) / vz;
    float rr2 = (ox + t_c2 * vx) * (ox + t_c2 * vx) + (oy + t_c2 * vy) * (oy + t_c2 * vy);
    float t = CUDART_INF_F;
    if (t_near > tmin && z_near > z1 && z_near < z2 && t_near < t) t = t_near;
    if (t_far > tmin && z_far > z1 && z_far < z2 && t_far < t) t = t_far;
    if (t_c1 > tmin && rr1 <= r2 && t_c1 < t) t = t_c1;
    if (t_c2 > tmin && rr2 <= r2 && t_c2 < t) t = t_c2;
    bool ok = t > tmin && t < CUDART_INF_F;
    if (ok) {
        bool sheet = (t == t_near || t == t_far);
        is.x = sheet ? (ox + t * vx) / r : 0.f;
        is.y = sheet ? (oy + t * vy) / r : 0.f;
        is.z = sheet ? 0.f : (t == t_c1 ? -1.f : 1.f);
        is.w = t;
    }
    return ok;
}

PHOX_D bool leaf_cone(float4& is, const float4& q0, float tmin, const float3& o, const float3& d) {
    float r1 = q0.x, z1 = q0.y, r2 = q0.z, z2 = q0.w;
    float r1r1 = r1 * r1, r2r2 = r2 * r2;
    float tth = (r2 - r1) / (z2 - z1);
    float tth2 = tth * tth;
    float z0 = (z2 * r1 - z1 * r2) / (r1 - r2);     // apex
    float idz = 1.f / d.z;
    float t_cap1 = d.z == 0.f ? kRtMax : (z1 - o.z) * idz;
    float t_cap2 = d.z == 0.f ? kRtMax : (z2 - o.z) * idz;
    float rr_cap1 = (o.x + t_cap1 * d.x) * (o.x + t_cap1 * d.x) + (o.y + t_cap1 * d.y) * (o.y + t_cap1 * d.y);
    float rr_cap2 = (o.x + t_cap2 * d.x) * (o.x + t_cap2 * d.x) + (o.y + t_cap2 * d.y) * (o.y + t_cap2 * d.y);
    t_cap1 = rr_cap1 < r1r1 && t_cap1 > tmin ? t_cap1 : kRtMax;
    t_cap2 = rr_cap2 < r2r2 && t_cap2 > tmin ? t_cap2 : kRtMax;
    float c2 = d.x * d.x + d.y * d.y - d.z * d.z * tth2;
    float c1 = o.x * d.x + o.y * d.y - (o.z - z0) * d.z * tth2;
    float c0 = o.x * o.x + o.y * o.y - (o.z - z0) * (o.z - z0) * tth2;
    float t_near, t_far, disc, sdisc;
    quad_roots_park(kRtMax, t_near, t_far, disc, sdisc, c2, c1, c0);
    float z_near = o.z + t_near * d.z;
    float z_far = o.z + t_far * d.z;
    t_near = z_near > z1 && z_near < z2 && t_near > tmin ? t_near : kRtMax;
    t_far = z_far > z1 && z_far < z2 && t_far > tmin ? t_far : kRtMax;
    float t = fminf(fminf(t_near, t_far), fminf(t_cap1, t_cap2));
    bool ok = t > tmin && t < kRtMax;
    if (ok) {
        if (t == t_cap1 || t == t_cap2) {
            is.x = 0.f; is.y = 0.f;
            is.z = t == t_cap2 ? 1.f : -1.f;
        } else {
            float3 n = normalize(f3(o.x + t * d.x, o.y + t * d.y, (z0 - (o.z + t * d.z)) * tth2));
            is.x = n.x; is.y = n.y; is.z = n.z;
        }
        is.w = t;
    }
    return ok;
}

PHOX_D bool leaf_convexpolyhedron(float4& is, const float4& q0, const float4* plan, float tmin, const float3& ro, const float3& rd) {
    float t0 = -CUDART_INF_F, t1 = CUDART_INF_F;
    float3 n0 = f3(0.f, 0.f, 0.f), n1 = f3(0.f, 0.f, 0.f);
    unsigned plane_idx = __float_as_uint(q0.x), plane_num = __float_as_uint(q0.y);
    for (unsigned i = 0; i < plane_num; i++) {
        float4 pl = __ldg(plan + plane_idx + i);
        float3 n = f3(pl.x, pl.y, pl.z);
        float nd = dot(n, rd);
        float no = dot(n, ro);
        float dist = no - pl.w;
        float tc = -dist / nd;
        bool par_in = nd == 0.f && dist < 0.f;
        bool par_out = nd == 0.f && dist > 0.f;
        if (par_in) continue;
        if (par_out) return false;
        if (nd < 0.f) { if (tc > t0) { t0 = tc; n0 = n; } }
        else          { if (tc < t1) { t1 = tc; n1 = n; } }
    }
    bool ok = t0 < t1;     // NB as in the reference: valid even when neither root exceeds tmin (t stays 0)
    if (ok) {
        if (t0 > tmin) { is.x = n0.x; is.y = n0.y; is.z = n0.z; is.w = t0; }
        else if (t1 > tmin) { is.x = n1.x; is.y = n1.y; is.z = n1.z; is.w = t1; }
    }
    return ok;
}

PHOX_D bool leaf_hyperboloid(float4& is, const float4& q0, float tmin, const float3& ro, const float3& rd) {
    float r0 = q0.x, zf = q0.y, z1 = q0.z, z2 = q0.w;
    float rr0 = r0 * r0;
    float z1s = z1 / zf, z2s = z2 / zf;
    float rr1 = rr0 * (z1s * z1s + 1.f);
    float rr2 = rr0 * (z2s * z2s + 1.f);
    float A = -rr0 / (zf * zf);
    float B = -rr0;
    float sx = rd.x, sy = rd.y, sz = rd.z, ox = ro.x, oy = ro.y, oz = ro.z;
    float d = sx * sx + sy * sy + A * sz * sz;
    float b = ox * sx + oy * sy + A * oz * sz;
    float c = ox * ox + oy * oy + A * oz * oz + B;
    float t1h, t2h, disc, sdisc;
    quad_roots(t1h, t2h, disc, sdisc, d, b, c);
    float h1z = oz + t1h * sz;
    float h2z = oz + t2h * sz;
    float osz = 1.f / sz;
    float t2c = (z2 - oz) * osz;
    float t1c = (z1 - oz) * osz;
    float3 c1 = ro + t1c * rd;
    float3 c2 = ro + t2c * rd;
    float crr1 = c1.x * c1.x + c1.y * c1.y;
    float crr2 = c2.x * c2.x + c2.y * c2.y;
    float ca = t1h > tmin && disc > 0.f && h1z > z1 && h1z < z2 ? t1h : kRtMax;
    float cb = t2h > tmin && disc > 0.f && h2z > z1 && h2z < z2 ? t2h : kRtMax;
    float cc = t2c > tmin && crr2 < rr2 ? t2c : kRtMax;
    float cd = t1c > tmin && crr1 < rr1 ? t1c : kRtMax;
    float t = fminf(fminf(ca, cb), fminf(cc, cd));
    bool ok = t > tmin && t < kRtMax;
    if (ok) {
        is.w = t;
        if (t == t1h || t == t2h) {
            float3 p = ro + t * rd;
            float3 n = normalize(f3(p.x, p.y, A * p.z));
            is.x = n.x; is.y = n.y; is.z = n.z;
        } else {
            is.x = 0.f; is.y = 0.f;
            is.z = t == t1c ? -1.f : 1.f;
        }
    }
    return ok;
}

PHOX_D bool leaf_halfspace(float4& is, const float4& q0, float tmin, const float3& o, const float3& d) {
    float3 n = f3(q0.x, q0.y, q0.z);
    float w = q0.w;
    float on = dot(o, n);
    float dn = dot(d, n);
    float on_w = on - w;
    float adn = fabsf(dn);
    bool inside = on_w < -1e-9f;
    float t = adn > 0.f ? -on_w / dn : tmin;
    bool ok = t > tmin;
    if (ok) { is.x = n.x; is.y = n.y; is.z = n.z; is.w = t; }
    else if (inside) is.y = -0.f;               // exit at infinity
    return ok;
}

PHOX_D bool leaf_phicut(float4& is, const float4& q0, float tmin, const float3& o, const float3& d) {
    float cosPhi0 = q0.x, sinPhi0 = q0.y, cosPhi1 = q0.z, sinPhi1 = q0.w;
    float d_n0 = d.x * sinPhi0 + d.y * (-cosPhi0);
    float d_n1 = d.x * (-sinPhi1) + d.y * (cosPhi1);
    float o_n0 = o.x * sinPhi0 + o.y * (-cosPhi0);
    float o_n1 = o.x * (-sinPhi1) + o.y * (cosPhi1);
    float t0 = d_n0 == 0.f ? tmin : -o_n0 / d_n0;
    float t1 = d_n1 == 0.f ? tmin : -o_n1 / d_n1;
    float PR = d_n0 == 0.f ? -o_n0 : -d_n0;
    float QR = d_n1 == 0.f ? -o_n1 : -d_n1;
    float PQ = cosPhi0 * sinPhi1 - cosPhi1 * sinPhi0;
    bool unbounded_exit = PQ >= 0.f ? (PR >= 0.f && QR <= 0.f) : (PR >= 0.f || QR <= 0.f);
    float side0 = o.x * cosPhi0 + o.y * sinPhi0 + (d.x * cosPhi0 + d.y * sinPhi0) * t0;
    float side1 = o.x * cosPhi1 + o.y * sinPhi1 + (d.x * cosPhi1 + d.y * sinPhi1) * t1;
    if (side0 < 0.f) t0 = tmin;
    if (side1 < 0.f) t1 = tmin;
    float t_near = fminf(t0, t1);
    float t_far = fmaxf(t0, t1);
    float t = t_near > tmin ? t_near : (t_far > tmin ? t_far : tmin);
    bool ok = t > tmin;
    if (ok) {
        is.x = t == t1 ? -sinPhi1 : sinPhi0;
        is.y = t == t1 ? cosPhi1 : -cosPhi0;
        is.z = 0.f;
        is.w = t;
    } else if (unbounded_exit) {
        is.y = -is.y;
    }
    return ok;
}

// One leaf node: optional inverse transform of the ray, shape switch, normal back-transform,
// complement handling.  `nd` points at the node's four float4.
PHOX_D bool intersect_leaf(float4& is, const float4* nd, const Geo& g, float tmin, const float3& ro, const float3& rd) {
    is = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 q0 = __ldg(nd + 0), q1 = __ldg(nd + 1), q3 = __ldg(nd + 3);
    unsigned typecode = __float_as_uint(q3.z);
    unsigned tw = __float_as_uint(q3.w);
    unsigned tidx = tw & 0x7fffffffu;
    bool complement = (tw & 0x80000000u) != 0u;

    float3 o = ro, d = rd;
    float4 r0, r1, r2;
    if (tidx > 0u) {
        const float4* m = g.itra + 4u * (tidx - 1u);
        r0 = __ldg(m + 0); r1 = __ldg(m + 1); r2 = __ldg(m + 2);
        float4 r3 = __ldg(m + 3);
        o = xform(r0, r1, r2, r3, ro, 1.f);
        d = xform(r0, r1, r2, r3, rd, 0.f);
    }
    bool ok = false;
    switch (typecode) {
        case CSG_SPHERE:           ok = leaf_sphere(is, q0, tmin, o, d); break;
        case CSG_ZSPHERE:          ok = leaf_zsphere(is, q0, q1, tmin, o, d); break;
        case CSG_CYLINDER:         ok = leaf_cylinder(is, q0, q1, tmin, o, d); break;
        case CSG_BOX3:             ok = leaf_box3(is, q0, tmin, o, d); break;
        case CSG_CONE:             ok = leaf_cone(is, q0, tmin, o, d); break;
        case CSG_CONVEXPOLYHEDRON: ok = leaf_convexpolyhedron(is, q0, g.plan, tmin, o, d); break;
        case CSG_HYPERBOLOID:      ok = leaf_hyperboloid(is, q0, tmin, o, d); break;
        case CSG_PHICUT:           ok = leaf_phicut(is, q0, tmin, o, d); break;
        case CSG_HALFSPACE:        ok = leaf_halfspace(is, q0, tmin, o, d); break;
        default: break;
    }
    if (ok && tidx > 0u) {
        float3 n = xform_normal(r0, r1, r2, f3(is.x, is.y, is.z));
        is.x = n.x; is.y = n.y; is.z = n.z;
    }
    if (complement) {
        is.x = ok ? -is.x : -0.f;
        is.y = ok ? -is.y : is.y;
        is.z = ok ? -is.z : is.z;
    }
    return ok;
}

// Out-of-line copy for the list-node and boolean-tree evaluators: they call the leaf dispatcher
// from many places and are the rare case, so they share ONE compiled body instead of inlining
// nine shapes at every call site (the hot single-leaf prim path inlines intersect_leaf directly).
__device__ __noinline__ bool intersect_leaf_cold(float4& is, const float4* nd, const Geo& g, float tmin, const float3& ro, const float3& rd) {
    return intersect_leaf(is, nd, g, tmin, ro, rd);
}

// ENTER / EXIT / MISS of an isect against the ray (csg_classify.h CSG_CLASSIFY)
enum : int { ST_ENTER = 0, ST_EXIT = 1, ST_MISS = 2 };
PHOX_D int classify(const float4& is, const float3& rd, float tmin) {
    return fabsf(is.w) > tmin ? ((is.x * rd.x + is.y * rd.y + is.z * rd.z < 0.f) ? ST_ENTER : ST_EXIT) : ST_MISS;
}

// ---- list nodes ------------------------------------------------------------------------------
__device__ __noinline__ bool list_discontiguous(float4& is, const float4* nd, const float4* root, const Geo& g, float tmin, const float3& ro, const float3& rd) {
    float4 h = __ldg(nd);
    unsigned num = __float_as_uint(h.x), off = __float_as_uint(h.y);
    float4 closest = make_float4(0.f, 0.f, 0.f, kRtMax);
    float4 sub;
    for (unsigned i = 0; i < num; i++) {
        if (intersect_leaf_cold(sub, root + 4u * (off + i), g, tmin, ro, rd)) {
            if (sub.w < closest.w) closest = sub;
        }
    }
    bool ok = closest.w < kRtMax;
    if (ok) is = closest;
    return ok;
}

__device__ __noinline__ bool list_overlap(float4& is, const float4* nd, const float4* root, const Geo& g, float tmin, const float3& ro, const float3& rd) {
    float4 h = __ldg(nd);
    unsigned num = __float_as_uint(h.x), off = __float_as_uint(h.y);
    float4 far_enter = make_float4(0.f, 0.f, 0.f, tmin);
    float4 near_exit = make_float4(0.f, 0.f, 0.f, kRtMax);
    float4 sub = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned n_enter = 0, n_exit = 0;
    for (unsigned i = 0; i < num; i++) {
        const float4* sn = root + 4u * (off + i);
        if (intersect_leaf_cold(sub, sn, g, tmin, ro, rd)) {
            int st = classify(sub, rd, tmin);
            if (st == ST_ENTER) {
                n_enter += 1;
                if (sub.w > far_enter.w) far_enter = sub;
                float tadv = sub.w + 0.0001f;
                if (intersect_leaf_cold(sub, sn, g, tadv, ro, rd)) {
                    if (classify(sub, rd, tadv) == ST_EXIT) {
                        n_exit += 1;
                        if (sub.w < near_exit.w) near_exit = sub;
                    }
                }
            } else if (st == ST_EXIT) {
                n_exit += 1;
                if (sub.w < near_exit.w) near_exit = sub;
            }
        }
    }
    bool ok = false;
    bool all = far_enter.w < near_exit.w && max(n_enter, n_exit) == num;
    if (all) {
        if (far_enter.w > tmin && far_enter.w < kRtMax) { ok = true; is = far_enter; }
        else if (near_exit.w > tmin && near_exit.w < kRtMax) { ok = true; is = near_exit; }
    }
    return ok;
}

__device__ __noinline__ bool list_contiguous(float4& is, const float4* nd, const float4* root, const Geo& g, float tmin, const float3& ro, const float3& rd) {
    float4 h = __ldg(nd);
    int num = (int)__float_as_uint(h.x), off = (int)__float_as_uint(h.y);
    float4 near_enter = make_float4(0.f, 0.f, 0.f, kRtMax);
    float4 far_exit = make_float4(0.f, 0.f, 0.f, tmin);
    float4 sub = make_float4(0.f, 0.f, 0.f, 0.f);
    int n_exit = 0;
    for (int i = 0; i < num; i++) {
        if (intersect_leaf_cold(sub, root + 4 * (off + i), g, tmin, ro, rd)) {
            int st = classify(sub, rd, tmin);
            if (st == ST_ENTER) { if (sub.w < near_enter.w) near_enter = sub; }
            else if (st == ST_EXIT) n_exit += 1;
        }
    }
    if (n_exit == 0) {                          // ray starts outside the compound
        bool ok = near_enter.w > tmin && near_enter.w < kRtMax;
        if (ok) is = near_enter;
        return ok;
    }
    // inside: walk outwards through the overlapping constituents to the farthest contiguous exit
    int n_enter = 0;
    float enter[8]; int aux[8]; int idx[8];
    for (int i = 0; i < num; i++) {
        if (intersect_leaf_cold(sub, root + 4 * (off + i), g, tmin, ro, rd)) {
            int st = classify(sub, rd, tmin);
            if (st == ST_ENTER) {
                aux[n_enter] = i; idx[n_enter] = n_enter; enter[n_enter] = sub.w;
                n_enter += 1;
            } else if (st == ST_EXIT) {
                n_exit += 1;
                if (sub.w > far_exit.w) far_exit = sub;
            }
        }
    }
    for (int i = 1; i < n_enter; i++) {          // insertion sort of the enter distances
        int key = idx[i];
        int j = i - 1;
        while (j >= 0 && enter[idx[j]] > enter[key]) { idx[j + 1] = idx[j]; j = j - 1; }
        idx[j + 1] = key;
    }
    for (int i = 0; i < n_enter; i++) {
        float tadv = enter[idx[i]] + 0.0001f;
        int isub = aux[idx[i]];
        if (tadv < far_exit.w) {
            if (intersect_leaf_cold(sub, root + 4 * (off + isub), g, tadv, ro, rd)) {
                if (classify(sub, rd, tadv) == ST_EXIT) {
                    n_exit += 1;
                    if (sub.w > far_exit.w) far_exit = sub;
                }
            }
        }
    }
    bool ok = n_exit > 0 && far_exit.w > tmin;
    if (ok) is = far_exit;
    return ok;
}

// a node of a tree is a leaf or a list (csg_intersect_node.h:913-938)
PHOX_D bool intersect_node(float4& is, const float4* nd, const float4* root, const Geo& g, float tmin, const float3& ro, const float3& rd) {
    unsigned typecode = __float_as_uint(__ldg(nd + 3).z);
    switch (typecode) {
        case CSG_CONTIGUOUS:    return list_contiguous(is, nd, root, g, tmin, ro, rd);
        case CSG_OVERLAP:       return list_overlap(is, nd, root, g, tmin, ro, rd);
        case CSG_DISCONTIGUOUS: return list_discontiguous(is, nd, root, g, tmin, ro, rd);
        default:                return intersect_leaf_cold(is, nd, g, tmin, ro, rd);
    }
}

// ---- boolean tree ----------------------------------------------------------------------------
// action codes of the packed tables (csg_classify.h:55-62)
enum : int { ACT_MISS = 0, ACT_A = 1, ACT_B = 2, ACT_FLIP_B = 3, ACT_LOOP_A = 4, ACT_LOOP_B = 5 };

PHOX_D int boolean_action(unsigned op, int stA, int stB, bool a_closer) {
    // rows: union, intersection, difference ; nibble index 3*stateA + stateB (csg_classify.h:86-101)
    unsigned tab = a_closer ? (op == CSG_UNION ? 0x22121141u : op == CSG_INTERSECTION ? 0x00014014u : op == CSG_DIFFERENCE ? 0x00141141u : 0u)
                            : (op == CSG_UNION ? 0x22115122u : op == CSG_INTERSECTION ? 0x00022055u : op == CSG_DIFFERENCE ? 0x00133155u : 0u);
    unsigned off = 3u * (unsigned)stA + (unsigned)stB;
    return off < 8u ? (int)((tab >> (off * 4u)) & 0xfu) : ACT_MISS;
}

constexpr int kCsgStack = 15;        // csg_stack.h CSG_STACK_SIZE
constexpr int kTrancheStack = 4;     // csg_tranche.h TRANCHE_STACK_SIZE

PHOX_D unsigned postorder_next(unsigned i, unsigned elevation) {
    return (i & 1u) ? (i >> 1) : ((i << elevation) + (1u << elevation));
}

// Complete binary tree in level order at `root`, numNode = root.subNum.  Evaluated in postorder
// slices ("tranches"); LOOP actions re-shoot one side with tmin advanced past its last hit.
__device__ __noinline__ bool intersect_tree(float4& isect, const float4* root, const Geo& g, float t_min, const float3& ro, const float3& rd) {
    int num_node = (int)__float_as_uint(__ldg(root).x);
    unsigned height = (unsigned)(__ffs(num_node + 1) - 2);
    bool err = false;

    float tr_tmin[kTrancheStack];
    unsigned tr_slice[kTrancheStack];
    int tr_top = -1;
    float4 st[kCsgStack];
    int top = -1;

    tr_top = 0;
    tr_slice[0] = ((1u << height) & 0xffu) << 16;            // begin = leftmost, end = 0 (parent of root)
    tr_tmin[0] = t_min;

    while (tr_top > -1) {
        unsigned slice = tr_slice[tr_top];
        float tmin = tr_tmin[tr_top];
        tr_top--;
        unsigned idx = (slice >> 16) & 0xffu;
        unsigned end = (slice >> 24) & 0xffu;

        while (idx != end) {
            unsigned depth = 31u - (unsigned)__clz(idx);
            unsigned elevation = height - depth;
            const float4* nd = root + 4u * (idx - 1u);
            unsigned typecode = __float_as_uint(__ldg(nd + 3).z);
            if (typecode == CSG_ZERO) { idx = postorder_next(idx, elevation); continue; }

            if (typecode >= CSG_NODE) {
                float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
                intersect_node(h, nd, root, g, tmin, ro, rd);
                h.w = copysignf(h.w, (idx % 2u == 0u) ? -1.f : 1.f);      // sign of t = which side, left -ve
                if (top >= kCsgStack - 1) { err = true; break; }
                st[++top] = h;
            } else {
                if (top < 1) { err = true; break; }
                bool first_left = signbit(st[top].w);
                bool second_left = signbit(st[top - 1].w);
                if (!(first_left ^ second_left)) { err = true; break; }
                int left = first_left ? top : top - 1;
                int right = first_left ? top - 1 : top;
                float4 L = st[left], R = st[right];

                int l_state = classify(L, rd, tmin);
                int r_state = classify(R, rd, tmin);
                float t_left = fabsf(L.w), t_right = fabsf(R.w);
                bool left_closer = t_left <= t_right;

                bool l_promote = l_state == ST_MISS && (signbit(L.x) || signbit(L.y));
                bool r_promote = r_state == ST_MISS && (signbit(R.x) || signbit(R.y));
                if (r_promote) { r_state = ST_EXIT; left_closer = true; }
                if (l_promote) { l_state = ST_EXIT; left_closer = false; }

                int act = boolean_action(typecode, l_state, r_state, left_closer);
                if (act < ACT_LOOP_A) {
                    float4 res = act == ACT_MISS ? make_float4(0.f, 0.f, 0.f, 0.f) : (act == ACT_A ? L : R);
                    if (act == ACT_FLIP_B) { res.x = -res.x; res.y = -res.y; res.z = -res.z; }
                    res.w = copysignf(res.w, (idx % 2u == 0u) ? -1.f : 1.f);
                    top -= 2;
                    st[++top] = res;
                } else {
                    bool loop_a = act == ACT_LOOP_A;
                    unsigned left_idx = 2u * idx, right_idx = left_idx + 1u;
                    float t_adv = fabsf(loop_a ? L.w : R.w) + 0.0001f;
                    float4 other = loop_a ? R : L;
                    top -= 2;
                    st[++top] = other;
                    unsigned end_tree = ((idx & 0xffu) << 16) | ((end & 0xffu) << 24);
                    unsigned left_tree = (((left_idx << (elevation - 1u)) & 0xffu) << 16) | (((right_idx << (elevation - 1u)) & 0xffu) << 24);
                    unsigned right_tree = (((right_idx << (elevation - 1u)) & 0xffu) << 16) | ((idx & 0xffu) << 24);
                    if (tr_top >= kTrancheStack - 2) { err = true; break; }
                    tr_top++; tr_slice[tr_top] = end_tree; tr_tmin[tr_top] = tmin;
                    tr_top++; tr_slice[tr_top] = loop_a ? left_tree : right_tree; tr_tmin[tr_top] = t_adv;
                    break;                                    // restart with the pushed sub-slices
                }
            }
            idx = postorder_next(idx, elevation);
        }
        if (err) break;
    }
    if (top == 0) isect = st[0];
    return isect.w > 0.f;
}

// nearest hit of one CSGPrim whose root node is at `root` (csg_intersect_tree.h:683-719).
// isect must come in zeroed.
PHOX_D bool intersect_prim(float4& isect, const float4* root, const Geo& g, float tmin, const float3& ro, const float3& rd) {
    unsigned typecode = __float_as_uint(__ldg(root + 3).z);
    if (typecode >= CSG_LEAF) return intersect_leaf(isect, root, g, tmin, ro, rd);
    if (typecode < CSG_NODE) return intersect_tree(isect, root, g, tmin, ro, rd);
    if (typecode == CSG_CONTIGUOUS) return list_contiguous(isect, root, root, g, tmin, ro, rd);
    if (typecode == CSG_DISCONTIGUOUS) return list_discontiguous(isect, root, root, g, tmin, ro, rd);
    if (typecode == CSG_OVERLAP) return list_overlap(isect, root, root, g, tmin, ro, rd);
    return false;
}

// fully out-of-line prim test for the validation paths (brute-force loop)
__device__ __noinline__ bool intersect_prim_cold(float4& isect, const float4* root, const Geo& g, float tmin, const float3& ro, const float3& rd) {
    unsigned typecode = __float_as_uint(__ldg(root + 3).z);
    if (typecode >= CSG_LEAF) return intersect_leaf_cold(isect, root, g, tmin, ro, rd);
    if (typecode < CSG_NODE) return intersect_tree(isect, root, g, tmin, ro, rd);
    if (typecode == CSG_CONTIGUOUS) return list_contiguous(isect, root, root, g, tmin, ro, rd);
    if (typecode == CSG_DISCONTIGUOUS) return list_discontiguous(isect, root, root, g, tmin, ro, rd);
    if (typecode == CSG_OVERLAP) return list_overlap(isect, root, root, g, tmin, ro, rd);
    return false;
}

}  // namespace phox
