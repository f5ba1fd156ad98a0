// oracle/phox_oracle.cpp : CPU restatement of the reference's simulate path.
//
// TEST INFRASTRUCTURE.  Nothing in the product (eic-opticks_b200/, include/) includes, links or
// calls this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs do, and only as the checker / the timed CPU baseline.
//
// Parity status: PINNED - (1) the CSG functions here are checked ray-by-ray against the
// reference's own headers compiled for the host (oracle/_ref/libcsgref.so, tests/test_oracle_csg.py),
// (2) the whole path is checked photon-by-photon on the B200 against the reference's device headers
// compiled unmodified (oracle/_ref/libphoxref_*.so, tests/test_parity_gpu.py), (3) the random
// stream is checked against curand's host implementation of Philox4_32_10 (tests/golden/).
// Floating point caveat: this code runs glibc logf/sinf/cosf and an emulation of the GPU's
// linear texture filter, so float results agree with the GPU to ~1e-6 relative, not bitwise.
//
// Each function names the reference lines it follows.  One scalar photon at a time, brute-force
// loop over instances and prims (use_boxes = 1: prim-box pre-test; use_boxes = 2: median-split box trees over instances and
// prims, which only cull and leave every output byte unchanged - that is the form bench.py times as the CPU arm).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct v3 { float x, y, z; };
struct v4 { float x, y, z, w; };
inline v3 mk(float x, float y, float z) { return {x, y, z}; }
inline v3 operator+(v3 a, v3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline v3 operator-(v3 a, v3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline v3 operator-(v3 a) { return {-a.x, -a.y, -a.z}; }
inline v3 operator*(v3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline v3 operator*(float s, v3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline v3 operator/(v3 a, float s) { float inv = 1.0f / s; return a * inv; }          // sysrap/scuda.h:544-548
inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline v3 cross(v3 a, v3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(v3 a) { return sqrtf(dot(a, a)); }
inline v3 normalize(v3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }   // scuda.h:606-610
const float PI_F = 3.14159265358979323846f;
const float RT_MAX = 1.e27f;
const float INF_F = INFINITY;

inline unsigned f2u(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float u2f(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline bool sgn(float f) { return (f2u(f) >> 31) != 0; }

// ---- Philox4_32_10, curand conventions (curand_philox4x32_x.h, curand_kernel.h:885-1040) ----
struct Rng {
    uint32_t ctr[4], key[2], out[4];
    int state;
    static void round_(uint32_t c[4], const uint32_t k[2]) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    void gen() {
        uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
        for (int r = 0; r < 10; r++) { round_(c, k); k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
        memcpy(out, c, 16);
    }
    void incr(uint64_t n) {                     // Philox_State_Incr(s, n)
        uint32_t lo = (uint32_t)n, hi = (uint32_t)(n >> 32);
        ctr[0] += lo; if (ctr[0] < lo) hi++;
        ctr[1] += hi; if (hi <= ctr[1]) return;
        if (++ctr[2]) return;
        ++ctr[3];
    }
    void skipahead(uint64_t n) {
        state += (int)(n & 3); n /= 4;
        if (state > 3) { n += 1; state -= 4; }
        incr(n); gen();
    }
    void init(uint64_t seed, uint64_t subsequence, uint64_t offset) {   // curand_init
        ctr[0] = ctr[1] = ctr[2] = ctr[3] = 0;
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        state = 0;
        uint32_t lo = (uint32_t)subsequence, hi = (uint32_t)(subsequence >> 32);   // skipahead_sequence
        ctr[2] += lo; if (ctr[2] < lo) hi++;
        ctr[3] += hi;
        gen();
        skipahead(offset);
    }
    uint32_t next() {                            // curand(state)
        uint32_t r = out[state++];
        if (state == 4) { incr(1); gen(); state = 0; }
        return r;
    }
    float uniform() { return next() * 2.3283064365386963e-10f + (2.3283064365386963e-10f / 2.0f); }   // curand_uniform
};

// stagr (sysrap/stag.h:231-262, set :333-337): 4-bit tag per random draw + the uniform itself, first 64 draws
struct OTagr {
    uint64_t* tag; float* flat; unsigned slot;
    void add(unsigned t, float f) {
        if (slot < 64u) { tag[slot / 16u] |= (uint64_t)(t & 0xfu) << (4u * (slot % 16u)); flat[slot] = f; }
        slot += 1u;
    }
};
static thread_local OTagr* g_tagr = nullptr;      // set per photon by oracle_simulate when tag / flat outputs were asked for
#define OTAG(t, u) do { if (g_tagr) g_tagr->add((t), (u)); } while (0)

// ---- geometry views --------------------------------------------------------------------------
struct Node { union { float f[16]; unsigned u[16]; int i[16]; }; };
struct Prim { union { float f[16]; unsigned u[16]; int i[16]; }; };
struct Inst { float inv[16]; int solid, identity, is_identity, prim_offset, num_prim; };

// Timing aid only (use_boxes == 2, the bench's CPU arm): median-split box trees over the instances and over the prims of each
// solid (padded boxes, near child first, enclosing volumes in a subtree of their own).  They only cull; the answer is the same as the plain loops' (nearest t, ties to the lower (instance, prim) pair).
struct CpuBvhNode { float bb[6]; int left, right, axis; };    // left < 0 : leaf holding item ~left ; bb already padded ; axis of the split
struct CpuBvh {
    std::vector<CpuBvhNode> nodes;
    int build(std::vector<int>& items, int lo, int hi, const std::vector<float>& boxes) {
        CpuBvhNode nd;
        for (int a = 0; a < 3; a++) { nd.bb[a] = INFINITY; nd.bb[a + 3] = -INFINITY; }
        for (int k = lo; k < hi; k++) for (int a = 0; a < 3; a++) {
            nd.bb[a] = fminf(nd.bb[a], boxes[6 * (size_t)items[k] + a]); nd.bb[a + 3] = fmaxf(nd.bb[a + 3], boxes[6 * (size_t)items[k] + 3 + a]);
        }
        int ax = 0; float ext = -1.f;
        for (int a = 0; a < 3; a++) { float e = nd.bb[a + 3] - nd.bb[a]; if (e > ext) { ext = e; ax = a; } }
        for (int a = 0; a < 3; a++) {          // the padding of box_hit, applied once here
            float pad = 4e-6f * fmaxf(1.f, fmaxf(fabsf(nd.bb[a]), fabsf(nd.bb[a + 3])));
            nd.bb[a] -= pad; nd.bb[a + 3] += pad;
        }
        nd.axis = ax; nd.left = nd.right = 0;
        int me = (int)nodes.size();
        nodes.push_back(nd);
        if (hi - lo == 1) { nodes[me].left = ~items[lo]; nodes[me].right = 0; return me; }
        int mid = (lo + hi) / 2;
        std::nth_element(items.begin() + lo, items.begin() + mid, items.begin() + hi, [&](int x, int y) {
            float cx = boxes[6 * (size_t)x + ax] + boxes[6 * (size_t)x + 3 + ax], cy = boxes[6 * (size_t)y + ax] + boxes[6 * (size_t)y + 3 + ax];
            return cx < cy || (cx == cy && x < y);
        });
        int l = build(items, lo, mid, boxes), r = build(items, mid, hi, boxes);
        nodes[me].left = l; nodes[me].right = r;
        return me;
    }
    // Geant4 geometries nest: a few enclosing volumes (world, mother boxes) span everything and would inflate every node of a
    // median-split tree.  Items much larger than the typical one get a subtree of their own next to the tree of the rest.
    void make(int n, const std::vector<float>& boxes) {
        nodes.clear();
        if (n <= 0) return;
        std::vector<int> items(n);
        for (int i = 0; i < n; i++) items[i] = i;
        nodes.reserve(2 * (size_t)n + 2);
        auto diag2 = [&](int i) { float d = 0.f; for (int a = 0; a < 3; a++) { float e = boxes[6 * (size_t)i + 3 + a] - boxes[6 * (size_t)i + a]; d += e * e; } return d; };
        std::vector<float> dd(n);
        for (int i = 0; i < n; i++) dd[i] = diag2(i);
        std::vector<float> sorted_dd(dd);
        std::nth_element(sorted_dd.begin(), sorted_dd.begin() + n / 2, sorted_dd.end());
        const float cut = 16.f * sorted_dd[n / 2];                       // diagonal more than 4 x the median one
        int nlarge = (int)(std::stable_partition(items.begin(), items.end(), [&](int i) { return dd[i] > cut; }) - items.begin());
        if (nlarge == 0 || nlarge == n || n < 8) { build(items, 0, n, boxes); return; }
        CpuBvhNode root;
        for (int a = 0; a < 3; a++) { root.bb[a] = -INFINITY; root.bb[a + 3] = INFINITY; }
        root.axis = 0; root.left = root.right = 0;
        nodes.push_back(root);
        int l = build(items, nlarge, n, boxes), r = build(items, 0, nlarge, boxes);     // the ordinary items are looked at first
        nodes[0].left = l; nodes[0].right = r;
    }
};

struct Scene {
    const Node* node; const v4* plan; std::vector<float> itra; const Prim* prim;
    std::vector<Inst> inst;
    bool use_boxes;
    bool use_bvh = false;
    CpuBvh tlas;                       // over the world boxes of the instances
    std::vector<CpuBvh> blas;          // per solid, over its prim boxes (items = prim index within the solid)
};

inline v3 right_multiply(const float* m, v3 v, float w) {        // sysrap/sqat4.h:45-52
    v3 r;
    r.x = m[0] * v.x + m[4] * v.y + m[8] * v.z + m[12] * w;
    r.y = m[1] * v.x + m[5] * v.y + m[9] * v.z + m[13] * w;
    r.z = m[2] * v.x + m[6] * v.y + m[10] * v.z + m[14] * w;
    return r;
}
inline v3 left_multiply(const float* m, v3 v, float w) {         // sqat4.h:85-104
    v3 r;
    r.x = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * w;
    r.y = m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * w;
    r.z = m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * w;
    return r;
}

// ---- CSG leaves ------------------------------------------------------------------------------
// CSG/csg_robust_quadratic_roots.h:95-111
void robust_quadratic_roots(float& t1, float& t2, float& disc, float& sdisc, float d, float b, float c) {
    disc = b * b - d * c;
    sdisc = disc > 0.f ? sqrtf(disc) : 0.f;
    float q = b > 0.f ? -(b + sdisc) : -(b - sdisc);
    float root1 = q / d, root2 = c / q;
    t1 = fminf(root1, root2); t2 = fmaxf(root1, root2);
}
// :122-135
void robust_quadratic_roots_disqualifying(float t_min, float& t1, float& t2, float& disc, float& sdisc, float d, float b, float c) {
    disc = b * b - d * c;
    sdisc = disc > 0.f ? sqrtf(disc) : 0.f;
    float q = b > 0.f ? -(b + sdisc) : -(b - sdisc);
    float root1 = sdisc > 0.f ? q / d : t_min;
    float root2 = sdisc > 0.f ? c / q : t_min;
    t1 = fminf(root1, root2); t2 = fmaxf(root1, root2);
}

// CSG/csg_intersect_leaf_sphere.h:15-55
bool intersect_leaf_sphere(v4& isect, const float* q0, float t_min, v3 ro, v3 rd) {
    v3 O = ro - mk(q0[0], q0[1], q0[2]);
    float radius = q0[3];
    float b = dot(O, rd), c = dot(O, O) - radius * radius, d = dot(rd, rd);
    float root1, root2, disc, sdisc;
    robust_quadratic_roots(root1, root2, disc, sdisc, d, b, c);
    float t_cand = sdisc > 0.f ? (root1 > t_min ? root1 : root2) : t_min;
    bool valid = t_cand > t_min;
    if (valid) {
        isect.x = (O.x + t_cand * rd.x) / radius; isect.y = (O.y + t_cand * rd.y) / radius; isect.z = (O.z + t_cand * rd.z) / radius;
        isect.w = t_cand;
    }
    return valid;
}

// CSG/csg_intersect_leaf_zsphere.h:44-151
bool intersect_leaf_zsphere(v4& isect, const float* q0, const float* q1, float t_min, v3 ro, v3 rd) {
    v3 center = mk(q0[0], q0[1], q0[2]);
    v3 O = ro - center;
    float radius = q0[3];
    float b = dot(O, rd), c = dot(O, O) - radius * radius;
    if (c > 0.f && b > 0.f) return false;
    float zmax = center.z + q1[1], zmin = center.z + q1[0];
    float d = dot(rd, rd);
    float t1sph, t2sph, disc, sdisc;
    robust_quadratic_roots(t1sph, t2sph, disc, sdisc, d, b, c);
    float z1sph = ro.z + t1sph * rd.z, z2sph = ro.z + t2sph * rd.z;
    float idz = 1.f / rd.z;
    float t_QCAP = (zmax - ro.z) * idz, t_PCAP = (zmin - ro.z) * idz;
    float t1cap = fminf(t_QCAP, t_PCAP), t2cap = fmaxf(t_QCAP, t_PCAP);
    if (t1cap < t1sph || t1cap > t2sph) t1cap = t_min;
    if (t2cap < t1sph || t2cap > t2sph) t2cap = t_min;
    float t_cand = t_min;
    if (sdisc > 0.f) {
        if (t1sph > t_min && z1sph > zmin && z1sph <= zmax) t_cand = t1sph;
        else if (t1cap > t_min) t_cand = t1cap;
        else if (t2cap > t_min) t_cand = t2cap;
        else if (t2sph > t_min && z2sph > zmin && z2sph <= zmax) t_cand = t2sph;
    }
    bool valid = t_cand > t_min;
    if (valid) {
        isect.w = t_cand;
        if (t_cand == t1sph || t_cand == t2sph) {
            isect.x = (O.x + t_cand * rd.x) / radius; isect.y = (O.y + t_cand * rd.y) / radius; isect.z = (O.z + t_cand * rd.z) / radius;
        } else { isect.x = 0.f; isect.y = 0.f; isect.z = t_cand == t_PCAP ? -1.f : 1.f; }
    }
    return valid;
}

// CSG/csg_intersect_leaf_box3.h:64-155
bool intersect_leaf_box3(v4& isect, const float* q0, float t_min, v3 ro, v3 rd) {
    v3 bmin = mk(-q0[0] / 2.f, -q0[1] / 2.f, -q0[2] / 2.f), bmax = mk(q0[0] / 2.f, q0[1] / 2.f, q0[2] / 2.f);
    v3 idir = mk(1.f / rd.x, 1.f / rd.y, 1.f / rd.z);
    v3 t0 = mk((bmin.x - ro.x) * idir.x, (bmin.y - ro.y) * idir.y, (bmin.z - ro.z) * idir.z);
    v3 t1 = mk((bmax.x - ro.x) * idir.x, (bmax.y - ro.y) * idir.y, (bmax.z - ro.z) * idir.z);
    float t_near = fmaxf(fmaxf(fminf(t0.x, t1.x), fminf(t0.y, t1.y)), fminf(t0.z, t1.z));
    float t_far = fminf(fminf(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y)), fmaxf(t0.z, t1.z));
    bool along_x = rd.x != 0.f && rd.y == 0.f && rd.z == 0.f;
    bool along_y = rd.x == 0.f && rd.y != 0.f && rd.z == 0.f;
    bool along_z = rd.x == 0.f && rd.y == 0.f && rd.z != 0.f;
    bool in_x = ro.x > bmin.x && ro.x < bmax.x, in_y = ro.y > bmin.y && ro.y < bmax.y, in_z = ro.z > bmin.z && ro.z < bmax.z;
    bool has;
    if (along_x) has = in_y && in_z; else if (along_y) has = in_x && in_z; else if (along_z) has = in_x && in_y;
    else has = (t_far > t_near && t_far > 0.f);
    bool valid = false;
    if (has) {
        float t_cand = t_min < t_near ? t_near : (t_min < t_far ? t_far : t_min);
        v3 p = mk(ro.x + t_cand * rd.x, ro.y + t_cand * rd.y, ro.z + t_cand * rd.z);
        v3 pa = mk(fabsf(p.x) / (bmax.x - bmin.x), fabsf(p.y) / (bmax.y - bmin.y), fabsf(p.z) / (bmax.z - bmin.z));
        v3 n = mk(0.f, 0.f, 0.f);
        if (pa.x >= pa.y && pa.x >= pa.z) n.x = copysignf(1.f, p.x);
        else if (pa.y >= pa.x && pa.y >= pa.z) n.y = copysignf(1.f, p.y);
        else if (pa.z >= pa.x && pa.z >= pa.y) n.z = copysignf(1.f, p.z);
        if (t_cand > t_min) { valid = true; isect.x = n.x; isect.y = n.y; isect.z = n.z; isect.w = t_cand; }
    }
    return valid;
}

// CSG/csg_intersect_leaf_cylinder.h:34-83
bool intersect_leaf_cylinder(v4& isect, const float* q0, const float* q1, float t_min, v3 ro, v3 rd) {
    float r = q0[3], z1 = q1[0], z2 = q1[1];
    float ox = ro.x, oy = ro.y, oz = ro.z, vx = rd.x, vy = rd.y, vz = rd.z;
    float r2 = r * r, a = vx * vx + vy * vy, b = ox * vx + oy * vy, c = ox * ox + oy * oy - r2;
    float t_near, t_far, disc, sdisc;
    robust_quadratic_roots_disqualifying(t_min, t_near, t_far, disc, sdisc, a, b, c);
    float z_near = oz + t_near * vz, z_far = oz + t_far * vz;
    float t_z1cap = (z1 - oz) / vz;
    float r2_z1cap = (ox + t_z1cap * vx) * (ox + t_z1cap * vx) + (oy + t_z1cap * vy) * (oy + t_z1cap * vy);
    float t_z2cap = (z2 - oz) / vz;
    float r2_z2cap = (ox + t_z2cap * vx) * (ox + t_z2cap * vx) + (oy + t_z2cap * vy) * (oy + t_z2cap * vy);
    float t_cand = INF_F;
    if (t_near > t_min && z_near > z1 && z_near < z2 && t_near < t_cand) t_cand = t_near;
    if (t_far > t_min && z_far > z1 && z_far < z2 && t_far < t_cand) t_cand = t_far;
    if (t_z1cap > t_min && r2_z1cap <= r2 && t_z1cap < t_cand) t_cand = t_z1cap;
    if (t_z2cap > t_min && r2_z2cap <= r2 && t_z2cap < t_cand) t_cand = t_z2cap;
    bool valid = t_cand > t_min && t_cand < INF_F;
    if (valid) {
        bool sheet = (t_cand == t_near || t_cand == t_far);
        isect.x = sheet ? (ox + t_cand * vx) / r : 0.f;
        isect.y = sheet ? (oy + t_cand * vy) / r : 0.f;
        isect.z = sheet ? 0.f : (t_cand == t_z1cap ? -1.f : 1.f);
        isect.w = t_cand;
    }
    return valid;
}

// CSG/csg_intersect_leaf_newcone.h:50-121
bool intersect_leaf_newcone(v4& isect, const float* q0, float t_min, v3 o, v3 d) {
    float r1 = q0[0], z1 = q0[1], r2 = q0[2], z2 = q0[3];
    float r1r1 = r1 * r1, r2r2 = r2 * r2;
    float tth = (r2 - r1) / (z2 - z1), tth2 = tth * tth;
    float z0 = (z2 * r1 - z1 * r2) / (r1 - r2);
    float idz = 1.f / d.z;
    float t_cap1 = d.z == 0.f ? RT_MAX : (z1 - o.z) * idz;
    float t_cap2 = d.z == 0.f ? RT_MAX : (z2 - o.z) * idz;
    float rr_cap1 = (o.x + t_cap1 * d.x) * (o.x + t_cap1 * d.x) + (o.y + t_cap1 * d.y) * (o.y + t_cap1 * d.y);
    float rr_cap2 = (o.x + t_cap2 * d.x) * (o.x + t_cap2 * d.x) + (o.y + t_cap2 * d.y) * (o.y + t_cap2 * d.y);
    t_cap1 = rr_cap1 < r1r1 && t_cap1 > t_min ? t_cap1 : RT_MAX;
    t_cap2 = rr_cap2 < r2r2 && t_cap2 > t_min ? t_cap2 : RT_MAX;
    float c2 = d.x * d.x + d.y * d.y - d.z * d.z * tth2;
    float c1 = o.x * d.x + o.y * d.y - (o.z - z0) * d.z * tth2;
    float c0 = o.x * o.x + o.y * o.y - (o.z - z0) * (o.z - z0) * tth2;
    float t_near, t_far, disc, sdisc;
    robust_quadratic_roots_disqualifying(RT_MAX, t_near, t_far, disc, sdisc, c2, c1, c0);
    float z_near = o.z + t_near * d.z, z_far = o.z + t_far * d.z;
    t_near = z_near > z1 && z_near < z2 && t_near > t_min ? t_near : RT_MAX;
    t_far = z_far > z1 && z_far < z2 && t_far > t_min ? t_far : RT_MAX;
    float t_cand = fminf(fminf(t_near, t_far), fminf(t_cap1, t_cap2));
    bool valid = t_cand > t_min && t_cand < RT_MAX;
    if (valid) {
        if (t_cand == t_cap1 || t_cand == t_cap2) { isect.x = 0.f; isect.y = 0.f; isect.z = t_cand == t_cap2 ? 1.f : -1.f; }
        else {
            v3 n = normalize(mk(o.x + t_cand * d.x, o.y + t_cand * d.y, (z0 - (o.z + t_cand * d.z)) * tth2));
            isect.x = n.x; isect.y = n.y; isect.z = n.z;
        }
        isect.w = t_cand;
    }
    return valid;
}

// CSG/csg_intersect_leaf_convexpolyhedron.h:20-102
bool intersect_leaf_convexpolyhedron(v4& isect, const float* q0, const v4* plan, float t_min, v3 ro, v3 rd) {
    float t0 = -INF_F, t1 = INF_F;
    v3 t0_normal = mk(0, 0, 0), t1_normal = mk(0, 0, 0);
    unsigned planeIdx = f2u(q0[0]), planeNum = f2u(q0[1]);
    for (unsigned i = 0; i < planeNum; i++) {
        const v4& plane = plan[planeIdx + i];
        v3 n = mk(plane.x, plane.y, plane.z);
        float nd = dot(n, rd), no = dot(n, ro), dist = no - plane.w, t_cand = -dist / nd;
        bool parallel_inside = nd == 0.f && dist < 0.f, parallel_outside = nd == 0.f && dist > 0.f;
        if (parallel_inside) continue;
        if (parallel_outside) return false;
        if (nd < 0.f) { if (t_cand > t0) { t0 = t_cand; t0_normal = n; } }
        else { if (t_cand < t1) { t1 = t_cand; t1_normal = n; } }
    }
    bool valid = t0 < t1;
    if (valid) {
        if (t0 > t_min) { isect.x = t0_normal.x; isect.y = t0_normal.y; isect.z = t0_normal.z; isect.w = t0; }
        else if (t1 > t_min) { isect.x = t1_normal.x; isect.y = t1_normal.y; isect.z = t1_normal.z; isect.w = t1; }
    }
    return valid;
}

// CSG/csg_intersect_leaf_hyperboloid.h
bool intersect_leaf_hyperboloid(v4& isect, const float* q0, float t_min, v3 ro, v3 rd) {
    float r0 = q0[0], zf = q0[1], z1 = q0[2], z2 = q0[3];
    float rr0 = r0 * r0, z1s = z1 / zf, z2s = z2 / zf;
    float rr1 = rr0 * (z1s * z1s + 1.f), rr2 = rr0 * (z2s * z2s + 1.f);
    float A = -rr0 / (zf * zf), B = -rr0;
    float sx = rd.x, sy = rd.y, sz = rd.z, ox = ro.x, oy = ro.y, oz = ro.z;
    float d = sx * sx + sy * sy + A * sz * sz, b = ox * sx + oy * sy + A * oz * sz, c = ox * ox + oy * oy + A * oz * oz + B;
    float t1hyp, t2hyp, disc, sdisc;
    robust_quadratic_roots(t1hyp, t2hyp, disc, sdisc, d, b, c);
    float h1z = oz + t1hyp * sz, h2z = oz + t2hyp * sz;
    float osz = 1.f / sz;
    float t2cap = (z2 - oz) * osz, t1cap = (z1 - oz) * osz;
    v3 c1 = ro + t1cap * rd, c2 = ro + t2cap * rd;
    float crr1 = c1.x * c1.x + c1.y * c1.y, crr2 = c2.x * c2.x + c2.y * c2.y;
    float ca = t1hyp > t_min && disc > 0.f && h1z > z1 && h1z < z2 ? t1hyp : RT_MAX;
    float cb = t2hyp > t_min && disc > 0.f && h2z > z1 && h2z < z2 ? t2hyp : RT_MAX;
    float cc = t2cap > t_min && crr2 < rr2 ? t2cap : RT_MAX;
    float cd = t1cap > t_min && crr1 < rr1 ? t1cap : RT_MAX;
    float t_cand = fminf(fminf(ca, cb), fminf(cc, cd));
    bool valid = t_cand > t_min && t_cand < RT_MAX;
    if (valid) {
        isect.w = t_cand;
        if (t_cand == t1hyp || t_cand == t2hyp) {
            v3 p = ro + t_cand * rd;
            v3 n = normalize(mk(p.x, p.y, A * p.z));
            isect.x = n.x; isect.y = n.y; isect.z = n.z;
        } else { isect.x = 0.f; isect.y = 0.f; isect.z = t_cand == t1cap ? -1.f : 1.f; }
    }
    return valid;
}

// CSG/csg_intersect_leaf_halfspace.h:156-193
bool intersect_leaf_halfspace(v4& isect, const float* q0, float t_min, v3 o, v3 d) {
    v3 n = mk(q0[0], q0[1], q0[2]);
    float w = q0[3];
    float on = dot(o, n), dn = dot(d, n), on_w = on - w, adn = fabsf(dn);
    bool inside = on_w < -1e-9f;
    float t = adn > 0.f ? -on_w / dn : t_min;
    bool valid = t > t_min;
    if (valid) { isect.x = n.x; isect.y = n.y; isect.z = n.z; isect.w = t; }
    else if (inside) isect.y = -0.f;
    return valid;
}

// CSG/csg_intersect_leaf_phicut.h:580-635
bool intersect_leaf_phicut_simple(v4& isect, const float* q0, float t_min, v3 o, v3 d) {
    float cosPhi0 = q0[0], sinPhi0 = q0[1], cosPhi1 = q0[2], sinPhi1 = q0[3];
    float d_n0 = d.x * sinPhi0 + d.y * (-cosPhi0), d_n1 = d.x * (-sinPhi1) + d.y * (cosPhi1);
    float o_n0 = o.x * sinPhi0 + o.y * (-cosPhi0), o_n1 = o.x * (-sinPhi1) + o.y * (cosPhi1);
    float t0 = d_n0 == 0.f ? t_min : -o_n0 / d_n0;
    float t1 = d_n1 == 0.f ? t_min : -o_n1 / d_n1;
    float PR = d_n0 == 0.f ? -o_n0 : -d_n0, QR = d_n1 == 0.f ? -o_n1 : -d_n1;
    float PQ = cosPhi0 * sinPhi1 - cosPhi1 * sinPhi0;
    bool unbounded_exit = PQ >= 0.f ? (PR >= 0.f && QR <= 0.f) : (PR >= 0.f || QR <= 0.f);
    float side0 = o.x * cosPhi0 + o.y * sinPhi0 + (d.x * cosPhi0 + d.y * sinPhi0) * t0;
    float side1 = o.x * cosPhi1 + o.y * sinPhi1 + (d.x * cosPhi1 + d.y * sinPhi1) * t1;
    if (side0 < 0.f) t0 = t_min;
    if (side1 < 0.f) t1 = t_min;
    float t_near = fminf(t0, t1), t_far = fmaxf(t0, t1);
    float t_cand = t_near > t_min ? t_near : (t_far > t_min ? t_far : t_min);
    bool valid = t_cand > t_min;
    if (valid) {
        isect.x = t_cand == t1 ? -sinPhi1 : sinPhi0; isect.y = t_cand == t1 ? cosPhi1 : -cosPhi0; isect.z = 0.f; isect.w = t_cand;
    } else if (unbounded_exit) isect.y = -isect.y;
    return valid;
}

// CSG/csg_intersect_leaf.h:174-324
bool intersect_leaf(v4& isect, const Node* node, const Scene& sc, float t_min, v3 ray_origin, v3 ray_direction) {
    isect = {0.f, 0.f, 0.f, 0.f};
    unsigned typecode = node->u[14], gtransformIdx = node->u[15] & 0x7fffffffu;
    bool complement = (node->u[15] & 0x80000000u) != 0;
    const float* q = gtransformIdx > 0 ? &sc.itra[16 * (gtransformIdx - 1)] : nullptr;
    v3 origin = q ? right_multiply(q, ray_origin, 1.f) : ray_origin;
    v3 direction = q ? right_multiply(q, ray_direction, 0.f) : ray_direction;
    const float* q0 = node->f; const float* q1 = node->f + 4;
    bool valid = false;
    switch (typecode) {
        case 101: valid = intersect_leaf_sphere(isect, q0, t_min, origin, direction); break;
        case 103: valid = intersect_leaf_zsphere(isect, q0, q1, t_min, origin, direction); break;
        case 105: valid = intersect_leaf_cylinder(isect, q0, q1, t_min, origin, direction); break;
        case 110: valid = intersect_leaf_box3(isect, q0, t_min, origin, direction); break;
        case 108: valid = intersect_leaf_newcone(isect, q0, t_min, origin, direction); break;
        case 112: valid = intersect_leaf_convexpolyhedron(isect, q0, sc.plan, t_min, origin, direction); break;
        case 117: valid = intersect_leaf_hyperboloid(isect, q0, t_min, origin, direction); break;
        case 121: valid = intersect_leaf_phicut_simple(isect, q0, t_min, origin, direction); break;
        case 125: valid = intersect_leaf_halfspace(isect, q0, t_min, origin, direction); break;
        default: break;
    }
    if (valid && q) { v3 n = left_multiply(q, mk(isect.x, isect.y, isect.z), 0.f); isect.x = n.x; isect.y = n.y; isect.z = n.z; }
    if (complement) {
        isect.x = valid ? -isect.x : -0.f;
        isect.y = valid ? -isect.y : isect.y;
        isect.z = valid ? -isect.z : isect.z;
    }
    return valid;
}

// CSG/csg_classify.h:55-101
enum { State_Enter = 0, State_Exit = 1, State_Miss = 2 };
enum { CTRL_RETURN_MISS = 0, CTRL_RETURN_A = 1, CTRL_RETURN_B = 2, CTRL_RETURN_FLIP_B = 3, CTRL_LOOP_A = 4, CTRL_LOOP_B = 5 };
inline int CSG_CLASSIFY(const v4& ise, v3 dir, float tmin) {
    return fabsf(ise.w) > tmin ? ((ise.x * dir.x + ise.y * dir.y + ise.z * dir.z < 0.f) ? State_Enter : State_Exit) : State_Miss;
}
int lut_lookup(unsigned op, int stateA, int stateB, bool ACloser) {
    static const unsigned A[4] = {0x22121141, 0x00014014, 0x00141141, 0x00000000};
    static const unsigned B[4] = {0x22115122, 0x00022055, 0x00133155, 0x00000000};
    const unsigned* lut = ACloser ? A : B;
    unsigned offset = 3 * (unsigned)stateA + (unsigned)stateB, index = op - 1u;
    return offset < 8 ? ((lut[index] >> (offset * 4)) & 0xf) : CTRL_RETURN_MISS;
}

// CSG/csg_intersect_node.h:654-686
bool intersect_node_discontiguous(v4& isect, const Node* node, const Node* root, const Scene& sc, float t_min, v3 ro, v3 rd) {
    unsigned num_sub = node->u[0], offset_sub = node->u[1];
    v4 closest = {0, 0, 0, RT_MAX}, sub = {0, 0, 0, 0};
    for (unsigned i = 0; i < num_sub; i++)
        if (intersect_leaf(sub, root + offset_sub + i, sc, t_min, ro, rd)) { if (sub.w < closest.w) closest = sub; }
    bool valid = closest.w < RT_MAX;
    if (valid) isect = closest;
    return valid;
}
// :819-905
bool intersect_node_overlap(v4& isect, const Node* node, const Node* root, const Scene& sc, float t_min, v3 ro, v3 rd) {
    unsigned num_sub = node->u[0], offset_sub = node->u[1];
    v4 farthest_enter = {0, 0, 0, t_min}, nearest_exit = {0, 0, 0, RT_MAX}, sub = {0, 0, 0, 0};
    unsigned enter_count = 0, exit_count = 0;
    for (unsigned i = 0; i < num_sub; i++) {
        const Node* sn = root + offset_sub + i;
        if (intersect_leaf(sub, sn, sc, t_min, ro, rd)) {
            int st = CSG_CLASSIFY(sub, rd, t_min);
            if (st == State_Enter) {
                enter_count++;
                if (sub.w > farthest_enter.w) farthest_enter = sub;
                float tminAdvanced = sub.w + 0.0001f;
                if (intersect_leaf(sub, sn, sc, tminAdvanced, ro, rd)) {
                    if (CSG_CLASSIFY(sub, rd, tminAdvanced) == State_Exit) { exit_count++; if (sub.w < nearest_exit.w) nearest_exit = sub; }
                }
            } else if (st == State_Exit) { exit_count++; if (sub.w < nearest_exit.w) nearest_exit = sub; }
        }
    }
    bool valid = false;
    bool overlap_all = farthest_enter.w < nearest_exit.w && std::max(enter_count, exit_count) == num_sub;
    if (overlap_all) {
        if (farthest_enter.w > t_min && farthest_enter.w < RT_MAX) { valid = true; isect = farthest_enter; }
        else if (nearest_exit.w > t_min && nearest_exit.w < RT_MAX) { valid = true; isect = nearest_exit; }
    }
    return valid;
}
// :392-640
bool intersect_node_contiguous(v4& isect, const Node* node, const Node* root, const Scene& sc, float t_min, v3 ro, v3 rd) {
    int num_sub = (int)node->u[0], offset_sub = (int)node->u[1];
    v4 nearest_enter = {0, 0, 0, RT_MAX}, farthest_exit = {0, 0, 0, t_min}, sub = {0, 0, 0, 0};
    int exit_count = 0;
    for (int i = 0; i < num_sub; i++) {
        if (intersect_leaf(sub, root + offset_sub + i, sc, t_min, ro, rd)) {
            int st = CSG_CLASSIFY(sub, rd, t_min);
            if (st == State_Enter) { if (sub.w < nearest_enter.w) nearest_enter = sub; }
            else if (st == State_Exit) exit_count++;
        }
    }
    if (exit_count == 0) {
        bool valid = nearest_enter.w > t_min && nearest_enter.w < RT_MAX;
        if (valid) isect = nearest_enter;
        return valid;
    }
    int enter_count = 0; float enter[8]; int aux[8]; int idx[8];
    for (int isub = 0; isub < num_sub; isub++) {
        if (intersect_leaf(sub, root + offset_sub + isub, sc, t_min, ro, rd)) {
            int st = CSG_CLASSIFY(sub, rd, t_min);
            if (st == State_Enter) { aux[enter_count] = isub; idx[enter_count] = enter_count; enter[enter_count] = sub.w; enter_count++; }
            else if (st == State_Exit) { exit_count++; if (sub.w > farthest_exit.w) farthest_exit = sub; }
        }
    }
    for (int i = 1; i < enter_count; i++) {
        int key = idx[i], j = i - 1;
        while (j >= 0 && enter[idx[j]] > enter[key]) { idx[j + 1] = idx[j]; j--; }
        idx[j + 1] = key;
    }
    for (int i = 0; i < enter_count; i++) {
        float tminAdvanced = enter[idx[i]] + 0.0001f;
        int isub = aux[idx[i]];
        if (tminAdvanced < farthest_exit.w) {
            if (intersect_leaf(sub, root + offset_sub + isub, sc, tminAdvanced, ro, rd)) {
                if (CSG_CLASSIFY(sub, rd, tminAdvanced) == State_Exit) { exit_count++; if (sub.w > farthest_exit.w) farthest_exit = sub; }
            }
        }
    }
    bool valid = exit_count > 0 && farthest_exit.w > t_min;
    if (valid) isect = farthest_exit;
    return valid;
}
// :913-938
bool intersect_node(v4& isect, const Node* node, const Node* root, const Scene& sc, float t_min, v3 ro, v3 rd) {
    switch (node->u[14]) {
        case 11: return intersect_node_contiguous(isect, node, root, sc, t_min, ro, rd);
        case 13: return intersect_node_overlap(isect, node, root, sc, t_min, ro, rd);
        case 12: return intersect_node_discontiguous(isect, node, root, sc, t_min, ro, rd);
        default: return intersect_leaf(isect, node, sc, t_min, ro, rd);
    }
}

inline int ffs_(unsigned v) { return v ? __builtin_ffs((int)v) : 0; }
inline unsigned POSTORDER_NEXT(unsigned i, unsigned elev) { return (i & 1) ? i >> 1 : (i << elev) + (1u << elev); }   // csg_postorder.h:56-71

// CSG/csg_intersect_tree.h:276-672
bool intersect_tree(v4& isect, const Node* node, const Scene& sc, float t_min, v3 ro, v3 rd) {
    int numNode = (int)node->u[0];
    unsigned height = (unsigned)(ffs_((unsigned)numNode + 1) - 2);
    const float propagate_epsilon = 0.0001f;
    int ierr = 0;
    float tr_tmin[4]; unsigned tr_slice[4]; int tr_curr = -1;
    v4 csg[15]; int curr = -1;
    tr_curr = 0; tr_slice[0] = ((1u << height) & 0xff) << 16; tr_tmin[0] = t_min;
    while (tr_curr > -1) {
        unsigned slice = tr_slice[tr_curr]; float tmin = tr_tmin[tr_curr]; tr_curr--;
        unsigned nodeIdx = (slice >> 16) & 0xff, endIdx = (slice >> 24) & 0xff;
        while (nodeIdx != endIdx) {
            unsigned depth = 31 - __builtin_clz(nodeIdx), elevation = height - depth;
            const Node* nd = node + nodeIdx - 1;
            unsigned typecode = nd->u[14];
            if (typecode == 0) { nodeIdx = POSTORDER_NEXT(nodeIdx, elevation); continue; }
            if (typecode >= 11) {
                v4 nd_isect = {0, 0, 0, 0};
                intersect_node(nd_isect, nd, node, sc, tmin, ro, rd);
                nd_isect.w = copysignf(nd_isect.w, nodeIdx % 2 == 0 ? -1.f : 1.f);
                if (curr >= 14) { ierr = 1; break; }
                csg[++curr] = nd_isect;
            } else {
                if (curr < 1) { ierr = 1; break; }
                bool firstLeft = sgn(csg[curr].w), secondLeft = sgn(csg[curr - 1].w);
                if (!(firstLeft ^ secondLeft)) { ierr = 1; break; }
                int left = firstLeft ? curr : curr - 1, right = firstLeft ? curr - 1 : curr;
                int l_state = CSG_CLASSIFY(csg[left], rd, tmin), r_state = CSG_CLASSIFY(csg[right], rd, tmin);
                float t_left = fabsf(csg[left].w), t_right = fabsf(csg[right].w);
                bool leftIsCloser = t_left <= t_right;
                bool l_promote = l_state == State_Miss && (sgn(csg[left].x) || sgn(csg[left].y));
                bool r_promote = r_state == State_Miss && (sgn(csg[right].x) || sgn(csg[right].y));
                if (r_promote) { r_state = State_Exit; leftIsCloser = true; }
                if (l_promote) { l_state = State_Exit; leftIsCloser = false; }
                int ctrl = lut_lookup(typecode, l_state, r_state, leftIsCloser);
                if (ctrl < CTRL_LOOP_A) {
                    v4 result = ctrl == CTRL_RETURN_MISS ? v4{0, 0, 0, 0} : csg[ctrl == CTRL_RETURN_A ? left : right];
                    if (ctrl == CTRL_RETURN_FLIP_B) { result.x = -result.x; result.y = -result.y; result.z = -result.z; }
                    result.w = copysignf(result.w, nodeIdx % 2 == 0 ? -1.f : 1.f);
                    curr -= 2; csg[++curr] = result;
                } else {
                    int loopside = ctrl == CTRL_LOOP_A ? left : right, otherside = ctrl == CTRL_LOOP_A ? right : left;
                    unsigned leftIdx = 2 * nodeIdx, rightIdx = leftIdx + 1;
                    float tminAdvanced = fabsf(csg[loopside].w) + propagate_epsilon;
                    v4 other = csg[otherside];
                    curr -= 2; csg[++curr] = other;
                    unsigned endTree = ((nodeIdx & 0xff) << 16) | ((endIdx & 0xff) << 24);
                    unsigned leftTree = (((leftIdx << (elevation - 1)) & 0xff) << 16) | (((rightIdx << (elevation - 1)) & 0xff) << 24);
                    unsigned rightTree = (((rightIdx << (elevation - 1)) & 0xff) << 16) | ((nodeIdx & 0xff) << 24);
                    if (tr_curr >= 3) { ierr = 1; break; }
                    tr_curr++; tr_slice[tr_curr] = endTree; tr_tmin[tr_curr] = tmin;
                    if (tr_curr >= 3) { ierr = 1; break; }
                    tr_curr++; tr_slice[tr_curr] = ctrl == CTRL_LOOP_A ? leftTree : rightTree; tr_tmin[tr_curr] = tminAdvanced;
                    break;
                }
            }
            nodeIdx = POSTORDER_NEXT(nodeIdx, elevation);
        }
        if (ierr) break;
    }
    if (curr == 0) isect = csg[0];
    return isect.w > 0.f;
}

// CSG/csg_intersect_tree.h:683-719
bool intersect_prim(v4& isect, const Node* node, const Scene& sc, float t_min, v3 ro, v3 rd) {
    unsigned typecode = node->u[14];
    if (typecode >= 101) return intersect_leaf(isect, node, sc, t_min, ro, rd);
    if (typecode < 11) return intersect_tree(isect, node, sc, t_min, ro, rd);
    if (typecode == 11) return intersect_node_contiguous(isect, node, node, sc, t_min, ro, rd);
    if (typecode == 12) return intersect_node_discontiguous(isect, node, node, sc, t_min, ro, rd);
    if (typecode == 13) return intersect_node_overlap(isect, node, node, sc, t_min, ro, rd);
    return false;
}

// ---- trace : closest hit over all instances and prims (stands in for optixTrace + IS + CH + MS,
// CSGOptiX/CSGOptiX7.cu:110-216, 655-682, 749-847, 869-940) ------------------------------------
struct Prd { v3 normal; float t; float lposcost, lposfphi; unsigned iindex_identity, prim_boundary; };

bool box_hit(const float* bb, v3 o, v3 d, float tmin, float tbest) {
    float tn = tmin, tf = tbest;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    for (int a = 0; a < 3; a++) {
        float inv = 1.f / dd[a];
        float pad = 4e-6f * fmaxf(1.f, fmaxf(fabsf(bb[a]), fabsf(bb[a + 3])));
        float t0 = (bb[a] - pad - oo[a]) * inv, t1 = (bb[a + 3] + pad - oo[a]) * inv;
        if (t0 != t0 || t1 != t1) continue;         // 0*inf : ray lies in the slab plane
        tn = fmaxf(tn, fminf(t0, t1)); tf = fminf(tf, fmaxf(t0, t1));
    }
    return tn <= tf * 1.000001f;
}

// slab test with a precomputed 1/d (same padding and tolerance as box_hit)
inline bool box_hit_inv(const float* bb, const float* oo, const float* inv, float tmin, float tbest) {
    float tn = tmin, tf = tbest;
    for (int a = 0; a < 3; a++) {
        float t0 = (bb[a] - oo[a]) * inv[a], t1 = (bb[a + 3] - oo[a]) * inv[a];          // bb is padded at build time
        if (t0 != t0 || t1 != t1) continue;
        tn = fmaxf(tn, fminf(t0, t1)); tf = fminf(tf, fmaxf(t0, t1));
    }
    return tn <= tf * 1.000001f;
}

struct Best { float t; bool found; v3 n; int inst, prim; unsigned boundary; };

// prims of one instance through the solid's box tree; order independent: ties go to the lower (instance, prim) pair
static void trace_solid_bvh(Best& b, const Scene& sc, int i, v3 oo, v3 dd, float tmin) {
    const Inst& ri = sc.inst[i];
    const CpuBvh& bv = sc.blas[ri.solid];
    if (bv.nodes.empty()) return;
    const float o3[3] = {oo.x, oo.y, oo.z}, inv[3] = {1.f / dd.x, 1.f / dd.y, 1.f / dd.z};
    int stack[128], sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const CpuBvhNode& nd = bv.nodes[stack[--sp]];
        if (!box_hit_inv(nd.bb, o3, inv, tmin, b.t)) continue;
        if (nd.left < 0) {
            int pidx = ri.prim_offset + ~nd.left;
            const Prim& pr = sc.prim[pidx];
            const Node* root = sc.node + pr.i[1];
            v4 isect = {0, 0, 0, 0};
            bool valid = intersect_prim(isect, root, sc, tmin, oo, dd);
            if (valid && isect.w > tmin) {
                bool closer = isect.w < b.t || (isect.w == b.t && (!b.found || i < b.inst || (i == b.inst && pidx < b.prim)));
                if (closer) { b.t = isect.w; b.found = true; b.n = mk(isect.x, isect.y, isect.z); b.inst = i; b.prim = pidx; b.boundary = root->u[6]; }
            }
            continue;
        }
        if (sp + 2 <= 128) {                                   // near child on top: the split axis orders the children along the ray
            const bool left_first = inv[nd.axis] >= 0.f;
            stack[sp++] = left_first ? nd.right : nd.left; stack[sp++] = left_first ? nd.left : nd.right;
        }
    }
}

static void trace_bvh(Best& b, const Scene& sc, v3 o, v3 d, float tmin) {
    if (sc.tlas.nodes.empty()) return;
    const float o3[3] = {o.x, o.y, o.z}, inv[3] = {1.f / d.x, 1.f / d.y, 1.f / d.z};
    int stack[128], sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const CpuBvhNode& nd = sc.tlas.nodes[stack[--sp]];
        if (!box_hit_inv(nd.bb, o3, inv, tmin, b.t)) continue;
        if (nd.left < 0) {
            int i = ~nd.left;
            const Inst& ri = sc.inst[i];
            v3 oo = ri.is_identity ? o : right_multiply(ri.inv, o, 1.f);
            v3 dd = ri.is_identity ? d : right_multiply(ri.inv, d, 0.f);
            trace_solid_bvh(b, sc, i, oo, dd, tmin);
            continue;
        }
        if (sp + 2 <= 128) {
            const bool left_first = inv[nd.axis] >= 0.f;
            stack[sp++] = left_first ? nd.right : nd.left; stack[sp++] = left_first ? nd.left : nd.right;
        }
    }
}

bool trace(Prd& prd, const Scene& sc, v3 o, v3 d, float tmin, float tmax) {
    float best_t = tmax; bool found = false; v3 best_n = mk(0, 0, 0); int best_inst = 0, best_prim = 0; unsigned best_boundary = 0;
    if (sc.use_bvh) {
        Best b; b.t = tmax; b.found = false; b.n = mk(0, 0, 0); b.inst = 0; b.prim = 0; b.boundary = 0;
        trace_bvh(b, sc, o, d, tmin);
        best_t = b.t; found = b.found; best_n = b.n; best_inst = b.inst; best_prim = b.prim; best_boundary = b.boundary;
    } else
    for (size_t i = 0; i < sc.inst.size(); i++) {
        const Inst& ri = sc.inst[i];
        v3 oo = ri.is_identity ? o : right_multiply(ri.inv, o, 1.f);
        v3 dd = ri.is_identity ? d : right_multiply(ri.inv, d, 0.f);
        for (int k = 0; k < ri.num_prim; k++) {
            int pidx = ri.prim_offset + k;
            const Prim& pr = sc.prim[pidx];
            if (sc.use_boxes && !box_hit(pr.f + 8, oo, dd, tmin, best_t)) continue;
            const Node* root = sc.node + pr.i[1];
            v4 isect = {0, 0, 0, 0};
            bool valid = intersect_prim(isect, root, sc, tmin, oo, dd);
            if (valid && isect.w > tmin && (isect.w < best_t || (!found && isect.w == best_t))) {
                best_t = isect.w; found = true; best_n = mk(isect.x, isect.y, isect.z);
                best_inst = (int)i; best_prim = pidx; best_boundary = root->u[6];
            }
        }
    }
    if (!found) { prd.normal = mk(0, 0, 0); prd.t = 1.f; prd.lposcost = 0; prd.lposfphi = 0; prd.iindex_identity = 0xffffffffu; prd.prim_boundary = 0xffffffffu; return false; }
    const Inst& ri = sc.inst[best_inst];
    v3 oo = ri.is_identity ? o : right_multiply(ri.inv, o, 1.f);
    v3 dd = ri.is_identity ? d : right_multiply(ri.inv, d, 0.f);
    v3 n = ri.is_identity ? best_n : left_multiply(ri.inv, best_n, 0.f);
    v3 lpos = oo + best_t * dd;
    prd.normal = n; prd.t = best_t;
    prd.lposcost = lpos.z / sqrtf(dot(lpos, lpos));                                // scuda.h normalize_cost
    prd.lposfphi = (atan2f(lpos.y, lpos.x) + PI_F) / (2.0f * PI_F);                // normalize_fphi
    prd.iindex_identity = (((unsigned)best_inst & 0xffffu) << 16) | ((unsigned)ri.identity & 0xffffu);
    prd.prim_boundary = ((sc.prim[best_prim].u[15] & 0xffffu) << 16) | (best_boundary & 0xffffu);
    return true;
}

// ---- textures : CUDA linear filtering, normalized coordinates, wrap addressing -----------------
// (CUDA C Programming Guide, "Texture Fetching": xB = N*frac(x) - 0.5, i = floor(xB),
//  alpha = frac(xB) held in 9-bit fixed point with 8 fractional bits).  Calibrated against the B200's texture
//  unit (tests/test_parity_gpu.py::test_oracle_texture_emulation_vs_hardware): round-to-nearest of the 8-bit
//  fraction and the lerp form t0 + a (t1 - t0) reproduce 97.4 % of fetches bit for bit; the rest land on the
//  other side of a 1/256 fraction step because the hardware's x*N arithmetic is not public.)
struct Tex { const float* data; int nx, ny, nc; };
inline int wrapi(int i, int n) { i %= n; return i < 0 ? i + n : i; }
inline void tex_coord(float u, int n, int& i0, int& i1, float& a) {
    float s = (u - floorf(u)) * (float)n - 0.5f;
    float fl = floorf(s);
    a = s - fl;
    a = floorf(a * 256.f + 0.5f) / 256.f;
    i0 = wrapi((int)fl, n); i1 = wrapi((int)fl + 1, n);
}
v4 tex2D4(const Tex& t, float x, float y) {
    int i0, i1, j0, j1; float a, b;
    tex_coord(x, t.nx, i0, i1, a); tex_coord(y, t.ny, j0, j1, b);
    v4 r; float* rr = &r.x;
    for (int c = 0; c < 4; c++) {
        float t00 = t.data[((size_t)j0 * t.nx + i0) * 4 + c], t10 = t.data[((size_t)j0 * t.nx + i1) * 4 + c];
        float t01 = t.data[((size_t)j1 * t.nx + i0) * 4 + c], t11 = t.data[((size_t)j1 * t.nx + i1) * 4 + c];
        float r0 = t00 + a * (t10 - t00), r1 = t01 + a * (t11 - t01);      // lerp form: 97.4 % bit-identical with the B200 texture unit
        rr[c] = r0 + b * (r1 - r0);
    }
    return r;
}
float tex2D1(const Tex& t, float x, float y) {
    int i0, i1, j0, j1; float a, b;
    tex_coord(x, t.nx, i0, i1, a); tex_coord(y, t.ny, j0, j1, b);
    float t00 = t.data[(size_t)j0 * t.nx + i0], t10 = t.data[(size_t)j0 * t.nx + i1], t01 = t.data[(size_t)j1 * t.nx + i0], t11 = t.data[(size_t)j1 * t.nx + i1];
    float r0 = t00 + a * (t10 - t00), r1 = t01 + a * (t11 - t01);
    return r0 + b * (r1 - r0);
}

struct Tables { Tex bnd; Tex icdf; const unsigned* optical; float nm0, nms; int hd_factor; };

// qudarap/qbnd.h:103-125
v4 boundary_lookup(const Tables& tb, float nm, unsigned line, unsigned k) {
    float fx = (nm - tb.nm0) / tb.nms;
    float x = (fx + 0.5f) / float(tb.bnd.nx);
    unsigned iy = 2 * line + k;
    float y = (float(iy) + 0.5f) / float(tb.bnd.ny);
    return tex2D4(tb.bnd, x, y);
}
// qudarap/qscint.h:147-228
float scint_wavelength(const Tables& tb, float u0) {
    const float y0 = 0.5f / 3.f, y1 = 1.5f / 3.f, y2 = 2.5f / 3.f;
    switch (tb.hd_factor) {
        case 0: return tex2D1(tb.icdf, u0, y0);
        case 10: if (u0 < 0.1f) return tex2D1(tb.icdf, u0 * 10.f, y1); else if (u0 > 0.9f) return tex2D1(tb.icdf, (u0 - 0.9f) * 10.f, y2); else return tex2D1(tb.icdf, u0, y0);
        case 20: if (u0 < 0.05f) return tex2D1(tb.icdf, u0 * 20.f, y1); else if (u0 > 0.95f) return tex2D1(tb.icdf, (u0 - 0.95f) * 20.f, y2); else return tex2D1(tb.icdf, u0, y0);
    }
    return 0.f;
}

// ---- photon --------------------------------------------------------------------------------------
struct Photon {                    // sysrap/sphoton.h:171-193
    v3 pos; float time; v3 mom; unsigned hitcount_iindex; v3 pol; float wavelength;
    unsigned orient_boundary_flag, identity, index, flagmask;
    void zero_flags() { orient_boundary_flag = 0; identity = 0; index = 0; flagmask = 0; hitcount_iindex = 0; }
    void set_flag(unsigned f) { orient_boundary_flag = (orient_boundary_flag & 0xffff0000u) | (f & 0xffffu); flagmask |= f; }
    unsigned flag() const { return orient_boundary_flag & 0xffffu; }
    unsigned boundary() const { return (orient_boundary_flag & 0x7fff0000u) >> 16; }
    void set_index(uint64_t full) { index = (unsigned)(full & 0xffffffffu); identity = ((unsigned)((full >> 32) & 0xffu) << 24) | (identity & 0xffffffu); }
    void set_prd(unsigned b, unsigned id, float orient, unsigned ii) {   // sphoton.h:482-488
        orient_boundary_flag = (orient_boundary_flag & 0x8000ffffu) | ((b & 0x7fffu) << 16);
        identity = (identity & 0xff000000u) | (id & 0x00ffffffu);
        orient_boundary_flag = (orient_boundary_flag & 0x7fffffffu) | ((orient < 0.f ? 1u : 0u) << 31);
        hitcount_iindex = 0x00010000u | (ii & 0xffffu);
    }
};
static_assert(sizeof(Photon) == 64, "sphoton");
struct Seq { uint64_t seqhis[2], seqbnd[2]; };
void seq_add_nibble(Seq& s, unsigned slot, unsigned flag, unsigned boundary) {       // sysrap/sseq.h:174-184
    unsigned iseq = slot / 16, shift = 4 * (slot - iseq * 16);
    if (iseq < 2) { s.seqhis[iseq] |= ((uint64_t)(ffs_(flag) & 0xf)) << shift; s.seqbnd[iseq] |= ((uint64_t)(boundary & 0xf)) << shift; }
}

enum { CERENKOV = 1, SCINTILLATION = 2, TORCH = 4, BULK_ABSORB = 8, BULK_REEMIT = 16, BULK_SCATTER = 32, SURFACE_DETECT = 64, SURFACE_ABSORB = 128,
       SURFACE_DREFLECT = 256, SURFACE_SREFLECT = 512, BOUNDARY_REFLECT = 1024, BOUNDARY_TRANSMIT = 2048 };
enum { BREAK = 1, CONTINUE = 2, BOUNDARY = 3 };

// sysrap/smath.h:77-95
void rotateUz(v3& d, v3 u) {
    float up = u.x * u.x + u.y * u.y;
    if (up > 0.f) {
        up = sqrtf(up);
        float px = d.x, py = d.y, pz = d.z;
        d.x = (u.x * u.z * px - u.y * py) / up + u.x * pz;
        d.y = (u.y * u.z * px + u.x * py) / up + u.y * pz;
        d.z = -up * px + u.z * pz;
    } else if (u.z < 0.f) { d.x = -d.x; d.z = -d.z; }
}

union GS { float f[24]; unsigned u[24]; int i[24]; };

// sysrap/storch.h:189-516
void storch_generate(Photon& p, Rng& rng, const GS& gs, uint64_t photon_id) {
    const float* f = gs.f;
    unsigned numphoton = gs.u[3], type = gs.u[23];
    v3 gpos = mk(f[4], f[5], f[6]), gmom = mk(f[8], f[9], f[10]);
    float zx = f[16], zy = f[17], ax = f[18], ay = f[19], radius = f[20], distance = f[21];
    p.wavelength = f[15]; p.time = f[7];
    if (type == 1) {            // T_DISC
        p.mom = gmom;
        float u_zenith = zx + rng.uniform() * (zy - zx), u_azimuth = ax + rng.uniform() * (ay - ax);
        float r = radius * u_zenith, phi = 2.f * PI_F * u_azimuth, sinPhi = sinf(phi), cosPhi = cosf(phi);
        p.pos = mk(r * cosPhi, r * sinPhi, 0.f); rotateUz(p.pos, p.mom); p.pos = p.pos + gpos;
        p.pol = mk(sinPhi, -cosPhi, 0.f); rotateUz(p.pol, p.mom);
    } else if (type == 7) {     // T_SPHERE
        float u_zenith = zx + rng.uniform() * (zy - zx), u_azimuth = ax + rng.uniform() * (ay - ax);
        float phi = 2.f * PI_F * u_azimuth, sinPhi = sinf(phi), cosPhi = cosf(phi);
        float cosTheta = 1.f - 2.0f * u_zenith, sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
        float flip = copysignf(1.f, radius);
        p.mom = mk(flip * sinTheta * cosPhi, flip * sinTheta * sinPhi, flip * cosTheta);
        float ar = fabsf(radius);
        p.pos = mk(sinTheta * cosPhi * ar, sinTheta * sinPhi * ar, cosTheta * ar);
        float phase = 2.f * PI_F * distance;
        p.pol = mk(cosf(phase), sinf(phase), 0.f); rotateUz(p.pol, p.mom);
    } else if (type == 6) {     // T_SPHERE_MARSAGLIA
        float u, v, b, a;
        do { float u0 = zx + rng.uniform() * (zy - zx), u1 = ax + rng.uniform() * (ay - ax); u = 2.f * u0 - 1.f; v = 2.f * u1 - 1.f; b = u * u + v * v; } while (b > 1.f);
        a = 2.f * sqrtf(1.f - b);
        float ar = fabsf(radius), flip = copysignf(1.f, radius);
        p.mom = mk(flip * a * u, flip * a * v, flip * (2.f * b - 1.f));
        p.pos = mk(a * u * ar, a * v * ar, (2.f * b - 1.f) * ar);
        float phase = 2.f * PI_F * distance;
        p.pol = mk(cosf(phase), sinf(phase), 0.f); rotateUz(p.pol, p.mom);
    } else if (type == 2) {     // T_LINE
        p.mom = gmom;
        float frac = float(photon_id) / float(numphoton), sfrac = 2.f * (frac - 0.5f), r = radius * sfrac;
        p.pos = mk(r, 0.f, 0.f); rotateUz(p.pos, p.mom); p.pos = p.pos + gpos;
        p.pol = mk(0.f, -1.f, 0.f); rotateUz(p.pol, p.mom);
    } else if (type == 3) {     // T_POINT
        p.mom = gmom;
        p.pos = mk(0, 0, 0); rotateUz(p.pos, p.mom); p.pos = p.pos + gpos;
        p.pol = mk(0.f, -1.f, 0.f); rotateUz(p.pol, p.mom);
    } else if (type == 4) {     // T_CIRCLE
        float ff = float(photon_id) / float(numphoton), frac = ax * (1.f - ff) + ay * ff;
        float phi = 2.f * PI_F * frac, sinPhi = sinf(phi), cosPhi = cosf(phi);
        float r = radius < 0.f ? -radius : radius;
        p.mom = mk(radius < 0.f ? -cosPhi : cosPhi, 0.f, radius < 0.f ? -sinPhi : sinPhi);
        p.pos = mk(r * cosPhi, 0.f, r * sinPhi); p.pos = p.pos + gpos;
        p.pol = mk(0.f, -1.f, 0.f); rotateUz(p.pol, p.mom);
    } else if (type == 5) {     // T_RECTANGLE
        int side_size = (int)(numphoton / 4), side = (int)(photon_id / (uint64_t)side_size), side_offset = side * side_size;
        int side_index = (int)photon_id - side_offset;
        float frac = float(side_index) / float(side_size);
        if (side == 0 || side == 1) { p.pos = mk(side == 0 ? ax : ay, 0.f, (1.f - frac) * zx + frac * zy); p.mom = mk(side == 0 ? 1.f : -1.f, 0.f, 0.f); }
        else if (side == 2 || side == 3) { p.pos = mk((1.f - frac) * ax + frac * ay, 0.f, side == 2 ? zx : zy); p.mom = mk(0.f, 0.f, side == 2 ? 1.f : -1.f); }
        p.pos = p.pos + gpos;
        p.pol = mk(0.f, -1.f, 0.f); rotateUz(p.pol, p.mom);
    }
    p.zero_flags(); p.set_flag(TORCH);
}

// qudarap/qcerenkov.h:56-119, 139-163, 285-327
void cerenkov_generate(Photon& p, Rng& rng, const GS& gs, const Tables& tb) {
    const float* f = gs.f;
    unsigned matline = gs.u[2];
    v3 gpos = mk(f[4], f[5], f[6]), DeltaPosition = mk(f[8], f[9], f[10]);
    float gtime = f[7], step_length = f[11], preVelocity = f[15], BetaInverse = f[16], Wmin = f[17], Wmax = f[18], maxSin2 = f[20];
    float Mean1 = f[21], Mean2 = f[22], postVelocity = f[23];
    v3 p0 = normalize(DeltaPosition);
    float wavelength, cosTheta, sin2Theta, u_maxSin2; unsigned count = 0;
    do {
        float u0 = rng.uniform();
        float w = Wmin + u0 * (Wmax - Wmin);
        wavelength = Wmin * Wmax / w;
        float sampledRI = boundary_lookup(tb, wavelength, matline, 0).x;
        cosTheta = BetaInverse / sampledRI;
        sin2Theta = fmaxf(0.f, (1.f - cosTheta) * (1.f + cosTheta));
        float u1 = rng.uniform();
        u_maxSin2 = u1 * maxSin2;
        count += 1;
    } while (u_maxSin2 > sin2Theta && count < 100);
    float sinTheta = sqrtf(sin2Theta);
    float u0 = rng.uniform(), phi = 2.f * PI_F * u0, sinPhi = sinf(phi), cosPhi = cosf(phi);
    p.mom = mk(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta); rotateUz(p.mom, p0);
    p.pol = mk(cosTheta * cosPhi, cosTheta * sinPhi, -sinTheta); rotateUz(p.pol, p0);
    p.wavelength = wavelength;
    float fraction, delta, NumberOfPhotons, N;
    float MeanMax = fmaxf(Mean1, Mean2), DeltaN = Mean1 - Mean2;
    do { fraction = rng.uniform(); delta = fraction * step_length; NumberOfPhotons = Mean1 - fraction * DeltaN; float u = rng.uniform(); N = u * MeanMax; } while (N > NumberOfPhotons);
    float midVelocity = preVelocity + fraction * (postVelocity - preVelocity) * 0.5f;
    p.time = gtime + delta / midVelocity;
    p.pos = gpos + fraction * DeltaPosition;
    p.zero_flags(); p.set_flag(CERENKOV);
}

// qudarap/qscint.h:55-75, 117-145
void scint_generate(Photon& p, Rng& rng, const GS& gs, const Tables& tb) {
    const float* f = gs.f;
    float u0 = rng.uniform(), u1 = rng.uniform(), u2 = rng.uniform(), u3 = rng.uniform();
    float cost = 1.f - 2.f * u0, sint = sqrtf((1.f - cost) * (1.f + cost));
    float phi = 2.f * PI_F * u1, sinp = sinf(phi), cosp = cosf(phi);
    p.mom = mk(sint * cosp, sint * sinp, cost);
    p.pol = mk(cost * cosp, cost * sinp, -sint);
    phi = 2.f * PI_F * u2; sinp = sinf(phi); cosp = cosf(phi);
    p.pol = normalize(cosp * p.pol + sinp * cross(p.mom, p.pol));
    p.wavelength = scint_wavelength(tb, u3);
    float charge = f[13];
    float fraction = charge == 0.f ? 1.f : rng.uniform();
    p.pos = mk(f[4], f[5], f[6]) + fraction * mk(f[8], f[9], f[10]);
    float u4 = rng.uniform();
    float deltaTime = fraction * f[11] / f[15] - f[20] * logf(u4);
    p.time = f[7] + deltaTime;
    p.zero_flags(); p.set_flag(SCINTILLATION);
}

// qudarap/qsim.h:2521-2541
void generate_photon(Photon& p, Rng& rng, const GS& gs, const Tables& tb, const Photon* input, uint64_t input_base, uint64_t photon_id) {
    switch (gs.i[0]) {
        case 14: {   // CARRIER sysrap/scarrier.h:47-58
            memcpy(&p, gs.f + 8, 64); p.pos.y += float(photon_id) * 10.f; p.set_flag(TORCH); break; }
        case 6: storch_generate(p, rng, gs, photon_id); break;
        case 18: case 15: cerenkov_generate(p, rng, gs, tb); break;
        case 5: case 16: scint_generate(p, rng, gs, tb); break;
        case 19: p = input[photon_id - input_base]; p.set_flag(TORCH); break;
        default: { unsigned* q = (unsigned*)&p; for (int k = 0; k < 16; k++) q[k] = (k % 4) + 1; p.set_flag(TORCH); break; }
    }
    p.set_index(photon_id);
}

v3 uniform_sphere(float u0, float u1) {            // qsim.h uniform_sphere
    float phi = u0 * 2.f * PI_F, cosTheta = 2.f * u1 - 1.f, sinTheta = sqrtf(1.f - cosTheta * cosTheta);
    return mk(cosf(phi) * sinTheta, sinf(phi) * sinTheta, cosTheta);
}
void random_direction_marsaglia(v3& dir, Rng& rng) {   // qsim.h:551-572
    float u, v, b;
    do { float u0 = rng.uniform(), u1 = rng.uniform(); u = 2.f * u0 - 1.f; v = 2.f * u1 - 1.f; b = u * u + v * v; } while (b > 1.f);
    float a = 2.f * sqrtf(1.f - b);
    dir = mk(a * u, a * v, 2.f * b - 1.f);
}
void rayleigh_scatter(Photon& p, Rng& rng) {           // qsim.h:601-689
    v3 direction, polarization; bool looping = true;
    do {
        float u0 = rng.uniform(), u1 = rng.uniform(), u2 = rng.uniform(), u3 = rng.uniform(), u4 = rng.uniform();
        OTAG(8, u0); OTAG(8, u1); OTAG(8, u2); OTAG(8, u3); OTAG(8, u4);                 // stag_sc (qsim.h:618-622)
        float cosTheta = u0, sinTheta = sqrtf(1.0f - u0 * u0);
        if (u1 < 0.5f) cosTheta = -cosTheta;
        float ang = 2.f * PI_F * u2, sinPhi = sinf(ang), cosPhi = cosf(ang);
        direction = mk(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
        rotateUz(direction, p.mom);
        float constant = -dot(direction, p.pol);
        polarization = mk(p.pol.x + constant * direction.x, p.pol.y + constant * direction.y, p.pol.z + constant * direction.z);
        if (dot(polarization, polarization) == 0.f) {
            ang = 2.f * PI_F * u3; sinPhi = sinf(ang); cosPhi = cosf(ang);
            polarization = mk(cosPhi, sinPhi, 0.f); rotateUz(polarization, direction);
        } else if (u3 < 0.5f) polarization = -polarization;
        polarization = normalize(polarization);
        float doCosTheta = dot(polarization, p.pol);
        looping = doCosTheta * doCosTheta < u4;
    } while (looping);
    p.mom = direction; p.pol = polarization;
}

struct State { v4 material1, m1group2, material2, surface; unsigned optical[4]; };   // sysrap/sstate.h:25-46

// qudarap/qsim.h:2218-2327 with :718-863, :998-1205, :1677-1755, :1977-2082 inlined as blocks
int propagate(Photon& p, Rng& rng, const Prd& prd, const Tables& tb, bool debug_tag) {
    unsigned boundary = prd.prim_boundary & 0xffffu, identity = prd.iindex_identity & 0xffffu, iindex = prd.iindex_identity >> 16;
    v3 normal = prd.normal;
    float cosTheta = dot(p.mom, normal);
    p.set_prd(boundary, identity, cosTheta, iindex);
    // qbnd::fill_state qudarap/qbnd.h:184-214
    State s;
    int line = boundary * 4;
    int m1_line = cosTheta > 0.f ? line + 3 : line + 0, m2_line = cosTheta > 0.f ? line + 0 : line + 3, su_line = cosTheta > 0.f ? line + 2 : line + 1;
    s.material1 = boundary_lookup(tb, p.wavelength, m1_line, 0);
    s.m1group2 = boundary_lookup(tb, p.wavelength, m1_line, 1);
    s.material2 = boundary_lookup(tb, p.wavelength, m2_line, 0);
    s.surface = boundary_lookup(tb, p.wavelength, su_line, 0);
    memcpy(s.optical, tb.optical + 4 * su_line, 16);

    unsigned flag = 0; int command;
    {   // propagate_to_boundary
        float absorption_length = s.material1.y, scattering_length = s.material1.z, reemission_prob = s.material1.w, group_velocity = s.m1group2.x;
        float distance_to_boundary = prd.t;
        if (debug_tag) { float u_to_sci = rng.uniform(), u_to_bnd = rng.uniform(); OTAG(1, u_to_sci); OTAG(2, u_to_bnd); }
        float u_scattering = rng.uniform(), u_absorption = rng.uniform();
        OTAG(3, u_scattering); OTAG(4, u_absorption);                                    // qsim.h:739-742
        float scattering_distance = -scattering_length * logf(u_scattering), absorption_distance = -absorption_length * logf(u_absorption);
        command = BOUNDARY;
        if (absorption_distance <= scattering_distance) {
            if (absorption_distance <= distance_to_boundary) {
                p.time += absorption_distance / group_velocity;
                p.pos = p.pos + absorption_distance * p.mom;
                float u_reemit = reemission_prob == 0.f ? 2.f : rng.uniform();
                if (u_reemit != 2.f) OTAG(9, u_reemit);
                if (u_reemit < reemission_prob) {
                    float u_re_wavelength = rng.uniform(), u_re_mom_ph = rng.uniform(), u_re_mom_ct = rng.uniform(), u_re_pol_ph = rng.uniform(), u_re_pol_ct = rng.uniform();
                    OTAG(10, u_re_wavelength); OTAG(11, u_re_mom_ph); OTAG(12, u_re_mom_ct); OTAG(13, u_re_pol_ph); OTAG(14, u_re_pol_ct);
                    p.wavelength = scint_wavelength(tb, u_re_wavelength);
                    p.mom = uniform_sphere(u_re_mom_ph, u_re_mom_ct);
                    p.pol = normalize(cross(uniform_sphere(u_re_pol_ph, u_re_pol_ct), p.mom));
                    flag = BULK_REEMIT; command = CONTINUE;
                } else { flag = BULK_ABSORB; command = BREAK; }
            }
        } else if (scattering_distance <= distance_to_boundary) {
            p.time += scattering_distance / group_velocity;
            p.pos = p.pos + scattering_distance * p.mom;
            rayleigh_scatter(p, rng);
            flag = BULK_SCATTER; command = CONTINUE;
        }
        if (command == BOUNDARY) { p.pos = p.pos + distance_to_boundary * p.mom; p.time += distance_to_boundary / group_velocity; }
    }
    if (command == BOUNDARY) {
        int ems = (int)s.optical[1];
        bool at_surface = false;
        if (ems == 1) {   // propagate_at_boundary
            float n1 = s.material1.x, n2 = s.material2.x, eta = n1 / n2;
            float _c1 = -dot(p.mom, normal);
            v3 oriented_normal = _c1 < 0.f ? -normal : normal;
            v3 trans = cross(p.mom, oriented_normal);
            float trans_length = length(trans);
            bool normal_incidence = trans_length < 1e-6f;
            v3 A_trans = normal_incidence ? p.pol : trans / trans_length;
            float E1_perp = dot(p.pol, A_trans);
            float c1 = fabsf(_c1);
            float c2c2 = 1.f - eta * eta * (1.f - c1 * c1);
            bool tir = c2c2 < 0.f;
            float EdotN = dot(p.pol, oriented_normal);
            float c2 = tir ? 0.f : sqrtf(c2c2);
            float n1c1 = n1 * c1, n2c2 = n2 * c2, n2c1 = n2 * c1, n1c2 = n1 * c2;
            float E1x = normal_incidence ? 0.f : E1_perp, E1y = normal_incidence ? 1.f : length(p.pol - (E1_perp * A_trans));
            float E2tx = 2.f * n1c1 * E1x / (n1c1 + n2c2), E2ty = 2.f * n1c1 * E1y / (n2c1 + n1c2);
            float E2rx = E2tx - E1x, E2ry = (n2 * E2ty / n1) - E1y;
            float rinv = 1.0f / sqrtf(E2rx * E2rx + E2ry * E2ry), tinv = 1.0f / sqrtf(E2tx * E2tx + E2ty * E2ty);
            float RRx = E2rx * rinv, RRy = E2ry * rinv, TTx = E2tx * tinv, TTy = E2ty * tinv;
            float TransCoeff = (tir || n1c1 == 0.f) ? 0.f : n2c2 * (E2tx * E2tx + E2ty * E2ty) / n1c1;
            if (debug_tag) { float u_boundary_burn = rng.uniform(); OTAG(5, u_boundary_burn); }
            float u_reflect = rng.uniform();
            OTAG(6, u_reflect);
            bool reflect = u_reflect > TransCoeff;
            p.mom = reflect ? p.mom + 2.0f * c1 * oriented_normal : eta * p.mom + (eta * c1 - c2) * oriented_normal;
            v3 A_paral = normalize(cross(p.mom, A_trans));
            p.pol = normal_incidence ? (reflect ? p.pol * (n2 > n1 ? -1.f : 1.f) : p.pol)
                                     : (reflect ? (tir ? -p.pol + 2.f * EdotN * oriented_normal : RRx * A_trans + RRy * A_paral) : TTx * A_trans + TTy * A_paral);
            flag = reflect ? BOUNDARY_REFLECT : BOUNDARY_TRANSMIT;
            if (debug_tag && reflect) { float a0 = rng.uniform(), a1 = rng.uniform(), a2 = rng.uniform(), a3 = rng.uniform(); OTAG(1, a0); OTAG(2, a1); OTAG(3, a2); OTAG(4, a3); }
            command = CONTINUE;
        } else if (ems == 2) at_surface = true;
        else if (prd.lposcost < 0.f) at_surface = true;
        else if (ems == 3) { rng.uniform(); flag = SURFACE_DETECT; command = BREAK; }
        if (at_surface) {   // propagate_at_surface
            float detect = s.surface.x, absorb = s.surface.y, reflect_diffuse_ = s.surface.w;
            float u_surface = rng.uniform();
            OTAG(5, u_surface);
            if (debug_tag) { float u_surface_burn = rng.uniform(); OTAG(7, u_surface_burn); }
            command = u_surface < absorb + detect ? BREAK : CONTINUE;
            if (command == BREAK) flag = u_surface < absorb ? SURFACE_ABSORB : SURFACE_DETECT;
            else {
                flag = u_surface < absorb + detect + reflect_diffuse_ ? SURFACE_DREFLECT : SURFACE_SREFLECT;
                if (flag == SURFACE_DREFLECT) {          // reflect_diffuse + lambertian_direction
                    v3 old_mom = p.mom;
                    float orient = dot(old_mom, normal) > 0.f ? -1.f : 1.f;
                    float ndotv, u; int count = 0;
                    do {
                        count++;
                        random_direction_marsaglia(p.mom, rng);
                        ndotv = dot(p.mom, normal) * orient;
                        if (ndotv < 0.f) { p.mom = -1.f * p.mom; ndotv = -1.f * ndotv; }
                        u = rng.uniform();
                    } while (!(u < ndotv) && (count < 1024));
                    v3 facet_normal = normalize(p.mom - old_mom);
                    float EdotN = dot(p.pol, facet_normal);
                    p.pol = -1.f * p.pol + 2.f * EdotN * facet_normal;
                } else {                                 // reflect_specular
                    float PdotN = dot(p.mom, normal);
                    p.mom = p.mom - 2.f * PdotN * normal;
                    float EdotN = dot(p.pol, normal);
                    p.pol = -1.f * p.pol + 2.f * EdotN * normal;
                }
            }
        }
    }
    p.set_flag(flag);
    return command;
}

struct Config {
    int max_bounce, max_record, event_index, debug_tag;
    float tmin, tmin0, tmax, max_time;
    unsigned eps0mask, hit_mask;
    uint64_t seed, offset, skipahead, photon_offset;
    int use_boxes, nthreads;
    unsigned refine; float refine_distance;          // PropagateRefine / PropagateRefineDistance (CSGOptiX/CSGOptiX7.cu:454-458)
};

bool invert_affine(const float* m, float* out) {
    double a[3][3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) a[r][c] = m[4 * r + c];
    double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) + a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    if (det == 0.0) return false;
    double id = 1.0 / det, b[3][3];
    b[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) * id; b[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id; b[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
    b[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) * id; b[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id; b[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
    b[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) * id; b[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id; b[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
    double t[3] = {m[12], m[13], m[14]};
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) out[4 * r + c] = (float)b[r][c];
    for (int c = 0; c < 3; c++) out[12 + c] = (float)(-(t[0] * b[0][c] + t[1] * b[1][c] + t[2] * b[2][c]));
    out[3] = out[7] = out[11] = 0.f; out[15] = 1.f;
    return true;
}

int make_scene(Scene& sc, const int* solid, int nsolid, const void* prim, const void* node, const void* plan, const float* itra, int nitra,
               const float* inst, int ninst, int use_boxes) {
    sc.node = (const Node*)node; sc.prim = (const Prim*)prim; sc.plan = (const v4*)plan; sc.use_boxes = use_boxes != 0;
    sc.use_bvh = false;
    sc.itra.assign(itra, itra + (size_t)nitra * 16);
    for (int i = 0; i < nitra; i++) { sc.itra[16 * i + 3] = 0.f; sc.itra[16 * i + 7] = 0.f; sc.itra[16 * i + 11] = 0.f; sc.itra[16 * i + 15] = 1.f; }
    if (ninst <= 0 || !inst) {
        Inst r; memset(&r, 0, sizeof(r));
        r.inv[0] = r.inv[5] = r.inv[10] = r.inv[15] = 1.f; r.is_identity = 1; r.num_prim = solid[4]; r.prim_offset = solid[5];
        sc.inst.push_back(r);
        return 0;
    }
    const int* insti = (const int*)inst;
    for (int i = 0; i < ninst; i++) {
        Inst r; memset(&r, 0, sizeof(r));
        float m[16]; memcpy(m, inst + 16 * i, 64);
        int gas = insti[16 * i + 7];
        if (gas < 0 || gas >= nsolid) return -1;
        r.identity = insti[16 * i + 11];
        m[3] = m[7] = m[11] = 0.f; m[15] = 1.f;
        static const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        r.is_identity = memcmp(m, ident, 64) == 0;
        if (!invert_affine(m, r.inv)) return -1;
        r.solid = gas; r.num_prim = solid[12 * gas + 4]; r.prim_offset = solid[12 * gas + 5];
        sc.inst.push_back(r);
    }
    if (use_boxes == 2) {              // box trees for the timed CPU arm
        sc.blas.assign(nsolid, CpuBvh());
        std::vector<float> sbox((size_t)nsolid * 6);
        for (int s_ = 0; s_ < nsolid; s_++) {
            int np = solid[12 * s_ + 4], po = solid[12 * s_ + 5];
            std::vector<float> boxes((size_t)np * 6);
            for (int a = 0; a < 3; a++) { sbox[6 * s_ + a] = INFINITY; sbox[6 * s_ + 3 + a] = -INFINITY; }
            for (int k = 0; k < np; k++) for (int a = 0; a < 6; a++) {
                float v = sc.prim[po + k].f[8 + a];
                boxes[6 * (size_t)k + a] = v;
                if (a < 3) sbox[6 * s_ + a] = fminf(sbox[6 * s_ + a], v); else sbox[6 * s_ + a] = fmaxf(sbox[6 * s_ + a], v);
            }
            sc.blas[s_].make(np, boxes);
        }
        std::vector<float> ibox((size_t)ninst * 6);
        for (int i = 0; i < ninst; i++) {
            float m[16]; memcpy(m, inst + 16 * i, 64);
            m[3] = m[7] = m[11] = 0.f; m[15] = 1.f;
            const float* sb = &sbox[6 * (size_t)sc.inst[i].solid];
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int c = 0; c < 8; c++) {
                float v[3] = {(c & 1) ? sb[3] : sb[0], (c & 2) ? sb[4] : sb[1], (c & 4) ? sb[5] : sb[2]};
                for (int a = 0; a < 3; a++) {
                    float w = m[a] * v[0] + m[4 + a] * v[1] + m[8 + a] * v[2] + m[12 + a];
                    lo[a] = fminf(lo[a], w); hi[a] = fmaxf(hi[a], w);
                }
            }
            for (int a = 0; a < 3; a++) {
                float pad = 1e-4f * fmaxf(1.f, fmaxf(fabsf(lo[a]), fabsf(hi[a])));
                ibox[6 * (size_t)i + a] = lo[a] - pad; ibox[6 * (size_t)i + 3 + a] = hi[a] + pad;
            }
        }
        sc.tlas.make(ninst, ibox);
        sc.use_bvh = true;
    }
    return 0;
}

}  // namespace

// sphotonlite::set_lpos (sysrap/sphotonlite.h:234-245); the u16 conversion saturates like the device's cvt.rzi.u16.f32
static unsigned pack_lpos(float lposcost, float lposfphi) {
    auto u16 = [](float v) { float x = v * 65535.f + 0.5f; x = x < 0.f ? 0.f : (x > 65535.f ? 65535.f : x); return (unsigned)x; };
    return (u16(lposcost) << 16) | u16(lposfphi);
}
static unsigned* g_lite_out = nullptr;      // optional sphotonlite[n] output of the next oracle_simulate call
static uint64_t* g_tag_out = nullptr;       // optional stag[n] (4 u64) / sflat[n] (64 f32) outputs, zeroed by the caller
static float* g_flat_out = nullptr;

extern "C" {

void oracle_set_lite_out(unsigned* p) { g_lite_out = p; }
void oracle_set_tag_out(uint64_t* tag, float* flat) { g_tag_out = tag; g_flat_out = flat; }

// CSGOptiX/CSGOptiX7.cu:405-503 per photon; seeding = QEvt.cu:181-237 (seed[i] = owning genstep)
int oracle_simulate(const int* solid, int nsolid, const void* prim, int nprim, const void* node, int nnode, const void* plan, int nplan,
                    const float* itra, int nitra, const float* inst, int ninst,
                    const float* bnd, int nbnd, int nwl, float dom_low, float dom_step, const int* optical,
                    const float* icdf, int icdf_nx, int hd_factor,
                    const void* genstep, int ngs, const void* input_photon, int ninput, const Config* cfg,
                    void* photon_out, void* record_out, void* seq_out, void* prd_out, uint64_t* nray_out, uint64_t* nhit_out) {
    (void)nprim; (void)nnode; (void)nplan; (void)ninput;
    Scene sc;
    if (make_scene(sc, solid, nsolid, prim, node, plan, itra, nitra, inst, ninst, cfg->use_boxes)) return -1;
    Tables tb;
    tb.bnd = {bnd, nwl, nbnd * 8, 4}; tb.icdf = {icdf, icdf_nx, 3, 1}; tb.optical = (const unsigned*)optical;
    tb.nm0 = dom_low; tb.nms = dom_step; tb.hd_factor = hd_factor;
    const GS* gs = (const GS*)genstep;
    std::vector<int> seed;
    for (int g = 0; g < ngs; g++) for (unsigned k = 0; k < gs[g].u[3]; k++) seed.push_back(g);
    int64_t n = (int64_t)seed.size();
    Photon* pout = (Photon*)photon_out; Photon* rec = (Photon*)record_out; Seq* seqo = (Seq*)seq_out; Prd* prdo = (Prd*)prd_out;
    int mr = cfg->max_record;
    uint64_t nray = 0, nhit = 0;
#ifdef _OPENMP
    if (cfg->nthreads > 0) omp_set_num_threads(cfg->nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nray, nhit)
    for (int64_t idx = 0; idx < n; idx++) {
        uint64_t photon_idx = cfg->photon_offset + (uint64_t)idx;
        Rng rng;
        rng.init(cfg->seed, photon_idx, cfg->offset);                        // qrng<Philox>::init qudarap/qrng.h:131-137
        rng.skipahead(cfg->skipahead * (uint64_t)cfg->event_index);
        Photon p; memset(&p, 0, sizeof(p));
        generate_photon(p, rng, gs[seed[idx]], tb, (const Photon*)input_photon, cfg->photon_offset, photon_idx);
        Seq seq = {{0, 0}, {0, 0}};
        int bounce = 0;
        unsigned last_lpos = 0u;
        OTagr tagr = {g_tag_out ? g_tag_out + 4 * idx : nullptr, g_flat_out ? g_flat_out + 64 * idx : nullptr, 0u};
        g_tagr = g_tag_out ? &tagr : nullptr;              // generation draws are not tagged (no tagr.add in the generators)
        if (rec && 0 < mr) rec[(size_t)mr * idx] = p;                          // sctx::point sysrap/sctx.h:134-140
        if (seqo) seq_add_nibble(seq, 0, p.flag(), p.boundary());
        while (bounce < cfg->max_bounce && p.time < cfg->max_time) {
            float tmin = (p.orient_boundary_flag & cfg->eps0mask) ? cfg->tmin0 : cfg->tmin;
            Prd prd;
            bool ok = trace(prd, sc, p.pos, p.mom, tmin, cfg->tmax);
            nray++;
            if (cfg->refine) {                                                 // trace<true> CSGOptiX/CSGOptiX7.cu:146-185 (distance is 1 after a miss)
                float t_approx = 0.99f * prd.t;
                if (t_approx > cfg->refine_distance) {
                    v3 closer = p.pos + t_approx * p.mom;
                    ok = trace(prd, sc, closer, p.mom, tmin, cfg->tmax);
                    nray++;
                    prd.t += t_approx;
                }
            }
            last_lpos = ok ? pack_lpos(prd.lposcost, prd.lposfphi) : 0u;
            if (!ok) break;
            prd.normal = normalize(prd.normal);
            if (prdo && bounce < mr) prdo[(size_t)mr * idx + bounce] = prd;    // sctx::trace
            int command = propagate(p, rng, prd, tb, cfg->debug_tag != 0);
            bounce++;
            if (rec && bounce < mr) rec[(size_t)mr * idx + bounce] = p;
            if (seqo) seq_add_nibble(seq, (unsigned)bounce, p.flag(), p.boundary());
            if (command == BREAK) break;
        }
        g_tagr = nullptr;
        if (seqo) seqo[idx] = seq;
        if (pout) pout[idx] = p;
        if (g_lite_out) {                                                      // sphotonlite::init + set_lpos, CSGOptiX7.cu:455-463
            unsigned* l = g_lite_out + 4 * idx;
            l[0] = (1u << 16) | (p.identity & 0xffffu); memcpy(l + 1, &p.time, 4); l[2] = last_lpos; l[3] = p.flagmask;
        }
        if ((p.flagmask & cfg->hit_mask) == cfg->hit_mask) nhit++;
    }
    if (nray_out) *nray_out = nray;
    if (nhit_out) *nhit_out = nhit;
    return 0;
}

int oracle_intersect(const int* solid, int nsolid, const void* prim, const void* node, const void* plan, const float* itra, int nitra,
                     const float* inst, int ninst, const float* o_tmin, const float* dir, int nray, float tmax, void* prd_out, int use_boxes) {
    Scene sc;
    if (make_scene(sc, solid, nsolid, prim, node, plan, itra, nitra, inst, ninst, use_boxes)) return -1;
    Prd* out = (Prd*)prd_out;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nray; i++) {
        Prd prd;
        trace(prd, sc, mk(o_tmin[4 * i], o_tmin[4 * i + 1], o_tmin[4 * i + 2]), mk(dir[4 * i], dir[4 * i + 1], dir[4 * i + 2]), o_tmin[4 * i + 3], tmax);
        out[i] = prd;
    }
    return 0;
}

// simtrace raygen (CSGOptiX/CSGOptiX7.cu:536-577) with qsim::generate_photon_simtrace_frame (qudarap/qsim.h:2459-2511) and
// sevent::add_simtrace (sysrap/sevent.h:670-697).  gensteps: FRAME (17) or INPUT_PHOTON_SIMTRACE (20); out: quad4 per slot.
int oracle_simtrace(const int* solid, int nsolid, const void* prim, const void* node, const void* plan, const float* itra, int nitra,
                    const float* inst, int ninst, const float* genstep, int ngs, const float* input, float tmin, float tmax,
                    uint64_t seed, uint64_t offset, float* out, int use_boxes) {
    Scene sc;
    if (make_scene(sc, solid, nsolid, prim, node, plan, itra, nitra, inst, ninst, use_boxes)) return -1;
    std::vector<int64_t> prefix(ngs + 1, 0);
    for (int g = 0; g < ngs; g++) { unsigned n; memcpy(&n, genstep + 24 * g + 3, 4); prefix[g + 1] = prefix[g] + n; }
    int64_t total = prefix[ngs];
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < total; idx++) {
        int g = int(std::upper_bound(prefix.begin(), prefix.end(), idx) - prefix.begin()) - 1;
        const float* gs = genstep + 24 * g;
        int gencode, gridaxes; memcpy(&gencode, gs, 4); memcpy(&gridaxes, gs + 1, 4);
        v3 pos = mk(0, 0, 0), mom = mk(0, 0, 1);
        if (gencode == 20) {
            pos = mk(input[16 * idx], input[16 * idx + 1], input[16 * idx + 2]);
            mom = mk(input[16 * idx + 4], input[16 * idx + 5], input[16 * idx + 6]);
        } else {
            Rng rng; rng.init(seed, (uint64_t)idx, offset);
            float u0 = rng.uniform();
            float sinPhi = sinf(2.f * PI_F * u0), cosPhi = cosf(2.f * PI_F * u0);
            float u1 = rng.uniform();
            float cosTheta = 2.f * u1 - 1.f, sinTheta = sqrtf(1.f - cosTheta * cosTheta);
            v3 l = mk(gs[4], gs[5], gs[6]), m;
            switch (gridaxes) {                                            // sxyz.h: XYZ 0, YZ 1, XZ 2, XY 3
                case 1: m = mk(0.f, cosPhi, sinPhi); break;
                case 2: m = mk(cosPhi, 0.f, sinPhi); break;
                case 3: m = mk(cosPhi, sinPhi, 0.f); break;
                default: m = mk(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta); break;
            }
            pos = right_multiply(gs + 8, l, 1.f);                          // qat4(gs) = rows q2..q5
            mom = right_multiply(gs + 8, m, 0.f);
        }
        Prd prd;
        bool ok = trace(prd, sc, pos, mom, tmin, tmax);
        if (!ok) { prd.normal = mk(0.6f, 0.6f, 0.6f); prd.t = 1.f; }       // __miss__ms: background colour, t = 1
        float* o = out + 16 * idx;
        o[0] = prd.normal.x; o[1] = prd.normal.y; o[2] = prd.normal.z; o[3] = prd.t;
        o[4] = pos.x + prd.t * mom.x; o[5] = pos.y + prd.t * mom.y; o[6] = pos.z + prd.t * mom.z; o[7] = tmin;
        o[8] = pos.x; o[9] = pos.y; o[10] = pos.z; memcpy(o + 11, &prd.prim_boundary, 4);
        o[12] = mom.x; o[13] = mom.y; o[14] = mom.z; memcpy(o + 15, &prd.iindex_identity, 4);
    }
    return (int)total;
}

// hit merging: SPM::merge_partial_select (sysrap/SPM.cu:153-290) with sphoton::select_pred / key_functor / reduce_op
// (sysrap/sphoton.h:277-304).  stable sort by key, then a left fold of each group.  Returns the merged count.
int oracle_merge(const float* photons, int n, unsigned mask, float tw, float* out) {
    struct P { float f[16]; };
    const P* in = (const P*)photons;
    auto u = [](const P& p, int k) { unsigned v; memcpy(&v, &p.f[k], 4); return v; };
    std::vector<int> sel;
    for (int i = 0; i < n; i++) if (mask == 0u || (u(in[i], 15) & mask) != 0u) sel.push_back(i);
    P* o = (P*)out;
    if (tw == 0.f) { for (size_t k = 0; k < sel.size(); k++) o[k] = in[sel[k]]; return (int)sel.size(); }
    auto key = [&](int i) { unsigned id = u(in[i], 13) & 0x00ffffffu; unsigned bucket = (unsigned)(in[i].f[3] / tw); return ((uint64_t)id << 48) | (uint64_t)bucket; };
    std::stable_sort(sel.begin(), sel.end(), [&](int a, int b) { return key(a) < key(b); });
    int m = 0;
    for (size_t k = 0; k < sel.size();) {
        uint64_t k0 = key(sel[k]);
        P r = in[sel[k]];
        unsigned hc = u(r, 7) >> 16;
        size_t j = k + 1;
        for (; j < sel.size() && key(sel[j]) == k0; j++) {
            const P& q = in[sel[j]];
            unsigned fm = u(r, 15) | u(q, 15);
            hc += u(q, 7) >> 16;
            bool r_first = fminf(r.f[3], q.f[3]) == r.f[3];
            if (!r_first) r = q;
            memcpy(&r.f[15], &fm, 4);
        }
        unsigned hi = (u(r, 7) & 0x0000ffffu) | ((hc & 0xffffu) << 16);
        memcpy(&r.f[7], &hi, 4);
        o[m++] = r;
        k = j;
    }
    return m;
}

// sphotonlite flavour of the merge (sysrap/sphotonlite.h key_functor / reduce_op): records are 4 x u32
int oracle_merge_lite(const unsigned* lite, int n, unsigned mask, float tw, unsigned* out) {
    auto tm = [&](int i) { float t; memcpy(&t, lite + 4 * i + 1, 4); return t; };
    std::vector<int> sel;
    for (int i = 0; i < n; i++) if (mask == 0u || (lite[4 * i + 3] & mask) != 0u) sel.push_back(i);
    if (tw == 0.f) { for (size_t k = 0; k < sel.size(); k++) memcpy(out + 4 * k, lite + 4 * sel[k], 16); return (int)sel.size(); }
    auto key = [&](int i) { unsigned id = lite[4 * i] & 0xffffu; unsigned bucket = (unsigned)(tm(i) / tw); return ((uint64_t)id << 48) | (uint64_t)bucket; };
    std::stable_sort(sel.begin(), sel.end(), [&](int a, int b) { return key(a) < key(b); });
    int m = 0;
    for (size_t k = 0; k < sel.size();) {
        uint64_t k0 = key(sel[k]);
        unsigned r[4]; memcpy(r, lite + 4 * sel[k], 16);
        float t = tm(sel[k]);
        unsigned hc = r[0] >> 16;
        size_t j = k + 1;
        for (; j < sel.size() && key(sel[j]) == k0; j++) { t = fminf(t, tm(sel[j])); r[3] |= lite[4 * sel[j] + 3]; hc += lite[4 * sel[j]] >> 16; }
        r[0] = ((hc & 0xffffu) << 16) | (r[0] & 0xffffu);
        memcpy(r + 1, &t, 4);
        memcpy(out + 4 * m, r, 16); m++;
        k = j;
    }
    return m;
}

// one prim, one ray: used to compare against the reference CSG headers compiled for the host
int oracle_intersect_prim(const void* node, int node_offset, const void* plan, const float* itra, int nitra, const float* o, const float* d, float tmin,
                          float* isect_out) {
    Scene sc; sc.node = (const Node*)node; sc.plan = (const v4*)plan; sc.prim = nullptr; sc.use_boxes = false;
    sc.itra.assign(itra, itra + (size_t)nitra * 16);
    v4 is = {0, 0, 0, 0};
    bool valid = intersect_prim(is, sc.node + node_offset, sc, tmin, mk(o[0], o[1], o[2]), mk(d[0], d[1], d[2]));
    isect_out[0] = is.x; isect_out[1] = is.y; isect_out[2] = is.z; isect_out[3] = is.w;
    return valid ? 1 : 0;
}

int oracle_intersect_prim_batch(const void* node, int node_offset, const void* plan, const float* itra, int nitra, const float* o, const float* d,
                                const float* tmin, int n, float* isect_out, int* valid_out) {
    Scene sc; sc.node = (const Node*)node; sc.plan = (const v4*)plan; sc.prim = nullptr; sc.use_boxes = false;
    sc.itra.assign(itra, itra + (size_t)nitra * 16);
    for (int i = 0; i < n; i++) {
        v4 is = {0, 0, 0, 0};
        bool valid = intersect_prim(is, sc.node + node_offset, sc, tmin[i], mk(o[3 * i], o[3 * i + 1], o[3 * i + 2]), mk(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
        isect_out[4 * i] = is.x; isect_out[4 * i + 1] = is.y; isect_out[4 * i + 2] = is.z; isect_out[4 * i + 3] = is.w;
        valid_out[i] = valid ? 1 : 0;
    }
    return 0;
}

// first nv uniforms of subsequences [id0, id0+ni)  (qudarap/QSim.cu:43-68)
void oracle_rng_sequence(float* out, int ni, int nv, uint64_t id0, uint64_t seed, uint64_t offset) {
    for (int i = 0; i < ni; i++) {
        Rng r; r.init(seed, id0 + (uint64_t)i, offset);
        for (int k = 0; k < nv; k++) out[(size_t)i * nv + k] = r.uniform();
    }
}

// the texture filter emulation on its own, for calibration against the hardware
void oracle_tex2d4(const float* data, int nx, int ny, const float* xy, int n, float* out) {
    Tex t = {data, nx, ny, 4};
    for (int i = 0; i < n; i++) { v4 r = tex2D4(t, xy[2 * i], xy[2 * i + 1]); memcpy(out + 4 * i, &r, 16); }
}

int oracle_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
