// oracle/ref_gpu_harness.cu : TEST INFRASTRUCTURE - never linked into or called by the product.
//
// Compiles the REFERENCE's own device headers (qudarap/qsim.h and friends, CSG/csg_intersect_*.h,
// included from where they lie under /root/reference - nothing is copied) into a plain CUDA kernel
// and exposes it through a small C ABI, so that tests can run the reference physics + CSG code
// photon-by-photon on the B200 next to libphox.so.  The only part of the reference path that is
// not the reference's code is ray traversal: OptiX (absent here) is replaced by a brute-force
// loop over instances and prims with closest-hit semantics; everything a ray does once it has a
// candidate prim (intersect_prim, normal transform, qsim::propagate ...) is the reference's.
//
// The driver below restates CSGOptiX/CSGOptiX7.cu:simulate (405-503), __intersection__is
// (869-940) and __closesthit__ch (749-847, custom-primitive branch) around those calls.
// Built by oracle/Makefile into oracle/_ref/libphoxref_<variant>.so, once with the reference's
// as-built flags (DEBUG_TAG, non-PRODUCTION: CSGOptiX/CMakeLists.txt:54-56) and once with
// -DPRODUCTION.
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include "scuda.h"
#include "squad.h"
#include "sqat4.h"
#include "sphoton.h"
#include "sphotonlite.h"
#include "scerenkov.h"
#include "sstate.h"
#ifndef PRODUCTION
#include "stag.h"
#include "sseq.h"
#include "srec.h"
#endif
#include "sevent.h"
#include "sctx.h"
#include "qrng.h"
#include "qsim.h"
#include "csg_intersect_leaf.h"
#include "csg_intersect_node.h"
#include "csg_intersect_tree.h"

struct RefInst {
    float inv[16];          // world -> object, row-vector convention, w column cleared
    int solid, identity, is_identity, prim_offset, num_prim, pad0, pad1, pad2;
};

struct RefParams {
    const CSGNode* node;
    const float4* plan;
    const qat4* itra;
    const CSGPrim* prim;
    const RefInst* inst;
    int ninst;
    qsim* sim;
    sevent* evt;
    unsigned long long photon_slot_offset;
    float tmin, tmin0, tmax, max_time;
    unsigned eps0mask;
    unsigned refine;                 // params.PropagateRefine
    float refine_distance;           // params.PropagateRefineDistance
    unsigned long long* nray;
};

// brute-force stand-in for optixTrace + IS + CH + MS
static __device__ void ref_trace(const RefParams& P, const float3& o, const float3& d, float tmin, float tmax, quad2* prd) {
    float best_t = tmax;
    bool found = false;
    float3 best_n = make_float3(0.f, 0.f, 0.f);
    int best_inst = 0, best_prim = 0;
    unsigned best_boundary = 0;
    for (int i = 0; i < P.ninst; i++) {
        const RefInst& ri = P.inst[i];
        const qat4* q = (const qat4*)ri.inv;
        float3 oo = ri.is_identity ? o : q->right_multiply(o, 1.f);
        float3 dd = ri.is_identity ? d : q->right_multiply(d, 0.f);
        for (int k = 0; k < ri.num_prim; k++) {
            int pidx = ri.prim_offset + k;
            const CSGPrim& pr = P.prim[pidx];
            const CSGNode* root = P.node + pr.nodeOffset();
            float4 isect = make_float4(0.f, 0.f, 0.f, 0.f);
            bool valid = intersect_prim(isect, root, P.plan, P.itra, tmin, oo, dd, false);
            if (valid && isect.w > tmin && (isect.w < best_t || (!found && isect.w == best_t))) {
                best_t = isect.w; found = true;
                best_n = make_float3(isect.x, isect.y, isect.z);
                best_inst = i; best_prim = pidx; best_boundary = root->boundary();
            }
        }
    }
    if (!found) {
        prd->q0.f.x = 0.f; prd->q0.f.y = 0.f; prd->q0.f.z = 0.f; prd->q0.f.w = 1.f;
        prd->q1.u.x = 0u; prd->q1.u.y = 0u;
        prd->set_iindex_identity_(0xffffffffu);
        prd->set_globalPrimIdx_boundary_(0xffffffffu);
        prd->set_lpos(0.f, 0.f);
        return;
    }
    const RefInst& ri = P.inst[best_inst];
    const qat4* q = (const qat4*)ri.inv;
    float3 oo = ri.is_identity ? o : q->right_multiply(o, 1.f);
    float3 dd = ri.is_identity ? d : q->right_multiply(d, 0.f);
    float3 n = best_n;
    if (!ri.is_identity) n = q->left_multiply(best_n, 0.f);       // optixTransformNormalFromObjectToWorldSpace
    const float3 lpos = oo + best_t * dd;
    prd->q0.f.x = n.x; prd->q0.f.y = n.y; prd->q0.f.z = n.z; prd->q0.f.w = best_t;
    prd->set_lpos(normalize_cost(lpos), normalize_fphi(lpos));
    prd->set_iindex_identity_((((unsigned)best_inst & 0xffffu) << 16) | ((unsigned)ri.identity & 0xffffu));
    unsigned gpi = P.prim[best_prim].globalPrimIdx();
    prd->set_globalPrimIdx_boundary_(((gpi & 0xffffu) << 16) | (best_boundary & 0xffffu));
}

__global__ void ref_simulate(RefParams P) {
    sevent* evt = P.evt;
    unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= evt->num_seed) return;
    unsigned genstep_idx = evt->seed[idx];
    const quad6& gs = evt->genstep[genstep_idx];
    unsigned long long photon_idx = P.photon_slot_offset + idx;
    qsim* sim = P.sim;

    RNG rng;
    sim->rng->init(rng, sim->evt->index, photon_idx);

    quad2 prd_;
    quad2* prd = &prd_;
    sctx ctx = {};
    ctx.evt = evt;
    ctx.prd = prd;
    ctx.idx = idx;
    ctx.pidx = photon_idx;

    sim->generate_photon(ctx.p, rng, gs, photon_idx, genstep_idx);

    int command = START;
    int bounce = 0;
    unsigned nray = 0;
#ifndef PRODUCTION
    ctx.point(bounce);
#endif
    while (bounce < evt->max_bounce && ctx.p.time < P.max_time) {
        float tmin = (ctx.p.orient_boundary_flag & P.eps0mask) ? P.tmin0 : P.tmin;
        ref_trace(P, ctx.p.pos, ctx.p.mom, tmin, P.tmax, prd);
        nray++;
        if (P.refine) {                                  // trace<true> (CSGOptiX7.cu:146-185), the same statements around ref_trace
            float t_approx = 0.99f * prd->distance();
            if (t_approx > P.refine_distance) {
                float3 closer_ray_origin = ctx.p.pos + t_approx * ctx.p.mom;
                ref_trace(P, closer_ray_origin, ctx.p.mom, tmin, P.tmax, prd);
                nray++;
                prd->distance_add(t_approx);
            }
        }
        if (prd->boundary() == 0xffffu) break;
        float3* normal = prd->normal();
        *normal = normalize(*normal);
#ifndef PRODUCTION
        ctx.trace(bounce);
#endif
        command = sim->propagate(bounce, rng, ctx);
        bounce++;
#ifndef PRODUCTION
        ctx.point(bounce);
#endif
        if (command == BREAK) break;
    }
#ifndef PRODUCTION
    ctx.end();
#endif
    if (evt->photon) evt->photon[idx] = ctx.p;
    atomicAdd(P.nray, (unsigned long long)nray);
}

__global__ void ref_intersect(RefParams P, const float4* o_tmin, const float4* dir, unsigned n, quad2* out) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    quad2 prd;
    ref_trace(P, make_float3(o_tmin[i].x, o_tmin[i].y, o_tmin[i].z), make_float3(dir[i].x, dir[i].y, dir[i].z), o_tmin[i].w, P.tmax, &prd);
    out[i] = prd;
}

// simtrace raygen (CSGOptiX7.cu:536-577) around the reference's own qsim::generate_photon_simtrace_frame
// (qudarap/qsim.h:2459-2511) and sevent::add_simtrace (sysrap/sevent.h:670-697)
__global__ void ref_simtrace(RefParams P, const quad6* genstep, const int* seed, unsigned n, quad4* out, unsigned long long seedv,
                             unsigned long long offset) {
    unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const quad6& gs = genstep[seed[idx]];
    RNG rng;
    curand_init(seedv, (unsigned long long)idx, offset, &rng);       // qrng<Philox>::init with event index 0 (qrng.h:131-137)
    qsim sim;
    quad4 p;
    sim.generate_photon_simtrace_frame(p, rng, gs, idx, (unsigned)seed[idx]);
    const float3& pos = (const float3&)p.q0.f;
    const float3& mom = (const float3&)p.q1.f;
    quad2 prd;
    prd.zero();
    ref_trace(P, pos, mom, P.tmin, P.tmax, &prd);
    if (prd.boundary() == 0xffffu) { prd.q0.f.x = 0.6f; prd.q0.f.y = 0.6f; prd.q0.f.z = 0.6f; prd.q0.f.w = 1.f; }   // __miss__ms: SBT.cc:181-193 background
    sevent evt;
    evt.simtrace = out;
    evt.add_simtrace(idx, p, &prd, P.tmin);
}

// ---- host side ---------------------------------------------------------------------------------
#define RCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(g_err, sizeof(g_err), "%s: %s", #call, cudaGetErrorString(e_)); return -1; } } while (0)
static char g_err[512];

static bool invert_affine(const float* m, float* out) {
    double a[3][3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) a[r][c] = m[4 * r + c];
    double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                 a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
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

static cudaError_t make_tex(cudaArray_t* arr, cudaTextureObject_t* tex, const void* src, size_t w, size_t h, int ch) {
    // qudarap/QTex.cc:225-275 : array upload + linear filter, normalized coordinates, wrap addressing
    cudaChannelFormatDesc desc = ch == 4 ? cudaCreateChannelDesc<float4>() : cudaCreateChannelDesc<float>();
    cudaError_t e = cudaMallocArray(arr, &desc, w, h);
    if (e != cudaSuccess) return e;
    size_t pitch = w * 4 * ch;
    e = cudaMemcpy2DToArray(*arr, 0, 0, src, pitch, pitch, h, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    cudaResourceDesc res; memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray; res.res.array.array = *arr;
    cudaTextureDesc td; memset(&td, 0, sizeof(td));
    td.addressMode[0] = cudaAddressModeWrap; td.addressMode[1] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 1;
    return cudaCreateTextureObject(tex, &res, &td, nullptr);
}

template <typename T> static cudaError_t up(T** d, const void* h, size_t bytes) {
    cudaError_t e = cudaMalloc((void**)d, bytes ? bytes : 16);
    if (e != cudaSuccess) return e;
    if (bytes && h) e = cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice);
    return e;
}

struct RefScene {
    CSGNode* d_node = nullptr; float4* d_plan = nullptr; qat4* d_itra = nullptr; CSGPrim* d_prim = nullptr; RefInst* d_inst = nullptr;
    int ninst = 0;
    int setup(const void* solid_, int nsolid, const void* prim, int nprim, const void* node, int nnode, const void* plan, int nplan,
              const void* itra, int nitra, const void* inst_, int ninst_) {
        RCK(up(&d_node, node, (size_t)nnode * 64));
        RCK(up(&d_prim, prim, (size_t)nprim * 64));
        RCK(up(&d_plan, plan, (size_t)nplan * 16));
        std::vector<float> it((size_t)nitra * 16);
        if (nitra) memcpy(it.data(), itra, (size_t)nitra * 64);
        for (int i = 0; i < nitra; i++) { it[16 * i + 3] = 0.f; it[16 * i + 7] = 0.f; it[16 * i + 11] = 0.f; it[16 * i + 15] = 1.f; }
        RCK(up(&d_itra, it.data(), (size_t)nitra * 64));
        const int* solid = (const int*)solid_;
        std::vector<RefInst> recs;
        if (ninst_ <= 0 || !inst_) {
            RefInst r; memset(&r, 0, sizeof(r));
            r.inv[0] = r.inv[5] = r.inv[10] = r.inv[15] = 1.f; r.is_identity = 1; r.solid = 0; r.identity = 0;
            r.num_prim = solid[4]; r.prim_offset = solid[5];
            recs.push_back(r);
        } else {
            const float* inst = (const float*)inst_;
            const int* insti = (const int*)inst_;
            for (int i = 0; i < ninst_; i++) {
                RefInst r; memset(&r, 0, sizeof(r));
                float m[16]; memcpy(m, inst + 16 * i, 64);
                int gas = insti[16 * i + 7];
                r.identity = insti[16 * i + 11];
                m[3] = m[7] = m[11] = 0.f; m[15] = 1.f;
                static const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
                r.is_identity = memcmp(m, ident, 64) == 0;
                if (!invert_affine(m, r.inv)) { snprintf(g_err, sizeof(g_err), "singular instance"); return -1; }
                if (gas < 0 || gas >= nsolid) { snprintf(g_err, sizeof(g_err), "bad gas_idx"); return -1; }
                r.solid = gas; r.num_prim = solid[12 * gas + 4]; r.prim_offset = solid[12 * gas + 5];
                recs.push_back(r);
            }
        }
        ninst = (int)recs.size();
        RCK(up(&d_inst, recs.data(), recs.size() * sizeof(RefInst)));
        return 0;
    }
    void release() { cudaFree(d_node); cudaFree(d_plan); cudaFree(d_itra); cudaFree(d_prim); cudaFree(d_inst); }
};

static void* g_tag_out = nullptr;      // optional stag[n] / sflat[n] host outputs of the next phoxref_simulate (DEBUG_TAG build)
static void* g_flat_out = nullptr;
extern "C" void phoxref_set_tag_out(void* tag, void* flat) { g_tag_out = tag; g_flat_out = flat; }
extern "C" const char* phoxref_last_error() { return g_err; }
extern "C" int phoxref_is_production() {
#ifdef PRODUCTION
    return 1;
#else
    return 0;
#endif
}

struct RefConfig {
    int max_bounce, max_record, event_index, pad;
    float tmin, tmin0, tmax, max_time;
    unsigned eps0mask, pad1;
    unsigned long long seed, offset, skipahead, photon_offset;
    unsigned refine; float refine_distance;
};

extern "C" int phoxref_simulate(const void* solid, int nsolid, const void* prim, int nprim, const void* node, int nnode, const void* plan, int nplan,
                                const void* itra, int nitra, const void* inst, int ninst,
                                const float* bnd, int nbnd, int nwl, float dom_low, float dom_step, const int* optical,
                                const float* icdf, int icdf_nx, int hd_factor,
                                const void* genstep, int ngs, const void* input_photon, int ninput, const RefConfig* cfg,
                                void* photon_out, void* record_out, void* seq_out, void* prd_out, unsigned long long* nray_out) {
    g_err[0] = 0;
    RefScene sc;
    if (sc.setup(solid, nsolid, prim, nprim, node, nnode, plan, nplan, itra, nitra, inst, ninst)) return -1;

    // seeds exactly like iexpand: seed[i] = index of the genstep that owns photon slot i
    const quad6* gs = (const quad6*)genstep;
    std::vector<int> seed;
    for (int g = 0; g < ngs; g++) for (unsigned k = 0; k < gs[g].q0.u.w; k++) seed.push_back(g);
    size_t n = seed.size();
    if (n == 0) { snprintf(g_err, sizeof(g_err), "no photons"); return -1; }

    cudaArray_t bnd_arr = nullptr, icdf_arr = nullptr;
    cudaTextureObject_t bnd_tex = 0, icdf_tex = 0;
    RCK(make_tex(&bnd_arr, &bnd_tex, bnd, nwl, (size_t)nbnd * 8, 4));
    if (icdf) RCK(make_tex(&icdf_arr, &icdf_tex, icdf, icdf_nx, 3, 1));

    quad4 meta; memset(&meta, 0, sizeof(meta));
    meta.q0.u.x = nwl; meta.q0.u.y = nbnd * 8; meta.q1.f.x = dom_low; meta.q1.f.z = dom_step;
    quad4* d_meta; RCK(up(&d_meta, &meta, sizeof(meta)));
    quad* d_optical; RCK(up(&d_optical, optical, (size_t)nbnd * 4 * 16));

    qbase h_base; h_base.pidx = 0xffffffffffffffffull;
    qbase* d_base; RCK(up(&d_base, &h_base, sizeof(h_base)));
    qbnd h_bnd; memset(&h_bnd, 0, sizeof(h_bnd));
    h_bnd.boundary_tex = bnd_tex; h_bnd.boundary_meta = d_meta; h_bnd.optical = d_optical;
    qbnd* d_bnd; RCK(up(&d_bnd, &h_bnd, sizeof(h_bnd)));
    qscint h_scint; memset(&h_scint, 0, sizeof(h_scint));
    h_scint.scint_tex = icdf_tex; h_scint.scint_meta = nullptr; h_scint.hd_factor = hd_factor;
    qscint* d_scint; RCK(up(&d_scint, &h_scint, sizeof(h_scint)));
    qcerenkov h_ck; memset(&h_ck, 0, sizeof(h_ck));
    h_ck.base = d_base; h_ck.bnd = d_bnd; h_ck.prop = nullptr;
    qcerenkov* d_ck; RCK(up(&d_ck, &h_ck, sizeof(h_ck)));
    qrng<RNG> h_rng; h_rng.seed = cfg->seed; h_rng.offset = cfg->offset; h_rng.skipahead_event_offset = cfg->skipahead;
    qrng<RNG>* d_rng; RCK(up(&d_rng, &h_rng, sizeof(h_rng)));

    sevent h_evt; memset(&h_evt, 0, sizeof(h_evt));
    h_evt.max_bounce = cfg->max_bounce;
    h_evt.index = cfg->event_index;
    h_evt.num_genstep = ngs; h_evt.num_seed = n; h_evt.num_photon = n;
    RCK(up(&h_evt.genstep, genstep, (size_t)ngs * sizeof(quad6)));
    RCK(up(&h_evt.seed, seed.data(), n * sizeof(int)));
    RCK(up(&h_evt.photon, (const void*)nullptr, n * sizeof(sphoton)));
    if (input_photon && ninput > 0) RCK(cudaMemcpy(h_evt.photon, input_photon, (size_t)ninput * sizeof(sphoton), cudaMemcpyHostToDevice));
#ifndef PRODUCTION
    if (record_out && cfg->max_record > 0) { h_evt.max_record = cfg->max_record; RCK(up(&h_evt.record, (const void*)nullptr, n * cfg->max_record * sizeof(sphoton))); RCK(cudaMemset(h_evt.record, 0, n * cfg->max_record * sizeof(sphoton))); }
    if (prd_out && cfg->max_record > 0) { h_evt.max_prd = cfg->max_record; RCK(up(&h_evt.prd, (const void*)nullptr, n * cfg->max_record * sizeof(quad2))); RCK(cudaMemset(h_evt.prd, 0, n * cfg->max_record * sizeof(quad2))); }
    if (seq_out) { h_evt.max_seq = 1; RCK(up(&h_evt.seq, (const void*)nullptr, n * sizeof(sseq))); }
#ifdef DEBUG_TAG
    if (g_tag_out && g_flat_out) {      // sctx::end writes evt->tag / evt->flat (sysrap/sctx.h:181-182)
        h_evt.max_tag = 1; h_evt.max_flat = 1;
        RCK(up(&h_evt.tag, (const void*)nullptr, n * sizeof(stag))); RCK(cudaMemset(h_evt.tag, 0, n * sizeof(stag)));
        RCK(up(&h_evt.flat, (const void*)nullptr, n * sizeof(sflat))); RCK(cudaMemset(h_evt.flat, 0, n * sizeof(sflat)));
    }
#endif
#endif
    sevent* d_evt; RCK(up(&d_evt, &h_evt, sizeof(h_evt)));

    qsim h_sim; memset(&h_sim, 0, sizeof(h_sim));
    h_sim.base = d_base; h_sim.evt = d_evt; h_sim.rng = d_rng; h_sim.bnd = d_bnd; h_sim.multifilm = nullptr;
    h_sim.cerenkov = d_ck; h_sim.scint = d_scint; h_sim.pmt = nullptr;
    qsim* d_sim; RCK(up(&d_sim, &h_sim, sizeof(h_sim)));

    unsigned long long* d_nray; RCK(up(&d_nray, (const void*)nullptr, 8)); RCK(cudaMemset(d_nray, 0, 8));

    RefParams P; memset(&P, 0, sizeof(P));
    P.node = sc.d_node; P.plan = sc.d_plan; P.itra = sc.d_itra; P.prim = sc.d_prim; P.inst = sc.d_inst; P.ninst = sc.ninst;
    P.sim = d_sim; P.evt = d_evt; P.photon_slot_offset = cfg->photon_offset;
    P.tmin = cfg->tmin; P.tmin0 = cfg->tmin0; P.tmax = cfg->tmax; P.max_time = cfg->max_time; P.eps0mask = cfg->eps0mask;
    P.nray = d_nray;
    P.refine = cfg->refine; P.refine_distance = cfg->refine_distance;

    RCK(cudaDeviceSetLimit(cudaLimitStackSize, 8192));
    const int T = 64;
    ref_simulate<<<(unsigned)((n + T - 1) / T), T>>>(P);
    RCK(cudaGetLastError());
    RCK(cudaDeviceSynchronize());

    if (photon_out) RCK(cudaMemcpy(photon_out, h_evt.photon, n * sizeof(sphoton), cudaMemcpyDeviceToHost));
#ifndef PRODUCTION
    if (record_out && h_evt.record) RCK(cudaMemcpy(record_out, h_evt.record, n * cfg->max_record * sizeof(sphoton), cudaMemcpyDeviceToHost));
    if (prd_out && h_evt.prd) RCK(cudaMemcpy(prd_out, h_evt.prd, n * cfg->max_record * sizeof(quad2), cudaMemcpyDeviceToHost));
    if (seq_out && h_evt.seq) RCK(cudaMemcpy(seq_out, h_evt.seq, n * sizeof(sseq), cudaMemcpyDeviceToHost));
#ifdef DEBUG_TAG
    if (g_tag_out && h_evt.tag) RCK(cudaMemcpy(g_tag_out, h_evt.tag, n * sizeof(stag), cudaMemcpyDeviceToHost));
    if (g_flat_out && h_evt.flat) RCK(cudaMemcpy(g_flat_out, h_evt.flat, n * sizeof(sflat), cudaMemcpyDeviceToHost));
    if (h_evt.tag) cudaFree(h_evt.tag);
    if (h_evt.flat) cudaFree(h_evt.flat);
#endif
#endif
    if (nray_out) RCK(cudaMemcpy(nray_out, d_nray, 8, cudaMemcpyDeviceToHost));

    cudaFree(d_nray); cudaFree(d_sim); cudaFree(d_evt); cudaFree(h_evt.genstep); cudaFree(h_evt.seed); cudaFree(h_evt.photon);
    if (h_evt.record) cudaFree(h_evt.record);
    if (h_evt.prd) cudaFree(h_evt.prd);
    if (h_evt.seq) cudaFree(h_evt.seq);
    cudaFree(d_rng); cudaFree(d_ck); cudaFree(d_scint); cudaFree(d_bnd); cudaFree(d_base); cudaFree(d_optical); cudaFree(d_meta);
    cudaDestroyTextureObject(bnd_tex); cudaFreeArray(bnd_arr);
    if (icdf_tex) { cudaDestroyTextureObject(icdf_tex); cudaFreeArray(icdf_arr); }
    sc.release();
    return 0;
}

extern "C" int phoxref_intersect(const void* solid, int nsolid, const void* prim, int nprim, const void* node, int nnode, const void* plan, int nplan,
                                 const void* itra, int nitra, const void* inst, int ninst,
                                 const float* o_tmin, const float* dir, int nray, float tmax, void* prd_out) {
    g_err[0] = 0;
    RefScene sc;
    if (sc.setup(solid, nsolid, prim, nprim, node, nnode, plan, nplan, itra, nitra, inst, ninst)) return -1;
    float4 *d_o, *d_d; quad2* d_out;
    RCK(up(&d_o, o_tmin, (size_t)nray * 16));
    RCK(up(&d_d, dir, (size_t)nray * 16));
    RCK(up(&d_out, (const void*)nullptr, (size_t)nray * 32));
    RefParams P; memset(&P, 0, sizeof(P));
    P.node = sc.d_node; P.plan = sc.d_plan; P.itra = sc.d_itra; P.prim = sc.d_prim; P.inst = sc.d_inst; P.ninst = sc.ninst; P.tmax = tmax;
    RCK(cudaDeviceSetLimit(cudaLimitStackSize, 8192));
    const int T = 64;
    ref_intersect<<<(unsigned)((nray + T - 1) / T), T>>>(P, d_o, d_d, (unsigned)nray, d_out);
    RCK(cudaGetLastError());
    RCK(cudaDeviceSynchronize());
    RCK(cudaMemcpy(prd_out, d_out, (size_t)nray * 32, cudaMemcpyDeviceToHost));
    cudaFree(d_o); cudaFree(d_d); cudaFree(d_out);
    sc.release();
    return 0;
}

extern "C" int phoxref_simtrace(const void* solid, int nsolid, const void* prim, int nprim, const void* node, int nnode, const void* plan, int nplan,
                                const void* itra, int nitra, const void* inst, int ninst, const void* genstep, int ngs, float tmin, float tmax,
                                unsigned long long seedv, unsigned long long offset, void* out) {
    g_err[0] = 0;
    RefScene sc;
    if (sc.setup(solid, nsolid, prim, nprim, node, nnode, plan, nplan, itra, nitra, inst, ninst)) return -1;
    const quad6* gs = (const quad6*)genstep;
    std::vector<int> seed;
    for (int g = 0; g < ngs; g++) for (unsigned k = 0; k < gs[g].q0.u.w; k++) seed.push_back(g);
    size_t n = seed.size();
    quad6* d_gs; int* d_seed; quad4* d_out;
    RCK(up(&d_gs, genstep, (size_t)ngs * sizeof(quad6)));
    RCK(up(&d_seed, seed.data(), n * sizeof(int)));
    RCK(up(&d_out, (const void*)nullptr, n * sizeof(quad4)));
    RefParams P; memset(&P, 0, sizeof(P));
    P.node = sc.d_node; P.plan = sc.d_plan; P.itra = sc.d_itra; P.prim = sc.d_prim; P.inst = sc.d_inst; P.ninst = sc.ninst; P.tmin = tmin; P.tmax = tmax;
    RCK(cudaDeviceSetLimit(cudaLimitStackSize, 8192));
    const int T = 64;
    ref_simtrace<<<(unsigned)((n + T - 1) / T), T>>>(P, d_gs, d_seed, (unsigned)n, d_out, seedv, offset);
    RCK(cudaGetLastError());
    RCK(cudaDeviceSynchronize());
    RCK(cudaMemcpy(out, d_out, n * sizeof(quad4), cudaMemcpyDeviceToHost));
    cudaFree(d_gs); cudaFree(d_seed); cudaFree(d_out);
    sc.release();
    return (int)n;
}

// hit merging with the reference's own functors (sysrap/sphoton.h:277-304, sysrap/sphotonlite.h), driven on the host the way
// SPM::merge_partial_select drives them on the device (sysrap/SPM.cu:153-290): select, key, stable sort by key, reduce by key.
template <typename T> static int ref_merge_t(const T* in, int n, unsigned mask, float tw, T* out) {
    typename T::select_pred sel{mask};
    typename T::key_functor key{tw};
    typename T::reduce_op red;
    std::vector<T> v;
    for (int i = 0; i < n; i++) if (mask == 0u || sel(in[i])) v.push_back(in[i]);
    if (tw == 0.f) { for (size_t k = 0; k < v.size(); k++) out[k] = v[k]; return (int)v.size(); }
    std::vector<int> order(v.size());
    for (size_t k = 0; k < v.size(); k++) order[k] = (int)k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(v[a]) < key(v[b]); });
    int m = 0;
    for (size_t k = 0; k < order.size();) {
        uint64_t k0 = key(v[order[k]]);
        T r = v[order[k]];
        size_t j = k + 1;
        for (; j < order.size() && key(v[order[j]]) == k0; j++) r = red(r, v[order[j]]);
        out[m++] = r;
        k = j;
    }
    return m;
}
extern "C" int phoxref_merge(const void* photons, int n, unsigned mask, float tw, void* out) {
    return ref_merge_t<sphoton>((const sphoton*)photons, n, mask, tw, (sphoton*)out);
}
extern "C" int phoxref_merge_lite(const void* lite, int n, unsigned mask, float tw, void* out) {
    return ref_merge_t<sphotonlite>((const sphotonlite*)lite, n, mask, tw, (sphotonlite*)out);
}
