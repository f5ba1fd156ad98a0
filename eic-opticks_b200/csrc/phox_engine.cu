// phox_engine.cu : host side of libphox.so - the context object behind the C ABI of include/phox.h.
//
// Plays the role of CSGOptiX (launch driver, CSGOptiX/CSGOptiX.cc:1122-1198), QSim::simulate
// (per-event loop with genstep slicing, qudarap/QSim.cc:428-617), QEvt (event buffers, genstep
// upload, hit gathering: qudarap/QEvt.cc:332-441, 934-963, 1113-1133), QBnd/QScint texture setup
// (qudarap/QBnd.cc:130-194, QTex.cc:260-275, QScint.cc:84-120) and CSGFoundry::upload
// (CSG/CSGFoundry.cc:3377-3405) for the one path this library covers.
//
// Design: one context = one GPU = one stream.  Geometry, BVH and tables are uploaded once and stay
// resident.  Event buffers grow on demand and are reused.  A launch is: (prefix of
// genstep.numphoton) -> k_simulate -> k_hit_offsets -> [sync: read hit total] -> k_hit_compact.
// There is no CPU fallback: without a CUDA device phox_create fails.
#include "../../include/phox.h"
#include "phox_kernels.cuh"
#include "phox_merge.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace phox;

namespace {
thread_local std::string g_create_error;

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;          // elements
    cudaError_t reserve(size_t n, bool keep = false, cudaStream_t st = 0) {
        if (n <= cap) return cudaSuccess;
        // growing buffers get headroom from the first allocation on: a cudaMalloc + cudaFree pair inside an event costs
        // ~1 ms alone but was measured at 40-700 ms when another thread of the process is inside the driver (NVML polling)
        size_t want = keep ? std::max(n + n / 4 + 1024, cap + cap / 2) : n;
        T* q = nullptr;
        cudaError_t e = cudaMalloc(&q, want * sizeof(T));
        if (e != cudaSuccess) return e;
        if (keep && p && cap) {
            e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { cudaFree(q); return e; }
        }
        if (p) cudaFree(p);
        p = q; cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
// per-call device scratch of the query entry points: freed on every return path
template <typename T>
struct TmpBuf {
    T* p = nullptr;
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
    ~TmpBuf() { if (p) cudaFree(p); }
    TmpBuf() = default;
    TmpBuf(const TmpBuf&) = delete;
    TmpBuf& operator=(const TmpBuf&) = delete;
};
}  // namespace

struct phox_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double last_bounces_per_photon = 0.;    // rays per photon of the previous launch (PHOX_KERNEL_AUTO heuristic)
    unsigned long long event_rays_seen = 0; // rays of the current event counted before this launch
    bool profiling = false;                 // phox_set_profiling: events between the kernels of the wavefront form
    std::vector<cudaEvent_t> prof_ev;
    std::string err;
    std::string description;
    phox_config cfg;
    size_t vram_total = 0;

    // geometry
    bool have_geometry = false;
    DevBuf<float4> d_node, d_plan, d_itra, d_prim;
    DevBuf<InstanceRec> d_inst;
    DevBuf<BvhNode> d_bvh;
    DevBuf<float> d_boxes;
    BvhScratch bvh_scratch;
    int nprim = 0, nnode = 0, nplan = 0, nitra = 0, ninst = 0, nsolid = 0;
    int tlas_root = 0;
    int build_kernels = 0;
    int sim_grid[2] = {0, 0};               // persistent grid size of k_simulate<false/true>

    // tables
    bool have_tables = false;
    cudaArray_t bnd_array = nullptr, icdf_array = nullptr;
    cudaTextureObject_t bnd_tex = 0, icdf_tex = 0;
    DevBuf<uint4> d_optical;
    unsigned nx = 0, ny = 0;
    float nm0 = 60.f, nms = 1.f;
    unsigned hd_factor = 0;
    float inv_ny = 0.f;
    unsigned y_fast = 0, need_lposcost = 1;

    // event
    DevBuf<Genstep> d_genstep;
    DevBuf<unsigned long long> d_prefix;
    DevBuf<Photon> d_input, d_photon, d_record, d_hit;
    DevBuf<Seq> d_seq;
    DevBuf<Prd> d_prd;
    DevBuf<unsigned> d_block_hits;
    DevBuf<unsigned> d_active[2], d_ndraw, d_wave_count;        // wavefront form: live-photon lists, draw counts, list lengths
    DevBuf<Prd> d_wave_hits;
    int wave_grid[3][2] = {{0, 0}, {0, 0}, {0, 0}};             // persistent grid sizes of generate/trace/propagate, <false/true>
    DevBuf<unsigned long long> d_block_off;
    DevBuf<unsigned long long> d_tag;          // DebugHeavy: stag per photon (4 u64)
    DevBuf<float> d_flat;                      // DebugHeavy: sflat per photon (64 floats)
    DevBuf<unsigned> d_tagslot;
    std::vector<unsigned long long> h_tag;
    std::vector<float> h_flat;
    DevBuf<unsigned> d_lpos;                   // lite mode: packed local position of each photon's last intersect
    DevBuf<PhotonLite> d_hitlite;              // lite mode: sphotonlite of every hit, same order as d_hit
    DevBuf<PhotonLite> d_merged_lite;
    DevBuf<Photon> d_merged;                   // result of the last phox_merge_hits / phox_merge
    DevBuf<Photon> d_merge_in;
    MergeScratch merge_scratch;
    DevBuf<float4> d_exact;                    // per CSGPrim: (sizes ; translation) of prims that are exactly a box
    DevBuf<float4> d_home;                     // per CSGPrim: HomeRec (box, candidate count, offset), see traverse_bvh
    DevBuf<unsigned> d_prim_pb;                // per CSGPrim: prim/boundary word of the hit record (hit_finish_core)
    DevBuf<float4> d_cand;                     // candidate lists of the home cells, two float4 per candidate
    DevBuf<unsigned> d_home_state[2];          // wavefront form, per list position (double-buffered like the lists): home cell of the photon
    DevBuf<uint2> d_pending; DevBuf<unsigned> d_pending_count;   // wavefront form: list positions the home cells left to k_wf_trace, and their count per bounce
    DevBuf<Photon> d_hit_stage[2];                 // phox_get_hits_async: hits of the last two events, copied out while the next event runs
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_stage = nullptr, ev_copied[2] = {nullptr, nullptr};
    int stage_idx = 0;
    DevBuf<unsigned> d_gs_home;                    // per genstep of the launch: home cell its photons start with
    DevBuf<Prd> d_wave_hits2;                      // second hit buffer: the physics kernel fills the next bounce's records while it reads this bounce's
    int num_home = 0;                          // prims that have a candidate list
    int max_prim_nodes = 1;                    // largest CSGPrim of the geometry, in nodes (picks the BVH kernel instance, see run_launch)
    unsigned tail_photons = 131072;            // PHOX_KERNEL_AUTO: live photons at or below which the persistent kernel finishes a wavefront event (env PHOX_TAIL_PHOTONS, 0 = never)
    DevBuf<float> d_slack;                     // per CSGPrim: exit-bound slack of prims that are exactly a box (0 = not such a prim)
    DevBuf<unsigned long long> d_counters;     // [0] rays, [1] hit total of the launch, [2] rays settled by their home cell, [3] work counter, [4-5] genstep info
    unsigned long long* h_counters = nullptr;  // pinned mirror
    std::vector<Photon> h_photon;              // concatenated per-launch arrays in debug modes
    std::vector<Photon> h_record;
    std::vector<Seq> h_seq;
    std::vector<Prd> h_prd;
    int64_t num_photon = 0, num_hit = 0;
    int event_max_record = 0;
    bool have_event = false;
    phox_stats stats;

    int fail(int code, const std::string& m) { err = m; return code; }
    int cuda_fail(cudaError_t e, const char* what) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return PHOX_E_CUDA;
    }
};

#define CK(call)                                                   \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return ctx->cuda_fail(_e, #call);   \
    } while (0)

// what PHOX_KERNEL_AUTO resolves to (the faster form on the north-star workload, see DESIGN.md / profiles/)
#ifndef PHOX_KERNEL_AUTO_CHOICE
#define PHOX_KERNEL_AUTO_CHOICE PHOX_KERNEL_WAVEFRONT
#endif

// PHOX_KERNEL_AUTO: photons per launch from which the wavefront form is picked, by the bounces per photon of the context's previous
// launch.  Long histories (the 8x8 crystals: 20 bounces per photon) pay off from 250 k photons, histories of a few bounces (tank 4,
// boolean zoo 3) from 400 k now that the loop stops launching once the list is empty, histories of one or two bounces (raindrop,
// PMT wall) only from 2 M photons: the persistent kernel has no per-bounce launches at all (scripts/small_events2.py, profiles/r2_summary.md).
static const int64_t kAutoWavefrontMinPhotons = 250000;
static const int64_t kAutoWavefrontMinPhotonsShort = 400000;
static const int64_t kAutoWavefrontMinPhotonsVeryShort = 2000000;
static const double kAutoShortHistory = 6.0, kAutoVeryShortHistory = 2.5;

extern "C" void phox_default_config(phox_config* c) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    c->max_bounce = 31;
    c->event_mode = PHOX_MODE_MINIMAL;
    c->max_record = 32;
    c->rng_mode = PHOX_RNG_DEBUG_TAG;
    c->accel = PHOX_ACCEL_BVH;
    c->hit_mask = F_SURFACE_DETECT;
    c->epsilon0_mask = F_TORCH | F_CERENKOV | F_SCINTILLATION | F_BULK_SCATTER | F_BULK_REEMIT;
    c->propagate_refine = 0;
    c->propagate_epsilon = 0.05f;
    c->propagate_epsilon0 = 0.05f;
    c->refine_distance = 5000.f;
    c->tmax = 1000000.f;
    c->max_time = 1.e27f;
    c->rng_seed = 0;
    c->rng_offset = 0;
    c->skipahead_event_offset = 100000;
    c->max_slot = 0;
}

extern "C" phox_context* phox_create(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("phox_create: no CUDA device (") + cudaGetErrorString(e) + "); this engine has no CPU path";
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        g_create_error = "phox_create: device index out of range";
        return nullptr;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return nullptr; }
    phox_context* ctx = new phox_context();
    ctx->device = device;
    phox_default_config(&ctx->cfg);
    std::memset(&ctx->stats, 0, sizeof(ctx->stats));
    e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    ctx->stream = ctx->own_stream;
    for (int k = 0; k < 4 && e == cudaSuccess; k++) e = cudaEventCreate(&ctx->ev[k]);
    if (const char* t = getenv("PHOX_TAIL_PHOTONS")) ctx->tail_photons = (unsigned)strtoul(t, nullptr, 10);
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_counters, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = ctx->d_counters.reserve(8);
    if (e != cudaSuccess) {
        g_create_error = std::string("phox_create: ") + cudaGetErrorString(e);
        delete ctx;
        return nullptr;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    size_t free_b = 0;
    cudaMemGetInfo(&free_b, &ctx->vram_total);
    for (int dbg = 0; dbg < 2; dbg++) {
        int per_sm = 0;
        e = dbg ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_simulate<true>, kSimThreads, 0)
                : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_simulate<false>, kSimThreads, 0);
        if (e != cudaSuccess || per_sm < 1) per_sm = 1;
        ctx->sim_grid[dbg] = per_sm * prop.multiProcessorCount;
        int w[3] = {0, 0, 0};
        if (dbg) {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w[0], k_wf_generate<true, true>, kWaveThreads, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w[1], k_wf_trace<true>, kTraceThreads, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w[2], k_wf_propagate<true, true>, kPropThreads, 0);
        } else {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w[0], k_wf_generate<false, true>, kWaveThreads, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w[1], k_wf_trace<false>, kTraceThreads, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w[2], k_wf_propagate<false, true>, kPropThreads, 0);
        }
        for (int k = 0; k < 3; k++) ctx->wave_grid[k][dbg] = std::max(w[k], 1) * prop.multiProcessorCount;

    }
    cudaGetLastError();
    char buf[256];
    std::snprintf(buf, sizeof(buf), "phox: B200-native simulate engine on device %d (%s, sm_%d%d, %d SMs, %.1f GB)", device, prop.name,
                  prop.major, prop.minor, prop.multiProcessorCount, ctx->vram_total / 1e9);
    ctx->description = buf;
    return ctx;
}

static void free_tables(phox_context* ctx) {
    if (ctx->bnd_tex) cudaDestroyTextureObject(ctx->bnd_tex);
    if (ctx->icdf_tex) cudaDestroyTextureObject(ctx->icdf_tex);
    if (ctx->bnd_array) cudaFreeArray(ctx->bnd_array);
    if (ctx->icdf_array) cudaFreeArray(ctx->icdf_array);
    ctx->bnd_tex = ctx->icdf_tex = 0;
    ctx->bnd_array = ctx->icdf_array = nullptr;
    ctx->have_tables = false;
}

extern "C" void phox_destroy(phox_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    free_tables(ctx);
    ctx->d_node.release(); ctx->d_plan.release(); ctx->d_itra.release(); ctx->d_prim.release();
    ctx->d_inst.release(); ctx->d_bvh.release(); ctx->d_boxes.release(); ctx->d_optical.release();
    ctx->d_genstep.release(); ctx->d_prefix.release(); ctx->d_input.release(); ctx->d_photon.release();
    ctx->d_record.release(); ctx->d_hit.release(); ctx->d_seq.release(); ctx->d_prd.release();
    ctx->d_block_hits.release(); ctx->d_block_off.release(); ctx->d_counters.release();
    ctx->d_slack.release(); ctx->d_exact.release();
    ctx->d_home.release(); ctx->d_cand.release(); ctx->d_prim_pb.release(); ctx->d_home_state[0].release(); ctx->d_home_state[1].release(); ctx->d_pending.release(); ctx->d_pending_count.release(); ctx->d_wave_hits2.release(); ctx->d_gs_home.release();
    ctx->d_tag.release(); ctx->d_flat.release(); ctx->d_tagslot.release();
    ctx->d_lpos.release(); ctx->d_hitlite.release(); ctx->d_merged_lite.release();
    ctx->d_merged.release(); ctx->d_merge_in.release(); merge_scratch_free(ctx->merge_scratch);
    ctx->d_active[0].release(); ctx->d_active[1].release(); ctx->d_ndraw.release(); ctx->d_wave_count.release(); ctx->d_wave_hits.release();
    bvh_scratch_free(ctx->bvh_scratch);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    for (int k = 0; k < 4; k++) if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
    for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->ev_stage) cudaEventDestroy(ctx->ev_stage);
    for (int k = 0; k < 2; k++) { if (ctx->ev_copied[k]) cudaEventDestroy(ctx->ev_copied[k]); ctx->d_hit_stage[k].release(); }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" const char* phox_last_error(const phox_context* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
extern "C" const char* phox_desc(const phox_context* ctx) { return ctx ? ctx->description.c_str() : "phox: no context"; }

// ---- geometry -----------------------------------------------------------------------------------
static bool invert_affine(const float* m, double* inv) {
    // m : qat4 row-vector convention, rows 0..2 = linear part, row 3 = translation ; v' = v*M
    double a[3][3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) a[r][c] = m[4 * r + c];
    double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                 a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    if (det == 0.0 || !std::isfinite(det)) return false;
    double id = 1.0 / det;
    double b[3][3];
    b[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) * id;
    b[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id;
    b[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
    b[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) * id;
    b[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id;
    b[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
    b[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) * id;
    b[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id;
    b[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
    double t[3] = {m[12], m[13], m[14]};
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) inv[4 * r + c] = b[r][c];
    for (int c = 0; c < 3; c++) inv[12 + c] = -(t[0] * b[0][c] + t[1] * b[1][c] + t[2] * b[2][c]);
    inv[3] = inv[7] = inv[11] = 0.0;
    inv[15] = 1.0;
    return true;
}

extern "C" int phox_set_geometry(phox_context* ctx, const void* solid_, int64_t nsolid, const void* prim_, int64_t nprim, const void* node_,
                                 int64_t nnode, const void* plan_, int64_t nplan, const void* itra_, int64_t nitra, const void* inst_,
                                 int64_t ninst) {
    if (!ctx) return PHOX_E_ARG;
    if (!solid_ || !prim_ || !node_ || nsolid <= 0 || nprim <= 0 || nnode <= 0) return ctx->fail(PHOX_E_ARG, "phox_set_geometry: solid/prim/node arrays are required");
    if ((nitra > 0 && !itra_) || (nplan > 0 && !plan_)) return ctx->fail(PHOX_E_ARG, "phox_set_geometry: null itra/plan with non-zero count");
    if (nprim > 0xffff + 1) { /* globalPrimIdx is truncated to 16 bits in prd, like the reference (CSGOptiX7.cu:899) */ }
    if (nprim > (int64_t)kLeafItemMask) return ctx->fail(PHOX_E_ARG, "phox_set_geometry: more than 2^29 prims (BVH leaf items carry two flag bits)");
    CK(cudaSetDevice(ctx->device));
    const Solid* solid = (const Solid*)solid_;
    const Prim* prim = (const Prim*)prim_;
    const Node* node = (const Node*)node_;
    const Qat4* inst = (const Qat4*)inst_;

    // validate the index structure before anything reaches the device
    for (int64_t s = 0; s < nsolid; s++) {
        if (solid[s].num_prim < 0 || solid[s].prim_offset < 0 || (int64_t)solid[s].prim_offset + solid[s].num_prim > nprim)
            return ctx->fail(PHOX_E_ARG, "phox_set_geometry: solid prim range outside prim array");
    }
    for (int64_t p = 0; p < nprim; p++) {
        int nn = prim[p].num_node(), no = prim[p].node_offset();
        if (nn <= 0 || no < 0 || (int64_t)no + nn > nnode) return ctx->fail(PHOX_E_ARG, "phox_set_geometry: prim node range outside node array");
    }
    for (int64_t n = 0; n < nnode; n++) {
        unsigned ti = node[n].transform_idx();
        if (ti > (unsigned)nitra) return ctx->fail(PHOX_E_ARG, "phox_set_geometry: node transform index outside itra array");
        if (node[n].typecode() == CSG_CONVEXPOLYHEDRON && (int64_t)node[n].u[0] + node[n].u[1] > nplan)
            return ctx->fail(PHOX_E_ARG, "phox_set_geometry: convexpolyhedron planes outside plan array");
    }
    // The BVH builder needs finite, ordered boxes (a NaN / inf box never finds a merge partner), and the CSG evaluators index
    // nodes relative to the prim's root: check both here so that a malformed foundry is an error code, not a hung or faulting GPU.
    for (int64_t p = 0; p < nprim; p++) {
        for (int k = 0; k < 3; k++) {
            float lo = prim[p].f[8 + k], hi = prim[p].f[11 + k];
            if (!std::isfinite(lo) || !std::isfinite(hi) || !(lo <= hi))
                return ctx->fail(PHOX_E_ARG, "phox_set_geometry: prim bounding box is not finite or has lo > hi");
        }
        const int nn = prim[p].num_node(), no = prim[p].node_offset();
        const Node& root = node[no];
        const unsigned tc = root.typecode();
        if (tc < CSG_NODE && tc != CSG_ZERO) {                       // boolean tree: complete binary tree of subNum nodes, list nodes may follow it
            const unsigned sub = root.sub_num();
            if (sub < 1u || sub > (unsigned)nn || ((sub + 1u) & sub) != 0u || sub > 255u)
                return ctx->fail(PHOX_E_ARG, "phox_set_geometry: tree root subNum is not 2^h - 1 within the prim's nodes (h <= 7)");
        }
        for (int k = 0; k < nn; k++) {
            const Node& nd = node[no + k];
            const unsigned t = nd.typecode();
            if (t == CSG_CONTIGUOUS || t == CSG_DISCONTIGUOUS || t == CSG_OVERLAP) {
                if ((uint64_t)nd.sub_offset() + nd.sub_num() > (uint64_t)nn || nd.sub_num() == 0u)
                    return ctx->fail(PHOX_E_ARG, "phox_set_geometry: list node sub range outside the prim's nodes");
                if (t == CSG_CONTIGUOUS && nd.sub_num() > 8u)
                    return ctx->fail(PHOX_E_ARG, "phox_set_geometry: contiguous list node with more than 8 subs (csg_intersect_node.h enter sort)");
            }
        }
    }

    // instances: default to one identity instance of solid 0 (what CSGFoundry::addInstance does for
    // the global remainder solid, sysrap/stree.h:6737-6747)
    std::vector<Qat4> inst_default;
    if (ninst <= 0 || !inst) {
        Qat4 q; std::memset(&q, 0, sizeof(q));
        q.f[0] = q.f[5] = q.f[10] = 1.f; q.f[15] = 1.f;
        q.i[3] = 0; q.i[7] = 0; q.i[11] = 0; q.i[15] = 0;
        inst_default.push_back(q);
        inst = inst_default.data();
        ninst = 1;
    }

    ctx->have_geometry = false;
    ctx->nsolid = (int)nsolid; ctx->nprim = (int)nprim; ctx->nnode = (int)nnode; ctx->nplan = (int)nplan; ctx->nitra = (int)nitra; ctx->ninst = (int)ninst;

    CK(ctx->d_node.reserve((size_t)nnode * 4));
    CK(ctx->d_prim.reserve((size_t)nprim * 4));
    CK(ctx->d_plan.reserve(std::max<size_t>(1, (size_t)nplan)));
    CK(ctx->d_itra.reserve(std::max<size_t>(4, (size_t)nitra * 4)));
    CK(cudaMemcpyAsync(ctx->d_node.p, node_, (size_t)nnode * 64, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_prim.p, prim_, (size_t)nprim * 64, cudaMemcpyHostToDevice, ctx->stream));
    if (nplan > 0) CK(cudaMemcpyAsync(ctx->d_plan.p, plan_, (size_t)nplan * 16, cudaMemcpyHostToDevice, ctx->stream));
    if (nitra > 0) {
        // the 4th column of a transform may carry identity ints (sqat4.h) : clear it, it multiplies w = 0
        std::vector<Qat4> it((const Qat4*)itra_, (const Qat4*)itra_ + nitra);
        for (auto& q : it) { q.f[3] = 0.f; q.f[7] = 0.f; q.f[11] = 0.f; q.f[15] = 1.f; }
        CK(cudaMemcpyAsync(ctx->d_itra.p, it.data(), (size_t)nitra * 64, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }

    // boxes: per-prim boxes straight from CSGPrim, per-solid union, per-instance world boxes
    std::vector<float> boxes((size_t)(nprim + ninst) * 6);
    // Prim boxes are padded by a few ulps of their coordinates: a prim computes its hit distance in
    // its own frame, the box test works in the solid frame, and at coincident faces (crystal on
    // grease, window on SiPM ...) the two roundings must not let the box cull a prim whose own t
    // ties with or just undercuts the current nearest hit.
    for (int64_t p = 0; p < nprim; p++) {
        float m = 1.f;
        for (int k = 0; k < 6; k++) m = std::max(m, std::fabs(prim[p].f[8 + k]));
        float pad = 2e-6f * m;
        for (int k = 0; k < 3; k++) { boxes[6 * p + k] = prim[p].f[8 + k] - pad; boxes[6 * p + 3 + k] = prim[p].f[11 + k] + pad; }
    }
    // Prims that ARE their box (a single un-complemented box3 leaf, no rotation): from inside such a prim the hit
    // is the exit face, so the traversal may drop it once a nearer hit is known (phox_kernels.cuh, exit bound).
    // slack = distance by which the padded box must be shrunk to lie inside the true box with a pad to spare.
    ctx->max_prim_nodes = 1;
    for (int64_t p = 0; p < nprim; p++) ctx->max_prim_nodes = std::max(ctx->max_prim_nodes, prim[p].num_node());
    {
        std::vector<unsigned> pb((size_t)nprim, 0u);
        const Node* nodes_h = (const Node*)node_;
        for (int64_t p = 0; p < nprim; p++) {
            const int no = prim[p].node_offset();
            const unsigned boundary = (no >= 0 && no < nnode) ? nodes_h[no].u[6] : 0u;         // node q1.z
            const unsigned gpi = prim[p].u[15];                                                 // prim q3.w
            pb[p] = ((gpi & 0xffffu) << 16) | (boundary & 0xffffu);
        }
        CK(ctx->d_prim_pb.reserve(std::max<size_t>(1, (size_t)nprim)));
        CK(cudaMemcpyAsync(ctx->d_prim_pb.p, pb.data(), (size_t)nprim * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    std::vector<float> slack((size_t)nprim, 0.f);
    std::vector<float> exact((size_t)nprim * 8, 0.f);
    const Node* hnode = (const Node*)node_;
    const Qat4* hitra = (const Qat4*)itra_;
    for (int64_t p = 0; p < nprim; p++) {
        if (prim[p].num_node() != 1) continue;
        int no = prim[p].node_offset();
        if (no < 0 || no >= nnode) continue;
        const Node& nd = hnode[no];
        if (nd.typecode() >= CSG_LEAF) slack[p] = -1.f;      // a single leaf node: kLeafSingle (overwritten below when it is an exact box)
        if (nd.typecode() != CSG_BOX3 || nd.complement()) continue;
        float c[3] = {0.f, 0.f, 0.f};
        unsigned ti = nd.transform_idx();
        if (ti > 0) {
            if ((int64_t)ti > nitra) continue;
            const Qat4& q = hitra[ti - 1];                 // inverse transform, row-vector convention: translation in row 3
            bool pure = q.f[0] == 1.f && q.f[5] == 1.f && q.f[10] == 1.f && q.f[1] == 0.f && q.f[2] == 0.f && q.f[4] == 0.f &&
                        q.f[6] == 0.f && q.f[8] == 0.f && q.f[9] == 0.f;
            if (!pure) continue;
            for (int k = 0; k < 3; k++) c[k] = -q.f[12 + k];
        }
        float m = 1.f, worst = 0.f;
        bool inside = true;
        for (int k = 0; k < 3; k++) {
            float tlo = c[k] - 0.5f * nd.f[k], thi = c[k] + 0.5f * nd.f[k];
            float plo = boxes[6 * p + k], phi = boxes[6 * p + 3 + k];
            if (!(plo <= tlo && thi <= phi)) inside = false;
            worst = std::max(worst, std::max(tlo - plo, phi - thi));
            m = std::max(m, std::max(std::fabs(plo), std::fabs(phi)));
        }
        if (inside && nd.f[0] > 0.f && nd.f[1] > 0.f && nd.f[2] > 0.f) {
            slack[p] = worst + 4e-6f * m;
            float* e = &exact[8 * (size_t)p];
            e[0] = nd.f[0]; e[1] = nd.f[1]; e[2] = nd.f[2]; e[3] = 0.f;      // q0: full sizes, as leaf_box3 reads them
            e[4] = -c[0]; e[5] = -c[1]; e[6] = -c[2]; e[7] = 0.f;            // row 3 of the inverse transform (0 without transform)
        }
    }
    std::vector<float> solid_box((size_t)nsolid * 6);
    for (int64_t s = 0; s < nsolid; s++) {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int k = 0; k < solid[s].num_prim; k++) {
            const float* b = &boxes[6 * (size_t)(solid[s].prim_offset + k)];
            for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b[a]); hi[a] = std::max(hi[a], b[a + 3]); }
        }
        for (int a = 0; a < 3; a++) { solid_box[6 * s + a] = lo[a]; solid_box[6 * s + 3 + a] = hi[a]; }
    }

    // node pool layout: [instance tree : max(ninst-1,1)] [solid s tree : max(numPrim-1,1)] ...
    std::vector<int> solid_root(nsolid);
    int pool = std::max<int>((int)ninst - 1, 1);
    ctx->tlas_root = 0;
    for (int64_t s = 0; s < nsolid; s++) { solid_root[s] = pool; pool += std::max(solid[s].num_prim - 1, 1); }

    std::vector<InstanceRec> recs(ninst);
    for (int64_t i = 0; i < ninst; i++) {
        const Qat4& q = inst[i];
        int gas = q.i[7];
        if (gas < 0 || gas >= nsolid) return ctx->fail(PHOX_E_ARG, "phox_set_geometry: instance gas_idx outside solid array");
        if (solid[gas].num_prim <= 0) return ctx->fail(PHOX_E_ARG, "phox_set_geometry: instance of a solid without prims");
        for (int k = 0; k < 16; k++)
            if (k % 4 != 3 && !std::isfinite(q.f[k])) return ctx->fail(PHOX_E_ARG, "phox_set_geometry: instance transform is not finite");
        float m[16];
        std::memcpy(m, q.f, 64);
        m[3] = m[7] = m[11] = 0.f; m[15] = 1.f;
        static const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        bool is_ident = std::memcmp(m, ident, 64) == 0;
        double inv[16];
        if (!invert_affine(m, inv)) return ctx->fail(PHOX_E_ARG, "phox_set_geometry: singular instance transform");
        InstanceRec& r = recs[i];
        std::memset(&r, 0, sizeof(r));
        for (int k = 0; k < 4; k++) r.inv[k] = make_float4((float)inv[4 * k], (float)inv[4 * k + 1], (float)inv[4 * k + 2], (float)inv[4 * k + 3]);
        r.solid = gas;
        r.identity = q.i[11];
        r.is_identity = is_ident ? 1 : 0;
        r.bvh_root = solid_root[gas];
        r.prim_offset = solid[gas].prim_offset;
        r.num_prim = solid[gas].num_prim;
        // world box of the instance = box of the 8 transformed corners of the solid box
        const float* sb = &solid_box[6 * (size_t)gas];
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int c = 0; c < 8; c++) {
            float v[3] = {(c & 1) ? sb[3] : sb[0], (c & 2) ? sb[4] : sb[1], (c & 4) ? sb[5] : sb[2]};
            for (int a = 0; a < 3; a++) {
                float w = m[a] * v[0] + m[4 + a] * v[1] + m[8 + a] * v[2] + m[12 + a];
                lo[a] = std::min(lo[a], w); hi[a] = std::max(hi[a], w);
            }
        }
        float* ib = &boxes[6 * (size_t)(nprim + i)];
        for (int a = 0; a < 3; a++) {
            float pad = 1e-4f * std::max(1.f, std::max(std::fabs(lo[a]), std::fabs(hi[a])));   // rounding of the corner transform
            ib[a] = lo[a] - pad; ib[a + 3] = hi[a] + pad;
        }
    }

    // Home cells (phox_kernels.cuh, traverse_bvh): for every prim of a solid that is placed exactly once, untransformed
    // (the remainder solid 0 of a CSGFoundry), the list of all prims whose padded box comes within a pad of its box.
    // A prim has no list when more than kHomeMaxCand prims do, when one of them is not an exact box (the candidate pass
    // runs inside the physics kernel and is kept free of calls into the general CSG evaluators), or when the box of a
    // transformed instance does.
    std::vector<float> home((size_t)nprim * 8, 0.f);
    std::vector<float4> cand;                            // two float4 per candidate: half sizes | prim, translation | instance (home_search)
    ctx->num_home = 0;
    {
        std::vector<int> solid_inst_count((size_t)nsolid, 0), solid_inst((size_t)nsolid, -1);
        for (int64_t i = 0; i < ninst; i++) { solid_inst_count[recs[i].solid]++; solid_inst[recs[i].solid] = (int)i; }
        struct WorldPrim { int item, inst; };
        std::vector<WorldPrim> wp;                       // prims of untransformed instances: their boxes are world boxes
        std::vector<int> moved;                          // transformed instances
        for (int64_t i = 0; i < ninst; i++) {
            if (!recs[i].is_identity) { moved.push_back((int)i); continue; }
            for (int k = 0; k < recs[i].num_prim; k++) {
                int q = recs[i].prim_offset + k;
                wp.push_back({slack[q] > 0.f ? q : (q | kLeafSingle), (int)i});       // flag = not an exact box
            }
        }
        size_t nelig = 0;
        for (int64_t sidx = 0; sidx < nsolid; sidx++)
            if (solid_inst_count[sidx] == 1 && recs[solid_inst[sidx]].is_identity) nelig += (size_t)solid[sidx].num_prim;
        const bool affordable = (double)nelig * (double)(wp.size() + moved.size()) < 4e9;      // host loop, early exit at kHomeMaxCand + 1
        for (int64_t sidx = 0; sidx < nsolid && affordable; sidx++) {
            if (solid_inst_count[sidx] != 1 || !recs[solid_inst[sidx]].is_identity) continue;
            for (int k = 0; k < solid[sidx].num_prim; k++) {
                const int p = solid[sidx].prim_offset + k;
                float m = 1.f;
                for (int a = 0; a < 6; a++) m = std::max(m, std::fabs(prim[p].f[8 + a]));
                for (int a = 0; a < 3; a++) m = std::max(m, prim[p].f[11 + a] - prim[p].f[8 + a]);
                const float hp = 2e-5f * m;
                float in_lo[3], in_hi[3], out_lo[3], out_hi[3];
                for (int a = 0; a < 3; a++) {
                    in_lo[a] = prim[p].f[8 + a] - hp; in_hi[a] = prim[p].f[11 + a] + hp;
                    out_lo[a] = in_lo[a] - hp; out_hi[a] = in_hi[a] + hp;
                }
                auto overlaps = [&](const float* b) {
                    return b[0] <= out_hi[0] && b[3] >= out_lo[0] && b[1] <= out_hi[1] && b[4] >= out_lo[1] && b[2] <= out_hi[2] && b[5] >= out_lo[2];
                };
                bool ok = true;
                for (size_t j = 0; j < moved.size() && ok; j++) ok = !overlaps(&boxes[6 * (size_t)(nprim + moved[j])]);
                const size_t first = cand.size();
                // wp is in ascending (instance, prim) order and so are the lists: the first of several equal distances is
                // the one keep_nearest would keep, and the candidate loop needs no tie rule of its own
                for (size_t j = 0; j < wp.size() && ok; j++) {
                    const int q = wp[j].item & kLeafItemMask;
                    if (!overlaps(&boxes[6 * (size_t)q])) continue;
                    const float* e = &exact[8 * (size_t)q];
                    if (!(wp[j].item & kLeafSingle)) {
                        // An exact box that holds this home's outer box with another pad to spare (the world volume, mother
                        // volumes) can only answer with its exit face, and that lies beyond the exit from the home box by
                        // more than the 1e-6 relative margin of the settle test (each of its exit-side slab distances exceeds
                        // the home box's by >= 2 pads = 4e-5 relative, against < 1e-6 of rounding): whenever the candidates
                        // settle a ray, such a prim is strictly farther than their answer.  It is left off the list.
                        bool encloses = true;
                        for (int a = 0; a < 3; a++) {
                            const float tlo = -e[4 + a] - 0.5f * e[a], thi = -e[4 + a] + 0.5f * e[a];
                            if (!(tlo <= out_lo[a] - hp && thi >= out_hi[a] + hp)) encloses = false;
                        }
                        if (encloses) continue;
                    }
                    if ((cand.size() - first) / 2 == (size_t)kHomeMaxCand || (wp[j].item & kLeafSingle)) { ok = false; break; }
                    cand.push_back(make_float4(0.5f * e[0], 0.5f * e[1], 0.5f * e[2], __builtin_bit_cast(float, q)));
                    cand.push_back(make_float4(e[4], e[5], e[6], __builtin_bit_cast(float, wp[j].inst)));
                }
                if (!ok || cand.size() == first) { cand.resize(first); continue; }
                float* h = &home[8 * (size_t)p];
                h[0] = in_lo[0]; h[1] = in_lo[1]; h[2] = in_lo[2]; h[3] = in_hi[0]; h[4] = in_hi[1]; h[5] = in_hi[2];
                int cnt = (int)(cand.size() - first) / 2, off = (int)first;
                std::memcpy(&h[6], &cnt, 4); std::memcpy(&h[7], &off, 4);
                ctx->num_home++;
            }
        }
    }
    CK(ctx->d_home.reserve(std::max<size_t>(2, (size_t)nprim * 2)));
    CK(cudaMemcpyAsync(ctx->d_home.p, home.data(), (size_t)nprim * 8 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->d_cand.reserve(std::max<size_t>(1, cand.size())));
    if (!cand.empty()) CK(cudaMemcpyAsync(ctx->d_cand.p, cand.data(), cand.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));

    CK(ctx->d_inst.reserve((size_t)ninst));
    CK(ctx->d_boxes.reserve(boxes.size()));
    CK(ctx->d_bvh.reserve((size_t)pool));
    CK(cudaMemcpyAsync(ctx->d_inst.p, recs.data(), recs.size() * sizeof(InstanceRec), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_boxes.p, boxes.data(), boxes.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->d_slack.reserve(std::max<size_t>(1, (size_t)nprim)));
    CK(cudaMemcpyAsync(ctx->d_slack.p, slack.data(), (size_t)nprim * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->d_exact.reserve(std::max<size_t>(2, (size_t)nprim * 2)));
    CK(cudaMemcpyAsync(ctx->d_exact.p, exact.data(), (size_t)nprim * 8 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));

    ctx->build_kernels = 0;
    CK(bvh_build(ctx->d_boxes.p + 6 * (size_t)nprim, (int)ninst, 0, ctx->d_bvh.p + ctx->tlas_root, ctx->bvh_scratch, ctx->stream, &ctx->build_kernels));
    for (int64_t s = 0; s < nsolid; s++) {
        if (solid[s].num_prim == 0) continue;
        CK(bvh_build(ctx->d_boxes.p + 6 * (size_t)solid[s].prim_offset, solid[s].num_prim, solid[s].prim_offset, ctx->d_bvh.p + solid_root[s],
                     ctx->bvh_scratch, ctx->stream, &ctx->build_kernels));
#if PHOX_EXACT_BOX
        k_mark_exact_boxes<<<(std::max(solid[s].num_prim - 1, 1) + 127) / 128, 128, 0, ctx->stream>>>(
            ctx->d_bvh.p + solid_root[s], std::max(solid[s].num_prim - 1, 1), ctx->d_slack.p);
        CK(cudaGetLastError());
        ctx->build_kernels += 1;
#endif
        CK(cudaStreamSynchronize(ctx->stream));     // scratch is reused by the next build
    }
    CK(cudaStreamSynchronize(ctx->stream));
    {   // the traversal stack holds kBvhStack entries and drops pushes beyond that: refuse trees that could overflow it
        std::vector<BvhNode> hn((size_t)pool);
        CK(cudaMemcpyAsync(hn.data(), ctx->d_bvh.p, (size_t)pool * sizeof(BvhNode), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const int dt = bvh_tree_depth(hn.data() + ctx->tlas_root, std::max<int>((int)ninst - 1, 1));
        int ds = 0;
        bool known = dt >= 0;
        for (int64_t s = 0; s < nsolid && known; s++) {
            if (solid[s].num_prim == 0) continue;
            const int d = bvh_tree_depth(hn.data() + solid_root[s], std::max(solid[s].num_prim - 1, 1));
            if (d < 0) known = false; else ds = std::max(ds, d);
        }
        if (known && dt + 1 + ds > kBvhStack)
            return ctx->fail(PHOX_E_ARG, "phox_set_geometry: BVH deeper than the traversal stack (instance tree + solid tree > 63 levels)");
    }
    ctx->have_geometry = true;
    return PHOX_OK;
}

// ---- tables -------------------------------------------------------------------------------------
static cudaError_t make_tex(cudaArray_t* arr, cudaTextureObject_t* tex, const void* src, size_t width, size_t height, int channels) {
    cudaChannelFormatDesc desc = channels == 4 ? cudaCreateChannelDesc<float4>() : cudaCreateChannelDesc<float>();
    cudaError_t e = cudaMallocArray(arr, &desc, width, height);
    if (e != cudaSuccess) return e;
    size_t pitch = width * sizeof(float) * channels;
    e = cudaMemcpy2DToArray(*arr, 0, 0, src, pitch, pitch, height, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    cudaResourceDesc res;
    std::memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray;
    res.res.array.array = *arr;
    cudaTextureDesc td;
    std::memset(&td, 0, sizeof(td));
    td.addressMode[0] = cudaAddressModeWrap;        // qudarap/QTex.cc:260-275
    td.addressMode[1] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 1;
    return cudaCreateTextureObject(tex, &res, &td, nullptr);
}

extern "C" int phox_set_tables(phox_context* ctx, const float* bnd, int64_t nbnd, int64_t nwl, float domain_low, float domain_step,
                               const int32_t* optical, const float* icdf, int64_t icdf_ny, int64_t icdf_nx, int32_t hd_factor) {
    if (!ctx) return PHOX_E_ARG;
    if (!bnd || !optical || nbnd <= 0 || nwl <= 1) return ctx->fail(PHOX_E_ARG, "phox_set_tables: bnd and optical are required");
    if (nbnd * 8 > 65536) return ctx->fail(PHOX_E_ARG, "phox_set_tables: more than 8192 boundaries do not fit one 2D texture");
    if (!(domain_step > 0.f)) return ctx->fail(PHOX_E_ARG, "phox_set_tables: domain_step must be positive");
    if (icdf && (icdf_ny != 3 || icdf_nx < 2)) return ctx->fail(PHOX_E_ARG, "phox_set_tables: icdf must be 3 rows (hd layers) x nx");
    if (icdf && hd_factor != 0 && hd_factor != 10 && hd_factor != 20) return ctx->fail(PHOX_E_ARG, "phox_set_tables: hd_factor must be 0, 10 or 20");
    CK(cudaSetDevice(ctx->device));
    free_tables(ctx);
    ctx->nx = (unsigned)nwl; ctx->ny = (unsigned)(nbnd * 8);
    ctx->nm0 = domain_low; ctx->nms = domain_step;
    {   // bnd_y (phox_physics.cuh): the reciprocal form of (iy + 0.5) / ny is used only when it gives the bits of the division for every row
        const float c = (float)ctx->ny;
        volatile float inv = 1.0f / c;
        bool same = true;
        for (unsigned iy = 0; iy < ctx->ny && same; iy++) {
            const float a = (float)iy + 0.5f;
            volatile float q = a * inv;
            const float fast = std::fmaf(std::fmaf(-q, c, a), inv, q);
            same = fast == a / c;
        }
        ctx->inv_ny = inv; ctx->y_fast = same ? 1u : 0u;
    }
    CK(make_tex(&ctx->bnd_array, &ctx->bnd_tex, bnd, (size_t)nwl, (size_t)nbnd * 8, 4));
    if (icdf) {
        CK(make_tex(&ctx->icdf_array, &ctx->icdf_tex, icdf, (size_t)icdf_nx, (size_t)icdf_ny, 1));
        ctx->hd_factor = (unsigned)hd_factor;
    }
    CK(ctx->d_optical.reserve((size_t)nbnd * 4));
    CK(cudaMemcpy(ctx->d_optical.p, optical, (size_t)nbnd * 4 * 16, cudaMemcpyHostToDevice));
    ctx->need_lposcost = 0;                           // optical[row] = (index, ems, ..): propagate() reads HitInfo::lposcost only behind a SURFACE row
    for (int64_t r = 0; r < nbnd * 4; r++) {          // (osur / isur) whose ems is neither NoSurface nor Surface (qsim.h:2296-2312)
        const unsigned ems = (unsigned)optical[4 * r + 1];
        if ((r % 4 == SP_OSUR || r % 4 == SP_ISUR) && ems != EMS_NoSurface && ems != EMS_Surface) ctx->need_lposcost = 1;
    }
    ctx->have_tables = true;
    return PHOX_OK;
}

extern "C" int phox_set_stream(phox_context* ctx, void* cuda_stream) {
    if (!ctx) return PHOX_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return PHOX_OK;
}

extern "C" int phox_set_config(phox_context* ctx, const phox_config* cfg) {
    if (!ctx || !cfg) return PHOX_E_ARG;
    if (cfg->max_bounce < 0) return ctx->fail(PHOX_E_ARG, "phox_set_config: max_bounce < 0");
    if (cfg->max_record < 0 || cfg->max_record > 32) return ctx->fail(PHOX_E_ARG, "phox_set_config: max_record must be 0..32 (sseq::SLOTS)");
    if (cfg->event_mode < PHOX_MODE_MINIMAL || cfg->event_mode > PHOX_MODE_DEBUGHEAVY) return ctx->fail(PHOX_E_ARG, "phox_set_config: unknown event_mode");
    if (cfg->max_slot < 0) return ctx->fail(PHOX_E_ARG, "phox_set_config: max_slot < 0");
    if (cfg->kernel_mode > PHOX_KERNEL_WAVEFRONT) return ctx->fail(PHOX_E_ARG, "phox_set_config: unknown kernel_mode");
    if (cfg->mode_lite > 1) return ctx->fail(PHOX_E_ARG, "phox_set_config: mode_lite must be 0 or 1");
    ctx->cfg = *cfg;
    return PHOX_OK;
}

extern "C" int phox_set_profiling(phox_context* ctx, int on) {
    if (!ctx) return PHOX_E_ARG;
    ctx->profiling = on != 0;
    return PHOX_OK;
}

extern "C" int phox_get_config(const phox_context* ctx, phox_config* cfg) {
    if (!ctx || !cfg) return PHOX_E_ARG;
    *cfg = ctx->cfg;
    return PHOX_OK;
}

// ---- events ---------------------------------------------------------------------------------------
static int64_t effective_max_slot(const phox_context* ctx) {
    if (ctx->cfg.max_slot > 0) return ctx->cfg.max_slot;
    // SEventConfig::HeuristicMaxSlot : 0.87*VRAM / (64 B * 1.75)   (sysrap/SEventConfig.cc:1897-1903)
    double v = 0.87 * (double)ctx->vram_total / (64.0 * 1.75);
    int64_t s = (int64_t)v;
    return std::min<int64_t>(s, 0xffffff00ll);      // one launch indexes slots with 32 bits
}

struct Slice { int64_t gs_start, gs_stop, ph_offset, ph_count; };

// SGenstep::GetGenstepSlices (sysrap/SGenstep.h:249-323): greedy, whole gensteps, in order.
static void make_slices(std::vector<Slice>& out, const Genstep* gs, int64_t n, int64_t max_slot) {
    Slice sl = {0, 0, 0, 0};
    for (int64_t i = 0; i < n; i++) {
        int64_t num = gs[i].numphoton();
        if (sl.ph_count + num <= max_slot) { sl.gs_stop = i + 1; sl.ph_count += num; }
        else {
            sl.gs_stop = i;
            out.push_back(sl);
            sl.ph_count = num; sl.gs_start = i; sl.gs_stop = i + 1;
        }
        if (i == n - 1) out.push_back(sl);
    }
    int64_t off = 0;
    for (auto& s : out) { s.ph_offset = off; off += s.ph_count; }
}

static bool mode_keeps_photon(int m) { return m != PHOX_MODE_MINIMAL; }
static bool mode_keeps_seq(int m) { return m == PHOX_MODE_HITPHOTONSEQ || m == PHOX_MODE_DEBUGLITE || m == PHOX_MODE_DEBUGHEAVY; }
static bool mode_keeps_record(int m) { return m == PHOX_MODE_DEBUGLITE || m == PHOX_MODE_DEBUGHEAVY; }
static bool mode_keeps_prd(int m) { return m == PHOX_MODE_DEBUGHEAVY; }

// one launch over slots [0,n) ; gensteps + prefix already on the device
static int run_launch(phox_context* ctx, const Genstep* d_gs, const unsigned long long* d_prefix, int ngs, const Photon* d_input,
                      unsigned long long input_base, unsigned long long photon_offset, int64_t n, int event_id) {
    const int T = kHitTile;
    int nblock = (int)((n + T - 1) / T);                        // hit-compaction tiles
    const phox_config& c = ctx->cfg;
    int mode = c.event_mode;
    bool dbg = mode_keeps_seq(mode) || mode_keeps_record(mode) || mode_keeps_prd(mode);

    CK(ctx->d_photon.reserve((size_t)n));
    CK(ctx->d_block_hits.reserve((size_t)nblock));
    CK(ctx->d_block_off.reserve((size_t)nblock));
    if (mode_keeps_seq(mode)) CK(ctx->d_seq.reserve((size_t)n));
    if (mode_keeps_record(mode)) CK(ctx->d_record.reserve((size_t)n * c.max_record));
    if (mode_keeps_prd(mode)) CK(ctx->d_prd.reserve((size_t)n * c.max_record));
    if (mode_keeps_prd(mode)) {        // DebugHeavy also keeps the tag / flat arrays (SEventConfig.cc:1590-1592: MaxTag = MaxFlat = 1)
        CK(ctx->d_tag.reserve((size_t)n * 4)); CK(ctx->d_flat.reserve((size_t)n * 64)); CK(ctx->d_tagslot.reserve((size_t)n));
        CK(cudaMemsetAsync(ctx->d_tag.p, 0, (size_t)n * 32, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_flat.p, 0, (size_t)n * 256, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_tagslot.p, 0, (size_t)n * 4, ctx->stream));
    }
    const bool lite = c.mode_lite != 0;
    if (lite) CK(ctx->d_lpos.reserve((size_t)n));
    // record / prd slots past the end of a history read as zero (debug modes only, so not on the production path)
    if (mode_keeps_record(mode) && c.max_record) CK(cudaMemsetAsync(ctx->d_record.p, 0, (size_t)n * c.max_record * sizeof(Photon), ctx->stream));
    if (mode_keeps_prd(mode) && c.max_record) CK(cudaMemsetAsync(ctx->d_prd.p, 0, (size_t)n * c.max_record * sizeof(Prd), ctx->stream));

    SimParams P;
    std::memset(&P, 0, sizeof(P));
    P.scene.geo.node = ctx->d_node.p; P.scene.geo.plan = ctx->d_plan.p; P.scene.geo.itra = ctx->d_itra.p;
    P.scene.prim = ctx->d_prim.p; P.scene.exact = ctx->d_exact.p; P.scene.nodes = ctx->d_bvh.p; P.scene.inst = ctx->d_inst.p; P.scene.prim_pb = ctx->d_prim_pb.p;
    P.scene.ninst = ctx->ninst; P.scene.tlas_root = ctx->tlas_root;
    P.scene.accel = c.accel == PHOX_ACCEL_BVH_NOHOME ? PHOX_ACCEL_BVH : c.accel;
    P.scene.home = (c.accel == PHOX_ACCEL_BVH && ctx->num_home > 0) ? ctx->d_home.p : nullptr; P.scene.cand = ctx->d_cand.p;
    P.tables.bnd_tex = ctx->bnd_tex; P.tables.icdf_tex = ctx->icdf_tex; P.tables.optical = ctx->d_optical.p;
    P.tables.nx = ctx->nx; P.tables.ny = ctx->ny; P.tables.nm0 = ctx->nm0; P.tables.nms = ctx->nms; P.tables.hd_factor = ctx->hd_factor;
    P.tables.inv_ny = ctx->inv_ny; P.tables.y_fast = ctx->y_fast; P.tables.need_lposcost = ctx->need_lposcost;
    P.genstep = d_gs; P.gs_prefix = d_prefix; P.num_genstep = ngs;
    P.input_photon = d_input; P.input_base = input_base; P.photon_offset = photon_offset;
    P.num_photon = (unsigned)n; P.event_index = event_id;
    P.photon = ctx->d_photon.p;
    P.seq = mode_keeps_seq(mode) ? ctx->d_seq.p : nullptr;
    P.record = mode_keeps_record(mode) ? ctx->d_record.p : nullptr;
    P.prd = mode_keeps_prd(mode) ? ctx->d_prd.p : nullptr;
    P.lpos = lite ? ctx->d_lpos.p : nullptr;
    P.tag = mode_keeps_prd(mode) ? ctx->d_tag.p : nullptr;
    P.flat = mode_keeps_prd(mode) ? ctx->d_flat.p : nullptr;
    P.tagslot = mode_keeps_prd(mode) ? ctx->d_tagslot.p : nullptr;
    P.max_record = c.max_record;
    P.work_counter = reinterpret_cast<unsigned*>(ctx->d_counters.p + 3);
    P.counters = ctx->d_counters.p;
    P.max_bounce = c.max_bounce;
    P.tmin = c.propagate_epsilon; P.tmin0 = c.propagate_epsilon0; P.tmax = c.tmax; P.max_time = c.max_time;
    P.refine_distance = c.refine_distance; P.eps0_mask = c.epsilon0_mask; P.hit_mask = c.hit_mask; P.refine = c.propagate_refine;
    P.seed = c.rng_seed; P.rng_offset = c.rng_offset; P.skipahead = c.skipahead_event_offset;
    P.burn = c.rng_mode == PHOX_RNG_DEBUG_TAG ? 1 : 0;
    P.resume_list = nullptr; P.resume_count = nullptr; P.resume_bounce = 0;

    // AUTO: the wavefront form wins once a launch fills the machine several times over; below that its ~65 kernels each
    // run a single under-filled wave and the one persistent kernel is up to 2x quicker (profiles/r1_summary.md, small events)
    if (n > 0x7fffffffll) return ctx->fail(PHOX_E_NOMEM, "launch of more than 2^31 photons: lower max_slot");
    const double bpp = ctx->last_bounces_per_photon;
    const int64_t auto_min = !(bpp > 0.) ? kAutoWavefrontMinPhotons
                             : bpp < kAutoVeryShortHistory ? kAutoWavefrontMinPhotonsVeryShort
                             : bpp < kAutoShortHistory ? kAutoWavefrontMinPhotonsShort : kAutoWavefrontMinPhotons;
    const bool wavefront = (c.kernel_mode == PHOX_KERNEL_AUTO ? (n < auto_min ? PHOX_KERNEL_PERSISTENT : PHOX_KERNEL_AUTO_CHOICE)
                                                              : c.kernel_mode) == PHOX_KERNEL_WAVEFRONT;
    if (!wavefront) {
        CK(cudaMemsetAsync(ctx->d_counters.p + 3, 0, sizeof(unsigned long long), ctx->stream));
        // persistent grid: every SM gets as many resident blocks as the kernel's registers allow
        int sim_blocks = ctx->sim_grid[dbg ? 1 : 0];
        sim_blocks = (int)std::min<int64_t>(sim_blocks, (n + kSimThreads - 1) / kSimThreads);
        CK(cudaEventRecord(ctx->ev[0], ctx->stream));
        if (dbg) k_simulate<true><<<sim_blocks, kSimThreads, 0, ctx->stream>>>(P);
        else k_simulate<false><<<sim_blocks, kSimThreads, 0, ctx->stream>>>(P);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[1], ctx->stream));
        ctx->stats.num_kernel += 1;
    } else {
        // wavefront: generate, then per bounce a trace kernel and a physics kernel over the live list;
        // list lengths stay on the device, so there is no host synchronisation inside the loop
        CK(ctx->d_active[0].reserve((size_t)n)); CK(ctx->d_active[1].reserve((size_t)n));
        CK(ctx->d_wave_hits.reserve((size_t)n));                  // (the draw counts travel in the index word of the photon records)
        const bool homes = P.scene.home != nullptr;
        const bool home_pass = homes && !c.propagate_refine;       // PropagateRefine re-traces from 0.99 t: those rays take the tree
        if (home_pass) {
            CK(ctx->d_home_state[0].reserve((size_t)n)); CK(ctx->d_home_state[1].reserve((size_t)n));
            CK(ctx->d_wave_hits2.reserve((size_t)n));
            CK(ctx->d_pending.reserve((size_t)n));
            CK(ctx->d_pending_count.reserve((size_t)c.max_bounce + 2));
            CK(cudaMemsetAsync(ctx->d_pending_count.p, 0, ((size_t)c.max_bounce + 2) * sizeof(unsigned), ctx->stream));
        }
        CK(ctx->d_wave_count.reserve((size_t)c.max_bounce + 2));
        CK(cudaMemsetAsync(ctx->d_wave_count.p, 0, ((size_t)c.max_bounce + 2) * sizeof(unsigned), ctx->stream));
        ctx->h_counters[3] = (unsigned long long)n;              // pinned staging for the first list length (low word)
        CK(cudaMemcpyAsync(ctx->d_wave_count.p, &ctx->h_counters[3], sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
        WaveParams W;
        std::memset(&W, 0, sizeof(W));
        W.sim = P;
        W.ndraw = nullptr; W.hits = ctx->d_wave_hits.p;
        W.home = nullptr; W.home_next = home_pass ? ctx->d_home_state[0].p : nullptr;     // k_wf_generate fills the list of bounce 0
        const int d = dbg ? 1 : 0;
        auto grid = [&](int k, int threads = kWaveThreads) {
            return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ctx->wave_grid[k][d], (n + threads - 1) / threads));
        };
        CK(cudaEventRecord(ctx->ev[0], ctx->stream));
        W.active_out = ctx->d_active[0].p;
        if (home_pass) {       // home cell of each genstep, handed to its photons
            CK(ctx->d_gs_home.reserve((size_t)std::max(ngs, 1)));
            k_genstep_home<<<(ngs + 127) / 128, 128, 0, ctx->stream>>>(d_gs, ngs, P.scene.home, ctx->nprim, ctx->d_gs_home.p);
            W.gs_home = ctx->d_gs_home.p;
            ctx->stats.num_kernel += 1;
        }
        if (home_pass) {
            W.hits_next = ctx->d_wave_hits.p;                     // bounce 0 reads hit_buf[0]
            W.pending = ctx->d_pending.p; W.pending_count = ctx->d_pending_count.p;
            if (dbg) k_wf_generate<true, true><<<grid(0), kWaveThreads, 0, ctx->stream>>>(W);
            else k_wf_generate<false, true><<<grid(0), kWaveThreads, 0, ctx->stream>>>(W);
        } else {
            if (dbg) k_wf_generate<true, false><<<grid(0), kWaveThreads, 0, ctx->stream>>>(W);
            else k_wf_generate<false, false><<<grid(0), kWaveThreads, 0, ctx->stream>>>(W);
        }
        CK(cudaGetLastError());
        const bool prof = ctx->profiling;
        if (prof) while (ctx->prof_ev.size() < 2 * (size_t)c.max_bounce + 1) { cudaEvent_t e; CK(cudaEventCreate(&e)); ctx->prof_ev.push_back(e); }
        Prd* hit_buf[2] = {ctx->d_wave_hits.p, home_pass ? ctx->d_wave_hits2.p : ctx->d_wave_hits.p};
        // The loop ends early once the list has run empty (short histories: one or two bounces on a PMT wall or in a water box, of
        // max_bounce = 32).  The host must not hold the device up inside the loop, so every kLiveCheck bounces the BVH kernel posts the
        // list length it starts from into a pinned host word (a plain store over PCIe: no copy, no event, no gap in the stream), and
        // the host reads the word of the batch BEFORE the one it queued last.
        constexpr int kLiveCheck = 4;
        constexpr unsigned kLiveUnknown = 0xffffffffu;
        volatile unsigned* h_live = reinterpret_cast<volatile unsigned*>(ctx->h_counters + 7);       // two pinned words, one per batch in flight
        int launched = 0;
        for (int b = 0; b < c.max_bounce; b++) {
            W.active_in = ctx->d_active[b & 1].p; W.active_out = ctx->d_active[(b + 1) & 1].p;
            W.home = home_pass ? ctx->d_home_state[b & 1].p : nullptr; W.home_next = home_pass ? ctx->d_home_state[(b + 1) & 1].p : nullptr;
            W.count_in = ctx->d_wave_count.p + b; W.count_out = ctx->d_wave_count.p + b + 1;
            W.hits = hit_buf[b & 1]; W.hits_next = hit_buf[(b + 1) & 1];
            W.bounce = b;
            W.live_report = nullptr;
            if (b > 0 && b % kLiveCheck == 0) {
                const int k = b / kLiveCheck;
                if (k >= 2) {                                   // the word of batch k - 1 (posted by the BVH kernel of bounce b - kLiveCheck)
                    unsigned v;
                    for (unsigned spin = 0; (v = h_live[(k - 1) & 1]) == kLiveUnknown; spin++)
                        if ((spin & 0xfffffu) == 0xfffffu && cudaStreamQuery(ctx->stream) != cudaErrorNotReady) break;      // a dead stream must not hang the host
                    if (v == 0u) break;                         // nothing was alive when batch k - 1 began: batch k - 1 ran on empty lists, the rest is not launched
                    // PHOX_KERNEL_AUTO, production modes: a list of a few thousand photons costs every bounce the latency of one ray through
                    // the tree TWICE over plus two launches (tank: 28 % of the event in such launches, profiles/r2_summary.md).  The rest of
                    // the histories goes to the persistent kernel - same bodies, same bytes - which has no launch boundary between bounces.
                    if (v <= ctx->tail_photons && c.kernel_mode == PHOX_KERNEL_AUTO && !dbg && !lite && !prof) {
                        SimParams R = P;
                        R.resume_list = ctx->d_active[b & 1].p; R.resume_count = ctx->d_wave_count.p + b; R.resume_bounce = b;
                        CK(cudaMemsetAsync(ctx->d_counters.p + 3, 0, sizeof(unsigned long long), ctx->stream));
                        const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ctx->sim_grid[0], ((int64_t)v + kSimThreads - 1) / kSimThreads));
                        k_simulate<false><<<blocks, kSimThreads, 0, ctx->stream>>>(R);
                        ctx->stats.num_kernel += 1;
                        break;
                    }
                }
                h_live[k & 1] = kLiveUnknown;
                W.live_report = const_cast<unsigned*>(h_live + (k & 1));
            }
            // per bounce: k_wf_trace -> k_wf_propagate ; with profiling on, an event before each kernel.
            // With home cells the physics kernel of bounce b - 1 has already written the hit records of the rays their home
            // settled (for bounce 0: k_wf_generate); the trace kernel takes the rest, the pending list.
            if (prof) CK(cudaEventRecord(ctx->prof_ev[2 * b], ctx->stream));
            W.pending = home_pass ? ctx->d_pending.p : nullptr;
            W.pending_count = home_pass ? ctx->d_pending_count.p + b : nullptr;
            if (dbg) k_wf_trace<true><<<grid(1, kTraceThreads), kTraceThreads, 0, ctx->stream>>>(W);
            else if (ctx->max_prim_nodes <= 3) k_wf_trace<false, true><<<grid(1, kTraceThreads), kTraceThreads, 0, ctx->stream>>>(W);      // no boolean tree beyond one operator: hit_finish in place
            else k_wf_trace<false><<<grid(1, kTraceThreads), kTraceThreads, 0, ctx->stream>>>(W);
            if (prof) CK(cudaEventRecord(ctx->prof_ev[2 * b + 1], ctx->stream));
            W.pending = home_pass ? ctx->d_pending.p : nullptr;
            W.pending_count = home_pass ? ctx->d_pending_count.p + b + 1 : nullptr;
            if (home_pass) {
                if (dbg) k_wf_propagate<true, true><<<grid(2, kPropThreads), kPropThreads, 0, ctx->stream>>>(W);
                else k_wf_propagate<false, true><<<grid(2, kPropThreads), kPropThreads, 0, ctx->stream>>>(W);
            } else {
                if (dbg) k_wf_propagate<true, false><<<grid(2, kPropThreads), kPropThreads, 0, ctx->stream>>>(W);
                else k_wf_propagate<false, false><<<grid(2, kPropThreads), kPropThreads, 0, ctx->stream>>>(W);
            }
            if (prof && b == c.max_bounce - 1) CK(cudaEventRecord(ctx->prof_ev[2 * b + 2], ctx->stream));
            launched = b + 1;
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[1], ctx->stream));
        ctx->stats.num_kernel += 1 + 2 * (uint64_t)launched;
    }
    k_hit_count<<<nblock, T, 0, ctx->stream>>>(ctx->d_photon.p, (unsigned)n, c.hit_mask, ctx->d_block_hits.p);
    CK(cudaGetLastError());
    k_hit_offsets<<<1, 1024, 0, ctx->stream>>>(ctx->d_block_hits.p, nblock, ctx->d_block_off.p, ctx->d_counters.p + 1);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int64_t nhit = (int64_t)ctx->h_counters[1];
    ctx->stats.num_kernel += 3;
    ctx->last_bounces_per_photon = double(ctx->h_counters[0] - ctx->event_rays_seen) / double(n);     // feeds PHOX_KERNEL_AUTO
    ctx->event_rays_seen = ctx->h_counters[0];
    if (nhit > 0) {
        CK(ctx->d_hit.reserve((size_t)(ctx->num_hit + nhit), true, ctx->stream));
        if (lite) CK(ctx->d_hitlite.reserve((size_t)(ctx->num_hit + nhit), true, ctx->stream));
        k_hit_compact<<<nblock, T, 0, ctx->stream>>>(ctx->d_photon.p, (unsigned)n, c.hit_mask, ctx->d_block_off.p, ctx->d_hit.p + ctx->num_hit,
                                                     lite ? ctx->d_lpos.p : nullptr, lite ? ctx->d_hitlite.p + ctx->num_hit : nullptr);
        CK(cudaGetLastError());
        ctx->stats.num_kernel += 1;
    }
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    CK(cudaEventSynchronize(ctx->ev[2]));
    float ms_sim = 0.f, ms_rest = 0.f;
    CK(cudaEventElapsedTime(&ms_sim, ctx->ev[0], ctx->ev[1]));
    CK(cudaEventElapsedTime(&ms_rest, ctx->ev[1], ctx->ev[2]));
    ctx->stats.simulate_kernel_seconds += ms_sim * 1e-3;
    ctx->stats.compact_kernel_seconds += ms_rest * 1e-3;
    if (wavefront && ctx->profiling && c.max_bounce > 0) {
        std::vector<unsigned> live((size_t)c.max_bounce + 2);
        CK(cudaMemcpy(live.data(), ctx->d_wave_count.p, live.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
        for (int b = 0; b < c.max_bounce; b++) {
            if (live[b] == 0) break;                // later kernels found an empty list
            float ms_t = 0.f, ms_p = 0.f;
            CK(cudaEventElapsedTime(&ms_t, ctx->prof_ev[2 * b], ctx->prof_ev[2 * b + 1]));
            CK(cudaEventElapsedTime(&ms_p, ctx->prof_ev[2 * b + 1], ctx->prof_ev[2 * b + 2]));
            ctx->stats.trace_kernel_seconds += ms_t * 1e-3;
            ctx->stats.propagate_kernel_seconds += ms_p * 1e-3;
            ctx->stats.num_trace_launch += 1;
        }
    }
    ctx->num_hit += nhit;
    ctx->stats.num_launch += 1;
    return PHOX_OK;
}

static int gather_debug(phox_context* ctx, int64_t n) {
    int mode = ctx->cfg.event_mode;
    int mr = ctx->cfg.max_record;
    if (mode_keeps_photon(mode)) {
        size_t o = ctx->h_photon.size();
        ctx->h_photon.resize(o + n);
        CK(cudaMemcpyAsync(ctx->h_photon.data() + o, ctx->d_photon.p, (size_t)n * 64, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (mode_keeps_seq(mode)) {
        size_t o = ctx->h_seq.size();
        ctx->h_seq.resize(o + n);
        CK(cudaMemcpyAsync(ctx->h_seq.data() + o, ctx->d_seq.p, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (mode_keeps_record(mode)) {
        size_t o = ctx->h_record.size();
        ctx->h_record.resize(o + (size_t)n * mr);
        CK(cudaMemcpyAsync(ctx->h_record.data() + o, ctx->d_record.p, (size_t)n * mr * 64, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (mode_keeps_prd(mode)) {
        size_t o = ctx->h_prd.size();
        ctx->h_prd.resize(o + (size_t)n * mr);
        CK(cudaMemcpyAsync(ctx->h_prd.data() + o, ctx->d_prd.p, (size_t)n * mr * 32, cudaMemcpyDeviceToHost, ctx->stream));
        size_t ot = ctx->h_tag.size(), of = ctx->h_flat.size();
        ctx->h_tag.resize(ot + (size_t)n * 4);
        ctx->h_flat.resize(of + (size_t)n * 64);
        CK(cudaMemcpyAsync(ctx->h_tag.data() + ot, ctx->d_tag.p, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_flat.data() + of, ctx->d_flat.p, (size_t)n * 256, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return PHOX_OK;
}

static int begin_event(phox_context* ctx) {
    if (!ctx->have_geometry) return ctx->fail(PHOX_E_STATE, "phox_simulate: geometry not set");
    if (!ctx->have_tables) return ctx->fail(PHOX_E_STATE, "phox_simulate: tables not set");
    CK(cudaSetDevice(ctx->device));
    ctx->num_photon = 0; ctx->num_hit = 0;
    ctx->h_photon.clear(); ctx->h_record.clear(); ctx->h_seq.clear(); ctx->h_prd.clear(); ctx->h_tag.clear(); ctx->h_flat.clear();
    std::memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->event_max_record = ctx->cfg.max_record;
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
    ctx->event_rays_seen = 0;
    ctx->have_event = false;
    return PHOX_OK;
}

static int check_gensteps(phox_context* ctx, const Genstep* gs, int64_t ngs, int64_t ninput, bool have_input) {
    int64_t n_input_gs = 0;
    for (int64_t i = 0; i < ngs; i++) {
        int code = gs[i].gencode();
        if ((code == GS_SCINTILLATION || code == GS_DsG4Scintillation_r4695) && !ctx->icdf_tex)
            return ctx->fail(PHOX_E_STATE, "phox_simulate: scintillation genstep but no icdf table was set");
        if (code == GS_INPUT_PHOTON) {
            n_input_gs++;
            if (!have_input) return ctx->fail(PHOX_E_ARG, "phox_simulate: INPUT_PHOTON genstep without input photons");
            if ((int64_t)gs[i].numphoton() != ninput || ngs != 1)
                return ctx->fail(PHOX_E_ARG, "phox_simulate: input photons ride on exactly one INPUT_PHOTON genstep with numphoton == ninput");
        }
    }
    (void)n_input_gs;
    return PHOX_OK;
}

extern "C" int phox_simulate(phox_context* ctx, const void* genstep_, int64_t ngs, const void* input_photon, int64_t ninput, int32_t event_id,
                             uint64_t photon_offset, double* launch_seconds) {
    if (!ctx) return PHOX_E_ARG;
    if (ngs < 0 || (ngs > 0 && !genstep_)) return ctx->fail(PHOX_E_ARG, "phox_simulate: null genstep array");
    int rc = begin_event(ctx);
    if (rc) return rc;
    if (ngs == 0) {          // a valid empty event (e.g. the share of a rank when an event has fewer gensteps than ranks): no launch, no hits
        if (launch_seconds) *launch_seconds = 0.;
        ctx->have_event = true;
        return PHOX_OK;
    }
    const Genstep* gs = (const Genstep*)genstep_;
    rc = check_gensteps(ctx, gs, ngs, ninput, input_photon != nullptr);
    if (rc) return rc;

    double t0 = now_s();
    std::vector<Slice> slices;
    int64_t max_slot = effective_max_slot(ctx);
    for (int64_t i = 0; i < ngs; i++)
        if ((int64_t)gs[i].numphoton() > max_slot) return ctx->fail(PHOX_E_NOMEM, "phox_simulate: one genstep exceeds max_slot; split it (photon_offset lets callers shard input photons)");
    make_slices(slices, gs, ngs, max_slot);

    CK(ctx->d_genstep.reserve((size_t)ngs));
    CK(ctx->d_prefix.reserve((size_t)ngs + 1));
    CK(cudaMemcpyAsync(ctx->d_genstep.p, gs, (size_t)ngs * sizeof(Genstep), cudaMemcpyHostToDevice, ctx->stream));
    if (input_photon) {
        CK(ctx->d_input.reserve((size_t)ninput));
        CK(cudaMemcpyAsync(ctx->d_input.p, input_photon, (size_t)ninput * 64, cudaMemcpyHostToDevice, ctx->stream));
    }
    std::vector<unsigned long long> prefix;
    double t_up = now_s() - t0, t_launch = 0., t_gather = 0.;
    for (const Slice& sl : slices) {
        if (sl.ph_count == 0) continue;
        double ta = now_s();
        int n = (int)(sl.gs_stop - sl.gs_start);
        prefix.assign((size_t)n + 1, 0ull);
        for (int k = 0; k < n; k++) prefix[k + 1] = prefix[k] + gs[sl.gs_start + k].numphoton();
        CK(cudaMemcpyAsync(ctx->d_prefix.p, prefix.data(), prefix.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        double tb = now_s();
        rc = run_launch(ctx, ctx->d_genstep.p + sl.gs_start, ctx->d_prefix.p, n, input_photon ? ctx->d_input.p : nullptr,
                        photon_offset, photon_offset + (uint64_t)sl.ph_offset, sl.ph_count, event_id);
        if (rc) return rc;
        double tc = now_s();
        if (mode_keeps_photon(ctx->cfg.event_mode)) { rc = gather_debug(ctx, sl.ph_count); if (rc) return rc; }
        double td = now_s();
        t_up += tb - ta; t_launch += tc - tb; t_gather += td - tc;
        ctx->num_photon += sl.ph_count;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(ctx->h_counters, ctx->d_counters.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    ctx->stats.num_photon = (uint64_t)ctx->num_photon; ctx->stats.num_hit = (uint64_t)ctx->num_hit; ctx->stats.num_ray = ctx->h_counters[0]; ctx->stats.num_home_ray = ctx->h_counters[2];
    ctx->stats.launch_seconds = t_launch; ctx->stats.upload_seconds = t_up; ctx->stats.gather_seconds = t_gather;
    if (launch_seconds) *launch_seconds = t_launch;
    ctx->have_event = true;
    return PHOX_OK;
}

extern "C" int phox_simulate_device(phox_context* ctx, const void* d_genstep, int64_t ngs, const void* d_input_photon, int64_t ninput,
                                    int32_t event_id, uint64_t photon_offset, double* launch_seconds) {
    if (!ctx) return PHOX_E_ARG;
    if (ngs < 0 || (ngs > 0 && !d_genstep)) return ctx->fail(PHOX_E_ARG, "phox_simulate_device: null genstep array");
    int rc = begin_event(ctx);
    if (rc) return rc;
    if (ngs == 0) { if (launch_seconds) *launch_seconds = 0.; ctx->have_event = true; return PHOX_OK; }      // empty event, like phox_simulate
    double t0 = now_s();
    CK(ctx->d_prefix.reserve((size_t)ngs + 1));
    // the prefix kernel reads every genstep anyway: it also reports what check_gensteps looks at on the host path
    // ([0] bit 0 scintillation genstep, bit 1 input-photon genstep; [1] input-photon gensteps; [2] their numphoton)
    unsigned* d_gsinfo = reinterpret_cast<unsigned*>(ctx->d_counters.p + 4);       // slots 4-5; 0 rays, 1 hit total, 3 work counter
    CK(cudaMemsetAsync(d_gsinfo, 0, 16, ctx->stream));
    k_genstep_prefix<<<1, 1024, 0, ctx->stream>>>((const Genstep*)d_genstep, (int)ngs, ctx->d_prefix.p, d_gsinfo);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_counters + 4, d_gsinfo, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_counters + 6, ctx->d_prefix.p + ngs, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int64_t n = (int64_t)ctx->h_counters[6];
    unsigned gsinfo[4];
    std::memcpy(gsinfo, ctx->h_counters + 4, 16);
    ctx->stats.num_kernel += 1;
    if (n > effective_max_slot(ctx)) return ctx->fail(PHOX_E_NOMEM, "phox_simulate_device: event exceeds max_slot; the device-resident path is single-launch");
    if ((gsinfo[0] & 1u) && !ctx->icdf_tex) return ctx->fail(PHOX_E_STATE, "phox_simulate_device: scintillation genstep but no icdf table was set");
    if (gsinfo[0] & 2u) {
        if (!d_input_photon) return ctx->fail(PHOX_E_ARG, "phox_simulate_device: INPUT_PHOTON genstep without input photons");
        if (gsinfo[1] != 1u || ngs != 1 || (int64_t)gsinfo[2] != ninput)
            return ctx->fail(PHOX_E_ARG, "phox_simulate_device: input photons ride on exactly one INPUT_PHOTON genstep with numphoton == ninput");
    }
    if (n == 0) { if (launch_seconds) *launch_seconds = now_s() - t0; ctx->have_event = true; return PHOX_OK; }
    rc = run_launch(ctx, (const Genstep*)d_genstep, ctx->d_prefix.p, (int)ngs, (const Photon*)d_input_photon, photon_offset, photon_offset, n, event_id);
    if (rc) return rc;
    if (mode_keeps_photon(ctx->cfg.event_mode)) { rc = gather_debug(ctx, n); if (rc) return rc; }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(ctx->h_counters, ctx->d_counters.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    ctx->num_photon = n;
    double dt = now_s() - t0;
    ctx->stats.num_photon = (uint64_t)n; ctx->stats.num_hit = (uint64_t)ctx->num_hit; ctx->stats.num_ray = ctx->h_counters[0]; ctx->stats.num_home_ray = ctx->h_counters[2];
    ctx->stats.launch_seconds = dt;
    if (launch_seconds) *launch_seconds = dt;
    ctx->have_event = true;
    return PHOX_OK;
}

extern "C" int64_t phox_num_photon(const phox_context* ctx) { return ctx && ctx->have_event ? ctx->num_photon : 0; }
extern "C" int64_t phox_num_hit(const phox_context* ctx) { return ctx && ctx->have_event ? ctx->num_hit : 0; }
extern "C" const void* phox_hits_device(const phox_context* ctx) { return ctx && ctx->have_event && ctx->num_hit ? ctx->d_hit.p : nullptr; }

extern "C" int phox_get_hits(phox_context* ctx, void* dst) {
    if (!ctx) return PHOX_E_ARG;
    if (!ctx->have_event) return ctx->fail(PHOX_E_STATE, "phox_get_hits: no event");
    if (ctx->num_hit == 0) return PHOX_OK;
    if (!dst) return ctx->fail(PHOX_E_ARG, "phox_get_hits: null destination");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dst, ctx->d_hit.p, (size_t)ctx->num_hit * 64, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PHOX_OK;
}

extern "C" int phox_get_hits_device(phox_context* ctx, void* d_dst) {
    if (!ctx) return PHOX_E_ARG;
    if (!ctx->have_event) return ctx->fail(PHOX_E_STATE, "phox_get_hits_device: no event");
    if (ctx->num_hit == 0) return PHOX_OK;
    if (!d_dst) return ctx->fail(PHOX_E_ARG, "phox_get_hits_device: null destination");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(d_dst, ctx->d_hit.p, (size_t)ctx->num_hit * 64, cudaMemcpyDeviceToDevice, ctx->stream));
    return PHOX_OK;
}

// Hits of the last event to host memory WITHOUT blocking the next event: a device-to-device copy into one of two staging
// buffers on the launch stream, then the device-to-host copy on a second stream.  The caller may start the next
// phox_simulate at once; phox_hits_wait() returns when every copy posted so far has landed.
extern "C" int phox_get_hits_async(phox_context* ctx, void* dst) {
    if (!ctx) return PHOX_E_ARG;
    if (!ctx->have_event) return ctx->fail(PHOX_E_STATE, "phox_get_hits_async: no event");
    if (ctx->num_hit == 0) return PHOX_OK;
    if (!dst) return ctx->fail(PHOX_E_ARG, "phox_get_hits_async: null destination");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ev_stage, cudaEventDisableTiming));
        for (int k = 0; k < 2; k++) CK(cudaEventCreateWithFlags(&ctx->ev_copied[k], cudaEventDisableTiming));
    }
    const int k = ctx->stage_idx;
    CK(cudaEventSynchronize(ctx->ev_copied[k]));                  // the copy that last used this staging buffer (two events ago) is done
    if ((size_t)ctx->num_hit > ctx->d_hit_stage[k].cap) CK(ctx->d_hit_stage[k].reserve((size_t)ctx->num_hit + (size_t)ctx->num_hit / 4 + 1024));
    const size_t bytes = (size_t)ctx->num_hit * 64;
    CK(cudaMemcpyAsync(ctx->d_hit_stage[k].p, ctx->d_hit.p, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->ev_stage, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_stage, 0));
    CK(cudaMemcpyAsync(dst, ctx->d_hit_stage[k].p, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CK(cudaEventRecord(ctx->ev_copied[k], ctx->copy_stream));
    ctx->stage_idx ^= 1;
    return PHOX_OK;
}

extern "C" int phox_hits_wait(phox_context* ctx) {
    if (!ctx) return PHOX_E_ARG;
    if (!ctx->copy_stream) return PHOX_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    return PHOX_OK;
}

// page-locked host memory for the hit buffers of a multi-GPU host (any context's device may copy into it)
extern "C" void* phox_host_alloc(int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0 || cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void phox_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int phox_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int64_t phox_get_array(phox_context* ctx, const char* name, void* dst, int64_t dst_bytes) {
    if (!ctx || !name) return PHOX_E_ARG;
    if (!ctx->have_event) return ctx->fail(PHOX_E_STATE, "phox_get_array: no event");
    const void* src = nullptr;
    int64_t bytes = 0;
    std::string n(name);
    if (n == "photon") { src = ctx->h_photon.data(); bytes = (int64_t)ctx->h_photon.size() * 64; }
    else if (n == "record") { src = ctx->h_record.data(); bytes = (int64_t)ctx->h_record.size() * 64; }
    else if (n == "seq") { src = ctx->h_seq.data(); bytes = (int64_t)ctx->h_seq.size() * 32; }
    else if (n == "prd") { src = ctx->h_prd.data(); bytes = (int64_t)ctx->h_prd.size() * 32; }
    else if (n == "tag") { src = ctx->h_tag.data(); bytes = (int64_t)ctx->h_tag.size() * 8; }
    else if (n == "flat") { src = ctx->h_flat.data(); bytes = (int64_t)ctx->h_flat.size() * 4; }
    else if (n == "hit") {
        bytes = ctx->num_hit * 64;
        if (!dst) return bytes;
        if (dst_bytes < bytes) return ctx->fail(PHOX_E_ARG, "phox_get_array: destination too small");
        int rc = phox_get_hits(ctx, dst);
        return rc ? rc : bytes;
    } else return ctx->fail(PHOX_E_ARG, "phox_get_array: unknown array name");
    if (!dst) return bytes;
    if (dst_bytes < bytes) return ctx->fail(PHOX_E_ARG, "phox_get_array: destination too small");
    if (bytes) std::memcpy(dst, src, (size_t)bytes);
    return bytes;
}

extern "C" int phox_get_stats(const phox_context* ctx, phox_stats* st) {
    if (!ctx || !st) return PHOX_E_ARG;
    *st = ctx->stats;
    return PHOX_OK;
}

extern "C" void phox_reset(phox_context* ctx) {
    if (!ctx) return;
    ctx->have_event = false;
    ctx->num_photon = 0; ctx->num_hit = 0;
    ctx->h_photon.clear(); ctx->h_record.clear(); ctx->h_seq.clear(); ctx->h_prd.clear(); ctx->h_tag.clear(); ctx->h_flat.clear();
    ctx->h_photon.shrink_to_fit(); ctx->h_record.shrink_to_fit(); ctx->h_seq.shrink_to_fit(); ctx->h_prd.shrink_to_fit();
}

extern "C" int phox_intersect(phox_context* ctx, const float* ray_o_tmin, const float* ray_d, int64_t nray, void* dst_prd, int32_t accel) {
    if (!ctx) return PHOX_E_ARG;
    if (!ctx->have_geometry) return ctx->fail(PHOX_E_STATE, "phox_intersect: geometry not set");
    if (!ray_o_tmin || !ray_d || !dst_prd || nray < 0) return ctx->fail(PHOX_E_ARG, "phox_intersect: bad arguments");
    if (nray == 0) return PHOX_OK;
    CK(cudaSetDevice(ctx->device));
    TmpBuf<float4> b_o, b_d;
    TmpBuf<Prd> b_out;
    CK(b_o.alloc((size_t)nray)); CK(b_d.alloc((size_t)nray)); CK(b_out.alloc((size_t)nray));
    float4 *d_o = b_o.p, *d_d = b_d.p;
    Prd* d_out = b_out.p;
    CK(cudaMemcpyAsync(d_o, ray_o_tmin, (size_t)nray * 16, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_d, ray_d, (size_t)nray * 16, cudaMemcpyHostToDevice, ctx->stream));
    Scene sc;
    sc.geo.node = ctx->d_node.p; sc.geo.plan = ctx->d_plan.p; sc.geo.itra = ctx->d_itra.p;
    sc.prim = ctx->d_prim.p; sc.exact = ctx->d_exact.p; sc.nodes = ctx->d_bvh.p; sc.inst = ctx->d_inst.p; sc.prim_pb = ctx->d_prim_pb.p;
    sc.ninst = ctx->ninst; sc.tlas_root = ctx->tlas_root; sc.accel = accel == PHOX_ACCEL_BVH_NOHOME ? PHOX_ACCEL_BVH : accel;
    sc.home = nullptr; sc.cand = nullptr;            // single rays carry no home
    const int T = 128;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    k_intersect<<<(unsigned)((nray + T - 1) / T), T, 0, ctx->stream>>>(sc, d_o, d_d, (unsigned)nray, ctx->cfg.tmax, d_out);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CK(cudaMemcpyAsync(dst_prd, d_out, (size_t)nray * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->stats.simulate_kernel_seconds = ms * 1e-3;              // device time of the query kernel
    return PHOX_OK;
}

extern "C" int64_t phox_simtrace(phox_context* ctx, const void* genstep, int64_t num_genstep, const void* input_simtrace, int64_t num_input,
                                 void* dst_simtrace, int64_t capacity) {
    if (!ctx) return PHOX_E_ARG;
    if (!ctx->have_geometry) return ctx->fail(PHOX_E_STATE, "phox_simtrace: geometry not set");
    if (!genstep || num_genstep <= 0 || num_input < 0 || capacity < 0) return ctx->fail(PHOX_E_ARG, "phox_simtrace: bad arguments");
    const Genstep* gs = (const Genstep*)genstep;
    std::vector<unsigned long long> prefix((size_t)num_genstep + 1, 0ull);
    int64_t need_input = 0;
    for (int64_t i = 0; i < num_genstep; i++) {
        int code = gs[i].gencode();
        if (code != GS_FRAME && code != GS_INPUT_PHOTON_SIMTRACE) return ctx->fail(PHOX_E_ARG, "phox_simtrace: genstep is neither FRAME nor INPUT_PHOTON_SIMTRACE");
        prefix[i + 1] = prefix[i] + gs[i].numphoton();
        if (code == GS_INPUT_PHOTON_SIMTRACE) need_input = (int64_t)prefix[i + 1];      // slots index the input array directly (qsim.h:2455)
    }
    int64_t n = (int64_t)prefix[num_genstep];
    if (n > 0x7fffffffll) return ctx->fail(PHOX_E_ARG, "phox_simtrace: more than 2^31 rays in one call");
    if (need_input > 0 && (!input_simtrace || num_input < need_input)) return ctx->fail(PHOX_E_ARG, "phox_simtrace: input simtrace array shorter than the gensteps ask for");
    if (!dst_simtrace) return n;                                 // size query
    if (capacity < n) return ctx->fail(PHOX_E_ARG, "phox_simtrace: destination too small");
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    Genstep* d_gs = nullptr; unsigned long long* d_prefix = nullptr; float4 *d_in = nullptr, *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_gs, (size_t)num_genstep * sizeof(Genstep));
    if (e == cudaSuccess) e = cudaMalloc(&d_prefix, prefix.size() * 8);
    if (e == cudaSuccess && need_input > 0) e = cudaMalloc(&d_in, (size_t)need_input * 64);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)n * 64);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_gs, gs, (size_t)num_genstep * sizeof(Genstep), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_prefix, prefix.data(), prefix.size() * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && need_input > 0) e = cudaMemcpyAsync(d_in, input_simtrace, (size_t)need_input * 64, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const phox_config& c = ctx->cfg;
        SimtraceParams S;
        std::memset(&S, 0, sizeof(S));
        S.scene.geo.node = ctx->d_node.p; S.scene.geo.plan = ctx->d_plan.p; S.scene.geo.itra = ctx->d_itra.p;
        S.scene.prim = ctx->d_prim.p; S.scene.exact = ctx->d_exact.p; S.scene.nodes = ctx->d_bvh.p; S.scene.inst = ctx->d_inst.p; S.scene.prim_pb = ctx->d_prim_pb.p;
        S.scene.ninst = ctx->ninst; S.scene.tlas_root = ctx->tlas_root; S.scene.accel = c.accel == PHOX_ACCEL_BVH_NOHOME ? PHOX_ACCEL_BVH : c.accel;
        S.scene.home = nullptr; S.scene.cand = nullptr;
        S.genstep = d_gs; S.gs_prefix = d_prefix; S.num_genstep = (int)num_genstep;
        S.input = d_in; S.input_base = 0; S.photon_offset = 0; S.num = (unsigned)n;
        S.tmin = c.propagate_epsilon; S.tmax = c.tmax; S.refine_distance = c.refine_distance; S.refine = c.propagate_refine;
        S.seed = c.rng_seed; S.rng_offset = c.rng_offset;
        S.out = d_out;
        const int T = 128;
        k_simtrace<<<(unsigned)((n + T - 1) / T), T, 0, ctx->stream>>>(S);
        e = cudaGetLastError();
        ctx->stats.num_kernel += 1;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst_simtrace, d_out, (size_t)n * 64, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_gs); cudaFree(d_prefix); cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) return ctx->cuda_fail(e, "phox_simtrace");
    return n;
}

extern "C" int64_t phox_merge_hits(phox_context* ctx, float time_window, void* dst, int64_t capacity) {
    if (!ctx) return PHOX_E_ARG;
    if (!ctx->have_event) return ctx->fail(PHOX_E_STATE, "phox_merge_hits: no event");
    if (!(time_window >= 0.f) || capacity < 0) return ctx->fail(PHOX_E_ARG, "phox_merge_hits: bad arguments");
    int64_t n = ctx->num_hit, m = 0;
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_merged.reserve((size_t)n));
    int nk = 0;
    // the hits are already the hit-mask selection (all bits); the any-bit select of the reference's merge is a no-op on them
    CK(merge_photons(ctx->d_hit.p, n, 0u, time_window, ctx->d_merged.p, &m, ctx->merge_scratch, ctx->stream, &nk));
    ctx->stats.num_kernel += (uint64_t)nk;
    if (!dst) return m;
    if (capacity < m) return ctx->fail(PHOX_E_ARG, "phox_merge_hits: destination too small");
    if (m > 0) {
        CK(cudaMemcpyAsync(dst, ctx->d_merged.p, (size_t)m * sizeof(Photon), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return m;
}

extern "C" int phox_get_hits_lite(phox_context* ctx, void* dst) {
    if (!ctx) return PHOX_E_ARG;
    if (!ctx->have_event) return ctx->fail(PHOX_E_STATE, "phox_get_hits_lite: no event");
    if (!ctx->cfg.mode_lite) return ctx->fail(PHOX_E_STATE, "phox_get_hits_lite: mode_lite is off");
    if (ctx->num_hit == 0) return PHOX_OK;
    if (!dst) return ctx->fail(PHOX_E_ARG, "phox_get_hits_lite: null destination");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dst, ctx->d_hitlite.p, (size_t)ctx->num_hit * sizeof(PhotonLite), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PHOX_OK;
}

extern "C" int64_t phox_merge_hits_lite(phox_context* ctx, float time_window, void* dst, int64_t capacity) {
    if (!ctx) return PHOX_E_ARG;
    if (!ctx->have_event) return ctx->fail(PHOX_E_STATE, "phox_merge_hits_lite: no event");
    if (!ctx->cfg.mode_lite) return ctx->fail(PHOX_E_STATE, "phox_merge_hits_lite: mode_lite is off");
    if (!(time_window >= 0.f) || capacity < 0) return ctx->fail(PHOX_E_ARG, "phox_merge_hits_lite: bad arguments");
    int64_t n = ctx->num_hit, m = 0;
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_merged_lite.reserve((size_t)n));
    int nk = 0;
    CK(merge_photons_lite(ctx->d_hitlite.p, n, 0u, time_window, ctx->d_merged_lite.p, &m, ctx->merge_scratch, ctx->stream, &nk));
    ctx->stats.num_kernel += (uint64_t)nk;
    if (!dst) return m;
    if (capacity < m) return ctx->fail(PHOX_E_ARG, "phox_merge_hits_lite: destination too small");
    if (m > 0) {
        CK(cudaMemcpyAsync(dst, ctx->d_merged_lite.p, (size_t)m * sizeof(PhotonLite), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return m;
}

extern "C" int64_t phox_merge(phox_context* ctx, const void* photons, int64_t n, uint32_t select_mask, float time_window, void* dst, int64_t capacity) {
    if (!ctx) return PHOX_E_ARG;
    if (n < 0 || (n > 0 && !photons) || !(time_window >= 0.f) || !dst || capacity < 0) return ctx->fail(PHOX_E_ARG, "phox_merge: bad arguments");
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_merge_in.reserve((size_t)n));
    CK(ctx->d_merged.reserve((size_t)n));
    CK(cudaMemcpyAsync(ctx->d_merge_in.p, photons, (size_t)n * sizeof(Photon), cudaMemcpyHostToDevice, ctx->stream));
    int64_t m = 0;
    int nk = 0;
    CK(merge_photons(ctx->d_merge_in.p, n, select_mask, time_window, ctx->d_merged.p, &m, ctx->merge_scratch, ctx->stream, &nk));
    ctx->stats.num_kernel += (uint64_t)nk;
    if (capacity < m) return ctx->fail(PHOX_E_ARG, "phox_merge: destination too small");
    if (m > 0) {
        CK(cudaMemcpyAsync(dst, ctx->d_merged.p, (size_t)m * sizeof(Photon), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return m;
}

extern "C" int phox_boundary_lookup(phox_context* ctx, const float* nm, const uint32_t* line, const uint32_t* k, int64_t n, float* dst) {
    if (!ctx) return PHOX_E_ARG;
    if (!ctx->have_tables) return ctx->fail(PHOX_E_STATE, "phox_boundary_lookup: tables not set");
    if (!nm || !line || !k || !dst || n <= 0) return ctx->fail(PHOX_E_ARG, "phox_boundary_lookup: bad arguments");
    CK(cudaSetDevice(ctx->device));
    TmpBuf<float> b_nm; TmpBuf<unsigned> b_line, b_k; TmpBuf<float4> b_out;
    CK(b_nm.alloc((size_t)n)); CK(b_line.alloc((size_t)n)); CK(b_k.alloc((size_t)n)); CK(b_out.alloc((size_t)n));
    CK(cudaMemcpyAsync(b_nm.p, nm, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(b_line.p, line, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(b_k.p, k, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    Tables tb;
    tb.bnd_tex = ctx->bnd_tex; tb.icdf_tex = ctx->icdf_tex; tb.optical = ctx->d_optical.p;
    tb.nx = ctx->nx; tb.ny = ctx->ny; tb.nm0 = ctx->nm0; tb.nms = ctx->nms; tb.hd_factor = ctx->hd_factor;
    tb.inv_ny = ctx->inv_ny; tb.y_fast = ctx->y_fast; tb.need_lposcost = ctx->need_lposcost;
    k_boundary_lookup<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(tb, b_nm.p, b_line.p, b_k.p, (unsigned)n, b_out.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dst, b_out.p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PHOX_OK;
}

extern "C" int phox_rng_sequence(phox_context* ctx, float* dst, int64_t ni, int64_t nv, uint64_t id0, int32_t event_id) {
    if (!ctx) return PHOX_E_ARG;
    if (!dst || ni <= 0 || nv <= 0) return ctx->fail(PHOX_E_ARG, "phox_rng_sequence: bad arguments");
    CK(cudaSetDevice(ctx->device));
    float* d = nullptr;
    CK(cudaMalloc(&d, (size_t)ni * nv * 4));
    const int T = 128;
    k_rng_sequence<<<(unsigned)((ni + T - 1) / T), T, 0, ctx->stream>>>(d, (unsigned)ni, (unsigned)nv, id0, ctx->cfg.rng_seed,
                                                                          ctx->cfg.rng_offset + ctx->cfg.skipahead_event_offset * (uint64_t)event_id);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst, d, (size_t)ni * nv * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return ctx->cuda_fail(e, "phox_rng_sequence");
    return PHOX_OK;
}
