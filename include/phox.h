/* phox.h : C ABI of the B200-native optical-photon propagation engine (libphox.so).
 *
 * This is the drop-in boundary for the reference's simulate path.  Every entry point takes
 * plain pointers and sizes; the arrays are the reference's own array layouts so a binding on the
 * reference side passes its buffers through unchanged (see INTEGRATION.md for the SSimulator
 * adaptor and the ctypes stub).  All functions return 0 on success or a negative PHOX_E_* code;
 * phox_last_error() gives the message.  Nothing aborts or raises signals across this ABI
 * (the reference asserts / raises SIGINT instead: qudarap/QSim.cc:446, CSG/CUDA_CHECK.h).
 *
 * A context owns one GPU.  A context is not thread-safe; separate contexts are independent, which
 * is how the 8 GPUs of a box are driven (one process or thread per context).
 */
#ifndef PHOX_H
#define PHOX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct phox_context phox_context;

enum {
    PHOX_OK = 0,
    PHOX_E_ARG = -1,      /* bad argument / arrays inconsistent */
    PHOX_E_STATE = -2,    /* call order wrong: geometry/tables missing, no event */
    PHOX_E_CUDA = -3,     /* CUDA runtime error, message has the detail */
    PHOX_E_NOMEM = -4,    /* event does not fit the configured max_slot / device memory */
    PHOX_E_NODEVICE = -5  /* no usable CUDA device: there is NO CPU fallback */
};

/* Event modes follow SEventConfig (sysrap/SEventConfig.cc:1528-1642). */
enum {
    PHOX_MODE_MINIMAL = 0,    /* gather hits only                                   */
    PHOX_MODE_HITPHOTON = 1,  /* hits + photon array                                */
    PHOX_MODE_HITPHOTONSEQ = 2, /* hits + photon + seq                              */
    PHOX_MODE_DEBUGLITE = 3,  /* photon, record[max_record], seq, hit               */
    PHOX_MODE_DEBUGHEAVY = 4  /* DebugLite + prd[max_record]                        */
};

/* Random-number consumption pattern of the physics (SURVEY 8a "RNG draws per bounce").
 * The reference's as-built device code is always the DEBUG_TAG variant
 * (CSGOptiX/CMakeLists.txt:54-56) which consumes extra "burn" uniforms to stay aligned with
 * Geant4 (qudarap/qsim.h:730-733, 1088-1098, 1188-1201, 1687-1692). */
enum {
    PHOX_RNG_PRODUCTION = 0,  /* 2 / 1 / 1 / 1 draws: to_boundary / at_boundary / at_surface / detect */
    PHOX_RNG_DEBUG_TAG = 1    /* 4 / 2(+4 on reflect) / 2 / 1 : matches the reference build */
};

/* How the bounce loop is scheduled on the GPU.  Both forms call the same trace and propagate code and
 * give bit-identical results. */
enum {
    PHOX_KERNEL_AUTO = 0,       /* by launch size: persistent below 250 k photons (400 k when the previous launch
                                   averaged < 6 bounces per photon, 2 M below 2.5), wavefront above - and, in the
                                   production event modes, the persistent kernel takes the live list over once it
                                   holds <= 131072 photons (env PHOX_TAIL_PHOTONS, 0 = never); results do not depend on it */
    PHOX_KERNEL_PERSISTENT = 1, /* one fused kernel, persistent warps that refill idle lanes                  */
    PHOX_KERNEL_WAVEFRONT = 2   /* per bounce: trace kernel + physics kernel over the list of live photons    */
};

/* How rays find their nearest CSGPrim. */
enum {
    PHOX_ACCEL_BVH = 0,       /* two-level BVH built on the GPU (instances, then prims per solid) */
    PHOX_ACCEL_BRUTE = 1,     /* loop over every instance and prim: validation of the BVH only   */
    PHOX_ACCEL_BVH_NOHOME = 2 /* the BVH without the home-cell shortcut (validation / A-B timing) */
};

/* Defaults are the reference's: sysrap/SEventConfig.cc:37-115, CSGOptiX/CSGOptiX.cc:651-668,
 * qudarap/QRng.cc:52-54. phox_default_config() fills them in. */
typedef struct phox_config {
    int32_t  max_bounce;            /* OPTICKS_MAX_BOUNCE, 31                                   */
    int32_t  event_mode;            /* PHOX_MODE_*, Minimal                                     */
    int32_t  max_record;            /* record/prd slots per photon in debug modes, <= 32        */
    int32_t  rng_mode;              /* PHOX_RNG_*, DEBUG_TAG                                    */
    int32_t  accel;                 /* PHOX_ACCEL_*                                             */
    uint32_t hit_mask;              /* OPTICKS_HIT_MASK, SD = 0x40                              */
    uint32_t epsilon0_mask;         /* OPTICKS_PROPAGATE_EPSILON0_MASK: TO|CK|SI|SC|RE = 0x37   */
    uint32_t propagate_refine;      /* OPTICKS_PROPAGATE_REFINE, 0                              */
    float    propagate_epsilon;     /* tmin after boundary flags, 0.05 mm                       */
    float    propagate_epsilon0;    /* tmin after the epsilon0_mask flags, 0.05 mm              */
    float    refine_distance;       /* OPTICKS_PROPAGATE_REFINE_DISTANCE, 5000 mm               */
    float    tmax;                  /* ray tmax, 1e6 mm                                         */
    float    max_time;              /* OPTICKS_MAX_TIME, 1e27 ns                                */
    uint32_t kernel_mode;           /* PHOX_KERNEL_*: how the bounce loop is scheduled (results are identical) */
    uint64_t rng_seed;              /* curand_init seed, 0                                      */
    uint64_t rng_offset;            /* curand_init offset (QRng__SEED_OFFSET), 0                */
    uint64_t skipahead_event_offset;/* OPTICKS_EVENT_SKIPAHEAD, 100000 draws per event index    */
    int64_t  max_slot;              /* photons per launch; 0 = 0.87*VRAM/(64*1.75) heuristic
                                       (sysrap/SEventConfig.cc:1897-1903)                       */
    uint32_t mode_lite;             /* OPTICKS_MODE_LITE: 1 = also keep sphotonlite hits (16 B: identity, time, packed local hit
                                       position, flagmask; sysrap/sphotonlite.h) -> phox_get_hits_lite / phox_merge_hits_lite */
    uint32_t reserved0;
} phox_config;

void phox_default_config(phox_config* cfg);

/* Lifecycle.  Replaces CSGOptiX::Create / ~CSGOptiX (CSGOptiX/CSGOptiX.cc:367-392). */
phox_context* phox_create(int device);
void          phox_destroy(phox_context* ctx);
const char*   phox_last_error(const phox_context* ctx);   /* ctx may be NULL: creation errors */
const char*   phox_desc(const phox_context* ctx);         /* SSimulator::desc                 */

/* Geometry: the CSGFoundry arrays (CSG/CSGFoundry.h:253-263; upload CSG/CSGFoundry.cc:3377-3405).
 *   solid : CSGSolid[nsolid] 48 B     prim : CSGPrim[nprim] 64 B     node : CSGNode[nnode] 64 B
 *   plan  : float4[nplan]             itra : qat4[nitra] inverse node transforms
 *   inst  : qat4[ninst] instance transforms, 4th column ints = (ins_idx, gas_idx,
 *           sensor_identifier+1, sensor_index) (sysrap/sqat4.h:345-407)
 * Builds the two-level BVH on the device (replaces SBT::createGAS/createIAS,
 * CSGOptiX/SBT.cc:277-370).  PHOX_E_ARG for inconsistent arrays, a singular instance transform,
 * more than 2^29 prims, or trees deeper than the traversal stack (instance tree + deepest
 * solid tree > 63 levels) - never a silently truncated traversal. */
int phox_set_geometry(phox_context* ctx,
                      const void* solid, int64_t nsolid,
                      const void* prim,  int64_t nprim,
                      const void* node,  int64_t nnode,
                      const void* plan,  int64_t nplan,
                      const void* itra,  int64_t nitra,
                      const void* inst,  int64_t ninst);

/* Physics tables: the SSim arrays that QSim::UploadComponents consumes (qudarap/QSim.cc:134-180).
 *   bnd     : float32 [nbnd,4,2,nwl,4]  boundary texture payload (qudarap/QBnd.cc:130-194)
 *   domain  : wavelength of sample 0 and the step in nm (60, 1 for the 761-sample fine domain)
 *   optical : int32 [4*nbnd,4]          (sysrap/sstandard.h:311-441)
 *   icdf    : float32 [icdf_ny,icdf_nx] scintillation inverse-CDF rows (3 x 4096, hd_factor 20;
 *             qudarap/QScint.cc:84-120) or NULL when no scintillator                      */
int phox_set_tables(phox_context* ctx,
                    const float* bnd, int64_t nbnd, int64_t nwl,
                    float domain_low, float domain_step,
                    const int32_t* optical,
                    const float* icdf, int64_t icdf_ny, int64_t icdf_nx, int32_t hd_factor);

/* Run all work of this context on the caller's CUDA stream (cudaStream_t as void*; NULL restores the
 * context's own stream).  Lets a host framework order and time the engine with its own events -
 * the reference always uses the default stream (CSGOptiX/CSGOptiX.cc:1170-1176). */
int phox_set_stream(phox_context* ctx, void* cuda_stream);

int phox_set_config(phox_context* ctx, const phox_config* cfg);
int phox_get_config(const phox_context* ctx, phox_config* cfg);

/* One event: gensteps in (quad6[ngenstep], host memory), hits out.  Replaces
 * SSimulator::simulate(eventID) / NP* CSGOptiX::simulate(const NP* gs, int eventID)
 * (sysrap/SSimulator.h:30, CSGOptiX/CSGOptiX.cc:798-826, qudarap/QSim.cc:428-617).
 *   input_photon : sphoton[ninput] for an OpticksGenstep_INPUT_PHOTON genstep, else NULL
 *   photon_offset: absolute index of this call's first photon (RNG subsequence and sphoton.index
 *                  use absolute indices: CSGOptiX/CSGOptiX7.cu:415-419), so that a rank handling
 *                  a slice of a bigger event gives results identical to the whole-event run
 * Events bigger than max_slot run as several launches (sysrap/SGenstep.h:249-323).
 * ngenstep == 0 (or gensteps that hold no photons) is a valid EMPTY event - no launch, zero hits, PHOX_OK: what a rank
 * gets when an event has fewer gensteps than ranks.  (The SSimulator adaptor still answers -1. without gensteps like
 * QSim::simulate, qudarap/QSim.cc:446.)
 * On return the launch seconds are in *launch_seconds (may be NULL). */
int phox_simulate(phox_context* ctx,
                  const void* genstep, int64_t ngenstep,
                  const void* input_photon, int64_t ninput,
                  int32_t event_id, uint64_t photon_offset,
                  double* launch_seconds);

/* Same event, but gensteps / input photons are already resident in device memory and hits stay
 * on the device (phox_hits_device).  This is the HBM-resident path that bench.py times as
 * `value`; phox_simulate is the host-buffer path it times as `e2e`. */
int phox_simulate_device(phox_context* ctx,
                         const void* d_genstep, int64_t ngenstep,
                         const void* d_input_photon, int64_t ninput,
                         int32_t event_id, uint64_t photon_offset,
                         double* launch_seconds);

/* Results of the last event; buffers are valid until phox_reset (SEvt::getNumHit / getHit,
 * sysrap/SEvt.cc:4924-4991).  Hits are in ascending absolute photon index. */
int64_t phox_num_photon(const phox_context* ctx);
int64_t phox_num_hit(const phox_context* ctx);
int     phox_get_hits(phox_context* ctx, void* dst_sphoton);      /* host dst, 64 B * num_hit */
const void* phox_hits_device(const phox_context* ctx);             /* device pointer          */
int     phox_get_hits_device(phox_context* ctx, void* d_dst);      /* device dst (e.g. the send buffer
                                                                      of the NCCL hit gather), async on
                                                                      the context's stream            */

/* Hits to host memory without holding up the next event (the multi-GPU host's gather, include/PhoxMultiGPU.h): the
 * records are staged on the device (two buffers) and copied out on a second stream, so phox_simulate of the next event
 * may be called at once.  dst should be page-locked (phox_host_alloc) for the copy to be asynchronous; it must stay
 * valid until phox_hits_wait() - which returns once every copy posted so far has landed - or the second-next call. */
int   phox_get_hits_async(phox_context* ctx, void* dst_sphoton);
int   phox_hits_wait(phox_context* ctx);
void* phox_host_alloc(int64_t bytes);     /* page-locked, usable from every device; NULL on failure */
void  phox_host_free(void* p);
int   phox_device_count(void);            /* CUDA devices visible to this process (0 = none: no CPU path) */

/* Named arrays of the last event, when the event mode keeps them:
 * "photon" (64 B/photon), "record" (64 B * max_record), "seq" (32 B), "prd" (32 B * max_record),
 * "tag" (stag, 32 B: 4-bit consumption tag of each of the first 64 tagged random draws, sysrap/stag.h) and "flat"
 * (sflat, 256 B: those 64 uniforms) in the DebugHeavy mode, "hit".  Returns the byte size when dst is NULL. */
int64_t phox_get_array(phox_context* ctx, const char* name, void* dst, int64_t dst_bytes);

/* Counters of the last event: total bounces (= intersect queries), launches, kernels launched. */
typedef struct phox_stats {
    uint64_t num_photon, num_hit, num_ray, num_launch, num_kernel;
    double   launch_seconds, upload_seconds, gather_seconds;
    double   simulate_kernel_seconds;   /* device time of the simulate kernel(s), CUDA events on the launch stream */
    double   compact_kernel_seconds;    /* device time of hit offset scan + compaction */
    /* filled only while phox_set_profiling(ctx, 1) is in effect (wavefront form): per-kernel device times */
    double   trace_kernel_seconds;      /* sum over the trace kernels of the event */
    double   propagate_kernel_seconds;  /* sum over the physics kernels of the event */
    uint64_t num_trace_launch;          /* trace kernels that had live photons */
    uint64_t num_home_ray;              /* rays of the event settled by the candidate list of their home cell, without the BVH */
} phox_stats;
int phox_get_stats(const phox_context* ctx, phox_stats* st);

/* Per-kernel timing of the bounce loop (CUDA events between the kernels of the wavefront form).  Off by
 * default: the extra event records are kept out of production launches.  Used by bench.py for the roofline
 * of the dominant kernel. */
int phox_set_profiling(phox_context* ctx, int on);

void phox_reset(phox_context* ctx);   /* SSimulator::reset(eventID) */

/* Geometry queries without physics (the simtrace / CSGScan role, CSGOptiX/CSGOptiX7.cu:536-577,
 * CSG/CSGScan.cu:11): nray rays (origin xyz + tmin, direction xyz + unused) -> quad2 prd each. */
int phox_intersect(phox_context* ctx, const float* ray_o_tmin, const float* ray_d, int64_t nray,
                   void* dst_prd, int32_t accel);

/* SSimulator::simtrace (sysrap/SSimulator.h:26; raygen CSGOptiX/CSGOptiX7.cu:536-577): one ray per slot of the
 * gensteps, which must be FRAME (17: center-extent grid gensteps, sysrap/SFrameGenstep.cc:604-735; the ray starts at
 * gs.q1 and takes a random direction in the plane gs.q0.y names, both through the transform in gs.q2..q5,
 * qudarap/qsim.h:2459-2511) or INPUT_PHOTON_SIMTRACE (20: slot i takes position/direction from input_simtrace[i].q0/q1).
 * One trace each (tmin = propagate_epsilon, PropagateRefine honoured); results in sevent::add_simtrace layout
 * (sysrap/sevent.h:670-697), quad4 per slot: q0 normal + t, q1 intersect position + tmin, q2 origin + prim<<16|boundary,
 * q3 direction + iindex<<16|identity; a miss keeps the miss program's q0 = (0.6,0.6,0.6,1) and 0xffffffff in both ints.
 * Returns the number of records (= sum of numphoton), or a negative PHOX_E_* code.  dst_simtrace NULL = size query. */
int64_t phox_simtrace(phox_context* ctx, const void* genstep, int64_t num_genstep, const void* input_simtrace, int64_t num_input,
                      void* dst_simtrace, int64_t capacity);

/* Hit merging by (sensor identity, time bucket): SPM::merge_partial_select (sysrap/SPM.cu:153-290) with the sphoton
 * functors (sysrap/sphoton.h:277-304; driven by QEvt::PerLaunchMerge / QEvt::FinalMerge, qudarap/QEvt.cc:1170-1248).
 * key = (identity << 48) | uint(time / time_window); within a key the earliest photon survives with flagmask = OR and
 * hitcount = sum of the group; output in ascending key order.  time_window 0 = no merging (the selection as it is).
 *   phox_merge_hits : the hits of the current event, on the device (dst NULL = count only).
 *   phox_merge      : any host sphoton array (the FinalMerge of per-launch / per-rank results); select_mask is the
 *                     reference's any-bit flagmask selection, 0 = take every record.
 * Both return the number of merged records or a negative PHOX_E_* code. */
int64_t phox_merge_hits(phox_context* ctx, float time_window, void* dst, int64_t capacity);
int64_t phox_merge(phox_context* ctx, const void* photons, int64_t n, uint32_t select_mask, float time_window, void* dst, int64_t capacity);

/* sphotonlite hits (sysrap/sphotonlite.h: {hitcount<<16|identity, time, lposcost<<16|lposfphi as u16 fractions, flagmask},
 * written by the raygen at CSGOptiX7.cu:455-463 and selected like hits, QEvt::gatherHitLite_).  Needs mode_lite = 1.
 * phox_get_hits_lite copies num_hit records (16 B each) to host memory; phox_merge_hits_lite merges them per
 * (identity, time bucket) with the sphotonlite functors (time = min, flagmask = OR, hitcount = sum, identity and local
 * position of the first of the group in photon order) and returns the merged count (dst NULL = count only). */
int     phox_get_hits_lite(phox_context* ctx, void* dst);
int64_t phox_merge_hits_lite(phox_context* ctx, float time_window, void* dst, int64_t capacity);

/* Precooked random streams (qudarap/QSim.cu:43-68): first nv curand_uniform floats of
 * subsequences [id0, id0+ni). dst is host float32[ni*nv]. */
/* Boundary-table readback through the hardware texture path (qudarap/QSim.cu boundary_lookup_line /
 * QBnd.cu:6 test kernels): n lookups of (wavelength nm, line = 4*boundary + species, k = payload group)
 * -> float4 each.  Used to check the CPU oracle's emulation of the GPU's linear texture filter. */
int phox_boundary_lookup(phox_context* ctx, const float* nm, const uint32_t* line, const uint32_t* k, int64_t n, float* dst_float4);

int phox_rng_sequence(phox_context* ctx, float* dst, int64_t ni, int64_t nv, uint64_t id0,
                      int32_t event_id);

#ifdef __cplusplus
}
#endif
#endif
