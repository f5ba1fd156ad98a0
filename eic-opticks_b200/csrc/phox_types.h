// phox_types.h : plain-old-data layouts shared by the host engine and the sm_100a kernels.
//
// Every struct here is byte-compatible with an array the reference persists or uploads, so a
// caller can hand over the reference's own buffers unchanged:
//
//   Photon  (64 B)  <-> sphoton            sysrap/sphoton.h:171-193
//   Genstep (96 B)  <-> quad6 / storch / scerenkov / sscint   sysrap/squad.h, storch.h:43-70,
//                                           scerenkov.h:29-58, sscint.h:34-62
//   Seq     (32 B)  <-> sseq               sysrap/sseq.h:54-60
//   Prd     (32 B)  <-> quad2              sysrap/squad.h:180-255
//   Node    (64 B)  <-> CSGNode            CSG/CSGNode.h:67-98
//   Prim    (64 B)  <-> CSGPrim            CSG/CSGPrim.h:72-118
//   Solid   (48 B)  <-> CSGSolid           CSG/CSGSolid.h:37-55
//   Qat4    (64 B)  <-> qat4               sysrap/sqat4.h (translation in elements 12..14,
//                                           identity ints in the 4th column :345-407)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PHOX_HD __host__ __device__ __forceinline__
#else
#define PHOX_HD inline
#endif

namespace phox {

// ---- history flags : sysrap/OpticksPhoton.h:22-46 -------------------------------------------
enum : unsigned {
    F_CERENKOV = 1u << 0, F_SCINTILLATION = 1u << 1, F_TORCH = 1u << 2, F_BULK_ABSORB = 1u << 3,
    F_BULK_REEMIT = 1u << 4, F_BULK_SCATTER = 1u << 5, F_SURFACE_DETECT = 1u << 6,
    F_SURFACE_ABSORB = 1u << 7, F_SURFACE_DREFLECT = 1u << 8, F_SURFACE_SREFLECT = 1u << 9,
    F_BOUNDARY_REFLECT = 1u << 10, F_BOUNDARY_TRANSMIT = 1u << 11, F_NAN_ABORT = 1u << 12,
    F_EFFICIENCY_COLLECT = 1u << 13, F_EFFICIENCY_CULL = 1u << 14, F_MISS = 1u << 15
};

// ---- genstep codes : sysrap/OpticksGenstep.h:19-43 -------------------------------------------
enum : int {
    GS_DsG4Scintillation_r4695 = 5, GS_TORCH = 6, GS_CARRIER = 14, GS_CERENKOV = 15,
    GS_SCINTILLATION = 16, GS_G4Cerenkov_modified = 18, GS_INPUT_PHOTON = 19
};

// ---- torch source shapes : sysrap/storchtype.h:8-19 ------------------------------------------
enum : unsigned { T_UNDEF = 0, T_DISC, T_LINE, T_POINT, T_CIRCLE, T_RECTANGLE, T_SPHERE_MARSAGLIA, T_SPHERE };

// ---- CSG typecodes : sysrap/OpticksCSG.h:21-62 ------------------------------------------------
enum : unsigned {
    CSG_ZERO = 0, CSG_UNION = 1, CSG_INTERSECTION = 2, CSG_DIFFERENCE = 3,
    CSG_NODE = 11, CSG_CONTIGUOUS = 11, CSG_DISCONTIGUOUS = 12, CSG_OVERLAP = 13,
    CSG_LEAF = 101, CSG_SPHERE = 101, CSG_ZSPHERE = 103, CSG_CYLINDER = 105, CSG_CONE = 108,
    CSG_BOX3 = 110, CSG_CONVEXPOLYHEDRON = 112, CSG_HYPERBOLOID = 117, CSG_PHICUT = 121,
    CSG_HALFSPACE = 125
};

// ---- optical-buffer "ems" : sysrap/smatsur.h:8-16 ---------------------------------------------
enum : unsigned { EMS_Material = 0, EMS_NoSurface = 1, EMS_Surface = 2, EMS_SensorA = 3, EMS_CustomART = 4, EMS_ZMinus = 5 };

// ---- per-bounce control : sysrap/sflow.h:10-19 --------------------------------------------------
enum : int { FLOW_UNDEFINED = 0, FLOW_BREAK = 1, FLOW_CONTINUE = 2, FLOW_BOUNDARY = 3, FLOW_PASS = 4, FLOW_START = 5 };

// boundary table species order inside one boundary : qudarap/qbnd.h (OMAT OSUR ISUR IMAT)
enum : int { SP_OMAT = 0, SP_OSUR = 1, SP_ISUR = 2, SP_IMAT = 3 };

struct alignas(16) F4 { float x, y, z, w; };
struct alignas(16) U4 { unsigned x, y, z, w; };

struct alignas(16) Photon {           // sphoton
    float    px, py, pz, time;        // q0
    float    mx, my, mz; unsigned hitcount_iindex;      // q1 : hi16 hitcount, lo16 iindex
    float    ex, ey, ez, wavelength;  // q2 : polarization, wavelength nm
    unsigned orient_boundary_flag;    // q3.x : orient<<31 | boundary<<16 | flag
    unsigned identity;                // q3.y : hi8 = index bits 32..39, lo24 = sensor id+1
    unsigned index;                   // q3.z : low 32 bits of absolute photon index
    unsigned flagmask;                // q3.w : OR of every flag the photon carried
};
static_assert(sizeof(Photon) == 64, "Photon must match sphoton");

struct alignas(16) PhotonLite {       // sphotonlite (sysrap/sphotonlite.h)
    unsigned hitcount_identity;       // hi16 hitcount (starts at 1), lo16 sensor identity
    float    time;
    unsigned lposcost_lposfphi;       // hi16 / lo16 : local hit position cos(theta), phi/2pi as u16 fractions
    unsigned flagmask;
};
static_assert(sizeof(PhotonLite) == 16, "PhotonLite must match sphotonlite");

struct alignas(16) Genstep {          // quad6 ; q0.x gencode, q0.z matline, q0.w numphoton
    union { int i[24]; unsigned u[24]; float f[24]; };
    PHOX_HD int gencode() const { return i[0]; }
    PHOX_HD unsigned numphoton() const { return u[3]; }
};
static_assert(sizeof(Genstep) == 96, "Genstep must match quad6");

struct Seq { unsigned long long seqhis[2]; unsigned long long seqbnd[2]; };   // sseq
static_assert(sizeof(Seq) == 32, "Seq must match sseq");

struct alignas(16) Prd {              // quad2
    float nx, ny, nz, t;              // q0 : world-frame normal, distance
    float lposcost, lposfphi;         // q1.xy
    unsigned iindex_identity;         // q1.z : iindex<<16 | identity
    unsigned prim_boundary;           // q1.w : globalPrimIdx<<16 | boundary
};
static_assert(sizeof(Prd) == 32, "Prd must match quad2");

struct alignas(16) Node {             // CSGNode
    union { float f[16]; unsigned u[16]; int i[16]; };
    PHOX_HD unsigned typecode() const { return u[14]; }
    PHOX_HD unsigned boundary() const { return u[6]; }
    PHOX_HD unsigned transform_idx() const { return u[15] & 0x7fffffffu; }   // 1-based, 0 = none
    PHOX_HD bool complement() const { return (u[15] & 0x80000000u) != 0u; }
    PHOX_HD unsigned sub_num() const { return u[0]; }
    PHOX_HD unsigned sub_offset() const { return u[1]; }
};
static_assert(sizeof(Node) == 64, "Node must match CSGNode");

struct alignas(16) Prim {             // CSGPrim
    union { float f[16]; unsigned u[16]; int i[16]; };
    PHOX_HD int num_node() const { return i[0]; }
    PHOX_HD int node_offset() const { return i[1]; }
    PHOX_HD unsigned global_prim_idx() const { return u[15]; }
    // AABB : f[8..10] = min, f[11..13] = max
};
static_assert(sizeof(Prim) == 64, "Prim must match CSGPrim");

struct Solid {                        // CSGSolid
    char label[16];
    int num_prim, prim_offset, type;
    char intent, pad0, pad1, pad2;
    float cx, cy, cz, extent;
};
static_assert(sizeof(Solid) == 48, "Solid must match CSGSolid");

struct alignas(16) Qat4 {             // qat4 : row-vector convention, v' = v * M
    union { float f[16]; int i[16]; unsigned u[16]; };
};
static_assert(sizeof(Qat4) == 64, "Qat4 must match qat4");

}  // namespace phox
