// phox_math.cuh : small vector helpers for the device code.
//
// The arithmetic order of each helper is fixed on purpose: positions/times must agree with the
// reference to 1e-4 relative and history flags bit-exactly when both consume the same random
// stream, so e.g. normalize is v * (1/sqrt(dot)) exactly like sysrap/scuda.h:606-610, and dot is
// the left-to-right sum x*x + y*y + z*z like scuda.h.  nvcc contracts a*b+c into FMA for both code
// bases in the same way when the expression trees match.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace phox {

#define PHOX_D __device__ __forceinline__

constexpr float kPi = 3.14159265358979323846f;      // M_PIf in sysrap/scuda.h
constexpr float kRtMax = 1.e27f;                    // RT_DEFAULT_MAX, CSG/csg_intersect_leaf_head.h

PHOX_D float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
PHOX_D float3 operator+(const float3& a, const float3& b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
PHOX_D float3 operator-(const float3& a, const float3& b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
PHOX_D float3 operator-(const float3& a) { return f3(-a.x, -a.y, -a.z); }
PHOX_D float3 operator*(const float3& a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
PHOX_D float3 operator*(float s, const float3& a) { return f3(s * a.x, s * a.y, s * a.z); }
PHOX_D float3 operator/(const float3& a, float s) { float inv = 1.0f / s; return a * inv; }   // scuda.h float3/float
PHOX_D float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PHOX_D float3 cross(const float3& a, const float3& b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
PHOX_D float length(const float3& v) { return sqrtf(dot(v, v)); }
PHOX_D float3 normalize(const float3& v) { float inv = 1.0f / sqrtf(dot(v, v)); return v * inv; }

PHOX_D float dot2(const float2& a, const float2& b) { return a.x * b.x + a.y * b.y; }
PHOX_D float2 normalize2(const float2& v) { float inv = 1.0f / sqrtf(dot2(v, v)); return make_float2(v.x * inv, v.y * inv); }

// Rotate d so that its z axis lies along unit vector u : CLHEP Hep3Vector::rotateUz as restated
// in sysrap/smath.h:77-95.
PHOX_D void rotate_uz(float3& d, const float3& u) {
    float up = u.x * u.x + u.y * u.y;
    if (up > 0.f) {
        up = sqrtf(up);
        float px = d.x, py = d.y, pz = d.z;
        d.x = (u.x * u.z * px - u.y * py) / up + u.x * pz;
        d.y = (u.y * u.z * px + u.x * py) / up + u.y * pz;
        d.z = -up * px + u.z * pz;
    } else if (u.z < 0.f) {
        d.x = -d.x;
        d.z = -d.z;
    }
}

// 256-bit streaming accesses (sm_100: LDG.E.EF.256 / STG.E.EF.256; the address must be 32 B aligned).  The per-photon records
// are 64 B (sphoton) and 32 B (quad2) per lane at a 64 / 32 B stride: with 128-bit accesses every 32 B sector passes the L1 data
// pipe twice, and that pipe - not DRAM, not the issue slots - was the busiest unit of the physics kernel (l1tex lsu
// wavefronts 45 % of peak, profiles/r2_summary.md).
PHOX_D void ldcs256(const void* p, float4& a, float4& b) {
    asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p) : "memory");
}
PHOX_D void stcs256(void* p, const float4& a, const float4& b) {
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// point / direction through an inverse transform, row-vector convention (sysrap/sqat4.h:45-52)
PHOX_D float3 xform(const float4& r0, const float4& r1, const float4& r2, const float4& r3, const float3& v, float w) {
    float3 o;
    o.x = r0.x * v.x + r1.x * v.y + r2.x * v.z + r3.x * w;
    o.y = r0.y * v.x + r1.y * v.y + r2.y * v.z + r3.y * w;
    o.z = r0.z * v.x + r1.z * v.y + r2.z * v.z + r3.z * w;
    return o;
}
// normal back to the parent frame with the transpose of the same inverse (sqat4.h:96-104);
// w = 0 so the 4th column never contributes.
PHOX_D float3 xform_normal(const float4& r0, const float4& r1, const float4& r2, const float3& n) {
    float3 o;
    o.x = r0.x * n.x + r0.y * n.y + r0.z * n.z + r0.w * 0.f;
    o.y = r1.x * n.x + r1.y * n.y + r1.z * n.z + r1.w * 0.f;
    o.z = r2.x * n.x + r2.y * n.y + r2.z * n.z + r2.w * 0.f;
    return o;
}

}  // namespace phox
