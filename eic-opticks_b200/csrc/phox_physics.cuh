// phox_physics.cuh : photon generation from gensteps and the per-bounce physics.
//
// Restates, with the reference's arithmetic order and random-number consumption:
//   generate_photon dispatch                qudarap/qsim.h:2521-2541
//   torch source shapes                     sysrap/storch.h:189-516
//   Cerenkov sampling                       qudarap/qcerenkov.h:56-119, 139-163, 285-327
//   scintillation sampling / ICDF lookup    qudarap/qscint.h:55-75, 117-145, 207-228
//   carrier                                 sysrap/scarrier.h:47-58
//   boundary table lookup / state fill      qudarap/qbnd.h:103-125, 184-214
//   propagate                               qudarap/qsim.h:2218-2327
//   propagate_to_boundary                   qsim.h:718-863   (+ rayleigh_scatter :601-689)
//   propagate_at_boundary                   qsim.h:998-1205
//   propagate_at_surface / _Detect          qsim.h:1677-1755
//   reflect_diffuse / reflect_specular      qsim.h:1977-2082 (+ lambertian :430-483, marsaglia :551-572)
//
// Differences in HOW: the photon lives in registers as plain float3 fields, the random stream is
// the 8-register Philox of phox_philox.cuh, the material/surface state is fetched lazily (only the
// texture rows a branch really reads), and the DEBUG_TAG "burn" draws of the as-built reference
// kernel are a runtime-uniform switch instead of a compile-time one.
#pragma once
#include "phox_types.h"
#include "phox_math.cuh"
#include "phox_philox.cuh"

namespace phox {

struct Tables {
    cudaTextureObject_t bnd_tex;    // float4 [ny = 8*nbnd][nx = nwl], linear, normalized, wrap
    cudaTextureObject_t icdf_tex;   // float  [3][4096], linear, normalized, wrap
    const uint4* optical;           // [4*nbnd]
    unsigned nx, ny;
    float nm0, nms;                 // wavelength of sample 0, step
    unsigned hd_factor;
    float inv_ny;                   // 1 / float(ny), correctly rounded
    unsigned y_fast;                // 1: bnd_y's reciprocal form was checked against the division for every iy < ny
    unsigned need_lposcost;         // 1: some optical row has ems >= 2, i.e. propagate() may read HitInfo::lposcost
};

struct PhotonState {                // sphoton in registers
    float3 pos; float time;
    float3 mom; unsigned hitcount_iindex;
    float3 pol; float wavelength;
    unsigned obf;                   // orient<<31 | boundary<<16 | flag
    unsigned identity, index, flagmask;

    PHOX_D void zero_flags() { obf = 0u; identity = 0u; index = 0u; flagmask = 0u; hitcount_iindex = 0u; }
    PHOX_D void set_flag(unsigned f) { obf = (obf & 0xffff0000u) | (f & 0xffffu); flagmask |= f; }
    PHOX_D unsigned flag() const { return obf & 0xffffu; }
    PHOX_D unsigned boundary() const { return (obf & 0x7fff0000u) >> 16; }
    PHOX_D void set_index(unsigned long long full) {                     // sphoton.h:224-232
        index = (unsigned)(full & 0xffffffffull);
        identity = ((unsigned)((full >> 32) & 0xffull) << 24) | (identity & 0xffffffu);
    }
    PHOX_D void set_prd(unsigned bnd, unsigned ident, float cosTheta, unsigned iindex) {   // sphoton.h:482-488
        obf = (obf & 0x8000ffffu) | ((bnd & 0x7fffu) << 16);
        identity = (identity & 0xff000000u) | (ident & 0x00ffffffu);
        obf = (obf & 0x7fffffffu) | ((cosTheta < 0.f ? 1u : 0u) << 31);
        hitcount_iindex = 0x00010000u | (iindex & 0xffffu);
    }
    PHOX_D void load(const Photon* src) {
        const float4* s = reinterpret_cast<const float4*>(src);
        float4 a = __ldg(s), b = __ldg(s + 1), c = __ldg(s + 2), d = __ldg(s + 3);
        pos = f3(a.x, a.y, a.z); time = a.w;
        mom = f3(b.x, b.y, b.z); hitcount_iindex = __float_as_uint(b.w);
        pol = f3(c.x, c.y, c.z); wavelength = c.w;
        obf = __float_as_uint(d.x); identity = __float_as_uint(d.y); index = __float_as_uint(d.z); flagmask = __float_as_uint(d.w);
    }
    PHOX_D void load_rw(const Photon* src) {                 // for slots the same kernel also writes (no read-only path)
        const float4* s = reinterpret_cast<const float4*>(src);
        float4 a = s[0], b = s[1], c = s[2], d = s[3];
        pos = f3(a.x, a.y, a.z); time = a.w;
        mom = f3(b.x, b.y, b.z); hitcount_iindex = __float_as_uint(b.w);
        pol = f3(c.x, c.y, c.z); wavelength = c.w;
        obf = __float_as_uint(d.x); identity = __float_as_uint(d.y); index = __float_as_uint(d.z); flagmask = __float_as_uint(d.w);
    }
    // streaming variants (ld/st.global.cs): the wavefront kernels touch every record once per bounce; marking the
    // stream evict-first keeps L1/L2 for what is re-used (BVH nodes, tables, local-memory frames)
    PHOX_D void load_cs(const Photon* src) {
        const float4* s = reinterpret_cast<const float4*>(src);
        float4 a, b, c, d;
        ldcs256(s, a, b); ldcs256(s + 2, c, d);
        pos = f3(a.x, a.y, a.z); time = a.w;
        mom = f3(b.x, b.y, b.z); hitcount_iindex = __float_as_uint(b.w);
        pol = f3(c.x, c.y, c.z); wavelength = c.w;
        obf = __float_as_uint(d.x); identity = __float_as_uint(d.y); index = __float_as_uint(d.z); flagmask = __float_as_uint(d.w);
    }
    PHOX_D void store_cs(Photon* dst) const {
        float4* o = reinterpret_cast<float4*>(dst);
        stcs256(o, make_float4(pos.x, pos.y, pos.z, time), make_float4(mom.x, mom.y, mom.z, __uint_as_float(hitcount_iindex)));
        stcs256(o + 2, make_float4(pol.x, pol.y, pol.z, wavelength),
                make_float4(__uint_as_float(obf), __uint_as_float(identity), __uint_as_float(index), __uint_as_float(flagmask)));
    }
    PHOX_D void store(Photon* dst) const {
        float4* o = reinterpret_cast<float4*>(dst);
        o[0] = make_float4(pos.x, pos.y, pos.z, time);
        o[1] = make_float4(mom.x, mom.y, mom.z, __uint_as_float(hitcount_iindex));
        o[2] = make_float4(pol.x, pol.y, pol.z, wavelength);
        o[3] = make_float4(__uint_as_float(obf), __uint_as_float(identity), __uint_as_float(index), __uint_as_float(flagmask));
    }
};

struct HitInfo {                    // quad2 prd in registers
    float3 normal; float t;
    float lposcost, lposfphi;
    unsigned iindex_identity;
    unsigned prim_boundary;
    PHOX_D unsigned boundary() const { return prim_boundary & 0xffffu; }
    PHOX_D unsigned identity() const { return iindex_identity & 0xffffu; }
    PHOX_D unsigned iindex() const { return iindex_identity >> 16; }
};

// stagr (sysrap/stag.h:231-262): which random draw was consumed where, for aligning the stream with Geant4 (DebugHeavy mode,
// as-built DEBUG_TAG flag): 4-bit tag per draw, 16 per u64, 4 u64 (stag) + the 64 uniforms themselves (sflat).
enum : unsigned {
    TAG_to_sci = 1, TAG_to_bnd = 2, TAG_to_sca = 3, TAG_to_abs = 4, TAG_at_burn_sf_sd = 5, TAG_at_ref = 6, TAG_sf_burn = 7, TAG_sc = 8,
    TAG_to_ree = 9, TAG_re_wl = 10, TAG_re_mom_ph = 11, TAG_re_mom_ct = 12, TAG_re_pol_ph = 13, TAG_re_pol_ct = 14
};
struct Tagr {
    unsigned long long* tag;        // [4] of this photon (zeroed before the event)
    float* flat;                    // [64] of this photon
    unsigned slot;
    PHOX_D void add(unsigned t, float f) {
        if (slot < 64u) {
            tag[slot >> 4] |= (unsigned long long)(t & 0xfu) << (4u * (slot & 15u));
            flat[slot] = f;
        }
        slot += 1u;
    }
};

// ---- tables ----------------------------------------------------------------------------------
// qbnd::boundary_lookup (qudarap/qbnd.h:103-132) in two halves: the x coordinate depends on the wavelength only, so a
// bounce works it out once for its three or four fetches; the y coordinate is (iy + 0.5) / ny, for which phox_set_tables
// checks over every iy < ny that the three-instruction form below (one product with the correctly rounded reciprocal and
// one exact-residual correction) returns the bits of the IEEE division, and otherwise leaves y_fast at 0.
PHOX_D float bnd_x(const Tables& tb, float nm) {
    float fx = (nm - tb.nm0) / tb.nms;
    return (fx + 0.5f) / float(tb.nx);
}
PHOX_D float bnd_y(const Tables& tb, unsigned iy) {
    const float a = float(iy) + 0.5f, c = float(tb.ny);
    if (tb.y_fast) {
        const float q = __fmul_rn(a, tb.inv_ny);
        return __fmaf_rn(__fmaf_rn(-q, c, a), tb.inv_ny, q);
    }
    return a / c;
}
PHOX_D float4 bnd_fetch(const Tables& tb, float x, unsigned line, unsigned k) {
    return tex2D<float4>(tb.bnd_tex, x, bnd_y(tb, 2u * line + k));
}
PHOX_D float4 bnd_lookup(const Tables& tb, float nm, unsigned line, unsigned k) {
    return bnd_fetch(tb, bnd_x(tb, nm), line, k);
}

PHOX_D float icdf_wavelength(const Tables& tb, float u0) {       // qscint::wavelength
    constexpr float y0 = 0.5f / 3.f, y1 = 1.5f / 3.f, y2 = 2.5f / 3.f;
    float wl;
    switch (tb.hd_factor) {
        case 0: wl = tex2D<float>(tb.icdf_tex, u0, y0); break;
        case 10:
            if (u0 < 0.1f) wl = tex2D<float>(tb.icdf_tex, u0 * 10.f, y1);
            else if (u0 > 0.9f) wl = tex2D<float>(tb.icdf_tex, (u0 - 0.9f) * 10.f, y2);
            else wl = tex2D<float>(tb.icdf_tex, u0, y0);
            break;
        case 20:
            if (u0 < 0.05f) wl = tex2D<float>(tb.icdf_tex, u0 * 20.f, y1);
            else if (u0 > 0.95f) wl = tex2D<float>(tb.icdf_tex, (u0 - 0.95f) * 20.f, y2);
            else wl = tex2D<float>(tb.icdf_tex, u0, y0);
            break;
        default: wl = 0.f;
    }
    return wl;
}

PHOX_D float3 uniform_sphere(float u0, float u1) {               // qsim.h uniform_sphere(u0,u1)
    float phi = u0 * 2.f * kPi;
    float cosTheta = 2.f * u1 - 1.f;
    float sinTheta = sqrtf(1.f - cosTheta * cosTheta);
    return f3(cosf(phi) * sinTheta, sinf(phi) * sinTheta, cosTheta);
}

// ---- generators ------------------------------------------------------------------------------
PHOX_D void generate_torch(PhotonState& p, Philox& rng, const Genstep& gs, unsigned long long photon_id) {
    const float* f = gs.f;
    unsigned numphoton = gs.u[3];
    float3 gpos = f3(f[4], f[5], f[6]); float gtime = f[7];
    float3 gmom = f3(f[8], f[9], f[10]);
    float gwavelength = f[15];
    float2 zenith = make_float2(f[16], f[17]);
    float2 azimuth = make_float2(f[18], f[19]);
    float radius = f[20], distance = f[21];
    unsigned type = gs.u[23];

    p.wavelength = gwavelength;
    p.time = gtime;
    if (type == T_DISC) {
        p.mom = gmom;
        float u_zenith = zenith.x + rng.uniform() * (zenith.y - zenith.x);
        float u_azimuth = azimuth.x + rng.uniform() * (azimuth.y - azimuth.x);
        float r = radius * u_zenith;
        float phi = 2.f * kPi * u_azimuth;
        float sinPhi = sinf(phi), cosPhi = cosf(phi);
        p.pos = f3(r * cosPhi, r * sinPhi, 0.f);
        rotate_uz(p.pos, p.mom);
        p.pos = p.pos + gpos;
        p.pol = f3(sinPhi, -cosPhi, 0.f);
        rotate_uz(p.pol, p.mom);
    } else if (type == T_SPHERE) {
        float u_zenith = zenith.x + rng.uniform() * (zenith.y - zenith.x);
        float u_azimuth = azimuth.x + rng.uniform() * (azimuth.y - azimuth.x);
        float phi = 2.f * kPi * u_azimuth;
        float sinPhi = sinf(phi), cosPhi = cosf(phi);
        float cosTheta = 1.f - 2.0f * u_zenith;
        float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
        float flip = copysignf(1.f, radius);
        p.mom = f3(flip * sinTheta * cosPhi, flip * sinTheta * sinPhi, flip * cosTheta);
        float ar = fabsf(radius);
        p.pos = f3(sinTheta * cosPhi * ar, sinTheta * sinPhi * ar, cosTheta * ar);
        float phase = 2.f * kPi * distance;              // distance field doubles as polarization phase fraction
        p.pol = f3(cosf(phase), sinf(phase), 0.f);
        rotate_uz(p.pol, p.mom);
    } else if (type == T_SPHERE_MARSAGLIA) {
        float u, v, b, a;
        do {
            float u0 = zenith.x + rng.uniform() * (zenith.y - zenith.x);
            float u1 = azimuth.x + rng.uniform() * (azimuth.y - azimuth.x);
            u = 2.f * u0 - 1.f;
            v = 2.f * u1 - 1.f;
            b = u * u + v * v;
        } while (b > 1.f);
        a = 2.f * sqrtf(1.f - b);
        float ar = fabsf(radius);
        float flip = copysignf(1.f, radius);
        p.mom = f3(flip * a * u, flip * a * v, flip * (2.f * b - 1.f));
        p.pos = f3(a * u * ar, a * v * ar, (2.f * b - 1.f) * ar);
        float phase = 2.f * kPi * distance;
        p.pol = f3(cosf(phase), sinf(phase), 0.f);
        rotate_uz(p.pol, p.mom);
    } else if (type == T_LINE) {
        p.mom = gmom;
        float frac = float(photon_id) / float(numphoton);
        float sfrac = 2.f * (frac - 0.5f);
        float r = radius * sfrac;
        p.pos = f3(r, 0.f, 0.f);
        rotate_uz(p.pos, p.mom);
        p.pos = p.pos + gpos;
        p.pol = f3(0.f, -1.f, 0.f);
        rotate_uz(p.pol, p.mom);
    } else if (type == T_POINT) {
        p.mom = gmom;
        p.pos = f3(0.f, 0.f, 0.f);
        rotate_uz(p.pos, p.mom);
        p.pos = p.pos + gpos;
        p.pol = f3(0.f, -1.f, 0.f);
        rotate_uz(p.pol, p.mom);
    } else if (type == T_CIRCLE) {
        float ff = float(photon_id) / float(numphoton);
        float frac = azimuth.x * (1.f - ff) + azimuth.y * (ff);
        float phi = 2.f * kPi * frac;
        float sinPhi = sinf(phi), cosPhi = cosf(phi);
        float r = radius < 0.f ? -radius : radius;
        p.mom = f3(radius < 0.f ? -cosPhi : cosPhi, 0.f, radius < 0.f ? -sinPhi : sinPhi);
        p.pos = f3(r * cosPhi, 0.f, r * sinPhi);
        p.pos = p.pos + gpos;
        p.pol = f3(0.f, -1.f, 0.f);
        rotate_uz(p.pol, p.mom);
    } else if (type == T_RECTANGLE) {
        int side_size = (int)(numphoton / 4u);
        int side = (int)(photon_id / (unsigned long long)side_size);
        int side_offset = side * side_size;
        int side_index = (int)photon_id - side_offset;
        float frac = float(side_index) / float(side_size);
        if (side == 0 || side == 1) {
            p.pos = f3(side == 0 ? azimuth.x : azimuth.y, 0.f, (1.f - frac) * zenith.x + frac * zenith.y);
            p.mom = f3(side == 0 ? 1.f : -1.f, 0.f, 0.f);
        } else if (side == 2 || side == 3) {
            p.pos = f3((1.f - frac) * azimuth.x + frac * azimuth.y, 0.f, side == 2 ? zenith.x : zenith.y);
            p.mom = f3(0.f, 0.f, side == 2 ? 1.f : -1.f);
        }
        p.pos = p.pos + gpos;
        p.pol = f3(0.f, -1.f, 0.f);
        rotate_uz(p.pol, p.mom);
    }
    p.zero_flags();
    p.set_flag(F_TORCH);
}

PHOX_D void generate_cerenkov(PhotonState& p, Philox& rng, const Genstep& gs, const Tables& tb) {
    const float* f = gs.f;
    unsigned matline = gs.u[2];
    float3 gpos = f3(f[4], f[5], f[6]); float gtime = f[7];
    float3 delta_pos = f3(f[8], f[9], f[10]); float step_length = f[11];
    float preVelocity = f[15], BetaInverse = f[16], Wmin = f[17], Wmax = f[18];
    float maxSin2 = f[20], Mean1 = f[21], Mean2 = f[22], postVelocity = f[23];

    float3 p0 = normalize(delta_pos);

    // wavelength by rejection against the material RINDEX held in the boundary texture
    float wavelength, cosTheta, sin2Theta, u_maxSin2;
    unsigned count = 0;
    do {
        float u0 = rng.uniform();
        float w = Wmin + u0 * (Wmax - Wmin);
        wavelength = Wmin * Wmax / w;                      // flat in energy
        float sampledRI = bnd_lookup(tb, wavelength, matline, 0u).x;
        cosTheta = BetaInverse / sampledRI;
        sin2Theta = fmaxf(0.f, (1.f - cosTheta) * (1.f + cosTheta));
        float u1 = rng.uniform();
        u_maxSin2 = u1 * maxSin2;
        count += 1;
    } while (u_maxSin2 > sin2Theta && count < 100);

    float sinTheta = sqrtf(sin2Theta);
    float u0 = rng.uniform();
    float phi = 2.f * kPi * u0;
    float sinPhi = sinf(phi), cosPhi = cosf(phi);
    p.mom = f3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
    rotate_uz(p.mom, p0);
    p.pol = f3(cosTheta * cosPhi, cosTheta * sinPhi, -sinTheta);
    rotate_uz(p.pol, p0);
    p.wavelength = wavelength;

    // position along the step
    float fraction, delta, N, NumberOfPhotons;
    float MeanMax = fmaxf(Mean1, Mean2);
    float DeltaN = (Mean1 - Mean2);
    do {
        fraction = rng.uniform();
        delta = fraction * step_length;
        NumberOfPhotons = Mean1 - fraction * DeltaN;
        float u = rng.uniform();
        N = u * MeanMax;
    } while (N > NumberOfPhotons);

    float midVelocity = preVelocity + fraction * (postVelocity - preVelocity) * 0.5f;
    p.time = gtime + delta / midVelocity;
    p.pos = gpos + fraction * delta_pos;
    p.zero_flags();
    p.set_flag(F_CERENKOV);
}

PHOX_D void generate_scint(PhotonState& p, Philox& rng, const Genstep& gs, const Tables& tb) {
    const float* f = gs.f;
    float3 gpos = f3(f[4], f[5], f[6]); float gtime = f[7];
    float3 delta_pos = f3(f[8], f[9], f[10]); float step_length = f[11];
    float charge = f[13], meanVelocity = f[15], ScintillationTime = f[20];

    float u0 = rng.uniform(), u1 = rng.uniform(), u2 = rng.uniform(), u3 = rng.uniform();
    float cost = 1.f - 2.f * u0;
    float sint = sqrtf((1.f - cost) * (1.f + cost));
    float phi = 2.f * kPi * u1;
    float sinp = sinf(phi), cosp = cosf(phi);
    p.mom = f3(sint * cosp, sint * sinp, cost);
    p.pol = f3(cost * cosp, cost * sinp, -sint);
    phi = 2.f * kPi * u2;
    sinp = sinf(phi); cosp = cosf(phi);
    p.pol = normalize(cosp * p.pol + sinp * cross(p.mom, p.pol));
    p.wavelength = icdf_wavelength(tb, u3);

    float fraction = charge == 0.f ? 1.f : rng.uniform();
    p.pos = gpos + fraction * delta_pos;
    float u4 = rng.uniform();
    float deltaTime = fraction * step_length / meanVelocity - ScintillationTime * logf(u4);
    p.time = gtime + deltaTime;
    p.zero_flags();
    p.set_flag(F_SCINTILLATION);
}

PHOX_D void generate_carrier(PhotonState& p, const Genstep& gs, unsigned long long photon_id) {
    const float* f = gs.f;
    p.pos = f3(f[8], f[9] + float(photon_id) * 10.f, f[10]); p.time = f[11];
    p.mom = f3(f[12], f[13], f[14]); p.hitcount_iindex = gs.u[15];
    p.pol = f3(f[16], f[17], f[18]); p.wavelength = f[19];
    p.obf = gs.u[20]; p.identity = gs.u[21]; p.index = gs.u[22]; p.flagmask = gs.u[23];
    p.set_flag(F_TORCH);
}

// qsim::generate_photon : input photons are indexed by the absolute photon id like the reference
// out of line: runs once per photon, must not sit in the instruction stream of the bounce loop
__device__ __noinline__ void generate_photon(PhotonState& p, Philox& rng, const Genstep& gs, const Tables& tb,
                            const Photon* input_photon, unsigned long long input_base, unsigned long long photon_id) {
    switch (gs.gencode()) {
        case GS_CARRIER: generate_carrier(p, gs, photon_id); break;
        case GS_TORCH: generate_torch(p, rng, gs, photon_id); break;
        case GS_G4Cerenkov_modified:
        case GS_CERENKOV: generate_cerenkov(p, rng, gs, tb); break;
        case GS_DsG4Scintillation_r4695:
        case GS_SCINTILLATION: generate_scint(p, rng, gs, tb); break;
        case GS_INPUT_PHOTON: p.load(input_photon + (photon_id - input_base)); p.set_flag(F_TORCH); break;
        default:    // generate_photon_dummy
            p.pos = f3(__int_as_float(1), __int_as_float(2), __int_as_float(3)); p.time = __int_as_float(4);
            p.mom = p.pos; p.hitcount_iindex = 4u;
            p.pol = p.pos; p.wavelength = __int_as_float(4);
            p.obf = 1u; p.identity = 2u; p.index = 3u; p.flagmask = 4u;
            p.set_flag(F_TORCH);
            break;
    }
    p.set_index(photon_id);
}

// ---- bulk + surface physics --------------------------------------------------------------------
PHOX_D void marsaglia_direction(float3& dir, Philox& rng) {
    float u, v, b;
    do {
        float u0 = rng.uniform();
        float u1 = rng.uniform();
        u = 2.f * u0 - 1.f;
        v = 2.f * u1 - 1.f;
        b = u * u + v * v;
    } while (b > 1.f);
    float a = 2.f * sqrtf(1.f - b);
    dir = f3(a * u, a * v, 2.f * b - 1.f);
}

// Lambertian reflection about the geometric normal flipped against the incident direction (qsim::reflect_diffuse).
// Every product and sum is spelled out with the round-to-nearest intrinsics, which the compiler never fuses: written with
// plain operators, this block was seen to get a different a*b+c contraction in k_wf_propagate<false, ..> than in the other
// kernels (scripts/form_consistency.py: last-bit differences in mom / pol of exactly the photons that took this branch).
// Pinned like this it is the arithmetic of the -fmad=false build in every kernel of every build.
PHOX_D float dot_rn(const float3& a, const float3& b) { return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z)); }
PHOX_D void diffuse_reflect(float3& mom_io, float3& pol_io, const float3& normal, Philox& rng) {
    const float3 old_mom = mom_io, old_pol = pol_io;
    float3 mom = old_mom;
    const float orient = dot_rn(old_mom, normal) > 0.f ? -1.f : 1.f;
    float ndotv, u;
    int count = 0;
    do {
        count++;
        {   // marsaglia_direction
            float mu, mv, mb;
            do {
                const float u0 = rng.uniform();
                const float u1 = rng.uniform();
                mu = __fadd_rn(__fmul_rn(2.f, u0), -1.f);
                mv = __fadd_rn(__fmul_rn(2.f, u1), -1.f);
                mb = __fadd_rn(__fmul_rn(mu, mu), __fmul_rn(mv, mv));
            } while (mb > 1.f);
            const float ma = __fmul_rn(2.f, sqrtf(__fadd_rn(1.f, -mb)));
            mom = f3(__fmul_rn(ma, mu), __fmul_rn(ma, mv), __fadd_rn(__fmul_rn(2.f, mb), -1.f));
        }
        ndotv = __fmul_rn(dot_rn(mom, normal), orient);
        if (ndotv < 0.f) {
            mom = f3(__fmul_rn(-1.f, mom.x), __fmul_rn(-1.f, mom.y), __fmul_rn(-1.f, mom.z));
            ndotv = __fmul_rn(-1.f, ndotv);
        }
        u = rng.uniform();
    } while (!(u < ndotv) && (count < 1024));
    const float3 diff = f3(__fadd_rn(mom.x, -old_mom.x), __fadd_rn(mom.y, -old_mom.y), __fadd_rn(mom.z, -old_mom.z));
    const float inv = 1.0f / sqrtf(dot_rn(diff, diff));
    const float3 facet_normal = f3(__fmul_rn(diff.x, inv), __fmul_rn(diff.y, inv), __fmul_rn(diff.z, inv));
    const float two_edotn = __fmul_rn(2.f, dot_rn(old_pol, facet_normal));
    mom_io = mom;
    pol_io = f3(__fadd_rn(__fmul_rn(-1.f, old_pol.x), __fmul_rn(two_edotn, facet_normal.x)),
                __fadd_rn(__fmul_rn(-1.f, old_pol.y), __fmul_rn(two_edotn, facet_normal.y)),
                __fadd_rn(__fmul_rn(-1.f, old_pol.z), __fmul_rn(two_edotn, facet_normal.z)));
}

// Rayleigh scattering (qsim::rayleigh_scatter), plain operators like the reference so that nvcc fuses what it fuses there.
// ONE compiled body for every kernel (out of line, on private copies handed over by address: nothing of the caller's state gets
// its address taken): inlined, the persistent and the wavefront kernel were seen to fuse the a*b + c*d terms of the polarisation
// differently once the code around this block changed (scripts/form_diff.py on sphere_leak: last-bit differences in pol of exactly
// the BULK_SCATTER photons).  The path is cold on detector geometries; the call costs nothing where it is not taken.
struct ScatterIO { float3 mom, pol; };
template <bool TAG>
__device__ __noinline__ void rayleigh_scatter_cold(ScatterIO* io, Philox* rng_io, Tagr* tg) {
    const float3 mom = io->mom, pol = io->pol;
    Philox rng = *rng_io;
    float3 direction, polarization;
    bool looping = true;
    do {
        float u0 = rng.uniform(), u1 = rng.uniform(), u2 = rng.uniform(), u3 = rng.uniform(), u4 = rng.uniform();
        if (TAG) { tg->add(TAG_sc, u0); tg->add(TAG_sc, u1); tg->add(TAG_sc, u2); tg->add(TAG_sc, u3); tg->add(TAG_sc, u4); }
        float cosTheta = u0;
        float sinTheta = sqrtf(1.0f - u0 * u0);
        if (u1 < 0.5f) cosTheta = -cosTheta;
        float sinPhi, cosPhi;
        sincosf(2.f * kPi * u2, &sinPhi, &cosPhi);
        direction = f3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
        rotate_uz(direction, mom);
        float constant = -dot(direction, pol);
        polarization = f3(pol.x + constant * direction.x, pol.y + constant * direction.y, pol.z + constant * direction.z);
        if (dot(polarization, polarization) == 0.f) {
            sincosf(2.f * kPi * u3, &sinPhi, &cosPhi);
            polarization = f3(cosPhi, sinPhi, 0.f);
            rotate_uz(polarization, direction);
        } else {
            if (u3 < 0.5f) polarization = -polarization;
        }
        polarization = normalize(polarization);
        float doCosTheta = dot(polarization, pol);
        float doCosTheta2 = doCosTheta * doCosTheta;
        looping = doCosTheta2 < u4;
    } while (looping);
    io->mom = direction;
    io->pol = polarization;
    *rng_io = rng;
}
template <bool TAG>
PHOX_D void rayleigh_scatter(PhotonState& p, Philox& rng, Tagr* tg) {
    ScatterIO io;
    io.mom = p.mom; io.pol = p.pol;
    Philox rc = rng;
    rayleigh_scatter_cold<TAG>(&io, &rc, tg);
    p.mom = io.mom; p.pol = io.pol;
    rng = rc;
}

// One bounce: the photon has a ray hit `h` (normal normalised, world frame).  Returns the flow
// command; on BREAK the photon is finished.  `burn` selects the DEBUG_TAG consumption pattern.
// propagate_core is the physics; propagate_body wraps it so that it always works on private value copies, which makes
// the compiler see the same expression graph (hence the same FMA contraction) at every site it is compiled into:
// k_wf_propagate, k_simulate and their debug instantiations give bit-identical photons (tests/test_parity_gpu.py checks
// it on every build; measured: +23 % on k_wf_propagate against the former out-of-line call through local memory).
// TAG = true is a second body that also records every tagged draw (qsim.h tagr.add sites); it runs only in the
// DebugHeavy event mode, out of line (propagate_t<true>), and the production body carries no trace of it.
template <bool TAG>
PHOX_D int propagate_core(PhotonState& p, Philox& rng, const HitInfo& h, const Tables& tb, bool burn, Tagr* tg) {
    const unsigned boundary = h.boundary();
    const float3 normal = h.normal;
    float cosTheta = dot(p.mom, normal);
    p.set_prd(boundary, h.identity(), cosTheta, h.iindex());

    // qbnd::fill_state, fetched lazily
    const int line = boundary * 4;
    const int m1_line = cosTheta > 0.f ? line + SP_IMAT : line + SP_OMAT;
    const int m2_line = cosTheta > 0.f ? line + SP_OMAT : line + SP_IMAT;
    const int su_line = cosTheta > 0.f ? line + SP_ISUR : line + SP_OSUR;
    const float bx = bnd_x(tb, p.wavelength);       // every fetch of this bounce happens before the wavelength can change (re-emission ends it)
    const float4 material1 = bnd_fetch(tb, bx, m1_line, 0);
    const float group_velocity = bnd_fetch(tb, bx, m1_line, 1).x;
    // The fetch the boundary step will want - the far side's refractive index without a surface, the surface row with one - is
    // issued here, next to the material fetches, so that it travels during the bulk step instead of being waited for behind
    // it (7.9 % of the physics kernel's stall samples sat on that wait).  Same texel, same coordinates: same bits.
    const unsigned ems = __ldg(tb.optical + su_line).y;
    const float4 second = bnd_fetch(tb, bx, ems == EMS_NoSurface ? m2_line : su_line, 0);

    unsigned flag = 0u;
    int command;

    // ---- propagate_to_boundary ----
    {
        const float absorption_length = material1.y, scattering_length = material1.z, reemission_prob = material1.w;
        const float distance_to_boundary = h.t;
        if (burn) {
            if (TAG) {
                float u_to_sci = rng.uniform(), u_to_bnd = rng.uniform();
                tg->add(TAG_to_sci, u_to_sci); tg->add(TAG_to_bnd, u_to_bnd);
            } else rng.skip(2u);                                 // burns: stepped over, never generated
        }
        float u_scattering, u_absorption;
        rng.draw2_ahead(u_scattering, u_absorption);             // warp-converged refills; the block of the next draw is cached on return
        if (TAG) { tg->add(TAG_to_sca, u_scattering); tg->add(TAG_to_abs, u_absorption); }
        float scattering_distance = -scattering_length * logf(u_scattering);
        float absorption_distance = -absorption_length * logf(u_absorption);

        command = FLOW_BOUNDARY;
        if (absorption_distance <= scattering_distance) {
            if (absorption_distance <= distance_to_boundary) {
                p.time += absorption_distance / group_velocity;
                p.pos = p.pos + absorption_distance * p.mom;
                float u_reemit = reemission_prob == 0.f ? 2.f : rng.uniform();
                if (TAG) { if (u_reemit != 2.f) tg->add(TAG_to_ree, u_reemit); }
                if (u_reemit < reemission_prob) {
                    float u_re_wavelength = rng.uniform();
                    float u_re_mom_ph = rng.uniform(), u_re_mom_ct = rng.uniform();
                    float u_re_pol_ph = rng.uniform(), u_re_pol_ct = rng.uniform();
                    if (TAG) {
                        tg->add(TAG_re_wl, u_re_wavelength); tg->add(TAG_re_mom_ph, u_re_mom_ph); tg->add(TAG_re_mom_ct, u_re_mom_ct);
                        tg->add(TAG_re_pol_ph, u_re_pol_ph); tg->add(TAG_re_pol_ct, u_re_pol_ct);
                    }
                    p.wavelength = icdf_wavelength(tb, u_re_wavelength);
                    p.mom = uniform_sphere(u_re_mom_ph, u_re_mom_ct);
                    p.pol = normalize(cross(uniform_sphere(u_re_pol_ph, u_re_pol_ct), p.mom));
                    flag = F_BULK_REEMIT;
                    command = FLOW_CONTINUE;
                } else {
                    flag = F_BULK_ABSORB;
                    command = FLOW_BREAK;
                }
            }
        } else {
            if (scattering_distance <= distance_to_boundary) {
                p.time += scattering_distance / group_velocity;
                p.pos = p.pos + scattering_distance * p.mom;
                rayleigh_scatter<TAG>(p, rng, tg);
                flag = F_BULK_SCATTER;
                command = FLOW_CONTINUE;
            }
        }
        if (command == FLOW_BOUNDARY) {
            p.pos = p.pos + distance_to_boundary * p.mom;
            p.time += distance_to_boundary / group_velocity;
        }
    }

    if (command == FLOW_BOUNDARY) {
        bool at_surface = false;
        if (ems == EMS_NoSurface) {
            // ---- propagate_at_boundary : Fresnel reflect / refract ----
            const float n1 = material1.x;
            const float n2 = second.x;
            const float eta = n1 / n2;
            const float _c1 = -dot(p.mom, normal);
            const float3 on = _c1 < 0.f ? -normal : normal;          // oriented against the incident direction
            const float3 trans = cross(p.mom, on);
            const float trans_length = length(trans);
            const bool normal_incidence = trans_length < 1e-6f;
            const float3 A_trans = normal_incidence ? p.pol : trans / trans_length;
            const float E1_perp = dot(p.pol, A_trans);
            const float c1 = fabsf(_c1);
            const float c2c2 = 1.f - eta * eta * (1.f - c1 * c1);
            const bool tir = c2c2 < 0.f;
            const float EdotN = dot(p.pol, on);
            const float c2 = tir ? 0.f : sqrtf(c2c2);
            const float n1c1 = n1 * c1, n2c2 = n2 * c2, n2c1 = n2 * c1, n1c2 = n1 * c2;
            const float2 E1 = normal_incidence ? make_float2(0.f, 1.f) : make_float2(E1_perp, length(p.pol - (E1_perp * A_trans)));
            const float2 E2_t = make_float2(2.f * n1c1 * E1.x / (n1c1 + n2c2), 2.f * n1c1 * E1.y / (n2c1 + n1c2));
            const float2 E2_r = make_float2(E2_t.x - E1.x, (n2 * E2_t.y / n1) - E1.y);
            const float2 RR = normalize2(E2_r);
            const float2 TT = normalize2(E2_t);
            const float TransCoeff = (tir || n1c1 == 0.f) ? 0.f : n2c2 * dot2(E2_t, E2_t) / n1c1;

            if (burn) {
                if (TAG) { const float u_boundary_burn = rng.uniform(); tg->add(TAG_at_burn_sf_sd, u_boundary_burn); }
                else rng.skip(1u);
            }
            const float u_reflect = rng.uniform();
            if (TAG) tg->add(TAG_at_ref, u_reflect);
            const bool reflect = u_reflect > TransCoeff;

            p.mom = reflect ? p.mom + 2.0f * c1 * on : eta * (p.mom) + (eta * c1 - c2) * on;
            const float3 A_paral = normalize(cross(p.mom, A_trans));
            p.pol = normal_incidence
                        ? (reflect ? p.pol * (n2 > n1 ? -1.f : 1.f) : p.pol)
                        : (reflect ? (tir ? -p.pol + 2.f * EdotN * on : RR.x * A_trans + RR.y * A_paral)
                                   : TT.x * A_trans + TT.y * A_paral);
            flag = reflect ? F_BOUNDARY_REFLECT : F_BOUNDARY_TRANSMIT;
            if (burn && reflect) {
                if (TAG) {
                    const float a0 = rng.uniform(), a1 = rng.uniform(), a2 = rng.uniform(), a3 = rng.uniform();
                    tg->add(TAG_to_sci, a0); tg->add(TAG_to_bnd, a1); tg->add(TAG_to_sca, a2); tg->add(TAG_to_abs, a3);
                } else rng.skip(4u);
            }
            command = FLOW_CONTINUE;
        } else if (ems == EMS_Surface) {
            at_surface = true;
        } else if (h.lposcost < 0.f) {
            at_surface = true;
        } else if (ems == EMS_SensorA) {
            rng.skip(1u);
            flag = F_SURFACE_DETECT;
            command = FLOW_BREAK;
        }
        // EMS_CustomART needs the Custom4 PMT model, which the reference build does not enable either
        // (WITH_CUSTOM4 undefined): falls through with flag 0 like the reference.

        if (at_surface) {
            // ---- propagate_at_surface ----
            const float4 surface = second;                            // detect, absorb, specular, diffuse (ems != NoSurface: the surface row)
            const float detect = surface.x, absorb = surface.y, diffuse = surface.w;
            float u_surface = rng.uniform();
            if (TAG) tg->add(TAG_at_burn_sf_sd, u_surface);
            if (burn) {
                if (TAG) { const float u_surface_burn = rng.uniform(); tg->add(TAG_sf_burn, u_surface_burn); }
                else rng.skip(1u);
            }
            command = u_surface < absorb + detect ? FLOW_BREAK : FLOW_CONTINUE;
            if (command == FLOW_BREAK) {
                flag = u_surface < absorb ? F_SURFACE_ABSORB : F_SURFACE_DETECT;
            } else {
                flag = u_surface < absorb + detect + diffuse ? F_SURFACE_DREFLECT : F_SURFACE_SREFLECT;
                if (flag == F_SURFACE_DREFLECT) {
                    diffuse_reflect(p.mom, p.pol, normal, rng);
                } else {
                    const float PdotN = dot(p.mom, normal);
                    p.mom = p.mom - 2.f * PdotN * normal;
                    const float EdotN = dot(p.pol, normal);
                    p.pol = -1.f * (p.pol) + 2.f * EdotN * normal;
                }
            }
        }
    }
    p.set_flag(flag);
    return command;
}

// the production body
// The physics works on private copies of the photon, the random stream and the hit: whatever the caller hands in
// (references into local memory for the out-of-line form, registers for an inlined one), the body itself is compiled
// from the same value-only expression graph.
template <bool TAG>
PHOX_D int propagate_body(PhotonState& p_io, Philox& rng_io, const HitInfo& h_in, const Tables& tb, bool burn, Tagr* tg) {
#if PHOX_PROP_SSA
    PhotonState p = p_io;
    Philox rng = rng_io;
    const HitInfo h = h_in;
    const int command = propagate_core<TAG>(p, rng, h, tb, burn, tg);
    p_io = p;
    rng_io = rng;
    return command;
#else
    return propagate_core<TAG>(p_io, rng_io, h_in, tb, burn, tg);
#endif
}
template <bool TAG>
__device__ __noinline__ int propagate_t(PhotonState& p, Philox& rng, const HitInfo& h, const Tables& tb, bool burn, Tagr* tg) {
    return propagate_body<TAG>(p, rng, h, tb, burn, tg);
}
PHOX_D int propagate(PhotonState& p, Philox& rng, const HitInfo& h, const Tables& tb, bool burn) {
#if PHOX_PROP_INLINE_ALL
    return propagate_body<false>(p, rng, h, tb, burn, nullptr);
#else
    return propagate_t<false>(p, rng, h, tb, burn, nullptr);
#endif
}

}  // namespace phox
