// phox_philox.cuh : counter-based Philox4x32-10 stream, draw-for-draw identical to the curand
// generator the reference uses:
//
//   curand_init(seed, subsequence = absolute photon index, offset, &rng);
//   skipahead(skipahead_event_offset * event_index, &rng);            qudarap/qrng.h:124-148
//   u = curand_uniform(&rng)                                          every draw of the physics
//
// The algorithm is the published Philox4x32 with 10 rounds (Salmon et al., SC'11) with curand's
// conventions (CUDA toolkit 12.9 curand_philox4x32_x.h / curand_kernel.h, the third-party
// dependency named in SURVEY 8c): key = 64-bit seed, counter words (x,y) = 128-bit-block index,
// (z,w) = subsequence; the four 32-bit outputs of a block are handed out in x,y,z,w order and an
// element offset n selects block n/4, lane n%4; a uniform is  u32 * 2^-32 + 2^-33  in (0,1].
//
// Instead of the 64-byte curandStatePhilox4_32_10 the stream caches TWO consecutive blocks (8
// outputs) plus the block counter and a position, so that the refill (10 Philox rounds) can be done
// at points where the warp is converged (align()) instead of inside divergent physics branches.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace phox {

struct Philox {
    uint4    a, b;       // outputs of blocks `blk` and `blk + 1`
    uint32_t blk_lo, blk_hi;
    uint32_t sub_lo, sub_hi;
    uint32_t key_lo, key_hi;
    uint32_t pos;        // next draw, 0..7 into (a,b); 8 = both blocks used up

    // Out of line on purpose: uniform() is expanded at ~30 draw sites and each would otherwise carry its own
    // copy of the 10 rounds; the simulate kernel is instruction-fetch bound (profiles/), so code size matters.
    static __device__ __noinline__ uint4 block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
        constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
            uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
            uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            k0 += W0; k1 += W1;
        }
        return make_uint4(c0, c1, c2, c3);
    }

    // a <- b, b <- the block after it
    __device__ __forceinline__ void advance() {
        blk_lo += 1u;
        if (blk_lo == 0u) blk_hi += 1u;         // carry into the subsequence words cannot happen for < 2^66 draws
        uint32_t nlo = blk_lo + 1u, nhi = blk_hi + (nlo == 0u ? 1u : 0u);
        a = b;
        b = block(nlo, nhi, sub_lo, sub_hi, key_lo, key_hi);
    }

    // seed / subsequence / element offset (offset already includes any per-event skipahead)
    __device__ __forceinline__ void init(uint64_t seed, uint64_t subsequence, uint64_t element_offset) {
        key_lo = (uint32_t)seed; key_hi = (uint32_t)(seed >> 32);
        sub_lo = (uint32_t)subsequence; sub_hi = (uint32_t)(subsequence >> 32);
        uint64_t blk = element_offset >> 2;
        blk_lo = (uint32_t)blk; blk_hi = (uint32_t)(blk >> 32);
        pos = (uint32_t)(element_offset & 3u);
        a = block(blk_lo, blk_hi, sub_lo, sub_hi, key_lo, key_hi);
        uint32_t nlo = blk_lo + 1u, nhi = blk_hi + (nlo == 0u ? 1u : 0u);
        b = block(nlo, nhi, sub_lo, sub_hi, key_lo, key_hi);
    }

    // Call where the warp is converged (top of a bounce, before the surface/boundary draws): afterwards
    // at least 5 draws are cached, so the 4 + 2 draws of a typical bounce never refill inside divergent
    // code.  Has no effect on the sequence of numbers handed out.
    __device__ __forceinline__ void align() {
        if (pos >= 4u) { advance(); pos -= 4u; }
    }

    __device__ __forceinline__ uint32_t next_u32() {
        if (pos == 8u) { advance(); pos = 4u; }
        uint32_t p = pos;
        uint32_t r = p < 4u ? (p == 0u ? a.x : p == 1u ? a.y : p == 2u ? a.z : a.w)
                            : (p == 4u ? b.x : p == 5u ? b.y : p == 6u ? b.z : b.w);
        pos = p + 1u;
        return r;
    }

    // uniforms handed out so far, counted from element offset `base` (lets a stream be parked as 4 bytes
    // and re-created with init(seed, subsequence, base + consumed))
    __device__ __forceinline__ uint32_t consumed(uint64_t base) const {
        uint64_t blk = ((uint64_t)blk_hi << 32) | blk_lo;
        return (uint32_t)(blk * 4ull + pos - base);
    }

    // curand_uniform : (0,1]
    __device__ __forceinline__ float uniform() {
        return next_u32() * 2.3283064365386963e-10f + (2.3283064365386963e-10f / 2.0f);
    }
};

}  // namespace phox
