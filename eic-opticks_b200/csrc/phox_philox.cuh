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
// Instead of the 64-byte curandStatePhilox4_32_10 the stream caches ONE block (4 outputs), lazily: a block is only
// worked out when a draw that is actually USED falls into it.  The as-built reference (DEBUG_TAG) burns 3 - 7 of the
// 6 - 10 draws of a bounce; skip() steps over those without generating anything, and draw2_ahead() - the two draws
// every bounce starts with - refills at points where the whole warp is converged, so that a typical bounce costs two
// warp-level block computations (it was 3.3 with the former eager two-block cache, profiles/r2_summary.md).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace phox {

struct Philox {
    uint4    a;          // outputs of block `blk + cblk`
    uint32_t blk_lo, blk_hi;     // block of the element offset the stream was created with
    uint32_t sub_lo, sub_hi;
    uint32_t key_lo, key_hi;
    uint32_t rel;        // next draw, counted from the first output of block `blk`
    uint32_t cblk;       // which block `a` holds, relative to blk (rel >> 2 of its draws); kNone = nothing cached yet
    static constexpr uint32_t kNone = 0xffffffffu;

    // Out of line on purpose: uniform() is expanded at ~30 draw sites and each would otherwise carry its own
    // copy of the 10 rounds; the simulate kernel is instruction-fetch bound (profiles/), so code size matters.
    static __device__ __forceinline__ uint4 block_inline(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
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
    static __device__ __noinline__ uint4 block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
        return block_inline(c0, c1, c2, c3, k0, k1);
    }

    // a <- the block the next draw falls into
    __device__ __forceinline__ void fill() {
        const uint32_t b = rel >> 2;
        const uint32_t lo = blk_lo + b;
        const uint32_t hi = blk_hi + (lo < blk_lo ? 1u : 0u);      // carry into the subsequence words cannot happen for < 2^66 draws
        a = block(lo, hi, sub_lo, sub_hi, key_lo, key_hi);
        cblk = b;
    }

    // The same with the ten rounds compiled in place.  A call is a scoreboard boundary: the caller waits for every load and
    // texture fetch in flight before it branches, and the first refill of a bounce sits right behind the material fetches
    // (profiles/r2_summary.md: 8.9 % of the physics kernel's stall samples were on that one CALL).  In place, the rounds
    // run while the fetches travel.
    __device__ __forceinline__ void fill_inline() {
        const uint32_t b = rel >> 2;
        const uint32_t lo = blk_lo + b;
        const uint32_t hi = blk_hi + (lo < blk_lo ? 1u : 0u);
        a = block_inline(lo, hi, sub_lo, sub_hi, key_lo, key_hi);
        cblk = b;
    }

    // seed / subsequence / element offset (offset already includes any per-event skipahead)
    __device__ __forceinline__ void init(uint64_t seed, uint64_t subsequence, uint64_t element_offset) {
        key_lo = (uint32_t)seed; key_hi = (uint32_t)(seed >> 32);
        sub_lo = (uint32_t)subsequence; sub_hi = (uint32_t)(subsequence >> 32);
        uint64_t blk = element_offset >> 2;
        blk_lo = (uint32_t)blk; blk_hi = (uint32_t)(blk >> 32);
        rel = (uint32_t)(element_offset & 3u);
        cblk = kNone;
        a = make_uint4(0u, 0u, 0u, 0u);
    }

    // Call where the warp is converged: the block of the next draw is worked out here if it is not cached yet.
    // Has no effect on the sequence of numbers handed out.
    __device__ __forceinline__ void align() {
        if ((rel >> 2) != cblk) fill();
    }

    // draws whose value nobody reads (the burns of the DEBUG_TAG consumption pattern): nothing is generated
    __device__ __forceinline__ void skip(uint32_t n) { rel += n; }

    __device__ __forceinline__ uint32_t pick(uint32_t r) const {
        const uint32_t p = r & 3u;
        return p == 0u ? a.x : p == 1u ? a.y : p == 2u ? a.z : a.w;
    }

    __device__ __forceinline__ uint32_t next_u32() {
        if ((rel >> 2) != cblk) fill();
        const uint32_t r = pick(rel);
        rel += 1u;
        return r;
    }

    // uniforms handed out (or skipped) so far, counted from element offset `base` (lets a stream be parked as 4 bytes
    // and re-created with init(seed, subsequence, base + consumed))
    __device__ __forceinline__ uint32_t consumed(uint64_t base) const {
        uint64_t blk = ((uint64_t)blk_hi << 32) | blk_lo;
        return (uint32_t)(blk * 4ull + rel - base);
    }

    static __device__ __forceinline__ float to_uniform(uint32_t x) {       // curand_uniform : (0,1]
        return x * 2.3283064365386963e-10f + (2.3283064365386963e-10f / 2.0f);
    }
    __device__ __forceinline__ float uniform() { return to_uniform(next_u32()); }

    // u0 = uniform(); u1 = uniform(); and the block of the draw AFTER them is cached on return.  Both refills sit at
    // warp-converged points (every lane of a bounce comes through here), the second one only for the lanes whose next
    // two draws leave the block of the first.
    __device__ __forceinline__ void draw2_ahead(float& u0, float& u1) {
        if ((rel >> 2) != cblk) fill_inline();
        const uint32_t x0 = pick(rel);
        const uint32_t r1 = rel + 1u;
        const bool same = (r1 >> 2) == cblk;
        uint32_t x1 = pick(r1);                  // valid when `same`
        rel += 2u;
        if ((rel >> 2) != cblk) fill();          // r1 & 3 == 0 (then it holds draw r1 in .x) or rel & 3 == 0
        if (!same) x1 = a.x;
        u0 = to_uniform(x0); u1 = to_uniform(x1);
    }
};

}  // namespace phox
