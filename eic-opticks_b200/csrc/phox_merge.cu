// phox_merge.cu : hit merging by (sensor identity, time bucket).
//
// Replaces SPM::merge_partial_select (sysrap/SPM.cu:153-290: thrust count_if/copy_if, transform to keys,
// sort_by_key, reduce_by_key) with the sphoton functors of sysrap/sphoton.h:277-304:
//   select   (flagmask & mask) != 0                              any bit, unlike the all-bits hit selection
//   key      (u64(identity & 0xffffff) << 48) | u32(time / tw)
//   reduce   the earlier photon survives (ties: the one first in key order, i.e. lower photon index),
//            flagmask = a | b, hitcount = a + b
// Output order = ascending key, like reduce_by_key.  tw == 0 returns the selection unmerged, in input order.
//
// Own kernels: a stable LSD radix sort (8-bit digits; one warp owns one tile and ranks it row by row with
// __match_any_sync, so equal keys keep their input order), head flags + tile scan, one thread per group folds it.
#include <cuda_runtime.h>
#include <stdint.h>

#include "phox_merge.cuh"

namespace phox {

namespace {

constexpr int kTileRows = 64;                    // a warp's tile = 32 x 64 = 2048 keys
constexpr int kTile = 32 * kTileRows;
constexpr int kSortWarps = 4;                    // warps (tiles) per block
constexpr unsigned long long kNoKey = ~0ull;

__global__ void k_merge_keys(const Photon* __restrict__ in, unsigned n, unsigned mask, float tw, unsigned long long* __restrict__ key,
                             unsigned* __restrict__ idx) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Photon& p = in[i];
    bool sel = mask == 0u || (p.flagmask & mask) != 0u;
    unsigned id = p.identity & 0x00ffffffu;
    unsigned bucket = static_cast<unsigned>(p.time / tw);
    key[i] = sel ? (((unsigned long long)id << 48) | (unsigned long long)bucket) : kNoKey;
    idx[i] = i;
}

// sphotonlite::key_functor (sysrap/sphotonlite.h): identity = low 16 bits
__global__ void k_merge_keys_lite(const PhotonLite* __restrict__ in, unsigned n, unsigned mask, float tw, unsigned long long* __restrict__ key,
                                  unsigned* __restrict__ idx) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PhotonLite p = in[i];
    bool sel = mask == 0u || (p.flagmask & mask) != 0u;
    unsigned id = p.hitcount_identity & 0xffffu;
    unsigned bucket = static_cast<unsigned>(p.time / tw);
    key[i] = sel ? (((unsigned long long)id << 48) | (unsigned long long)bucket) : kNoKey;
    idx[i] = i;
}

// digit histogram of every warp tile: hist[digit * ntile + tile]
__global__ void __launch_bounds__(32 * kSortWarps) k_radix_hist(const unsigned long long* __restrict__ key, unsigned n, int shift, unsigned ntile,
                                                                unsigned* __restrict__ hist) {
    __shared__ unsigned cnt[kSortWarps][256];
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    unsigned tile = blockIdx.x * kSortWarps + w;
    for (int k = lane; k < 256; k += 32) cnt[w][k] = 0u;
    __syncwarp();
    if (tile < ntile) {
        unsigned base = tile * kTile;
        for (int r = 0; r < kTileRows; r++) {
            unsigned i = base + r * 32 + lane;
            if (i < n) atomicAdd(&cnt[w][(unsigned)(key[i] >> shift) & 0xffu], 1u);
        }
        __syncwarp();
        for (int k = lane; k < 256; k += 32) hist[(size_t)k * ntile + tile] = cnt[w][k];
    }
}

// exclusive scan of a u32 array in place (single block, any length)
__global__ void k_scan_u32(unsigned* __restrict__ a, unsigned n) {
    __shared__ unsigned s[1024];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (unsigned base = 0; base < n; base += 1024) {
        unsigned i = base + threadIdx.x;
        unsigned v = i < n ? a[i] : 0u;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            unsigned t = threadIdx.x >= (unsigned)off ? s[threadIdx.x - off] : 0u;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n) a[i] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
}

// stable scatter: rank of a key = scanned start of (digit, tile) + keys of that digit earlier in the tile
__global__ void __launch_bounds__(32 * kSortWarps) k_radix_scatter(const unsigned long long* __restrict__ key_in, const unsigned* __restrict__ idx_in,
                                                                   unsigned n, int shift, unsigned ntile, const unsigned* __restrict__ hist,
                                                                   unsigned long long* __restrict__ key_out, unsigned* __restrict__ idx_out) {
    __shared__ unsigned pos[kSortWarps][256];
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    unsigned tile = blockIdx.x * kSortWarps + w;
    if (tile >= ntile) return;
    for (int k = lane; k < 256; k += 32) pos[w][k] = hist[(size_t)k * ntile + tile];
    __syncwarp();
    unsigned base = tile * kTile;
    for (int r = 0; r < kTileRows; r++) {
        unsigned i = base + r * 32 + lane;
        bool live = i < n;
        unsigned long long k = live ? key_in[i] : 0ull;
        unsigned d = live ? ((unsigned)(k >> shift) & 0xffu) : 0x100u + lane;      // dead lanes match nobody
        unsigned active = __ballot_sync(0xffffffffu, live);
        unsigned peers = __match_any_sync(0xffffffffu, d) & active;
        unsigned dst = 0;
        if (live) dst = pos[w][d] + __popc(peers & lt);
        __syncwarp();
        if (live && (peers & lt) == 0u) pos[w][d] += __popc(peers);               // the first lane of each digit group advances it
        __syncwarp();
        if (live) { key_out[dst] = k; idx_out[dst] = idx_in[i]; }
    }
}

constexpr int kHeadTile = 256;
// group heads per tile of the sorted keys
__global__ void __launch_bounds__(kHeadTile) k_merge_heads(const unsigned long long* __restrict__ key, unsigned n, unsigned* __restrict__ tile_heads) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    bool head = i < n && key[i] != kNoKey && (i == 0 || key[i] != key[i - 1]);
    int c = __syncthreads_count(head);
    if (threadIdx.x == 0) tile_heads[blockIdx.x] = (unsigned)c;
}

// one thread per group: fold it left to right with sphoton::reduce_op
__global__ void __launch_bounds__(kHeadTile) k_merge_reduce(const Photon* __restrict__ in, const unsigned long long* __restrict__ key,
                                                            const unsigned* __restrict__ idx, unsigned n, const unsigned* __restrict__ tile_off,
                                                            Photon* __restrict__ out) {
    __shared__ unsigned wsum[kHeadTile / 32];
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    bool head = i < n && key[i] != kNoKey && (i == 0 || key[i] != key[i - 1]);
    unsigned b = __ballot_sync(0xffffffffu, head);
    if (lane == 0) wsum[w] = __popc(b);
    __syncthreads();
    unsigned before = 0;
    for (unsigned k = 0; k < w; k++) before += wsum[k];
    if (!head) return;
    unsigned o = tile_off[blockIdx.x] + before + __popc(b & ((1u << lane) - 1u));
    unsigned long long k0 = key[i];
    Photon r = in[idx[i]];
    unsigned hc = r.hitcount_iindex >> 16;
    for (unsigned j = i + 1; j < n && key[j] == k0; j++) {
        const Photon& q = in[idx[j]];
        unsigned fm = r.flagmask | q.flagmask;
        hc += q.hitcount_iindex >> 16;
        bool r_first = fminf(r.time, q.time) == r.time;
        if (!r_first) r = q;
        r.flagmask = fm;
    }
    r.hitcount_iindex = (r.hitcount_iindex & 0x0000ffffu) | ((hc & 0xffffu) << 16);      // sphoton::set_hitcount
    out[o] = r;
}

// sphotonlite::reduce_op: r = a; time = min; flagmask |= ; hitcount summed, identity of a (and a's local position)
__global__ void __launch_bounds__(kHeadTile) k_merge_reduce_lite(const PhotonLite* __restrict__ in, const unsigned long long* __restrict__ key,
                                                                 const unsigned* __restrict__ idx, unsigned n, const unsigned* __restrict__ tile_off,
                                                                 PhotonLite* __restrict__ out) {
    __shared__ unsigned wsum[kHeadTile / 32];
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    bool head = i < n && key[i] != kNoKey && (i == 0 || key[i] != key[i - 1]);
    unsigned b = __ballot_sync(0xffffffffu, head);
    if (lane == 0) wsum[w] = __popc(b);
    __syncthreads();
    unsigned before = 0;
    for (unsigned k = 0; k < w; k++) before += wsum[k];
    if (!head) return;
    unsigned o = tile_off[blockIdx.x] + before + __popc(b & ((1u << lane) - 1u));
    unsigned long long k0 = key[i];
    PhotonLite r = in[idx[i]];
    unsigned hc = r.hitcount_identity >> 16;
    for (unsigned j = i + 1; j < n && key[j] == k0; j++) {
        PhotonLite q = in[idx[j]];
        r.time = fminf(r.time, q.time);
        r.flagmask |= q.flagmask;
        hc += q.hitcount_identity >> 16;
    }
    r.hitcount_identity = ((hc & 0xffffu) << 16) | (r.hitcount_identity & 0xffffu);
    out[o] = r;
}

__global__ void __launch_bounds__(kHeadTile) k_select_count_lite(const PhotonLite* __restrict__ in, unsigned n, unsigned mask, unsigned* __restrict__ tile_cnt) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    bool sel = i < n && (mask == 0u || (in[i].flagmask & mask) != 0u);
    int c = __syncthreads_count(sel);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = (unsigned)c;
}
__global__ void __launch_bounds__(kHeadTile) k_select_copy_lite(const PhotonLite* __restrict__ in, unsigned n, unsigned mask,
                                                                const unsigned* __restrict__ tile_off, PhotonLite* __restrict__ out) {
    __shared__ unsigned wsum[kHeadTile / 32];
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    bool sel = i < n && (mask == 0u || (in[i].flagmask & mask) != 0u);
    unsigned b = __ballot_sync(0xffffffffu, sel);
    if (lane == 0) wsum[w] = __popc(b);
    __syncthreads();
    unsigned before = 0;
    for (unsigned k = 0; k < w; k++) before += wsum[k];
    if (sel) out[tile_off[blockIdx.x] + before + __popc(b & ((1u << lane) - 1u))] = in[i];
}

// tw == 0: the selection, in input order (tile counts + offsets + ordered copy)
__global__ void __launch_bounds__(kHeadTile) k_select_count(const Photon* __restrict__ in, unsigned n, unsigned mask, unsigned* __restrict__ tile_cnt) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    bool sel = i < n && (mask == 0u || (in[i].flagmask & mask) != 0u);
    int c = __syncthreads_count(sel);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = (unsigned)c;
}
__global__ void __launch_bounds__(kHeadTile) k_select_copy(const Photon* __restrict__ in, unsigned n, unsigned mask, const unsigned* __restrict__ tile_off,
                                                           Photon* __restrict__ out) {
    __shared__ unsigned wsum[kHeadTile / 32];
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    bool sel = i < n && (mask == 0u || (in[i].flagmask & mask) != 0u);
    unsigned b = __ballot_sync(0xffffffffu, sel);
    if (lane == 0) wsum[w] = __popc(b);
    __syncthreads();
    unsigned before = 0;
    for (unsigned k = 0; k < w; k++) before += wsum[k];
    if (sel) out[tile_off[blockIdx.x] + before + __popc(b & ((1u << lane) - 1u))] = in[i];
}

__global__ void k_total(const unsigned* __restrict__ off, const unsigned* __restrict__ cnt_last, unsigned ntile, unsigned* __restrict__ total) {
    total[0] = off[ntile - 1] + cnt_last[0];
}

}  // namespace

static cudaError_t merge_any(bool lite, const void* d_in_, int64_t n64, unsigned select_mask, float tw, void* d_out_, int64_t* n_out, MergeScratch& sc,
                             cudaStream_t stream, int* kernel_count) {
    const Photon* d_in = (const Photon*)d_in_; Photon* d_out = (Photon*)d_out_;
    const PhotonLite* l_in = (const PhotonLite*)d_in_; PhotonLite* l_out = (PhotonLite*)d_out_;
    *n_out = 0;
    if (n64 <= 0) return cudaSuccess;
    if (n64 > 0x7fffffffll) return cudaErrorInvalidValue;
    const unsigned n = (unsigned)n64;
    const unsigned ntile = (n + kTile - 1) / kTile;
    const unsigned nhead = (n + kHeadTile - 1) / kHeadTile;
    // scratch: keys x2, idx x2, hist (256 x ntile), tile counts, their scan, the last count, the total
    size_t need = (size_t)n * (8 + 8 + 4 + 4) + (size_t)256 * ntile * 4 + (size_t)nhead * 8 + 64;
    if (sc.bytes < need) {
        if (sc.buf) cudaFree(sc.buf);
        sc.buf = nullptr; sc.bytes = 0;
        cudaError_t e = cudaMalloc(&sc.buf, need);
        if (e != cudaSuccess) return e;
        sc.bytes = need;
    }
    char* p = (char*)sc.buf;
    unsigned long long* key0 = (unsigned long long*)p; p += (size_t)n * 8;
    unsigned long long* key1 = (unsigned long long*)p; p += (size_t)n * 8;
    unsigned* idx0 = (unsigned*)p; p += (size_t)n * 4;
    unsigned* idx1 = (unsigned*)p; p += (size_t)n * 4;
    unsigned* hist = (unsigned*)p; p += (size_t)256 * ntile * 4;
    unsigned* tcnt = (unsigned*)p; p += (size_t)nhead * 4;
    unsigned* toff = (unsigned*)p; p += (size_t)nhead * 4;
    unsigned* total = (unsigned*)p;
    int nk = 0;
    unsigned h_total = 0;
    cudaError_t e;

    if (tw == 0.f) {
        if (lite) k_select_count_lite<<<nhead, kHeadTile, 0, stream>>>(l_in, n, select_mask, tcnt);
        else k_select_count<<<nhead, kHeadTile, 0, stream>>>(d_in, n, select_mask, tcnt);
        cudaMemcpyAsync(toff, tcnt, (size_t)nhead * 4, cudaMemcpyDeviceToDevice, stream);
        k_scan_u32<<<1, 1024, 0, stream>>>(toff, nhead);
        k_total<<<1, 1, 0, stream>>>(toff, tcnt + nhead - 1, nhead, total);
        if (lite) k_select_copy_lite<<<nhead, kHeadTile, 0, stream>>>(l_in, n, select_mask, toff, l_out);
        else k_select_copy<<<nhead, kHeadTile, 0, stream>>>(d_in, n, select_mask, toff, d_out);
        nk += 4;
    } else {
        if (lite) k_merge_keys_lite<<<(n + 255) / 256, 256, 0, stream>>>(l_in, n, select_mask, tw, key0, idx0);
        else k_merge_keys<<<(n + 255) / 256, 256, 0, stream>>>(d_in, n, select_mask, tw, key0, idx0);
        nk += 1;
        const unsigned sblocks = (ntile + kSortWarps - 1) / kSortWarps;
        static const int shifts[6] = {0, 8, 16, 24, 48, 56};          // bits 32..47 of a key are always zero (or all ones for unselected entries)
        for (int s = 0; s < 6; s++) {
            k_radix_hist<<<sblocks, 32 * kSortWarps, 0, stream>>>(key0, n, shifts[s], ntile, hist);
            k_scan_u32<<<1, 1024, 0, stream>>>(hist, 256u * ntile);
            k_radix_scatter<<<sblocks, 32 * kSortWarps, 0, stream>>>(key0, idx0, n, shifts[s], ntile, hist, key1, idx1);
            unsigned long long* tk = key0; key0 = key1; key1 = tk;
            unsigned* ti = idx0; idx0 = idx1; idx1 = ti;
            nk += 3;
        }
        k_merge_heads<<<nhead, kHeadTile, 0, stream>>>(key0, n, tcnt);
        cudaMemcpyAsync(toff, tcnt, (size_t)nhead * 4, cudaMemcpyDeviceToDevice, stream);
        k_scan_u32<<<1, 1024, 0, stream>>>(toff, nhead);
        k_total<<<1, 1, 0, stream>>>(toff, tcnt + nhead - 1, nhead, total);
        if (lite) k_merge_reduce_lite<<<nhead, kHeadTile, 0, stream>>>(l_in, key0, idx0, n, toff, l_out);
        else k_merge_reduce<<<nhead, kHeadTile, 0, stream>>>(d_in, key0, idx0, n, toff, d_out);
        nk += 4;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(&h_total, total, 4, cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    *n_out = (int64_t)h_total;
    if (kernel_count) *kernel_count += nk;
    return cudaSuccess;
}

cudaError_t merge_photons(const Photon* d_in, int64_t n, unsigned select_mask, float tw, Photon* d_out, int64_t* n_out, MergeScratch& sc,
                          cudaStream_t stream, int* kernel_count) {
    return merge_any(false, d_in, n, select_mask, tw, d_out, n_out, sc, stream, kernel_count);
}

cudaError_t merge_photons_lite(const PhotonLite* d_in, int64_t n, unsigned select_mask, float tw, PhotonLite* d_out, int64_t* n_out, MergeScratch& sc,
                               cudaStream_t stream, int* kernel_count) {
    return merge_any(true, d_in, n, select_mask, tw, d_out, n_out, sc, stream, kernel_count);
}

void merge_scratch_free(MergeScratch& sc) {
    if (sc.buf) cudaFree(sc.buf);
    sc.buf = nullptr; sc.bytes = 0;
}

}  // namespace phox
