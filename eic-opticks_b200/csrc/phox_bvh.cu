// phox_bvh.cu : GPU construction of one binary LBVH (see phox_bvh.cuh for the role it plays).
//
// Pipeline, all on the device:
//   1. k_bounds     : block-reduce the union of the item boxes
//   2. k_morton     : 30-bit Morton code of each box centre, key = code<<32 | item (unique keys)
//   3. k_bitonic*   : sort the 64-bit keys (padded to a power of two)
//   4. k_hierarchy  : Karras 2012 "Maximizing parallelism in the construction of BVHs": each
//                     internal node finds its key range and split from common-prefix lengths
//   5. k_refit      : leaves walk up; the second arrival at a node merges the child boxes
//   6. k_emit       : write the 64 B two-child-box nodes
// Geometry is small (hundreds of prims, ~1e4 instances) so build time is irrelevant next to the
// per-event work; what matters is that nothing geometry-sized is built on the host.
#include "phox_bvh.cuh"
#include <math_constants.h>

namespace phox {

namespace {

__device__ __forceinline__ unsigned expand10(unsigned v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void k_bounds(const float* __restrict__ boxes, int n, float* __restrict__ out6) {
    __shared__ float s[6][256];
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(lo[k], boxes[6 * i + k]);
            hi[k] = fmaxf(hi[k], boxes[6 * i + 3 + k]);
        }
    }
    for (int k = 0; k < 3; k++) { s[k][threadIdx.x] = lo[k]; s[3 + k][threadIdx.x] = hi[k]; }
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) {
            for (int k = 0; k < 3; k++) {
                s[k][threadIdx.x] = fminf(s[k][threadIdx.x], s[k][threadIdx.x + w]);
                s[3 + k][threadIdx.x] = fmaxf(s[3 + k][threadIdx.x], s[3 + k][threadIdx.x + w]);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < 6) out6[threadIdx.x] = s[threadIdx.x][0];
}

__global__ void k_morton(const float* __restrict__ boxes, int n, int npad, const float* __restrict__ bounds6,
                         unsigned long long* __restrict__ keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    if (i >= n) { keys[i] = ~0ull; return; }
    unsigned q[3];
    for (int k = 0; k < 3; k++) {
        float lo = bounds6[k], hi = bounds6[3 + k];
        float c = 0.5f * (boxes[6 * i + k] + boxes[6 * i + 3 + k]);
        float ext = hi - lo;
        float u = ext > 0.f ? (c - lo) / ext : 0.5f;
        u = fminf(fmaxf(u * 1024.f, 0.f), 1023.f);
        q[k] = (unsigned)u;
    }
    unsigned code = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    keys[i] = ((unsigned long long)code << 32) | (unsigned)i;
}

__global__ void k_bitonic_step(unsigned long long* __restrict__ keys, int npad, int j, int k) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    int ixj = i ^ j;
    if (ixj > i) {
        unsigned long long a = keys[i], b = keys[ixj];
        bool up = (i & k) == 0;
        if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
    }
}

__device__ __forceinline__ int prefix_len(const unsigned long long* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll(keys[i] ^ keys[j]);       // keys are unique, so never 64
}

// internal node i in [0, n-1) ; children encoded: >=0 internal, <0 ~leaf_position
__global__ void k_hierarchy(const unsigned long long* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right,
                            int* __restrict__ parent_internal, int* __restrict__ parent_leaf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (prefix_len(keys, n, i, i + 1) - prefix_len(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = prefix_len(keys, n, i, i - d);
    int lmax = 2;
    while (prefix_len(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2) {
        if (prefix_len(keys, n, i, i + (l + t) * d) > dmin) l += t;
    }
    int j = i + l * d;
    int dnode = prefix_len(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (prefix_len(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int lc = (lo == gamma) ? ~gamma : gamma;
    int rc = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    left[i] = lc; right[i] = rc;
    if (lc >= 0) parent_internal[lc] = i; else parent_leaf[~lc] = i;
    if (rc >= 0) parent_internal[rc] = i; else parent_leaf[~rc] = i;
    if (i == 0) parent_internal[0] = -1;
}

__global__ void k_refit(const float* __restrict__ boxes, const unsigned long long* __restrict__ keys, int n,
                        const int* __restrict__ left, const int* __restrict__ right,
                        const int* __restrict__ parent_internal, const int* __restrict__ parent_leaf,
                        float* __restrict__ node_box, int* __restrict__ visit) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = parent_leaf[i];
    while (node >= 0) {
        if (atomicAdd(&visit[node], 1) == 0) return;      // first child to arrive stops; the second carries on
        __threadfence();
        float bx[6];
        for (int side = 0; side < 2; side++) {
            int c = side == 0 ? left[node] : right[node];
            const float* src = c >= 0 ? node_box + 6 * c : boxes + 6 * (int)(keys[~c] & 0xffffffffull);
            for (int k = 0; k < 3; k++) {
                float lo = __ldcg(src + k), hi = __ldcg(src + 3 + k);
                bx[k] = side == 0 ? lo : fminf(bx[k], lo);
                bx[3 + k] = side == 0 ? hi : fmaxf(bx[3 + k], hi);
            }
        }
        for (int k = 0; k < 6; k++) node_box[6 * node + k] = bx[k];
        __threadfence();
        node = parent_internal[node];
    }
}

__global__ void k_emit(const float* __restrict__ boxes, const unsigned long long* __restrict__ keys, int n, int base_item,
                       const int* __restrict__ left, const int* __restrict__ right, const float* __restrict__ node_box,
                       BvhNode* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    float bx[2][6];
    int ch[2];
    for (int side = 0; side < 2; side++) {
        int c = side == 0 ? left[i] : right[i];
        const float* src;
        if (c >= 0) { src = node_box + 6 * c; ch[side] = c; }
        else {
            int item = (int)(keys[~c] & 0xffffffffull);
            src = boxes + 6 * item;
            ch[side] = ~(base_item + item);
        }
        for (int k = 0; k < 6; k++) bx[side][k] = src[k];
    }
    BvhNode nd;
    nd.a = make_float4(bx[0][0], bx[0][1], bx[0][2], bx[0][3]);
    nd.b = make_float4(bx[0][4], bx[0][5], bx[1][0], bx[1][1]);
    nd.c = make_float4(bx[1][2], bx[1][3], bx[1][4], bx[1][5]);
    nd.d = make_int4(ch[0], ch[1], 0, 0);
    out[i] = nd;
}

__global__ void k_single(const float* __restrict__ boxes, int base_item, BvhNode* __restrict__ out) {
    BvhNode nd;
    nd.a = make_float4(boxes[0], boxes[1], boxes[2], boxes[3]);
    nd.b = make_float4(boxes[4], boxes[5], 0.f, 0.f);
    nd.c = make_float4(0.f, 0.f, 0.f, 0.f);
    nd.d = make_int4(~base_item, kBvhNoChild, 0, 0);            // one item: the second child does not exist
    out[0] = nd;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

void bvh_scratch_free(BvhScratch& s) {
    if (s.buf) cudaFree(s.buf);
    s.buf = nullptr; s.bytes = 0;
}

cudaError_t bvh_build(const float* d_boxes, int n, int base_item, BvhNode* d_out, BvhScratch& scratch, cudaStream_t stream, int* kernel_count) {
    if (n <= 0) return cudaSuccess;
    if (n == 1) {
        k_single<<<1, 1, 0, stream>>>(d_boxes, base_item, d_out);
        if (kernel_count) *kernel_count += 1;
        return cudaGetLastError();
    }
    int npad = 1;
    while (npad < n) npad <<= 1;

    size_t off = 0;
    size_t o_keys = off;    off = align_up(off + sizeof(unsigned long long) * npad, 256);
    size_t o_bounds = off;  off = align_up(off + sizeof(float) * 8, 256);
    size_t o_left = off;    off = align_up(off + sizeof(int) * n, 256);
    size_t o_right = off;   off = align_up(off + sizeof(int) * n, 256);
    size_t o_pint = off;    off = align_up(off + sizeof(int) * n, 256);
    size_t o_pleaf = off;   off = align_up(off + sizeof(int) * n, 256);
    size_t o_visit = off;   off = align_up(off + sizeof(int) * n, 256);
    size_t o_nbox = off;    off = align_up(off + sizeof(float) * 6 * n, 256);
    if (scratch.bytes < off) {
        bvh_scratch_free(scratch);
        cudaError_t e = cudaMalloc(&scratch.buf, off);
        if (e != cudaSuccess) return e;
        scratch.bytes = off;
    }
    char* base = (char*)scratch.buf;
    auto* keys = (unsigned long long*)(base + o_keys);
    auto* bounds = (float*)(base + o_bounds);
    int* left = (int*)(base + o_left);
    int* right = (int*)(base + o_right);
    int* pint = (int*)(base + o_pint);
    int* pleaf = (int*)(base + o_pleaf);
    int* visit = (int*)(base + o_visit);
    float* nbox = (float*)(base + o_nbox);

    const int T = 256;
    int kc = 0;
    k_bounds<<<1, 256, 0, stream>>>(d_boxes, n, bounds); kc++;
    k_morton<<<(npad + T - 1) / T, T, 0, stream>>>(d_boxes, n, npad, bounds, keys); kc++;
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) { k_bitonic_step<<<(npad + T - 1) / T, T, 0, stream>>>(keys, npad, j, k); kc++; }
    cudaMemsetAsync(visit, 0, sizeof(int) * n, stream);
    k_hierarchy<<<(n + T - 1) / T, T, 0, stream>>>(keys, n, left, right, pint, pleaf); kc++;
    k_refit<<<(n + T - 1) / T, T, 0, stream>>>(d_boxes, keys, n, left, right, pint, pleaf, nbox, visit); kc++;
    k_emit<<<(n + T - 1) / T, T, 0, stream>>>(d_boxes, keys, n, base_item, left, right, nbox, d_out); kc++;
    if (kernel_count) *kernel_count += kc;
    return cudaGetLastError();
}

}  // namespace phox
