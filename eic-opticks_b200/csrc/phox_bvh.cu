// phox_bvh.cu : GPU construction of one binary BVH (see phox_bvh.cuh for the role it plays).
//
// Pipeline, all on the device:
//   1. k_bounds     : block-reduce the union of the item boxes
//   2. k_morton     : 30-bit Morton code of each box centre, key = code<<32 | item (unique keys)
//   3. k_bitonic*   : sort the 64-bit keys (padded to a power of two)
//   4. k_ploc       : PLOC agglomerative clustering over the Morton order (surface-area driven), which
//                     writes the 64 B two-child-box nodes directly; the last merge is node 0, the root
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

__device__ __forceinline__ float merged_area(const float* a, const float* b) {
    float dx = fmaxf(a[3], b[3]) - fminf(a[0], b[0]);
    float dy = fmaxf(a[4], b[4]) - fminf(a[1], b[1]);
    float dz = fmaxf(a[5], b[5]) - fminf(a[2], b[2]);
    return dx * dy + dy * dz + dz * dx;
}

// PLOC - parallel locally-ordered clustering (Meister & Bittner 2018): bottom-up agglomeration of the
// Morton-ordered clusters.  Every round each cluster looks `radius` neighbours to either side for the
// partner giving the smallest merged surface area; mutual choices merge into a new node.  Large boxes
// (a world volume, slabs spanning the whole detector) only merge late, i.e. end up near the root,
// which is what keeps rays that start deep inside small volumes from visiting them at every level.
// One block builds one tree: geometry is small (<= ~1e4 items) and built once.
constexpr int kPlocRadius = 16;
__global__ void __launch_bounds__(1024) k_ploc(const float* __restrict__ boxes, const unsigned long long* __restrict__ keys, int n, int base_item,
                                               float* __restrict__ cbox0, float* __restrict__ cbox1, int* __restrict__ cid0, int* __restrict__ cid1,
                                               int* __restrict__ nn, int* __restrict__ keep, BvhNode* __restrict__ out) {
    __shared__ int s_scan[1024];
    __shared__ int s_count, s_next_node;
    const int T = blockDim.x, tid = threadIdx.x;
    float* cb = cbox0; float* cb2 = cbox1; int* id = cid0; int* id2 = cid1;
    for (int i = tid; i < n; i += T) {
        int item = (int)(keys[i] & 0xffffffffull);
        for (int k = 0; k < 6; k++) cb[6 * i + k] = boxes[6 * item + k];
        id[i] = ~(base_item + item);
    }
    if (tid == 0) { s_count = n; s_next_node = n - 2; }          // nodes are handed out downwards: the last merge is node 0 = root
    __syncthreads();
    int count = n;
    while (count > 1) {
        for (int i = tid; i < count; i += T) {                   // nearest neighbour in the window
            float best = CUDART_INF_F; int bj = -1;
            int j0 = max(0, i - kPlocRadius), j1 = min(count - 1, i + kPlocRadius);
            for (int j = j0; j <= j1; j++) {
                if (j == i) continue;
                float a = merged_area(cb + 6 * i, cb + 6 * j);
                if (a < best) { best = a; bj = j; }
            }
            // boxes are validated as finite before the build; should an area still come out inf / NaN, fall back to the
            // neighbour in Morton order so that every round merges at least one pair and the loop terminates
            if (bj < 0) bj = i + 1 < count ? i + 1 : i - 1;
            nn[i] = bj;
        }
        __syncthreads();
        for (int i = tid; i < count; i += T) {                   // mutual pairs merge (the lower index owns the new node)
            int j = nn[i];
            bool mutual = nn[j] == i;
            if (mutual && i < j) {
                int node = atomicSub(&s_next_node, 1);
                BvhNode nd;
                const float* A = cb + 6 * i; const float* B = cb + 6 * j;
                nd.a = make_float4(A[0], A[1], A[2], A[3]);
                nd.b = make_float4(A[4], A[5], B[0], B[1]);
                nd.c = make_float4(B[2], B[3], B[4], B[5]);
                nd.d = make_int4(id[i], id[j], 0, 0);
                out[node] = nd;
                for (int k = 0; k < 3; k++) { cb2[6 * i + k] = fminf(A[k], B[k]); cb2[6 * i + 3 + k] = fmaxf(A[3 + k], B[3 + k]); }
                id2[i] = node;
                keep[i] = 1;
            } else if (mutual) {
                keep[i] = 0;
            } else {
                for (int k = 0; k < 6; k++) cb2[6 * i + k] = cb[6 * i + k];
                id2[i] = id[i];
                keep[i] = 1;
            }
        }
        __syncthreads();
        // order-preserving compaction of the kept clusters
        int chunk = (count + T - 1) / T;
        int lo = min(count, tid * chunk), hi = min(count, lo + chunk);
        int local = 0;
        for (int i = lo; i < hi; i++) local += keep[i];
        s_scan[tid] = local;
        __syncthreads();
        for (int off = 1; off < T; off <<= 1) {
            int v = tid >= off ? s_scan[tid - off] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        int pos = s_scan[tid] - local;
        for (int i = lo; i < hi; i++) {
            if (keep[i]) {
                for (int k = 0; k < 6; k++) cb[6 * pos + k] = cb2[6 * i + k];
                id[pos] = id2[i];
                pos++;
            }
        }
        if (tid == T - 1) s_count = s_scan[T - 1];
        __syncthreads();
        count = s_count;
        __syncthreads();
    }
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
    size_t o_cb0 = off;     off = align_up(off + sizeof(float) * 6 * n, 256);
    size_t o_cb1 = off;     off = align_up(off + sizeof(float) * 6 * n, 256);
    size_t o_id0 = off;     off = align_up(off + sizeof(int) * n, 256);
    size_t o_id1 = off;     off = align_up(off + sizeof(int) * n, 256);
    size_t o_nn = off;      off = align_up(off + sizeof(int) * n, 256);
    size_t o_keep = off;    off = align_up(off + sizeof(int) * n, 256);
    if (scratch.bytes < off) {
        bvh_scratch_free(scratch);
        cudaError_t e = cudaMalloc(&scratch.buf, off);
        if (e != cudaSuccess) return e;
        scratch.bytes = off;
    }
    char* base = (char*)scratch.buf;
    auto* keys = (unsigned long long*)(base + o_keys);
    auto* bounds = (float*)(base + o_bounds);

    const int T = 256;
    int kc = 0;
    k_bounds<<<1, 256, 0, stream>>>(d_boxes, n, bounds); kc++;
    k_morton<<<(npad + T - 1) / T, T, 0, stream>>>(d_boxes, n, npad, bounds, keys); kc++;
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) { k_bitonic_step<<<(npad + T - 1) / T, T, 0, stream>>>(keys, npad, j, k); kc++; }
    k_ploc<<<1, 1024, 0, stream>>>(d_boxes, keys, n, base_item, (float*)(base + o_cb0), (float*)(base + o_cb1), (int*)(base + o_id0),
                                   (int*)(base + o_id1), (int*)(base + o_nn), (int*)(base + o_keep), d_out); kc++;
    if (kernel_count) *kernel_count += kc;
    return cudaGetLastError();
}

}  // namespace phox
