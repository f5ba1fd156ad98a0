// phox_bvh.cuh : node format and ray/box test of the two-level BVH that replaces the OptiX
// GAS/IAS/SBT machinery (CSGOptiX/SBT.cc:277-370, 427-557; GAS_Builder.cc; IAS_Builder.cc).
//
// Level 1 ("instance BVH") is built over the world-space boxes of the instances (inst qat4 x
// solid box); level 2 is one BVH per CSGSolid over the CSGPrim boxes (CSGPrim.h q2,q3: the same
// boxes the reference hands to optixAccelBuild as custom-primitive AABBs).  Both are binary BVHs
// built on the GPU (Morton codes -> sort -> PLOC surface-area clustering) and then laid out
// as 64-byte nodes that hold BOTH children's boxes, so one 64 B fetch decides two subtrees.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <utility>
#include <vector>

namespace phox {

struct alignas(16) BvhNode {
    // child 0 box: lo0.xyz, hi0.xyz ; child 1 box: lo1.xyz, hi1.xyz ; packed as 3 float4
    float4 a;   // lo0.x lo0.y lo0.z hi0.x
    float4 b;   // hi0.y hi0.z lo1.x lo1.y
    float4 c;   // lo1.z hi1.x hi1.y hi1.z
    int4   d;   // child0, child1 ( >= 0 : node index relative to the tree root ; < 0 : ~item ), 0, 0
};
static_assert(sizeof(BvhNode) == 64, "BvhNode is one 64 B line pair");
constexpr int kBvhNoChild = 0x7fffffff;     // child slot of a one-item tree that holds nothing

struct InstanceRec {                // one per instance, 128 B
    float4 inv[4];                  // inverse transform rows (world -> object), row-vector convention, 4th column cleared
    int    solid;                   // gas_idx
    int    identity;                // sensor_identifier + 1 (0 = not a sensor), low 16 bits go to prd
    int    is_identity;             // transform is exactly identity: skip the ray transform
    int    bvh_root;                // index of the solid's root BvhNode in the node pool
    int    prim_offset;             // first CSGPrim of the solid
    int    num_prim;
    int    pad0, pad1;
    float4 pad2[2];
};
static_assert(sizeof(InstanceRec) == 128, "InstanceRec is 128 B");

// Build a BVH over n boxes (6 floats each: lo.xyz hi.xyz, device memory) into out[0 .. max(n-1,1)).
// Item ids stored in leaves are base_item + i.  All work is enqueued on `stream`.
// Returns cudaSuccess or the first CUDA error.  scratch is grown as needed.
struct BvhScratch {
    void* buf = nullptr;
    size_t bytes = 0;
};
cudaError_t bvh_build(const float* d_boxes, int n, int base_item, BvhNode* d_out, BvhScratch& scratch, cudaStream_t stream, int* kernel_count);
void bvh_scratch_free(BvhScratch& scratch);

// Host: number of internal nodes on the longest root-to-leaf path of a tree laid out in nodes[0 .. nnode), or -1 when the
// array does not describe a tree (index out of range, more visits than nodes).  The traversal parks at most one entry per
// internal node of the current path, so `depth(instance tree) + 1 + depth(deepest solid tree)` bounds its stack.
inline int bvh_tree_depth(const BvhNode* nodes, int nnode) {
    if (!nodes || nnode <= 0) return -1;
    std::vector<std::pair<int, int>> todo;
    todo.emplace_back(0, 1);
    int deepest = 0;
    long long visited = 0;
    while (!todo.empty()) {
        const std::pair<int, int> e = todo.back();
        todo.pop_back();
        if (e.first < 0 || e.first >= nnode || ++visited > (long long)nnode) return -1;
        if (e.second > deepest) deepest = e.second;
        const int c0 = nodes[e.first].d.x, c1 = nodes[e.first].d.y;
        if (c0 >= 0 && c0 != kBvhNoChild) todo.emplace_back(c0, e.second + 1);
        if (c1 >= 0 && c1 != kBvhNoChild) todo.emplace_back(c1, e.second + 1);
    }
    return deepest;
}

#if defined(__CUDACC__)
// slab test against a box given as lo/hi ; returns entry distance, or +inf when missed, and the exit
// distance in texit.  (lo - o) * idir is kept in this exact-difference form on purpose: the cheaper
// fma(lo, idir, -o*idir) loses ~|o|/|lo-o| ulps to cancellation and would cull boxes far from the
// origin wrongly.  idir may hold +-inf for axis-parallel rays; fminf/fmaxf drop the NaN of 0*inf.
// The test only culls: prim boxes are padded at build time and tf is widened, so it is conservative
// against the prims' own arithmetic.
__device__ __forceinline__ float box_entry(float lox, float loy, float loz, float hix, float hiy, float hiz,
                                           const float3& o, const float3& idir, float tmin, float tbest, float& texit) {
    float tx0 = (lox - o.x) * idir.x, tx1 = (hix - o.x) * idir.x;
    float ty0 = (loy - o.y) * idir.y, ty1 = (hiy - o.y) * idir.y;
    float tz0 = (loz - o.z) * idir.z, tz1 = (hiz - o.z) * idir.z;
    float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), tmin));
    float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tbest));
    tf *= 1.0000004f;        // conservative: never cull a box whose prim would report a hit at its face
    texit = tf;
    return tn <= tf ? tn : CUDART_INF_F;
}
// same test, the verdict as a predicate (tentry is only meaningful for a hit)
__device__ __forceinline__ bool box_hit(float lox, float loy, float loz, float hix, float hiy, float hiz,
                                        const float3& o, const float3& idir, float tmin, float tbest, float& tentry, float& texit) {
    float tx0 = (lox - o.x) * idir.x, tx1 = (hix - o.x) * idir.x;
    float ty0 = (loy - o.y) * idir.y, ty1 = (hiy - o.y) * idir.y;
    float tz0 = (loz - o.z) * idir.z, tz1 = (hiz - o.z) * idir.z;
    float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), tmin));
    float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tbest));
    tf *= 1.0000004f;
    tentry = tn; texit = tf;
    return tn <= tf;
}

// For a box that IS the prim (box3 leaf without rotation) whose padded copy is (lo, hi): when the ray origin lies
// inside the box shrunk by `slack` on every entering side, the prim's own answer is its exit face, which cannot be
// nearer than the exit of the shrunk box.  Returns that bound (scaled down by a few ulps), else `entry`.
// Axis-parallel rays give inf - inf = NaN on the parallel axes, which fminf/fmaxf drop.
__device__ __forceinline__ float box_exit_bound(float lox, float loy, float loz, float hix, float hiy, float hiz,
                                                const float3& o, const float3& idir, float slack, float entry) {
    float sx = slack * fabsf(idir.x), sy = slack * fabsf(idir.y), sz = slack * fabsf(idir.z);
    float tx0 = (lox - o.x) * idir.x, tx1 = (hix - o.x) * idir.x;
    float ty0 = (loy - o.y) * idir.y, ty1 = (hiy - o.y) * idir.y;
    float tz0 = (loz - o.z) * idir.z, tz1 = (hiz - o.z) * idir.z;
    float near_max = fmaxf(fmaxf(fminf(tx0, tx1) + sx, fminf(ty0, ty1) + sy), fminf(tz0, tz1) + sz);
    float far_min = fminf(fminf(fmaxf(tx0, tx1) - sx, fmaxf(ty0, ty1) - sy), fmaxf(tz0, tz1) - sz);
    return near_max < 0.f ? fmaxf(entry, far_min * 0.999999f) : entry;
}
#endif

}  // namespace phox
