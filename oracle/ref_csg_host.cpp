// oracle/ref_csg_host.cpp : TEST INFRASTRUCTURE.  Compiles the REFERENCE's CSG intersect headers
// (CSG/csg_intersect_{leaf,node,tree}.h, read in place from /root/reference) for the host and
// exposes intersect_prim through a C ABI, so the CPU restatement in phox_oracle.cpp can be checked
// ray by ray against the reference's own code without a GPU.  Output: oracle/_ref/libcsgref.so.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector_types.h>
#include <vector_functions.h>
#include "scuda.h"
#include "squad.h"
#include "sqat4.h"
#include "csg_intersect_leaf.h"
#include "csg_intersect_node.h"
#include "csg_intersect_tree.h"

extern "C" int csgref_intersect_prim_batch(const void* node, int node_offset, const void* plan, const void* itra, const float* o, const float* d,
                                           const float* tmin, int n, float* isect_out, int* valid_out) {
    const CSGNode* nd = (const CSGNode*)node + node_offset;
    for (int i = 0; i < n; i++) {
        float4 isect = make_float4(0.f, 0.f, 0.f, 0.f);
        float3 ro = make_float3(o[3 * i], o[3 * i + 1], o[3 * i + 2]);
        float3 rd = make_float3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        bool valid = intersect_prim(isect, nd, (const float4*)plan, (const qat4*)itra, tmin[i], ro, rd, false);
        isect_out[4 * i] = isect.x; isect_out[4 * i + 1] = isect.y; isect_out[4 * i + 2] = isect.z; isect_out[4 * i + 3] = isect.w;
        valid_out[i] = valid ? 1 : 0;
    }
    return 0;
}
