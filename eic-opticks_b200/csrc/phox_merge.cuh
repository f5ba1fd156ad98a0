// phox_merge.cuh : hit merging by (sensor identity, time bucket) - see phox_merge.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "phox_types.h"

namespace phox {

struct MergeScratch {
    void* buf = nullptr;
    size_t bytes = 0;
};

// d_in[n] -> d_out (capacity n), *n_out records.  select_mask: any-bit flagmask selection (0 = take all).
// tw > 0: merge per (identity, time / tw) group; tw == 0: selection only, input order.  Synchronises the stream.
cudaError_t merge_photons(const Photon* d_in, int64_t n, unsigned select_mask, float tw, Photon* d_out, int64_t* n_out, MergeScratch& scratch,
                          cudaStream_t stream, int* kernel_count);
cudaError_t merge_photons_lite(const PhotonLite* d_in, int64_t n, unsigned select_mask, float tw, PhotonLite* d_out, int64_t* n_out,
                               MergeScratch& scratch, cudaStream_t stream, int* kernel_count);
void merge_scratch_free(MergeScratch& scratch);

}  // namespace phox
