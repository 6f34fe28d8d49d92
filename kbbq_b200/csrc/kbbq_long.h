// kbbq_long.h -- the long-read repair kernel's launcher (kbbq_long.cu), called by kbbq_capi.cu
#pragma once
#include <algorithm>

#include "kbbq_views.cuh"

namespace kbbq {

// pass3_repair_long_kernel over the reads of the repair list; max_len = the batch's longest read (> KBBQ_MAX_READ_LEN)
cudaError_t launch_repair_long(const BatchView& b, const BloomView& trusted, int k, const uint32_t* tflags, uint32_t* err_out,
                               uint32_t* mod_out, const uint32_t* repair_list, const unsigned int* repair_count,
                               unsigned int* work_counter, uint32_t max_len, int sm_count, cudaStream_t stream);

}  // namespace kbbq
