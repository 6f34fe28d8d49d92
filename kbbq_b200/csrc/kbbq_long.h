// kbbq_long.h -- the long-read repair kernel's launcher (kbbq_long.cu), called by kbbq_capi.cu
#pragma once
#include <algorithm>

#include "kbbq_views.cuh"

namespace kbbq {

// the repair list of a long batch, split: short reads to the front of out[n_reads], long ones to its end (counts[0], counts[2])
cudaError_t launch_split_repair_list(const uint32_t* len, const uint32_t* list, const unsigned int* n_list, uint32_t n_reads,
                                     uint32_t* out, unsigned int* counts, int sm_count, cudaStream_t stream);
// pass3_repair_long_kernel over the reads repair_list[0], repair_list[-1], ... (the END of a split list);
// max_len = the batch's longest read (> KBBQ_MAX_READ_LEN)
cudaError_t launch_repair_long(const BatchView& b, const BloomView& trusted, int k, const uint32_t* tflags, uint32_t* err_out,
                               uint32_t* mod_out, const uint32_t* repair_list, const unsigned int* repair_count,
                               unsigned int* work_counter, uint32_t max_len, int sm_count, cudaStream_t stream);

}  // namespace kbbq
