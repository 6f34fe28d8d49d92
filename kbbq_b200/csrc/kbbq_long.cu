// kbbq_long.cu -- pass 3's repairs for batches that hold a read of more than KBBQ_MAX_READ_LEN (512) bases.
//
// The reference has no read-length limit (readutils.cc:64-104).  The short-read kernels keep their per-warp scratch in
// statically sized shared memory; here the same ErrCaller (kbbq_errcall.cuh) runs over ErrScratchDyn, thirteen arrays
// of words + 1 entries cut from dynamic shared memory sized for the batch's longest read.  A slow path on purpose; a
// translation unit of its own so that the common repair kernel's code does not depend on this one being there.
#include "kbbq_long.h"

#include "kbbq_errcall.cuh"

namespace kbbq {

// the error caller's warp-wide questions, one lane per probe (kbbq_kernels.cuh: DevQuery)
struct LongQuery {
    const BloomView* f;
    template <class F>
    __device__ uint32_t query(int count, F fn) {
        const int lane = threadIdx.x & 31;
        uint64_t key = 0;
        const bool valid = lane < count && fn(lane, key);
        return __ballot_sync(KBBQ_FULL, valid && bloom_contains_dev(*f, key));
    }
    template <class F>
    __device__ uint32_t map32(F fn) { return __ballot_sync(KBBQ_FULL, fn((int)(threadIdx.x & 31))); }
    template <class F>
    __device__ void par(int n, F fn) {
        for (int i = (int)(threadIdx.x & 31); i < n; i += 32) fn(i);
        __syncwarp();
    }
};

// The same for a batch whose longest read exceeds KBBQ_MAX_READ_LEN: the scratch is cut from dynamic shared memory,
// ErrScratchDyn::bytes(maxw) per warp, the block has as many warps as fit (the launch decides: 8 down to 1), and the
// per-word set-up walks the read 32 words at a time.  A slow path: 198 registers, one block per SM.
__global__ void __launch_bounds__(KBBQ_WARPS_PER_BLOCK * 32, 1)
pass3_repair_long_kernel(BatchView b, BloomView trusted, int k, const uint32_t* __restrict__ tflags, uint32_t* __restrict__ err_out,
                         uint32_t* __restrict__ mod_out, const uint32_t* __restrict__ repair_list,
                         const unsigned int* __restrict__ repair_count, unsigned int* __restrict__ work_counter, int maxw) {
    extern __shared__ __align__(16) unsigned char repair_dyn[];  // [warps of the block][ErrScratchDyn::bytes(maxw)]
    __shared__ BloomView filter;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (threadIdx.x == 0) filter = trusted;
    __syncthreads();
    LongQuery q{&filter};
    ErrScratchDyn sc;
    sc.carve(repair_dyn + (size_t)wib * ErrScratchDyn::bytes(maxw), maxw);
    const uint32_t n_list = *repair_count;
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = atomicAdd(work_counter, 1u);
        idx = __shfl_sync(KBBQ_FULL, idx, 0);
        if (idx >= n_list) break;
        const uint32_t r = repair_list[-(long long)idx];  // the long reads sit at the END of the split list, last first
        const int len = (int)b.len[r];
        const uint32_t wo = b.word_off[r];
        const int nw = (len + 31) >> 5;
        const uint8_t* qual = b.qual + (size_t)wo * 32;
        bool any_n = false;
        for (int w = lane; w <= nw; w += 32) {
            sc.P_w[w] = w < nw ? b.bases[wo + w] : 0ull;
            const uint32_t nm = w < nw ? b.nmask[wo + w] : 0u;
            sc.P_nm[w] = nm;
            any_n = any_n || nm != 0u;
            sc.E[w] = 0;
            sc.T_orig[w] = 32 * w < len - k + 1 ? tflags[wo + w] : 0u;
            const int left = len - 32 * w;
            sc.QL[w] = w < nw ? b.qmask[wo + w].y & (left >= 32 ? ~0u : ((1u << left) - 1u)) : 0u;
        }
        __syncwarp();
        const bool has_n = __any_sync(KBBQ_FULL, any_n);
        ErrCaller<LongQuery, ErrScratchDyn> ec(q, k, qual, sc, has_n);
        const ErrResult res = ec.run(len, true);
        __syncwarp();
        for (int w = lane; w < nw; w += 32) err_out[wo + w] = sc.E[w];
        if (lane == 0) mod_out[r] = res.mod_idx >= 0 ? (uint32_t)(4 * res.mod_idx + res.mod_base) : 0xFFFFFFFFu;
        __syncwarp();
    }
}


// Splits the scan's repair list of a long batch: reads of up to KBBQ_MAX_READ_LEN bases to the front of `out` (they go
// through the common repair kernel at its full speed), longer ones to its end; counts[0] / counts[2] = how many of each
// (counts[1], counts[3]: the two kernels' work counters, zeroed by the caller).
__global__ void __launch_bounds__(256) split_repair_list_kernel(const uint32_t* __restrict__ len, const uint32_t* __restrict__ list,
                                                                const unsigned int* __restrict__ n_list, uint32_t n_reads,
                                                                uint32_t* __restrict__ out, unsigned int* __restrict__ counts) {
    const uint32_t n = *n_list;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t r = list[i];
        if (len[r] <= KBBQ_MAX_READ_LEN) out[atomicAdd(counts, 1u)] = r;
        else out[n_reads - 1 - atomicAdd(counts + 2, 1u)] = r;
    }
}
cudaError_t launch_split_repair_list(const uint32_t* len, const uint32_t* list, const unsigned int* n_list, uint32_t n_reads,
                                     uint32_t* out, unsigned int* counts, int sm_count, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(counts, 0, 4 * sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    const unsigned blocks = (unsigned)std::max(1ll, std::min(((long long)n_reads + 255) / 256, (long long)sm_count * 8));
    split_repair_list_kernel<<<blocks, 256, 0, stream>>>(len, list, n_list, n_reads, out, counts);
    return cudaGetLastError();
}

cudaError_t launch_repair_long(const BatchView& b, const BloomView& trusted, int k, const uint32_t* tflags, uint32_t* err_out,
                               uint32_t* mod_out, const uint32_t* repair_list, const unsigned int* repair_count,
                               unsigned int* work_counter, uint32_t max_len, int sm_count, cudaStream_t stream) {
    const int maxw = ((int)max_len + 31) / 32;
    const size_t per_warp = ErrScratchDyn::bytes(maxw);
    const int warps = (int)std::max<size_t>(1, std::min<size_t>(KBBQ_WARPS_PER_BLOCK, (200 * 1024) / per_warp));
    const size_t smem = per_warp * warps;
    if (smem > 220 * 1024) return cudaErrorInvalidValue;  // beyond KBBQ_ABI_MAX_READ_LEN
    cudaError_t e = cudaFuncSetAttribute(pass3_repair_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long blocks = std::max(1ll, std::min(((long long)b.n_reads + warps - 1) / warps, (long long)sm_count * 4));
    pass3_repair_long_kernel<<<(unsigned)blocks, warps * 32, smem, stream>>>(b, trusted, k, tflags, err_out, mod_out, repair_list,
                                                                            repair_count, work_counter, maxw);
    return cudaGetLastError();
}

}  // namespace kbbq
