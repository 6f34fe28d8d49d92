// tools/probe_ceiling.cu -- the honest ceiling for Bloom probes on this GPU: random 64-byte
// block reads and random 64-byte block atomicOr updates over tables of several sizes (L2-resident to HBM-resident),
// plus a streaming copy for scale.  Not part of the product; prints one JSON line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_ceiling tools/probe_ceiling.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint64_t mix(uint64_t x) {  // splitmix64
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// four lanes per block, one 16-byte quarter each (the pass-2 probe shape)
__global__ void gather_coop(const uint4* __restrict__ t, uint64_t n_blocks, uint64_t probes, unsigned long long* sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    const int sub = threadIdx.x & 3;
    unsigned long long acc = 0;
    for (uint64_t p = tid >> 2; p < probes; p += nthr >> 2) {
        const uint64_t blk = mix(p) % n_blocks;
        const uint4 v = __ldg(t + blk * 4 + sub);
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}
// one lane per block, two 32-byte loads (the pass-3 probe shape)
__global__ void gather_lane(const unsigned long long* __restrict__ t, uint64_t n_blocks, uint64_t probes, unsigned long long* sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (uint64_t p = tid; p < probes; p += nthr) {
        const uint64_t blk = mix(p) % n_blocks;
        unsigned long long a, b, c, d, e, f, g, h;
        asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(t + blk * 8));
        asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(e), "=l"(f), "=l"(g), "=l"(h) : "l"(t + blk * 8 + 4));
        acc += a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}
// insert shape: four lanes per block, atomicOr on `words_per_quarter` of each lane's two 64-bit words
__global__ void or_coop(unsigned long long* t, uint64_t n_blocks, uint64_t probes, int both) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    const int sub = threadIdx.x & 3;
    for (uint64_t p = tid >> 2; p < probes; p += nthr >> 2) {
        const uint64_t h = mix(p), blk = h % n_blocks;
        unsigned long long* B = t + blk * 8 + 2 * sub;
        atomicOr(B, 1ull << (h >> 58));
        if (both) atomicOr(B + 1, 1ull << ((h >> 52) & 63));
    }
}
__global__ void copy_k(const uint4* __restrict__ a, uint4* __restrict__ b, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <class F>
static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const uint64_t sizes[] = {64ull << 20, 512ull << 20, 1ull << 30, 8ull << 30};
    const uint64_t probes = 1ull << 28;
    unsigned long long* sink;
    cudaMalloc(&sink, 8);
    printf("{\"sms\": %d, \"probes\": %llu, \"tables\": [", sms, (unsigned long long)probes);
    for (int s = 0; s < 4; ++s) {
        unsigned long long* t = nullptr;
        if (cudaMalloc(&t, sizes[s]) != cudaSuccess) break;
        cudaMemset(t, 0, sizes[s]);
        const uint64_t nb = sizes[s] / 64;
        const int grid = sms * 8;
        const float g1 = time_ms([&] { gather_coop<<<grid, 256>>>((const uint4*)t, nb, probes, sink); }, 3);
        const float g2 = time_ms([&] { gather_lane<<<grid, 256>>>(t, nb, probes, sink); }, 3);
        const float o1 = time_ms([&] { or_coop<<<grid, 256>>>(t, nb, probes, 0); }, 3);
        const float o2 = time_ms([&] { or_coop<<<grid, 256>>>(t, nb, probes, 1); }, 3);
        printf("%s{\"table_MB\": %llu, \"gather_coop_Gprobes_s\": %.2f, \"gather_coop_GBps\": %.0f, \"gather_lane_Gprobes_s\": %.2f, "
               "\"or4_Gblocks_s\": %.2f, \"or8_Gblocks_s\": %.2f, \"or8_GBps_rmw128\": %.0f}",
               s ? ", " : "", (unsigned long long)(sizes[s] >> 20), probes / g1 / 1e6, probes * 64.0 / g1 / 1e6, probes / g2 / 1e6,
               probes / o1 / 1e6, probes / o2 / 1e6, probes * 128.0 / o2 / 1e6);
        cudaFree(t);
    }
    {
        const uint64_t n = (2ull << 30) / 16;
        uint4 *a, *b;
        cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16);
        cudaMemset(a, 1, n * 16);
        const float c = time_ms([&] { copy_k<<<sms * 16, 256>>>(a, b, n); }, 5);
        printf("], \"copy_GBps\": %.0f}\n", 2.0 * n * 16 / c / 1e6);
    }
    return 0;
}
