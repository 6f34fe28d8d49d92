// tools/probe_lab.cu -- round-2 measurements behind the kernel choices (not part of the product; prints JSON lines).
//
//  1. Bloom probe variants against HBM-resident (1 GiB) and L2-resident (64 MiB) tables:
//       blk          four lanes per 64-byte block, block only                                  (the floor)
//       full         ... + the 64-byte pattern from a 4 MiB table (what round 1 shipped)
//       full_hint    ... blocks loaded evict_first, patterns evict_last
//       compact      ... + a 16-byte compact pattern (<= 14 nine-bit positions) from a 1 MiB table
//       compact_hint ... with the same cache hints
//       lane_full    one lane per probe, 2 x 32-byte loads of block and pattern (round 1's error-caller probe)
//       lane_compact one lane per probe, block 2 x 32 bytes + compact pattern
//  2. atomic ceilings for the covariate kernel: shared-memory atomicAdd (spread / 5 hot addresses), plain
//     lane-private shared read-modify-write, global RED (spread over 6000 cells / 20 hot cells).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/probe_lab tools/probe_lab.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ uint64_t policy_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 ld16(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld16_hint(const void* p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}

enum { V_BLK = 0, V_FULL, V_FULL_HINT, V_COMPACT, V_COMPACT_HINT };

// quarter mask of a compact pattern: positions whose (pos >> 7) == sub, as a 128-bit mask {lo, hi}
__device__ __forceinline__ void compact_quarter(uint4 c, int sub, int npos, uint64_t& lo, uint64_t& hi) {
    const uint64_t w0 = ((uint64_t)c.y << 32) | c.x, w1 = ((uint64_t)c.w << 32) | c.z;  // 7 positions each
    lo = 0; hi = 0;
#pragma unroll
    for (int i = 0; i < 14; ++i) {
        if (i < npos) {
            const uint32_t pos = (uint32_t)((i < 7 ? w0 >> (9 * i) : w1 >> (9 * (i - 7))) & 511u);
            const uint64_t bit = ((pos >> 7) == (uint32_t)sub) ? 1ull << (pos & 63) : 0ull;
            if (pos & 64) hi |= bit; else lo |= bit;
        }
    }
}

template <int V>
__global__ void __launch_bounds__(256) probe_coop(const uint4* __restrict__ t, uint64_t n_blocks, const uint4* __restrict__ pat,
                                                  const uint4* __restrict__ cpat, uint64_t probes, int npos, unsigned long long* sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    const int sub = threadIdx.x & 3;
    const uint64_t pf = policy_first(), pl = policy_last();
    unsigned long long acc = 0;
    // four probes per lane group in flight, like the product kernels
    for (uint64_t p0 = (tid >> 2) * 4; p0 < probes; p0 += (nthr >> 2) * 4) {
        uint4 x[4], y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint64_t h = mix(p0 + i);
            const uint64_t blk = __umul64hi(h, n_blocks);
            const uint32_t pn = (uint32_t)h & 0xFFFFu;
            if (V == V_FULL_HINT || V == V_COMPACT_HINT) x[i] = ld16_hint(t + blk * 4 + sub, pf);
            else x[i] = ld16(t + blk * 4 + sub);
            if (V == V_FULL) y[i] = ld16(pat + (uint64_t)pn * 4 + sub);
            else if (V == V_FULL_HINT) y[i] = ld16_hint(pat + (uint64_t)pn * 4 + sub, pl);
            else if (V == V_COMPACT) y[i] = ld16(cpat + pn);
            else if (V == V_COMPACT_HINT) y[i] = ld16_hint(cpat + pn, pl);
            else y[i] = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bool ok;
            if (V == V_COMPACT || V == V_COMPACT_HINT) {
                uint64_t lo, hi;
                compact_quarter(y[i], sub, npos, lo, hi);
                const uint64_t blo = ((uint64_t)x[i].y << 32) | x[i].x, bhi = ((uint64_t)x[i].w << 32) | x[i].z;
                ok = ((~blo & lo) | (~bhi & hi)) == 0;
            } else {
                ok = ((~x[i].x & y[i].x) | (~x[i].y & y[i].y) | (~x[i].z & y[i].z) | (~x[i].w & y[i].w)) == 0;
            }
            uint32_t m = __ballot_sync(0xffffffffu, ok);
            m &= (m >> 1) & (m >> 2) & (m >> 3) & 0x11111111u;
            acc += m;
        }
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}

template <bool COMPACT>
__global__ void __launch_bounds__(256) probe_lane(const unsigned long long* __restrict__ t, uint64_t n_blocks,
                                                  const unsigned long long* __restrict__ pat, const uint4* __restrict__ cpat,
                                                  uint64_t probes, int npos, unsigned long long* sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (uint64_t p = tid; p < probes; p += nthr) {
        const uint64_t h = mix(p);
        const uint64_t blk = __umul64hi(h, n_blocks);
        const uint32_t pn = (uint32_t)h & 0xFFFFu;
        unsigned long long b[8];
        asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(b[0]), "=l"(b[1]), "=l"(b[2]), "=l"(b[3]) : "l"(t + blk * 8));
        asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(b[4]), "=l"(b[5]), "=l"(b[6]), "=l"(b[7]) : "l"(t + blk * 8 + 4));
        if (COMPACT) {
            const uint4 c = ld16(cpat + pn);
            const uint64_t w0 = ((uint64_t)c.y << 32) | c.x, w1 = ((uint64_t)c.w << 32) | c.z;
            bool ok = true;
#pragma unroll
            for (int i = 0; i < 14; ++i) {
                if (i < npos) {
                    const uint32_t pos = (uint32_t)((i < 7 ? w0 >> (9 * i) : w1 >> (9 * (i - 7))) & 511u);
                    // select word pos >> 6 of eight without dynamic register indexing
                    const uint32_t wi = pos >> 6;
                    unsigned long long w = b[0];
#pragma unroll
                    for (int j = 1; j < 8; ++j) w = wi == (uint32_t)j ? b[j] : w;
                    ok = ok && ((w >> (pos & 63)) & 1ull);
                }
            }
            acc += ok;
        } else {
            unsigned long long q[8];
            asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(q[0]), "=l"(q[1]), "=l"(q[2]), "=l"(q[3]) : "l"(pat + (uint64_t)pn * 8));
            asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(q[4]), "=l"(q[5]), "=l"(q[6]), "=l"(q[7]) : "l"(pat + (uint64_t)pn * 8 + 4));
            unsigned long long miss = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) miss |= ~b[j] & q[j];
            acc += miss == 0;
        }
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}

// ---- atomic ceilings -------------------------------------------------------------------------------------------------
// every lane adds to shared cell idx(lane, step); mode 0: spread (lane-distinct cells), 1: five hot cells,
// 2: lane-private plain read-modify-write (no atomic), 3: match_any leaders add popc to five hot cells
__global__ void __launch_bounds__(512) smem_atomics(int mode, int iters, unsigned long long* sink) {
    extern __shared__ uint32_t cells[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) cells[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = (uint32_t)mix(blockIdx.x * 1024 + threadIdx.x);
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        if (mode == 0) atomicAdd(&cells[(x >> 8) & 8191], 1u);
        else if (mode == 1) atomicAdd(&cells[((x >> 8) % 5u) * 37u], 1u);
        else if (mode == 2) cells[warp * 512 + ((x >> 8) & 15) * 32 + lane] += 1u;
        else {
            const uint32_t key = (x >> 8) % 5u;
            const uint32_t same = __match_any_sync(0xffffffffu, key);
            if (lane == __ffs(same) - 1) atomicAdd(&cells[key * 37u], (uint32_t)__popc(same));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && cells[1] == 0xFFFFFFFFu) *sink = cells[1];
}
__global__ void __launch_bounds__(256) gmem_red(unsigned long long* cells, int n_cells, int iters) {
    uint32_t x = (uint32_t)mix((uint64_t)blockIdx.x * 1024 + threadIdx.x);
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        atomicAdd(&cells[(x >> 8) % (uint32_t)n_cells], 1ull);
    }
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
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms / reps;
}

int main(int argc, char** argv) {
    const bool quick = argc > 1 && !strcmp(argv[1], "quick");  // one launch per variant (for an ncu pass)
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const uint64_t probes = quick ? 1ull << 26 : 1ull << 28;
    const int reps = quick ? 1 : 3;
    unsigned long long* sink;
    cudaMalloc(&sink, 8);
    // pattern tables: 65536 patterns of 11 random bits, full (64 B) and compact (16 B)
    const int npos = 11;
    std::vector<uint64_t> full((size_t)65536 * 8, 0), comp((size_t)65536 * 2, 0);
    uint64_t s = 12345;
    for (int p = 0; p < 65536; ++p)
        for (int i = 0; i < 14; ++i) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            static uint32_t first;
            uint32_t pos = (uint32_t)(s >> 33) & 511u;
            if (i == 0) first = pos;
            if (i >= npos) pos = first;  // padding repeats a position: harmless
            full[(size_t)p * 8 + (pos >> 6)] |= 1ull << (pos & 63);
            comp[(size_t)p * 2 + (i / 7)] |= (uint64_t)pos << (9 * (i % 7));
        }
    uint4 *d_full, *d_comp;
    cudaMalloc(&d_full, full.size() * 8);
    cudaMalloc(&d_comp, comp.size() * 8);
    cudaMemcpy(d_full, full.data(), full.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_comp, comp.data(), comp.size() * 8, cudaMemcpyHostToDevice);
    const uint64_t sizes[] = {1ull << 30, 64ull << 20};
    for (int si = 0; si < 2; ++si) {
        unsigned long long* t = nullptr;
        if (cudaMalloc(&t, sizes[si]) != cudaSuccess) break;
        cudaMemset(t, 0xFF, sizes[si]);  // every probe hits: all four quarters are compared
        const uint64_t nb = sizes[si] / 64;
        const int grid = sms * 8;
        const uint4* tt = (const uint4*)t;
        float r[7];
        r[0] = time_ms([&] { probe_coop<V_BLK><<<grid, 256>>>(tt, nb, d_full, d_comp, probes, npos, sink); }, reps);
        r[1] = time_ms([&] { probe_coop<V_FULL><<<grid, 256>>>(tt, nb, d_full, d_comp, probes, npos, sink); }, reps);
        r[2] = time_ms([&] { probe_coop<V_FULL_HINT><<<grid, 256>>>(tt, nb, d_full, d_comp, probes, npos, sink); }, reps);
        r[3] = time_ms([&] { probe_coop<V_COMPACT><<<grid, 256>>>(tt, nb, d_full, d_comp, probes, npos, sink); }, reps);
        r[4] = time_ms([&] { probe_coop<V_COMPACT_HINT><<<grid, 256>>>(tt, nb, d_full, d_comp, probes, npos, sink); }, reps);
        r[5] = time_ms([&] { probe_lane<false><<<grid, 256>>>(t, nb, (const unsigned long long*)d_full, d_comp, probes, npos, sink); }, reps);
        r[6] = time_ms([&] { probe_lane<true><<<grid, 256>>>(t, nb, (const unsigned long long*)d_full, d_comp, probes, npos, sink); }, reps);
        const char* names[7] = {"blk", "full", "full_hint", "compact", "compact_hint", "lane_full", "lane_compact"};
        printf("{\"table_MB\": %llu, \"probes\": %llu", (unsigned long long)(sizes[si] >> 20), (unsigned long long)probes);
        for (int i = 0; i < 7; ++i) printf(", \"%s_Gprobes_s\": %.2f", names[i], probes / r[i] / 1e6);
        printf("}\n");
        cudaFree(t);
    }
    {
        const int iters = quick ? 2000 : 20000;
        cudaFuncSetAttribute(smem_atomics, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
        printf("{\"smem_ops_G_per_s\": {");
        const char* names[4] = {"atomic_spread", "atomic_5hot", "private_rmw", "match_any_5hot"};
        for (int mode = 0; mode < 4; ++mode) {
            const float ms = time_ms([&] { smem_atomics<<<sms * 2, 512, 32768>>>(mode, iters, sink); }, reps);
            printf("%s\"%s\": %.1f", mode ? ", " : "", names[mode], (double)sms * 2 * 512 * iters / ms / 1e6);
        }
        unsigned long long* cells;
        cudaMalloc(&cells, 8 * 8192);
        cudaMemset(cells, 0, 8 * 8192);
        const int it2 = quick ? 200 : 2000;
        const float g1 = time_ms([&] { gmem_red<<<sms * 8, 256>>>(cells, 6000, it2); }, reps);
        const float g2 = time_ms([&] { gmem_red<<<sms * 8, 256>>>(cells, 20, it2); }, reps);
        printf("}, \"gmem_red_G_per_s\": {\"spread_6000\": %.1f, \"hot_20\": %.1f}}\n", (double)sms * 8 * 256 * it2 / g1 / 1e6,
               (double)sms * 8 * 256 * it2 / g2 / 1e6);
    }
    return 0;
}
