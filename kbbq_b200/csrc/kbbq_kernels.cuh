// kbbq_kernels.cuh -- the sm_100a kernels of the recalibration hot path.
//
// Mapping used throughout: ONE WARP PER READ (reads are 100-250 bases: a handful of 32-base words), lanes fan out
// over k-mer start positions or base positions, per-read intermediates are 32-bit masks produced by __ballot_sync
// and kept in a few words of shared memory.  Filter probes are 64-byte block reads / atomicOr on 64-bit words of a
// table that lives in HBM (or L2 when it fits); the 4 MiB pattern table stays L2-resident.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kbbq_core.cuh"
#include "kbbq_errcall.cuh"

namespace kbbq {

struct BatchView {
    const unsigned long long* bases;
    const uint32_t* nmask;
    const uint8_t* qual;
    const uint32_t* word_off;
    const uint32_t* len;
    const uint16_t* rg;
    const uint8_t* second;
    const unsigned long long* slot_off;  // exclusive prefix of max(0, len-k+1)
    uint32_t n_reads;
};

struct CountView {  // covariate count tables, {errors,total} pairs of u64
    unsigned long long* rgc;
    unsigned long long* qc;
    unsigned long long* cyc;
    unsigned long long* dic;
    int max_cycle;
};

struct ModelView {  // dq_t folded into lookup tables (readutils.cc:579-587)
    const int16_t* base;  // [R][94]  meanq + rgdq + qscoredq
    const int8_t* cy;     // [R][94][2][C]
    const int8_t* di;     // [R][94][16]
    int max_cycle;
};

// read-only view of a packed read in global memory, usable with kmer_key()
struct CSeqRef {
    const unsigned long long* w;
    const uint32_t* nm;
    int len;
};

#define KBBQ_WARPS_PER_BLOCK 8
#define KBBQ_FULL 0xffffffffu

// ---- filter probes ------------------------------------------------------------------------------------------------
// 32 bytes in one LDG.E.256 (sm_100): a 64-byte block or pattern is two instructions
__device__ __forceinline__ void ldg256(const unsigned long long* p, unsigned long long& a, unsigned long long& b,
                                       unsigned long long& c, unsigned long long& d) {
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
// pattern_blocked_bf::contains (bloom.hh:276-292): (block & pattern) == pattern over 512 bits, one lane per key
__device__ __forceinline__ bool bloom_contains_dev(const BloomView& f, uint64_t key) {
    const uint64_t blk = bloom_block(f, key);
    const uint32_t pat = bloom_pattern(f, key);
    const unsigned long long* B = f.words + blk * 8;
    const unsigned long long* P = f.patterns + (uint64_t)pat * 8;
    unsigned long long b0, b1, b2, b3, b4, b5, b6, b7, p0, p1, p2, p3, p4, p5, p6, p7;
    ldg256(B, b0, b1, b2, b3);
    ldg256(B + 4, b4, b5, b6, b7);
    ldg256(P, p0, p1, p2, p3);
    ldg256(P + 4, p4, p5, p6, p7);
    const unsigned long long miss = (~b0 & p0) | (~b1 & p1) | (~b2 & p2) | (~b3 & p3) | (~b4 & p4) | (~b5 & p5) | (~b6 & p6) | (~b7 & p7);
    return miss == 0;
}

// pattern_blocked_bf::insert (bloom.hh:255-267): block |= pattern.  OR is order independent, so atomicOr on the
// non-zero 64-bit words gives the reference's bits whatever the interleaving.  The caller counts the insert.
__device__ __forceinline__ void bloom_insert_dev(const BloomView& f, uint64_t key) {
    const uint64_t blk = bloom_block(f, key);
    const uint32_t pat = bloom_pattern(f, key);
    const ulonglong2* P = reinterpret_cast<const ulonglong2*>(f.patterns + (uint64_t)pat * 8);
    unsigned long long* B = f.words + blk * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const ulonglong2 p = __ldg(P + i);
        if (p.x) atomicOr(B + 2 * i, p.x);
        if (p.y) atomicOr(B + 2 * i + 1, p.y);
    }
}

// ---- warp-cooperative probes ---------------------------------------------------------------------------------------
// A probe touches one 64-byte block at a random address plus one 64-byte pattern.  With one lane per probe a warp
// load instruction hits 32 different cache lines and uses 16 of each line's 128 bytes: 4x the L1/L2 wavefronts the
// data needs.  Here FOUR lanes share a probe, each moving one 16-byte quarter, so a warp instruction covers 8 probes
// as 8 fully used 64-byte segments; a 32-probe query is 4 rounds, all of whose loads are issued before any is used.

// eight flags spaced 4 bits apart (bits 0,4,..,28) -> bits 0..7
__device__ __forceinline__ uint32_t gather_nibble_flags(uint32_t t) {
    t = (t | (t >> 3)) & 0x03030303u;
    t = (t | (t >> 6)) & 0x000F000Fu;
    return (t | (t >> 12)) & 0xFFu;
}

// The probe proper: every lane brings the (block, pattern) of its own key; returns the mask of lanes that hit.
template <int ROUNDS_AT_ONCE>
__device__ __forceinline__ uint32_t coop_probe_impl(const unsigned long long* __restrict__ words,
                                                    const unsigned long long* __restrict__ patterns, uint64_t blk,
                                                    uint32_t pat, uint32_t vmask) {
    const int lane = threadIdx.x & 31, sub = lane & 3, grp = lane >> 2;
    uint32_t result = 0;
#pragma unroll
    for (int r0 = 0; r0 < 4; r0 += ROUNDS_AT_ONCE) {
        if (!((vmask >> (8 * r0)) & (ROUNDS_AT_ONCE == 4 ? 0xFFFFFFFFu : ((1u << (8 * ROUNDS_AT_ONCE)) - 1)))) continue;
        uint4 x[ROUNDS_AT_ONCE], y[ROUNDS_AT_ONCE];
        bool v[ROUNDS_AT_ONCE];
#pragma unroll
        for (int i = 0; i < ROUNDS_AT_ONCE; ++i) {  // issue every load of the group before using any
            const int src = 8 * (r0 + i) + grp;
            const uint64_t b = __shfl_sync(KBBQ_FULL, blk, src);
            const uint32_t p = __shfl_sync(KBBQ_FULL, pat, src);
            v[i] = (vmask >> src) & 1u;
            x[i] = make_uint4(0, 0, 0, 0);
            y[i] = make_uint4(0, 0, 0, 0);
            if (v[i]) {
                x[i] = __ldg(reinterpret_cast<const uint4*>(words + b * 8) + sub);
                y[i] = __ldg(reinterpret_cast<const uint4*>(patterns + (uint64_t)p * 8) + sub);
            }
        }
#pragma unroll
        for (int i = 0; i < ROUNDS_AT_ONCE; ++i) {
            const bool ok = v[i] && ((~x[i].x & y[i].x) | (~x[i].y & y[i].y) | (~x[i].z & y[i].z) | (~x[i].w & y[i].w)) == 0;
            uint32_t m = __ballot_sync(KBBQ_FULL, ok);
            m &= (m >> 1) & (m >> 2) & (m >> 3) & 0x11111111u;  // all four quarters of a probe agree
            result |= gather_nibble_flags(m) << (8 * (r0 + i));
        }
    }
    return result;
}
// out-of-line copy for the error-calling kernel, which asks from a dozen places
__device__ __noinline__ uint32_t coop_probe_call(const unsigned long long* words, const unsigned long long* patterns,
                                                 uint64_t blk, uint32_t pat, uint32_t vmask) {
    return coop_probe_impl<2>(words, patterns, blk, pat, vmask);
}

// mask of lanes whose key is valid and contained in f (pattern_blocked_bf::contains, bloom.hh:276-292).
// All 32 lanes must call.
template <bool INLINE>
__device__ __forceinline__ uint32_t warp_bloom_query(const BloomView& f, bool valid, uint64_t key) {
    const uint32_t vmask = __ballot_sync(KBBQ_FULL, valid);
    if (vmask == 0) return 0;
    const uint64_t blk = valid ? bloom_block(f, key) : 0;
    const uint32_t pat = valid ? bloom_pattern(f, key) : 0;
    if (INLINE) return coop_probe_impl<4>(f.words, f.patterns, blk, pat, vmask);
    return coop_probe_call(f.words, f.patterns, blk, pat, vmask);
}

// pattern_blocked_bf::insert (bloom.hh:255-267) for the lanes with `ins` set; returns how many were inserted.
// The (block, pattern) pairs are first compacted through `stage` (32 x 12 bytes of per-warp shared memory) so that
// sparse inserts (pass 1 samples ~14-35 % of the k-mers) still fill every round.  All 32 lanes must call.
struct InsertStage {
    unsigned long long blk[32];
    uint32_t pat[32];
};
__device__ __forceinline__ int warp_bloom_insert(const BloomView& f, bool ins, uint64_t key, InsertStage& stage) {
    const uint32_t imask = __ballot_sync(KBBQ_FULL, ins);
    if (imask == 0) return 0;
    const int lane = threadIdx.x & 31, sub = lane & 3, grp = lane >> 2;
    const int n = __popc(imask);
    if (ins) {
        const int rank = __popc(imask & ((1u << lane) - 1));
        stage.blk[rank] = bloom_block(f, key);
        stage.pat[rank] = bloom_pattern(f, key);
    }
    __syncwarp();
    for (int base = 0; base < n; base += 8) {
        const int idx = base + grp;
        if (idx < n) {
            const ulonglong2 p = __ldg(reinterpret_cast<const ulonglong2*>(f.patterns + (uint64_t)stage.pat[idx] * 8) + sub);
            unsigned long long* B = f.words + stage.blk[idx] * 8 + 2 * sub;
            if (p.x) atomicOr(B, p.x);
            if (p.y) atomicOr(B + 1, p.y);
        }
    }
    __syncwarp();
    return n;
}

// ---- pass 1a: the sampling stream ---------------------------------------------------------------------------------
// One serial xoshiro256** stream decides every k-mer slot of the file (htsiter.cc:113-129).  The generator is linear
// over GF(2): thread t jumps straight to draw 1024*t with precomputed 256x256 bit matrices of 2^e steps
// (columns of 4 words; jump[b] = 2^(10+b) steps), then produces 1024 consecutive decisions as 32 mask words.
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

__global__ void __launch_bounds__(128) rng_sample_kernel(uint64_t s0, uint64_t s1, uint64_t s2, uint64_t s3,
                                                         const unsigned long long* __restrict__ jump,
                                                         unsigned long long n_slots, unsigned long long threshold,
                                                         int take_all, uint32_t* __restrict__ bits) {
    const unsigned long long chunk = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long first = chunk * 1024ull;
    if (first >= n_slots) return;
    uint64_t s[4] = {s0, s1, s2, s3};
    for (int b = 0; (chunk >> b) != 0; ++b) {
        if (!((chunk >> b) & 1)) continue;
        const unsigned long long* cols = jump + (size_t)b * 1024;
        uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint64_t m = s[w];
            while (m) {
                const int c = w * 64 + __ffsll((long long)m) - 1;
                m &= m - 1;
                const ulonglong2* col = reinterpret_cast<const ulonglong2*>(cols + (size_t)c * 4);
                const ulonglong2 x = __ldg(col), y = __ldg(col + 1);
                a0 ^= x.x; a1 ^= x.y; a2 ^= y.x; a3 ^= y.y;
            }
        }
        s[0] = a0; s[1] = a1; s[2] = a2; s[3] = a3;
    }
    for (int w = 0; w < 32; ++w) {
        uint32_t word = 0;
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            // Xoshiro256StarStarEngine::Next (minion.hpp:103-118)
            const uint64_t x = rotl64(s[1] * 5, 7) * 9;
            const uint64_t t = s[1] << 17;
            s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
            s[2] ^= t;
            s[3] = rotl64(s[3], 45);
            // std::bernoulli_distribution: generate_canonical<double,53> = double(x)/2^64 (clamped below 1) < p.
            // u64 -> double rounding is monotone, so the test is "x below a threshold" that the host found with the
            // very same double arithmetic (kbbq_sampler_seed): one integer compare instead of an I2F + DMUL + DSETP.
            if ((take_all || x < threshold) && first + (unsigned)(w * 32 + i) < n_slots) word |= 1u << i;
        }
        bits[chunk * 32 + w] = word;
    }
}

// ---- pass 1b: insert the sampled k-mers (recalibrateutils.cc:7-13) ---------------------------------------------------
__global__ void __launch_bounds__(KBBQ_WARPS_PER_BLOCK * 32) pass1_insert_kernel(BatchView b, BloomView f, int k,
                                                                                  const uint32_t* __restrict__ sample_bits) {
    __shared__ InsertStage stage_all[KBBQ_WARPS_PER_BLOCK];
    InsertStage& stage = stage_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long inserted = 0;
    for (uint32_t r = warp; r < b.n_reads; r += nwarps) {
        const int len = (int)b.len[r];
        const int nk = len - k + 1;
        if (nk <= 0) continue;  // fewer than k bases: no slot, no draw (htsiter.cc:100-105)
        const uint32_t wo = b.word_off[r];
        const CSeqRef s{b.bases + wo, b.nmask + wo, len};
        const unsigned long long so = b.slot_off[r];
        for (int j0 = 0; j0 < nk; j0 += 32) {
            const int j = j0 + lane;
            bool ins = false;
            uint64_t key = 0;
            if (j < nk) {
                const unsigned long long slot = so + (unsigned)j;
                if ((__ldg(sample_bits + (slot >> 5)) >> (slot & 31)) & 1u)
                    ins = kmer_key(s, j, k, 0, -1, key);  // an invalid k-mer still consumed its draw
            }
            inserted += warp_bloom_insert(f, ins, key, stage);
        }
    }
    if (lane == 0 && inserted) atomicAdd(f.inserted, inserted);
}

// ---- pass 2: infer_read_errors + find_trusted_kmers (readutils.cc:173-193, bloom.cc:28-67, recalibrateutils.cc:15-40)
__global__ void __launch_bounds__(KBBQ_WARPS_PER_BLOCK * 32)
pass2_trusted_kernel(BatchView b, BloomView sampled, BloomView trusted, int k, const int* __restrict__ thresholds,
                     uint32_t* __restrict__ err_out) {
    __shared__ int thr[KBBQ_MAX_K + 1];
    __shared__ uint32_t Pm_all[KBBQ_WARPS_PER_BLOCK][KBBQ_MAXW + 1];
    __shared__ uint32_t Em_all[KBBQ_WARPS_PER_BLOCK][KBBQ_MAXW + 1];
    __shared__ InsertStage stage_all[KBBQ_WARPS_PER_BLOCK];
    InsertStage& stage = stage_all[threadIdx.x >> 5];
    if (threadIdx.x <= k) thr[threadIdx.x] = thresholds[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* Pm = Pm_all[wib];
    uint32_t* Em = Em_all[wib];
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long inserted = 0;
    for (uint32_t r = warp; r < b.n_reads; r += nwarps) {
        const int len = (int)b.len[r];
        const int nk = len - k + 1;
        if (nk <= 0) continue;  // out of contract in the reference (bloom.cc:30 underflows)
        const uint32_t wo = b.word_off[r];
        const CSeqRef s{b.bases + wo, b.nmask + wo, len};
        const uint8_t* q = b.qual + (size_t)wo * 32;
        // kmer_present (bloom.cc:35-47)
        for (int j0 = 0; j0 < nk; j0 += 32) {
            const int j = j0 + lane;
            uint64_t key = 0;
            const bool valid = j < nk && kmer_key(s, j, k, 0, -1, key);
            const uint32_t m = warp_bloom_query<true>(sampled, valid, key);
            if (lane == 0) Pm[j0 >> 5] = m;
        }
        __syncwarp();
        // errors[i] = in[i] <= thresholds[possible[i]] || qual[i] <= 2 (readutils.cc:178)
        for (int i0 = 0; i0 < len; i0 += 32) {
            const int i = i0 + lane;
            bool e = false;
            if (i < len) {
                const int lo = i - k + 1 > 0 ? i - k + 1 : 0;
                const int hi = (i < nk - 1 ? i : nk - 1) + 1;
                const int in = count_range(Pm, lo, hi);
                e = in <= thr[hi - lo] || q[i] <= 2;
            }
            const uint32_t m = __ballot_sync(KBBQ_FULL, e);
            if (lane == 0) {
                Em[i0 >> 5] = m;
                if (err_out) err_out[wo + (i0 >> 5)] = m;
            }
        }
        __syncwarp();
        // a k-mer is trusted when it is valid and none of its k bases is flagged (recalibrateutils.cc:26-36)
        for (int j0 = 0; j0 < nk; j0 += 32) {
            const int j = j0 + lane;
            uint64_t key = 0;
            const bool ins = j < nk && count_range(Em, j, j + k) == 0 && kmer_key(s, j, k, 0, -1, key);
            inserted += warp_bloom_insert(trusted, ins, key, stage);
        }
        __syncwarp();
    }
    if (lane == 0 && inserted) atomicAdd(trusted.inserted, inserted);
}

// ---- covariate counting (covariateutils.cc:30-42,65-76,102-116,147-164) for one read, by one warp -----------------
// E = read.errors as bit masks; w/nm = the read as consume_read sees it.
__device__ __forceinline__ void consume_read_warp(const CountView& cv, const unsigned long long* w, const uint32_t* nm,
                                                  const uint8_t* q, const uint32_t* E, int len, int rg, int sec, int lane) {
    const int nw = (len + 31) >> 5;
    int nerr = lane < nw ? __popc(E[lane]) : 0;
#pragma unroll
    for (int d = 16; d; d >>= 1) nerr += __shfl_xor_sync(KBBQ_FULL, nerr, d);
    if (lane == 0) {  // CRGCovariate: {#errors, #bases}; skips are never set (covariateutils.cc:194-197)
        if (nerr) atomicAdd(cv.rgc + 2 * rg, (unsigned long long)nerr);
        atomicAdd(cv.rgc + 2 * rg + 1, (unsigned long long)len);
    }
    for (int i0 = 0; i0 < len; i0 += 32) {
        const int i = i0 + lane;
        const bool in = i < len;
        const int qi = in ? q[i] : 255;
        const bool e = in && ((E[i0 >> 5] >> lane) & 1u);
        const uint32_t emask = __ballot_sync(KBBQ_FULL, e);
        // CQCovariate: lanes with the same quality add through one leader
        const uint32_t same_q = __match_any_sync(KBBQ_FULL, qi);
        const size_t row = (size_t)rg * KBBQ_NQ + qi;
        if (in && lane == __ffs(same_q) - 1) {
            atomicAdd(cv.qc + row * 2 + 1, (unsigned long long)__popc(same_q));
            const int ne = __popc(same_q & emask);
            if (ne) atomicAdd(cv.qc + row * 2, (unsigned long long)ne);
        }
        // CCycleCovariate: cycle = index in read orientation, split by `second`
        if (in) {
            unsigned long long* c = cv.cyc + ((row * 2 + sec) * (size_t)cv.max_cycle + i) * 2;
            atomicAdd(c + 1, 1ull);
            if (e) atomicAdd(c, 1ull);
        }
        // CDinucCovariate: i >= 1, both bases ACGT, qual >= 6
        int dkey = -1;
        if (in && i >= 1 && qi >= 6) {
            const bool n1 = (nm[i >> 5] >> (i & 31)) & 1u, n0 = (nm[(i - 1) >> 5] >> ((i - 1) & 31)) & 1u;
            if (!n1 && !n0) {
                const int c1 = (int)((w[i >> 5] >> (2 * (i & 31))) & 3), c0 = (int)((w[(i - 1) >> 5] >> (2 * ((i - 1) & 31))) & 3);
                dkey = qi * 16 + ((c0 << 2) | c1);
            }
        }
        const uint32_t same_d = __match_any_sync(KBBQ_FULL, dkey);
        if (dkey >= 0 && lane == __ffs(same_d) - 1) {
            unsigned long long* d = cv.dic + ((size_t)rg * KBBQ_NQ * 16 + dkey) * 2;
            atomicAdd(d + 1, (unsigned long long)__popc(same_d));
            const int ne = __popc(same_d & emask);
            if (ne) atomicAdd(d, (unsigned long long)ne);
        }
    }
}

// ---- pass 3: get_errors + consume_read (recalibrateutils.cc:42-89) ------------------------------------------------
// one copy of hash + 4 loads + compare for every question the error caller asks (it asks from five places)
__device__ __noinline__ uint32_t probe_ballot(const BloomView* f, bool valid, uint64_t key) {
    return __ballot_sync(KBBQ_FULL, valid && bloom_contains_dev(*f, key));
}
struct DevQuery {
    const BloomView* f;  // in shared memory: a kernel parameter has no address an out-of-line function could take
    template <class F>
    KBBQ_HD uint32_t query(int count, F fn) {
#if defined(__CUDA_ARCH__)
        const int lane = threadIdx.x & 31;
        uint64_t key = 0;
        const bool valid = lane < count && fn(lane, key);
#if defined(KBBQ_P3_PROBE_INLINE)
        return __ballot_sync(KBBQ_FULL, valid && bloom_contains_dev(*f, key));
#else
        return probe_ballot(f, valid, key);
#endif
#else
        (void)count; (void)fn;
        return 0;
#endif
    }
    template <class F>
    KBBQ_HD uint32_t map32(F fn) {
#if defined(__CUDA_ARCH__)
        return __ballot_sync(KBBQ_FULL, fn((int)(threadIdx.x & 31)));
#else
        (void)fn;
        return 0;
#endif
    }
};

// 4 blocks/SM = 64 registers, 32 warps/SM: measured 21.2 ms vs 23.0 (unbounded, 80 regs) and 23.8 (5 blocks, spills)
#ifndef KBBQ_P3_MIN_BLOCKS
#define KBBQ_P3_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(KBBQ_WARPS_PER_BLOCK * 32, KBBQ_P3_MIN_BLOCKS)
pass3_errors_kernel(BatchView b, BloomView trusted, CountView cv, int k, uint32_t* __restrict__ err_out, int do_consume,
                    unsigned int* __restrict__ work_counter) {
    __shared__ ErrScratch scratch[KBBQ_WARPS_PER_BLOCK];
    __shared__ BloomView filter;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    ErrScratch& sc = scratch[wib];
    if (threadIdx.x == 0) filter = trusted;
    __syncthreads();
    DevQuery q{&filter};
    for (;;) {
        // reads differ a lot in cost (most need only the mandatory scan, a few need ~1000 probes): pull, don't stride
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(work_counter, 1u);
        r = __shfl_sync(KBBQ_FULL, r, 0);
        if (r >= b.n_reads) break;
        const int len = (int)b.len[r];
        const uint32_t wo = b.word_off[r];
        const int nw = (len + 31) >> 5;
        const uint8_t* qual = b.qual + (size_t)wo * 32;
        if (lane <= nw) {
            sc.P_w[lane] = lane < nw ? b.bases[wo + lane] : 0ull;
            sc.P_nm[lane] = lane < nw ? b.nmask[wo + lane] : 0u;
            sc.E[lane] = 0;
        }
        for (int i0 = 0; i0 <= len; i0 += 32) {  // low-quality flags for the overcorrection half weight
            const int i = i0 + lane;
            const uint32_t m = __ballot_sync(KBBQ_FULL, i < len && qual[i] <= 6);
            if (lane == 0) sc.QL[i0 >> 5] = m;
        }
        __syncwarp();
        ErrCaller<DevQuery> ec(q, k, qual, sc);
#if defined(KBBQ_P3_EXPERIMENT)
        ErrResult res;
        res.mod_idx = -1; res.mod_base = -1;
        if (do_consume & 2) { if (len >= k) { SeqRef P{sc.P_w, sc.P_nm, len}; ec.scan(P, sc.T_orig); } }
        else res = ec.run(len);
#else
        const ErrResult res = ec.run(len);
#endif
        __syncwarp();
        if (err_out && lane < nw) err_out[wo + lane] = sc.E[lane];
        if (do_consume & 1) {
            // consume_read sees the original bases (readutils.cc:568) except for the early-return quirk
            if (lane < nw) {
                sc.A_w[lane] = b.bases[wo + lane];
                sc.A_nm[lane] = b.nmask[wo + lane];
            }
            __syncwarp();
            if (res.mod_idx >= 0) {
                SeqRef a{sc.A_w, sc.A_nm, len};
                seq_set(a, res.mod_idx, res.mod_base);
                __syncwarp();
            }
            consume_read_warp(cv, sc.A_w, sc.A_nm, qual, sc.E, len, (int)b.rg[r], b.second[r] ? 1 : 0, lane);
        }
        __syncwarp();
    }
}

// ---- --fixed mode: caller-supplied error flags, then consume_read (kbbq.cc:367-378) ----------------------------------
__global__ void __launch_bounds__(KBBQ_WARPS_PER_BLOCK * 32)
consume_kernel(BatchView b, CountView cv, const uint32_t* __restrict__ err_bits) {
    __shared__ uint32_t E_all[KBBQ_WARPS_PER_BLOCK][KBBQ_MAXW + 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < b.n_reads; r += nwarps) {
        const int len = (int)b.len[r];
        const uint32_t wo = b.word_off[r];
        const int nw = (len + 31) >> 5;
        if (lane < nw) E_all[wib][lane] = err_bits[wo + lane];
        __syncwarp();
        consume_read_warp(cv, b.bases + wo, b.nmask + wo, b.qual + (size_t)wo * 32, E_all[wib], len, (int)b.rg[r],
                          b.second[r] ? 1 : 0, lane);
        __syncwarp();
    }
}

// ---- pass 4: CReadData::recalibrate (readutils.cc:572-595) as a fused table lookup ---------------------------------
__global__ void __launch_bounds__(KBBQ_WARPS_PER_BLOCK * 32)
pass4_apply_kernel(BatchView b, ModelView mv, uint8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < b.n_reads; r += nwarps) {
        const int len = (int)b.len[r];
        const uint32_t wo = b.word_off[r];
        const unsigned long long* w = b.bases + wo;
        const uint32_t* nm = b.nmask + wo;
        const uint8_t* q = b.qual + (size_t)wo * 32;
        const size_t rgrow = (size_t)b.rg[r] * KBBQ_NQ;
        const int sec = b.second[r] ? 1 : 0;
        for (int i = lane; i < len; i += 32) {
            const int qi = q[i];
            int v = qi;
            if (qi >= 6) {  // minqual: lower qualities pass through unchanged
                const size_t row = rgrow + qi;
                v = mv.base[row] + mv.cy[(row * 2 + sec) * (size_t)mv.max_cycle + i];
                if (i > 0) {
                    const bool n1 = (nm[i >> 5] >> (i & 31)) & 1u, n0 = (nm[(i - 1) >> 5] >> ((i - 1) & 31)) & 1u;
                    if (!n1 && !n0) {
                        const int c1 = (int)((w[i >> 5] >> (2 * (i & 31))) & 3), c0 = (int)((w[(i - 1) >> 5] >> (2 * ((i - 1) & 31))) & 3);
                        v += mv.di[row * 16 + ((c0 << 2) | c1)];
                    }
                }
                v = v < 0 ? 0 : (v > KBBQ_MAXQ ? KBBQ_MAXQ : v);
            }
            out[(size_t)wo * 32 + i] = (uint8_t)v;
        }
    }
}

// ---- raw-key filter access (bloom::Bloom::insert / query on a valid Kmer's word) ------------------------------------
__global__ void bloom_insert_keys_kernel(BloomView f, const unsigned long long* __restrict__ keys, unsigned long long n) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bloom_insert_dev(f, keys[i]);
    if (i == 0) atomicAdd(f.inserted, n);
}
__global__ void bloom_query_keys_kernel(BloomView f, const unsigned long long* __restrict__ keys, unsigned long long n,
                                        uint8_t* __restrict__ hits) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hits[i] = bloom_contains_dev(f, keys[i]) ? 1 : 0;
}

// ---- dst |= src over 128-bit words: the local half of the multi-GPU filter merge -------------------------------------
__global__ void or_words_kernel(ulonglong2* __restrict__ dst, const ulonglong2* __restrict__ src, unsigned long long n2) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        ulonglong2 a = dst[i];
        const ulonglong2 s = src[i];
        a.x |= s.x;
        a.y |= s.y;
        dst[i] = a;
    }
}

}  // namespace kbbq
