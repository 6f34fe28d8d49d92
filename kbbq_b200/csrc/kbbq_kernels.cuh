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
#include "kbbq_views.cuh"

namespace kbbq {

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
            const uint64_t b = __shfl_sync(KBBQ_FULL, (uint32_t)blk, src);  // a block number is a 32-bit hash reduced (bloom_block)
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
// TWO 32-probe queries at once: the block loads of all 64 probes are issued before anything is consumed, the patterns
// (L2 hits) follow in two halves.  ncu on the config-3 scan: 14 of 18 cycles per issue are long-scoreboard stalls and
// the DRAM system is 40 % busy -- the probes are latency bound, what helps is more of them in flight per warp.
__device__ __forceinline__ void coop_probe_pair(const unsigned long long* __restrict__ words, const unsigned long long* __restrict__ patterns,
                                                uint64_t blkA, uint32_t patA, uint32_t vmaskA, uint64_t blkB, uint32_t patB, uint32_t vmaskB,
                                                uint32_t& resA, uint32_t& resB) {
    const int lane = threadIdx.x & 31, sub = lane & 3, grp = lane >> 2;
    uint4 x[8];
    uint32_t pn[8];
    bool v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int src = 8 * (i & 3) + grp;
        const uint64_t b = __shfl_sync(KBBQ_FULL, (uint32_t)(i < 4 ? blkA : blkB), src);
        pn[i] = __shfl_sync(KBBQ_FULL, i < 4 ? patA : patB, src);
        v[i] = ((i < 4 ? vmaskA : vmaskB) >> src) & 1u;
        x[i] = make_uint4(0, 0, 0, 0);
        if (v[i]) x[i] = __ldg(reinterpret_cast<const uint4*>(words + b * 8) + sub);
    }
    uint32_t res[2] = {0, 0};
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint4 y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            y[i] = make_uint4(0, 0, 0, 0);
            if (v[4 * half + i]) y[i] = __ldg(reinterpret_cast<const uint4*>(patterns + (uint64_t)pn[4 * half + i] * 8) + sub);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 xx = x[4 * half + i];
            const bool ok = v[4 * half + i] && ((~xx.x & y[i].x) | (~xx.y & y[i].y) | (~xx.z & y[i].z) | (~xx.w & y[i].w)) == 0;
            uint32_t m = __ballot_sync(KBBQ_FULL, ok);
            m &= (m >> 1) & (m >> 2) & (m >> 3) & 0x11111111u;  // all four quarters of a probe agree
            res[half] |= gather_nibble_flags(m) << (8 * i);
        }
    }
    resA = res[0];
    resB = res[1];
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

// the same for lanes that bring their (block, pattern) themselves
__device__ __forceinline__ uint32_t warp_bloom_query_bp(const BloomView& f, bool valid, uint64_t blk, uint32_t pat) {
    const uint32_t vmask = __ballot_sync(KBBQ_FULL, valid);
    if (vmask == 0) return 0;
    return coop_probe_impl<4>(f.words, f.patterns, valid ? blk : 0, valid ? pat : 0, vmask);
}

// pattern_blocked_bf::insert (bloom.hh:255-267) for the lanes with `ins` set; returns how many were inserted.
// The (block, pattern) pairs are first compacted through `stage` (32 x 12 bytes of per-warp shared memory) so that
// sparse inserts (pass 1 samples ~14-35 % of the k-mers) still fill every round.  All 32 lanes must call.
struct InsertStage {
    unsigned long long blk[32];
    uint32_t pat[32];
};
// TEST_FIRST: read the block's quarter before touching it and OR only the words that still miss bits.  At 30-50X
// coverage nine inserts in ten find all their bits already set; a read leaves the line clean where an atomic dirties
// it, which halves the DRAM traffic of a filter that does not fit the L2 (the launch picks the variant by table size;
// for an L2-resident filter the extra load only costs issue slots).  Same bits either way: OR is idempotent, and a
// stale read can only cause a redundant atomic, never a missed one.
// Insert modes of the pass-1 and pass-2 kernels (a template parameter; the launch picks by filter size):
//   INS_DIRECT      atomicOr straight into the table -- filters that sit in the L2
//   INS_TEST_FIRST  the same, reading the block first (above)
//   INS_STAGED      do not touch the table: append (block, pattern) to the staging region of its 16 MiB slice of the
//                   table (below); bloom_apply_staged_kernel then applies region after region,
//                   so that at any moment all SMs update one L2-sized slice of an HBM-sized table instead of scattering
//                   128-byte line fills and dirty write-backs over all of it.  Same bits: OR is order independent.
enum { INS_DIRECT = 0, INS_TEST_FIRST = 1, INS_STAGED = 2 };
#define KBBQ_STAGE_SENTINEL 0xFFFFFFFFFFFFFFFFull  // filler of unused granule tails; applies nothing
#define KBBQ_STAGE_BUCKET_SHIFT 18                 // a bucket = 2^18 blocks = 16 MiB of table
#define KBBQ_STAGE_MAX_BUCKETS 4096

// INS_STAGED without a sort: the staging array is cut into one REGION per bucket (16 MiB slice of the table), and items
// go straight to their bucket's region.  Each thread block owns, per bucket, one granule of the region at a time
// (`state` in shared memory: granule base + 1 in the high bits, items used in the low 16); an item takes the next
// place of its block's granule with one shared-memory atomic, and the thread that finds the granule full reserves the
// next one from the region's cursor in global memory (one global atomic per `granule` items).  A region that is full --
// a skewed input: millions of copies of one k-mer -- sends its bucket's remaining items straight into the table
// (atomicOr, unsorted), so any input stays exact.  Nothing is counted on the host: bloom_apply_staged_kernel reads the
// cursors itself.  (Round 1 staged in arrival order, then radix-sorted by bucket with cub and read the item count
// back in the middle of the pass.)
struct StageSink {
    unsigned long long* items;      // [n_buckets * cap]: (block << 16) | pattern
    unsigned long long* cursor;     // [n_buckets] items reserved in each region (a multiple of the granule)
    unsigned long long cap;         // items per region (a multiple of the granule)
    uint32_t n_buckets, granule;
    int shift;                      // bucket = block >> shift (KBBQ_STAGE_BUCKET_SHIFT; smaller in tests)
};
#define KBBQ_STAGE_USED(s) ((uint32_t)((s) & 0xFFFFull))
// shared memory of a staging block: state[n_buckets] then overflow flags[n_buckets] (u32)
__device__ __forceinline__ void block_stage_open(const StageSink& sink, unsigned long long* state, uint32_t* ovf) {
    for (uint32_t b = threadIdx.x; b < sink.n_buckets; b += blockDim.x) {
        state[b] = (unsigned long long)sink.granule;  // no granule yet, and "full": the first item reserves one
        ovf[b] = 0;
    }
    __syncthreads();
}
// pattern_blocked_bf::insert straight into the table, one thread (the overflow path)
__device__ __noinline__ void bloom_insert_bp_thread(const BloomView& f, uint64_t blk, uint32_t pat) {
    const ulonglong2* P = reinterpret_cast<const ulonglong2*>(f.patterns + (uint64_t)pat * 8);
    unsigned long long* B = f.words + blk * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const ulonglong2 p = __ldg(P + i);
        if (p.x) atomicOr(B + 2 * i, p.x);
        if (p.y) atomicOr(B + 2 * i + 1, p.y);
    }
}
__device__ __forceinline__ void thread_stage_append(const BloomView& f, const StageSink& sink, unsigned long long* state, uint32_t* ovf,
                                                    uint64_t blk, uint32_t pat) {
    const uint32_t b = (uint32_t)(blk >> sink.shift);
    const unsigned long long item = ((unsigned long long)blk << 16) | pat;
    for (;;) {
        if (*(volatile uint32_t*)(ovf + b)) {  // the region is full: straight into the table
            bloom_insert_bp_thread(f, blk, pat);
            return;
        }
        const unsigned long long old = atomicAdd(state + b, 1ull);
        const uint32_t used = KBBQ_STAGE_USED(old);
        if (used < sink.granule) {
            sink.items[(old >> 16) - 1 + used] = item;
            return;
        }
        if (used == sink.granule) {  // this thread reserves the block's next granule of the region
            const unsigned long long at = atomicAdd(sink.cursor + b, (unsigned long long)sink.granule);
            if (at + sink.granule <= sink.cap) {
                const unsigned long long base = (unsigned long long)b * sink.cap + at;
                sink.items[base] = item;
                __threadfence_block();
                atomicExch(state + b, ((base + 1) << 16) | 1ull);
            } else {
                atomicExch(ovf + b, 1u);
                bloom_insert_bp_thread(f, blk, pat);
            }
            return;
        }
        // somebody is reserving: wait for the new granule (or the overflow flag), then try again
        while (KBBQ_STAGE_USED(*(volatile unsigned long long*)(state + b)) > sink.granule && !*(volatile uint32_t*)(ovf + b)) __nanosleep(40);
    }
}
// all 32 lanes call; returns how many lanes had an item
__device__ __forceinline__ int warp_stage_insert_bp(const BloomView& f, const StageSink& sink, unsigned long long* state, uint32_t* ovf, bool ins,
                                                    uint64_t blk, uint32_t pat) {
    const uint32_t imask = __ballot_sync(KBBQ_FULL, ins);
    if (ins) thread_stage_append(f, sink, state, ovf, blk, pat);
    __syncwarp();
    return __popc(imask);
}
__device__ __forceinline__ int warp_stage_insert(const BloomView& f, const StageSink& sink, unsigned long long* state, uint32_t* ovf, bool ins,
                                                 uint64_t key) {
    return warp_stage_insert_bp(f, sink, state, ovf, ins, ins ? bloom_block(f, key) : 0, ins ? bloom_pattern(f, key) : 0);
}
// fill the unused tail of every granule this block still holds
__device__ __forceinline__ void block_stage_close(const StageSink& sink, const unsigned long long* state) {
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    for (uint32_t b = wib; b < sink.n_buckets; b += nwb) {
        const unsigned long long s = state[b];
        if ((s >> 16) == 0) continue;  // never reserved a granule of this region
        const uint32_t used = KBBQ_STAGE_USED(s) < sink.granule ? KBBQ_STAGE_USED(s) : sink.granule;
        for (uint32_t i = used + lane; i < sink.granule; i += 32) sink.items[(s >> 16) - 1 + i] = KBBQ_STAGE_SENTINEL;
    }
}

template <bool TEST_FIRST>
__device__ __forceinline__ int warp_bloom_insert_bp(const BloomView& f, bool ins, uint64_t blk, uint32_t pat, InsertStage& stage) {
    const uint32_t imask = __ballot_sync(KBBQ_FULL, ins);
    if (imask == 0) return 0;
    const int lane = threadIdx.x & 31, sub = lane & 3, grp = lane >> 2;
    const int n = __popc(imask);
    if (ins) {
        const int rank = __popc(imask & ((1u << lane) - 1));
        stage.blk[rank] = blk;
        stage.pat[rank] = pat;
    }
    __syncwarp();
    for (int base = 0; base < n; base += 8) {
        const int idx = base + grp;
        if (idx < n) {
            ulonglong2 p = __ldg(reinterpret_cast<const ulonglong2*>(f.patterns + (uint64_t)stage.pat[idx] * 8) + sub);
            unsigned long long* B = f.words + stage.blk[idx] * 8 + 2 * sub;
            if (TEST_FIRST) {
                const ulonglong2 have = __ldcg(reinterpret_cast<const ulonglong2*>(B));
                p.x &= ~have.x;
                p.y &= ~have.y;
            }
            if (p.x) atomicOr(B, p.x);
            if (p.y) atomicOr(B + 1, p.y);
        }
    }
    __syncwarp();
    return n;
}
template <bool TEST_FIRST>
__device__ __forceinline__ int warp_bloom_insert(const BloomView& f, bool ins, uint64_t key, InsertStage& stage) {
    return warp_bloom_insert_bp<TEST_FIRST>(f, ins, ins ? bloom_block(f, key) : 0, ins ? bloom_pattern(f, key) : 0, stage);
}

// ---- batch set-up: quality masks from the quality bytes when the host did not bring them (one warp per word) ---------
__global__ void qual_masks_kernel(const uint8_t* __restrict__ qual, unsigned long long n_words, uint2* __restrict__ out) {
    const unsigned long long w = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_words) return;
    const uint8_t q = qual[w * 32 + (threadIdx.x & 31)];
    const uint32_t m2 = __ballot_sync(KBBQ_FULL, q <= 2), m6 = __ballot_sync(KBBQ_FULL, q <= 6);
    if ((threadIdx.x & 31) == 0) out[w] = make_uint2(m2, m6);
}

// ---- batch set-up: any quality above KBBQ_MAXQ?  (the count and apply kernels skip such bases; the host is told)
__global__ void __launch_bounds__(256) qual_range_kernel(const uint4* __restrict__ q, unsigned long long n16, int* __restrict__ flag) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    bool bad = false;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = __ldg(q + i);
        const uint32_t lim = 0x01010101u * KBBQ_MAXQ;
        bad = bad || __vmaxu4(v.x, lim) != lim || __vmaxu4(v.y, lim) != lim || __vmaxu4(v.z, lim) != lim || __vmaxu4(v.w, lim) != lim;
    }
    if (__any_sync(KBBQ_FULL, bad) && (threadIdx.x & 31) == 0) *flag = 1;
}

// ---- batch set-up: k-mer slots per read (max(0, len-k+1)) and a trailing 0, scanned in place into slot_off ---------
__global__ void slot_counts_kernel(const uint32_t* __restrict__ len, uint32_t n, int k, unsigned long long* __restrict__ out) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    const uint32_t l = r < n ? len[r] : 0u;
    out[r] = l >= (uint32_t)k ? (unsigned long long)(l - (uint32_t)k + 1) : 0ull;
}

// ---- pass 1a: the sampling stream ---------------------------------------------------------------------------------
// One serial xoshiro256** stream decides every k-mer slot of the file (htsiter.cc:113-129).  The generator is linear
// over GF(2): thread t jumps straight to draw 1024*t with precomputed 256x256 bit matrices of 2^e steps
// (columns of 4 words; jump[b] = 2^(10+b) steps), then produces 1024 consecutive decisions as 32 mask words.
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

// The jump is shared work: the 32 lanes of a warp own consecutive chunks, so they first advance the stream TOGETHER to
// the warp's first chunk (one matrix-vector product per set high bit of the chunk index, 8 state bits per lane, XOR
// reduced by shuffles: ~150 instructions instead of ~2000 per product), then every lane applies its own matrix of
// lane x 1024 steps (`lane_jump`, 32 precomputed matrices) -- one product with uniform control flow, because all lanes
// hold the same state at that point -- instead of up to five divergent ones.
__global__ void __launch_bounds__(128) rng_sample_kernel(uint64_t s0, uint64_t s1, uint64_t s2, uint64_t s3,
                                                         const unsigned long long* __restrict__ jump,
                                                         const unsigned long long* __restrict__ lane_jump,
                                                         unsigned long long n_slots, unsigned long long threshold,
                                                         int take_all, uint32_t* __restrict__ bits) {
    const int lane = threadIdx.x & 31;
    const unsigned long long chunk = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long base = chunk - lane;  // the warp's first chunk
    if (base * 1024ull >= n_slots) return;         // whole warp past the end
    uint64_t s[4] = {s0, s1, s2, s3};
    // together: to chunk `base` (its low five bits are zero)
    for (int b = 5; (base >> b) != 0; ++b) {
        if (!((base >> b) & 1)) continue;
        const unsigned long long* cols = jump + (size_t)b * 1024 + (size_t)(8 * lane) * 4;
        const uint64_t word = lane < 8 ? s[0] : lane < 16 ? s[1] : lane < 24 ? s[2] : s[3];
        uint32_t m = (uint32_t)(word >> (8 * (lane & 7))) & 0xFFu;
        uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        while (m) {
            const int c = __ffs((int)m) - 1;
            m &= m - 1;
            const ulonglong2* col = reinterpret_cast<const ulonglong2*>(cols + (size_t)c * 4);
            const ulonglong2 x = __ldg(col), y = __ldg(col + 1);
            a0 ^= x.x; a1 ^= x.y; a2 ^= y.x; a3 ^= y.y;
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            a0 ^= __shfl_xor_sync(KBBQ_FULL, a0, d);
            a1 ^= __shfl_xor_sync(KBBQ_FULL, a1, d);
            a2 ^= __shfl_xor_sync(KBBQ_FULL, a2, d);
            a3 ^= __shfl_xor_sync(KBBQ_FULL, a3, d);
        }
        s[0] = a0; s[1] = a1; s[2] = a2; s[3] = a3;
    }
    // each lane: lane x 1024 further steps
    if (lane) {
        const unsigned long long* cols = lane_jump + (size_t)lane * 1024;
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
    const unsigned long long first = chunk * 1024ull;
    if (first >= n_slots) return;
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
// LONG: the batch holds a read of more than KBBQ_MAX_READ_LEN bases, whose sampled positions go through the list in pieces
template <int MODE, bool LONG = false>
__global__ void __launch_bounds__(KBBQ_WARPS_PER_BLOCK * 32) pass1_insert_kernel(BatchView b, BloomView f, int k,
                                                                                  const uint32_t* __restrict__ sample_bits,
                                                                                  StageSink sink) {
    extern __shared__ unsigned long long p1_stage_state[];  // INS_STAGED: state[n_buckets], then flags[n_buckets]
    __shared__ InsertStage stage_all[KBBQ_WARPS_PER_BLOCK];
    __shared__ uint16_t pos_all[KBBQ_WARPS_PER_BLOCK][KBBQ_MAX_READ_LEN];  // start positions of a read's sampled k-mers
    InsertStage& stage = stage_all[threadIdx.x >> 5];
    uint16_t* pos = pos_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long inserted = 0;
    unsigned long long* sstate = p1_stage_state;
    uint32_t* sovf = reinterpret_cast<uint32_t*>(p1_stage_state + sink.n_buckets);
    if (MODE == INS_STAGED) block_stage_open(sink, sstate, sovf);
    for (uint32_t r = warp; r < b.n_reads; r += nwarps) {
        const int len = (int)b.len[r];
        const int nk = len - k + 1;
        if (nk <= 0) continue;  // fewer than k bases: no slot, no draw (htsiter.cc:100-105)
        const uint32_t wo = b.word_off[r];
        const CSeqRef s{b.bases + wo, b.nmask + wo, len};
        const unsigned long long so = b.slot_off[r];
        // the sampled positions first (a bit test and a ballot per 32 slots), then keys for those only, 32 at a time:
        // at alpha = 7/30 a 150-base read has 28 of them -- one dense round instead of four quarter-full ones
        // (a read longer than the list goes through in pieces of KBBQ_MAX_READ_LEN slots)
        for (int c0 = 0; c0 < (LONG ? nk : 1); c0 += KBBQ_MAX_READ_LEN) {
            const int c1 = !LONG || nk - c0 < KBBQ_MAX_READ_LEN ? nk : c0 + KBBQ_MAX_READ_LEN;
            int total = 0;
            for (int j0 = c0; j0 < c1; j0 += 32) {
                const int j = j0 + lane;
                bool hit = false;
                if (j < c1) {
                    const unsigned long long slot = so + (unsigned)j;
                    hit = (__ldg(sample_bits + (slot >> 5)) >> (slot & 31)) & 1u;
                }
                const uint32_t m = __ballot_sync(KBBQ_FULL, hit);
                if (hit) pos[total + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(j - c0);
                total += __popc(m);
            }
            __syncwarp();
            for (int t0 = 0; t0 < total; t0 += 32) {
                const int t = t0 + lane;
                uint64_t key = 0;
                const bool ins = t < total && kmer_key(s, c0 + (int)pos[t], k, 0, -1, key);  // an invalid k-mer still consumed its draw
                if (MODE == INS_STAGED) inserted += warp_stage_insert(f, sink, sstate, sovf, ins, key);
                else inserted += warp_bloom_insert<MODE == INS_TEST_FIRST>(f, ins, key, stage);
            }
            __syncwarp();
        }
    }
    if (MODE == INS_STAGED) block_stage_close(sink, sstate);
    if (lane == 0 && inserted) atomicAdd(f.inserted, inserted);
}

// ---- pass 2: infer_read_errors + find_trusted_kmers (readutils.cc:173-193, bloom.cc:28-67, recalibrateutils.cc:15-40)
#ifndef KBBQ_P2_MIN_BLOCKS
#define KBBQ_P2_MIN_BLOCKS 4  // 64 registers, 32 warps/SM: 5.05 ms vs 6.61 unbounded (78 registers, 24 warps)
#endif
// SHARED_HASH: both filters were built with the same seed, so salt_[0] and salt_[1] agree (the reference's filters
// always do: bloom.hh:389, kbbq.cc:265-266) and a k-mer's block hash and pattern number are computed once, for the
// query of `sampled`, and kept in shared memory for the insert into `trusted` (dynamic: 8 bytes per k-mer and warp).
// QMODE (with SHARED_HASH; `sampled` lives in HBM): 0 = one 32-k-mer query at a time; 2 = two at a time, all 64 block loads
// in flight together (coop_probe_pair); 1 = hash every k-mer and prefetch its block first, probe afterwards (prefetch_l2:
// measured, loses, kept as a dead end).
// LONG: a batch whose longest read exceeds KBBQ_MAX_READ_LEN (maxw > KBBQ_MAXW words): the presence and error masks
// then live at the end of the dynamic shared memory, maxw + 1 words each per warp (launched without SHARED_HASH).
template <int MODE, bool SHARED_HASH, int QMODE, bool LONG = false>
__global__ void __launch_bounds__(KBBQ_WARPS_PER_BLOCK * 32, LONG ? 1 : KBBQ_P2_MIN_BLOCKS)
pass2_trusted_kernel(BatchView b, BloomView sampled, BloomView trusted, int k, const int* __restrict__ thresholds,
                     uint32_t* __restrict__ err_out, StageSink sink, int max_kmers, const uint32_t* __restrict__ known_bits,
                     uint32_t* __restrict__ ins_out, int maxw) {
    extern __shared__ uint2 p2_hash[];  // [warps][max_kmers]: .x = hash_ap8(key, salt0), .y = pattern (0xFFFFFFFF: invalid k-mer)
    __shared__ int thr[KBBQ_MAX_K + 1];
    __shared__ uint32_t Pm_all[LONG ? 1 : KBBQ_WARPS_PER_BLOCK][KBBQ_MAXW + 1];
    __shared__ uint32_t Em_all[LONG ? 1 : KBBQ_WARPS_PER_BLOCK][KBBQ_MAXW + 1];
    __shared__ InsertStage stage_all[KBBQ_WARPS_PER_BLOCK];
    InsertStage& stage = stage_all[threadIdx.x >> 5];
    if (threadIdx.x <= k) thr[threadIdx.x] = thresholds[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* Pm = Pm_all[LONG ? 0 : wib];
    uint32_t* Em = Em_all[LONG ? 0 : wib];
    if (LONG) {  // behind the hashes (none: no SHARED_HASH) and the staging state
        uint32_t* masks = reinterpret_cast<uint32_t*>(p2_hash + (SHARED_HASH ? (size_t)KBBQ_WARPS_PER_BLOCK * max_kmers : 0)) +
                          (MODE == INS_STAGED ? 3 * (size_t)sink.n_buckets : 0);
        Pm = masks + (size_t)wib * 2 * (maxw + 1);
        Em = Pm + (maxw + 1);
    }
    uint2* HP = p2_hash + (size_t)wib * (SHARED_HASH ? max_kmers : 0);
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long inserted = 0;
    // INS_STAGED: the staging state follows the hashes in dynamic shared memory: state[n_buckets], then flags[n_buckets]
    unsigned long long* sstate = reinterpret_cast<unsigned long long*>(p2_hash + (SHARED_HASH ? (size_t)KBBQ_WARPS_PER_BLOCK * max_kmers : 0));
    uint32_t* sovf = reinterpret_cast<uint32_t*>(sstate + sink.n_buckets);
    if (MODE == INS_STAGED) block_stage_open(sink, sstate, sovf);
    for (uint32_t r = warp; r < b.n_reads; r += nwarps) {
        const int len = (int)b.len[r];
        const int nk = len - k + 1;
        if (nk <= 0) continue;  // out of contract in the reference (bloom.cc:30 underflows)
        const uint32_t wo = b.word_off[r];
        const CSeqRef s{b.bases + wo, b.nmask + wo, len};
        // kmer_present (bloom.cc:35-47).  A valid k-mer that pass 1 sampled from this very batch was inserted into
        // `sampled` then (known_bits = that pass's sample bits, given only while the filter has not been cleared since):
        // it is present without asking -- a Bloom filter has no false negatives -- and its 64-byte probe is saved.
        const unsigned long long so = b.slot_off[r];
        if (SHARED_HASH && QMODE == 1) {
            for (int j0 = 0; j0 < nk; j0 += 32) {
                const int j = j0 + lane;
                uint64_t key = 0;
                const bool valid = j < nk && kmer_key(s, j, k, 0, -1, key);
                bool known = false;
                if (known_bits && valid) {
                    const unsigned long long slot = so + (unsigned)j;
                    known = (__ldg(known_bits + (slot >> 5)) >> (slot & 31)) & 1u;
                }
                const uint32_t h = hash_ap8(key, sampled.salt0);
                if (valid && !known) prefetch_l2(sampled.words + bloom_block_of_hash(sampled, h) * 8);
                // the pattern number is kept for the insert into `trusted`; bit 31 marks "present without asking"
                if (j < nk) HP[j] = make_uint2(h, valid ? bloom_pattern(sampled, key) | (known ? 0x80000000u : 0u) : KBBQ_HP_INVALID);
            }
            __syncwarp();
            for (int j0 = 0; j0 < nk; j0 += 32) {
                const int j = j0 + lane;
                uint2 hp = j < nk ? HP[j] : make_uint2(0u, KBBQ_HP_INVALID);
                const bool valid = hp.y != KBBQ_HP_INVALID, known = valid && (hp.y & 0x80000000u);
                uint32_t m = warp_bloom_query_bp(sampled, valid && !known, bloom_block_of_hash(sampled, hp.x), hp.y & 0xFFFFu);
                m |= __ballot_sync(KBBQ_FULL, known);
                if (known) HP[j].y = hp.y & 0xFFFFu;  // back to a plain pattern number for the insert loop
                if (lane == 0) Pm[j0 >> 5] = m;
            }
        } else if (SHARED_HASH && QMODE == 2) {
            for (int j0 = 0; j0 < nk; j0 += 64) {
                uint64_t blk[2] = {0, 0};
                uint32_t pat[2] = {0, 0}, vm[2], kn[2], m[2];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int j = j0 + 32 * h2 + lane;
                    uint64_t key = 0;
                    const bool valid = j < nk && kmer_key(s, j, k, 0, -1, key);
                    bool known = false;
                    if (known_bits && valid) {
                        const unsigned long long slot = so + (unsigned)j;
                        known = (__ldg(known_bits + (slot >> 5)) >> (slot & 31)) & 1u;
                    }
                    const uint32_t h = hash_ap8(key, sampled.salt0);
                    pat[h2] = valid ? bloom_pattern(sampled, key) : 0xFFFFFFFFu;
                    if (j < nk) HP[j] = make_uint2(h, pat[h2]);
                    blk[h2] = bloom_block_of_hash(sampled, h);
                    vm[h2] = __ballot_sync(KBBQ_FULL, valid && !known);
                    kn[h2] = __ballot_sync(KBBQ_FULL, known);
                }
                coop_probe_pair(sampled.words, sampled.patterns, blk[0], pat[0], vm[0], blk[1], pat[1], vm[1], m[0], m[1]);
                if (lane == 0) {
                    Pm[j0 >> 5] = m[0] | kn[0];
                    if (j0 + 32 < nk) Pm[(j0 >> 5) + 1] = m[1] | kn[1];
                }
            }
        } else
        for (int j0 = 0; j0 < nk; j0 += 32) {
            const int j = j0 + lane;
            uint64_t key = 0;
            const bool valid = j < nk && kmer_key(s, j, k, 0, -1, key);
            bool known = false;
            if (known_bits && valid) {
                const unsigned long long slot = so + (unsigned)j;
                known = (__ldg(known_bits + (slot >> 5)) >> (slot & 31)) & 1u;
            }
            uint32_t m;
            if (SHARED_HASH) {
                const uint32_t h = hash_ap8(key, sampled.salt0), pat = valid ? bloom_pattern(sampled, key) : 0xFFFFFFFFu;
                if (j < nk) HP[j] = make_uint2(h, pat);
                m = warp_bloom_query_bp(sampled, valid && !known, bloom_block_of_hash(sampled, h), pat);
            } else {
                m = warp_bloom_query<true>(sampled, valid && !known, key);
            }
            m |= __ballot_sync(KBBQ_FULL, known);
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
                const int in = count_window32(Pm, lo, hi);
                e = in <= thr[hi - lo] || ((b.qmask[wo + (i0 >> 5)].x >> lane) & 1u);  // qual[i] <= 2
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
            if (SHARED_HASH) {
                const uint2 hp = j < nk ? HP[j] : make_uint2(0u, 0xFFFFFFFFu);
                const bool ins = j < nk && hp.y != 0xFFFFFFFFu && count_window32(Em, j, j + k) == 0;
                const uint64_t blk = bloom_block_of_hash(trusted, hp.x);
                if (ins_out) {  // which k-mers went into `trusted`: pass 3's scan need not ask about them
                    const uint32_t im = __ballot_sync(KBBQ_FULL, ins);
                    if (lane == 0) ins_out[wo + (j0 >> 5)] = im;
                }
                if (MODE == INS_STAGED) inserted += warp_stage_insert_bp(trusted, sink, sstate, sovf, ins, blk, hp.y);
                else inserted += warp_bloom_insert_bp<MODE == INS_TEST_FIRST>(trusted, ins, blk, hp.y, stage);
            } else {
                uint64_t key = 0;
                const bool ins = j < nk && count_window32(Em, j, j + k) == 0 && kmer_key(s, j, k, 0, -1, key);
                if (ins_out) {
                    const uint32_t im = __ballot_sync(KBBQ_FULL, ins);
                    if (lane == 0) ins_out[wo + (j0 >> 5)] = im;
                }
                if (MODE == INS_STAGED) inserted += warp_stage_insert(trusted, sink, sstate, sovf, ins, key);
                else inserted += warp_bloom_insert<MODE == INS_TEST_FIRST>(trusted, ins, key, stage);
            }
        }
        __syncwarp();
    }
    if (MODE == INS_STAGED) block_stage_close(sink, sstate);
    if (lane == 0 && inserted) atomicAdd(trusted.inserted, inserted);
}

// ---- the flat traversal shared by covariate counting and pass 4 -----------------------------------------------------
// Both kernels are streaming passes over every base with a little per-read context.  One THREAD takes one half word --
// 16 consecutive bases of one read: a 16-byte load of qualities, the word's 2-bit bases and N flags, the read's group
// and pair flag -- so a warp instruction covers 512 bases whatever the read length (one warp per read left 13 of 32
// lanes idle on 150-base reads and a third of the issue slots in shuffles).  The read a word belongs to is word / wpr
// for batches whose reads all occupy the same number of words (the launch checks), else looked up in word_read.
struct FlatMap {
    const uint32_t* word_read;  // [n_words] read index of every word, or nullptr: uniform batch
    uint32_t wpr;               // words per read of a uniform batch
};
struct HalfWord {
    uint32_t r, wo, gw;
    int i0, len, sec, o;        // first base of the half word in the read, read length, pair flag, offset in the word (0/16)
    uint32_t rgrow;             // rg * 94
    unsigned long long codes;   // 17 bases: the one before the half word, then its 16 (2 bits each from bit 0)
    uint32_t nflags;            // 17 N flags likewise; bit 0 also set when there is no base before (i0 == 0)
    uint4 q;                    // 16 quality bytes
};
__device__ __forceinline__ bool load_half_word(const BatchView& b, const FlatMap& fm, unsigned long long h, unsigned long long n_half,
                                               HalfWord& hw) {
    if (h >= n_half) return false;
    hw.gw = (uint32_t)(h >> 1);
    hw.o = (int)(h & 1) * 16;
    hw.r = fm.word_read ? __ldg(fm.word_read + hw.gw) : hw.gw / fm.wpr;
    hw.wo = b.word_off[hw.r];
    hw.len = (int)b.len[hw.r];
    hw.i0 = (int)(hw.gw - hw.wo) * 32 + hw.o;
    if (hw.i0 >= hw.len) return false;
    hw.rgrow = (uint32_t)b.rg[hw.r] * KBBQ_NQ;
    hw.sec = b.second[hw.r] ? 1 : 0;
    hw.q = __ldg(reinterpret_cast<const uint4*>(b.qual + (size_t)hw.gw * 32 + hw.o));
    const unsigned long long bw = b.bases[hw.gw];
    const uint32_t nw = b.nmask[hw.gw];
    if (hw.o) {
        hw.codes = (bw >> 30) & 0x3FFFFFFFFull;
        hw.nflags = (nw >> 15) & 0x1FFFFu;
    } else {
        const unsigned long long pw = hw.i0 ? b.bases[hw.gw - 1] : 0ull;
        const uint32_t pn = hw.i0 ? b.nmask[hw.gw - 1] : 1u << 31;  // no base before the first: no dinucleotide
        hw.codes = ((bw << 2) | (pw >> 62)) & 0x3FFFFFFFFull;
        hw.nflags = ((nw << 1) | (pn >> 31)) & 0x1FFFFu;
    }
    return true;
}
__device__ __forceinline__ int half_word_qual(const HalfWord& hw, int j) {
    const uint32_t w = j < 4 ? hw.q.x : j < 8 ? hw.q.y : j < 12 ? hw.q.z : hw.q.w;
    return (int)((w >> (8 * (j & 3))) & 255u);
}
__global__ void __launch_bounds__(256) word_read_kernel(const uint32_t* __restrict__ word_off, uint32_t n_reads, uint32_t* __restrict__ word_read) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    for (uint32_t w = word_off[r]; w < word_off[r + 1]; ++w) word_read[w] = r;
}

// ---- covariate counting: CCovariateData::consume_read (covariateutils.cc:30-42,65-76,102-116,147-164,193-202) -------
// Every base adds to its (read group, quality) row's cycle cell and dinucleotide cell; the row's {errors,total} pair
// and the read group's are sums of the cycle cells (skips are never set, covariateutils.cc:194-197: all three tables
// count every base) and are formed when a block flushes.  The tables are small and hot -- binned qualities put half of
// all bases on one row -- so as global RED operations they serialise per address in the L2 (tools/probe_lab.cu: 3.6 G/s
// on 20 hot cells).  Each block keeps an associative set of S rows in shared memory and counts there with shared-memory
// atomics, which the SM retires at ~16 per clock even on a handful of hot addresses (probe_lab: 2.3 T/s chip-wide,
// the same as conflict-free ones): two per base.  (Round 1 elected one leader per distinct key with __match_any_sync to
// save atomics: the match alone costs three times the atomics it saved.)  A row takes the next free slot the first
// time any thread of the block meets it; a row that finds none -- or belongs to a read group beyond the direct table --
// falls through to global atomics, so any input stays exact.  The launch bounds the words per block so that no
// 32-bit cell can overflow.
#define KBBQ_HIST_THREADS 512
#define KBBQ_HIST_EMPTY 0xFFFFFFFFu
#define KBBQ_HIST_DIRECT_RG 16
#define KBBQ_HIST_UNASSIGNED 0xFFFFFFFFu
#define KBBQ_HIST_CLAIMING 0xFFFFFFFEu
#define KBBQ_HIST_NOSLOT 0xFFFFFFFDu
// row layout: [0,2C) cycle totals (second*C + i), [2C,4C) cycle errors, [4C,4C+16) dinucleotide totals, then 16 errors
__device__ __forceinline__ int hist_slot(uint32_t* slot_of, uint32_t* tags, uint32_t* n_used, int S, uint32_t key) {
    if (key >= KBBQ_HIST_DIRECT_RG * KBBQ_NQ) return -1;
    volatile uint32_t* cell = slot_of + key;
    uint32_t v = *cell;
    if (v < (uint32_t)S) return (int)v;
    if (v == KBBQ_HIST_UNASSIGNED && atomicCAS(slot_of + key, KBBQ_HIST_UNASSIGNED, KBBQ_HIST_CLAIMING) == KBBQ_HIST_UNASSIGNED) {
        const uint32_t n = atomicAdd(n_used, 1u);
        v = n < (uint32_t)S ? n : KBBQ_HIST_NOSLOT;
        if (n < (uint32_t)S) tags[n] = key;
        __threadfence_block();
        *cell = v;
    } else {
        // another thread is claiming the row (independent thread scheduling lets it finish even from this warp)
        while ((v = *cell) == KBBQ_HIST_CLAIMING || v == KBBQ_HIST_UNASSIGNED) __nanosleep(20);
    }
    return v < (uint32_t)S ? (int)v : -1;
}

// err_bits: read.errors as bit masks (kbbq_batch word layout).  mod: per read, 4*index+base of the one base that
// get_errors left modified in seq (ref quirk, readutils.cc:263-265 vs 568), or 0xFFFFFFFF; nullptr = none.
// Counts half words [first_half, end_half); reads at or past n_first are not counted (--fixed mode, kbbq.cc:370).
__global__ void __launch_bounds__(KBBQ_HIST_THREADS, 2)
covariate_count_kernel(BatchView b, FlatMap fm, CountView cv, const uint32_t* __restrict__ err_bits, const uint32_t* __restrict__ mod,
                       unsigned long long first_half, unsigned long long end_half, uint32_t n_first, int S) {
    extern __shared__ uint32_t hist_smem[];
    const int C = cv.max_cycle, row_cells = 4 * C + 32;
    uint32_t* slot_of = hist_smem;                                    // [KBBQ_HIST_DIRECT_RG * 94]
    uint32_t* tags = slot_of + KBBQ_HIST_DIRECT_RG * KBBQ_NQ;         // [S] key of each slot in use (for the flush)
    uint32_t* n_used = tags + S;                                      // [1] (+1 pad)
    uint32_t* rows = n_used + 2;
    for (int i = threadIdx.x; i < KBBQ_HIST_DIRECT_RG * KBBQ_NQ; i += blockDim.x) slot_of[i] = KBBQ_HIST_UNASSIGNED;
    for (int i = threadIdx.x; i < S; i += blockDim.x) tags[i] = KBBQ_HIST_EMPTY;
    for (int i = threadIdx.x; i < 2 + S * row_cells; i += blockDim.x) n_used[i] = 0;
    __syncthreads();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long h = first_half + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; h < end_half; h += stride) {
        HalfWord hw;
        if (!load_half_word(b, fm, h, end_half, hw) || hw.r >= n_first) continue;
        const uint32_t e16 = (err_bits[hw.gw] >> hw.o) & 0xFFFFu;
        if (mod) {  // the base get_errors left modified takes part in the dinucleotide keys
            const uint32_t md = mod[hw.r];
            const int j = (int)(md >> 2) - (hw.i0 - 1);  // position among the 17 bases this thread looks at
            if (md != 0xFFFFFFFFu && j >= 0 && j <= 16) {
                hw.codes = (hw.codes & ~(3ull << (2 * j))) | ((unsigned long long)(md & 3u) << (2 * j));
                hw.nflags &= ~(1u << j);
            }
        }
        const int n = hw.len - hw.i0 < 16 ? hw.len - hw.i0 : 16;
        const uint32_t qw[4] = {hw.q.x, hw.q.y, hw.q.z, hw.q.w};
        const int cyc0 = hw.sec * C + hw.i0;  // the half word's first cycle cell within a row
        // fully unrolled: the quality byte of step j is a fixed byte of a fixed register, and no branch depends on what
        // the neighbouring lanes hold except the (rare) error and the (rare) first meeting with a row
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (j >= n) break;
            const uint32_t qi = (qw[j >> 2] >> (8 * (j & 3))) & 255u;
            if (qi > KBBQ_MAXQ) continue;  // a quality the tables cannot index is skipped (and reported: qual_range_kernel)
            const uint32_t key = hw.rgrow + qi;
            uint32_t v = key < KBBQ_HIST_DIRECT_RG * KBBQ_NQ ? slot_of[key] : KBBQ_HIST_NOSLOT;
            if (v >= (uint32_t)S && v != KBBQ_HIST_NOSLOT) {  // first meeting with the row (or somebody is claiming it)
                const int slot = hist_slot(slot_of, tags, n_used, S, key);
                v = slot >= 0 ? (uint32_t)slot : KBBQ_HIST_NOSLOT;
            }
            const bool er = (e16 >> j) & 1u;
            // CDinucCovariate: i >= 1, both bases ACGT, qual >= 6
            const bool di = qi >= 6 && ((hw.nflags >> j) & 3u) == 0;
            const uint32_t x = (uint32_t)(hw.codes >> (2 * j)) & 15u;  // previous base in the low pair
            const uint32_t d = ((x & 3u) << 2) | (x >> 2);
            if (v != KBBQ_HIST_NOSLOT) {
                uint32_t* row = rows + (size_t)v * row_cells;
                atomicAdd(row + cyc0 + j, 1u);
                if (di) atomicAdd(row + 4 * C + d, 1u);
                if (er) {
                    atomicAdd(row + 2 * C + cyc0 + j, 1u);
                    if (di) atomicAdd(row + 4 * C + 16 + d, 1u);
                }
            } else {  // no private row: straight to the tables, all four of them
                unsigned long long* cc = cv.cyc + ((size_t)key * 2 * (size_t)C + cyc0 + j) * 2;
                atomicAdd(cc + 1, 1ull);
                atomicAdd(cv.qc + (size_t)key * 2 + 1, 1ull);
                atomicAdd(cv.rgc + 2 * (hw.rgrow / KBBQ_NQ) + 1, 1ull);
                if (er) {
                    atomicAdd(cc, 1ull);
                    atomicAdd(cv.qc + (size_t)key * 2, 1ull);
                    atomicAdd(cv.rgc + 2 * (hw.rgrow / KBBQ_NQ), 1ull);
                }
                if (di) {
                    unsigned long long* dd = cv.dic + ((size_t)key * 16 + d) * 2;
                    atomicAdd(dd + 1, 1ull);
                    if (er) atomicAdd(dd, 1ull);
                }
            }
        }
    }
    __syncthreads();
    // flush what this block counted: one warp per row; the row's q pair and its read group's pair are sums of its cycle cells
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    for (int s = wib; s < S; s += nwb) {
        const uint32_t key = tags[s];
        if (key == KBBQ_HIST_EMPTY) continue;
        const uint32_t* row = rows + (size_t)s * row_cells;
        unsigned long long tot = 0, err = 0;
        for (int j = lane; j < row_cells; j += 32) {
            const uint32_t v = row[j];
            if (!v) continue;
            unsigned long long* dst;
            if (j < 2 * C) { dst = cv.cyc + ((size_t)key * 2 * C + j) * 2 + 1; tot += v; }
            else if (j < 4 * C) { dst = cv.cyc + ((size_t)key * 2 * C + (j - 2 * C)) * 2; err += v; }
            else if (j < 4 * C + 16) dst = cv.dic + ((size_t)key * 16 + (j - 4 * C)) * 2 + 1;
            else dst = cv.dic + ((size_t)key * 16 + (j - 4 * C - 16)) * 2;
            atomicAdd(dst, (unsigned long long)v);
        }
#pragma unroll
        for (int dlt = 16; dlt; dlt >>= 1) {
            tot += __shfl_xor_sync(KBBQ_FULL, tot, dlt);
            err += __shfl_xor_sync(KBBQ_FULL, err, dlt);
        }
        if (lane == 0) {
            if (tot) { atomicAdd(cv.qc + (size_t)key * 2 + 1, tot); atomicAdd(cv.rgc + 2 * (key / KBBQ_NQ) + 1, tot); }
            if (err) { atomicAdd(cv.qc + (size_t)key * 2, err); atomicAdd(cv.rgc + 2 * (key / KBBQ_NQ), err); }
        }
    }
}

// ---- pass 3: get_errors + consume_read (recalibrateutils.cc:42-89) ------------------------------------------------
// Pass 3 runs as two kernels.
//
// pass3_scan_kernel -- the mandatory part of get_errors (bloom::find_longest_trusted_seq's walk, bloom.cc:96-128): the
// trusted flag of EVERY k-mer of every read, with the warp-cooperative probe pass 2 uses (four lanes per 64-byte block,
// all loads of a 32-k-mer query in flight together).  One lane per probe -- what the error caller uses for its sparse,
// dependent questions -- touches 32 lines per warp load and reaches half the probe rate of the cooperative form once
// the filter lives in HBM (tools/probe_ceiling.cu: 21.8 vs 42.6 G probes/s from a 1 GB table), and these 119 probes per
// read are 2/3 of all of pass 3's.  A read whose k-mers are all trusted (readutils.cc:263-265) is finished here: its
// flags are zero.  Every other read gets its k-mer flags written (tflags, one word per 32 k-mers in the batch's word
// layout) and its index appended to the repair list.
//
// pass3_repair_kernel -- the rest of get_errors for the listed reads, one warp per read pulled from the list: anchors,
// repairs, overcorrection check, the one-level recursion; it starts from the scan's flags instead of re-scanning.
#ifndef KBBQ_P3_MIN_BLOCKS
#define KBBQ_P3_MIN_BLOCKS 4
#endif
// PREFETCH (filters that live in HBM): first hash every k-mer of the read, prefetch its block and park (hash, pattern)
// in shared memory (8 bytes per k-mer and warp, dynamic); then probe round by round from there.
// DEEP (filters that live in HBM): two 32-k-mer rounds per step, all 64 block loads in flight together
// (coop_probe_pair).
#ifndef KBBQ_P3D_MIN_BLOCKS
#define KBBQ_P3D_MIN_BLOCKS 4
#endif
template <bool PREFETCH, bool DEEP>
__global__ void __launch_bounds__(KBBQ_WARPS_PER_BLOCK * 32, DEEP ? KBBQ_P3D_MIN_BLOCKS : KBBQ_P3_MIN_BLOCKS)
pass3_scan_kernel(BatchView b, BloomView trusted, int k, uint32_t* __restrict__ tflags, uint32_t* __restrict__ err_out,
                  uint32_t* __restrict__ mod_out, uint32_t* __restrict__ repair_list, unsigned int* __restrict__ repair_count,
                  const uint32_t* __restrict__ known_bits, int max_kmers) {
    extern __shared__ uint2 scan_hash[];  // PREFETCH: [warps][max_kmers]
    const int lane = threadIdx.x & 31;
    uint2* HP = scan_hash + (size_t)(threadIdx.x >> 5) * (PREFETCH ? max_kmers : 0);
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < b.n_reads; r += nwarps) {
        const int len = (int)b.len[r];
        const int nk = len - k + 1;
        const uint32_t wo = b.word_off[r];
        const int nw = (len + 31) >> 5;
        bool all_trusted = true;
        if (nk > 0) {
            const CSeqRef s{b.bases + wo, b.nmask + wo, len};
            if (PREFETCH) {
                for (int j0 = 0; j0 < nk; j0 += 32) {
                    const int j = j0 + lane;
                    uint64_t key = 0;
                    const bool valid = j < nk && kmer_key(s, j, k, 0, -1, key);
                    const uint32_t known = known_bits ? __ldg(known_bits + wo + (j0 >> 5)) : 0u;
                    const bool ask = valid && !((known >> lane) & 1u);
                    const uint32_t h = hash_ap8(key, trusted.salt0);
                    if (ask) prefetch_l2(trusted.words + bloom_block_of_hash(trusted, h) * 8);
                    if (j < nk) HP[j] = make_uint2(h, ask ? bloom_pattern(trusted, key) : valid ? KBBQ_HP_KNOWN : KBBQ_HP_INVALID);
                }
                __syncwarp();
            }
            // k-mers pass 2 inserted from this very batch (known_bits, per k-mer in the word layout; given only while
            // `trusted` has not been cleared since) are present without asking: a third of all k-mers at 30X
            if (DEEP && !PREFETCH) {
                for (int j0 = 0; j0 < nk; j0 += 64) {
                    uint64_t blk[2] = {0, 0};
                    uint32_t pat[2] = {0, 0}, vm[2], kn[2], m[2];
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int j = j0 + 32 * h2 + lane;
                        uint64_t key = 0;
                        const bool valid = j < nk && kmer_key(s, j, k, 0, -1, key);
                        kn[h2] = known_bits && j0 + 32 * h2 < nk ? __ldg(known_bits + wo + ((j0 + 32 * h2) >> 5)) : 0u;
                        const bool ask = valid && !((kn[h2] >> lane) & 1u);
                        vm[h2] = __ballot_sync(KBBQ_FULL, ask);
                        if (ask) {
                            blk[h2] = bloom_block(trusted, key);
                            pat[h2] = bloom_pattern(trusted, key);
                        }
                    }
                    coop_probe_pair(trusted.words, trusted.patterns, blk[0], pat[0], vm[0], blk[1], pat[1], vm[1], m[0], m[1]);
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int jj = j0 + 32 * h2;
                        if (jj >= nk) break;
                        const uint32_t mm = m[h2] | kn[h2];
                        const int cnt = nk - jj < 32 ? nk - jj : 32;
                        all_trusted = all_trusted && mm == (cnt < 32 ? (1u << cnt) - 1u : ~0u);
                        if (lane == 0) tflags[wo + (jj >> 5)] = mm;
                    }
                }
            } else
            for (int j0 = 0; j0 < nk; j0 += 32) {
                const int j = j0 + lane;
                uint32_t m;
                if (PREFETCH) {
                    const uint2 hp = j < nk ? HP[j] : make_uint2(0u, KBBQ_HP_INVALID);
                    m = warp_bloom_query_bp(trusted, hp.y < KBBQ_HP_KNOWN, bloom_block_of_hash(trusted, hp.x), hp.y) |
                        __ballot_sync(KBBQ_FULL, hp.y == KBBQ_HP_KNOWN);
                } else {
                    uint64_t key = 0;
                    const bool valid = j < nk && kmer_key(s, j, k, 0, -1, key);
                    const uint32_t known = known_bits ? __ldg(known_bits + wo + (j0 >> 5)) : 0u;
                    m = warp_bloom_query<true>(trusted, valid && !((known >> lane) & 1u), key) | known;
                }
                const int cnt = nk - j0 < 32 ? nk - j0 : 32;
                all_trusted = all_trusted && m == (cnt < 32 ? (1u << cnt) - 1u : ~0u);
                if (lane == 0) tflags[wo + (j0 >> 5)] = m;
            }
            __syncwarp();
        }
        // fewer than k bases (out of contract in the reference, bloom.cc:30): untouched, like a clean read
        if (all_trusted || nk <= 0) {
            for (int w = lane; w < nw; w += 32) err_out[wo + w] = 0u;
            if (lane == 0) mod_out[r] = 0xFFFFFFFFu;
        } else if (lane == 0) {
            repair_list[atomicAdd(repair_count, 1u)] = r;
        }
    }
}

// one copy of hash + 4 loads + compare for every question the error caller asks (it asks from five places)
__device__ __noinline__ uint32_t probe_ballot(const BloomView* f, bool valid, uint64_t key) {
    return __ballot_sync(KBBQ_FULL, valid && bloom_contains_dev(*f, key));
}
// Wide questions (a walk over up to 32 consecutive k-mers) may use the four-lanes-per-probe form instead: half the L1
// wavefronts and DRAM sectors in flight per probe once the filter lives in HBM; slower while it sits in the L2 (round 1).
// KBBQ_P3R_COOP: 0 = one lane per probe always, N > 0 = cooperative for questions of more than N k-mers.
#ifndef KBBQ_P3R_COOP
#define KBBQ_P3R_COOP 0
#endif
__device__ __noinline__ uint32_t probe_coop_ballot(const BloomView* f, bool valid, uint64_t key) {
    const uint32_t vmask = __ballot_sync(KBBQ_FULL, valid);
    if (vmask == 0) return 0;
#ifndef KBBQ_P3R_COOP_ROUNDS
#define KBBQ_P3R_COOP_ROUNDS 2
#endif
    return coop_probe_impl<KBBQ_P3R_COOP_ROUNDS>(f->words, f->patterns, valid ? bloom_block(*f, key) : 0, valid ? bloom_pattern(*f, key) : 0, vmask);
}
struct DevQuery {
    const BloomView* f;  // in shared memory: a kernel parameter has no address an out-of-line function could take
    template <class F>
    KBBQ_HD uint32_t query(int count, F fn) {
#if defined(__CUDA_ARCH__)
        const int lane = threadIdx.x & 31;
        uint64_t key = 0;
        const bool valid = lane < count && fn(lane, key);
        if (KBBQ_P3R_COOP > 0 && count > KBBQ_P3R_COOP) return probe_coop_ballot(f, valid, key);
        return probe_ballot(f, valid, key);
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
    // fn(i) for i in [0, n), one lane each: per-word loops of the error caller (a read is a handful of 32-base words)
    // become one step of the warp instead of a chain of dependent shared-memory round trips
    template <class F>
    KBBQ_HD void par(int n, F fn) {
#if defined(__CUDA_ARCH__)
        for (int i = (int)(threadIdx.x & 31); i < n; i += 32) fn(i);
        __syncwarp();
#else
        (void)n; (void)fn;
#endif
    }
};

// Writes read.errors (err_out, kbbq_batch word layout) and, per read, the base get_errors left modified (mod_out, see
// covariate_count_kernel) for the reads of the repair list; the counting itself is covariate_count_kernel's.
#ifndef KBBQ_P3R_MIN_BLOCKS
#define KBBQ_P3R_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(KBBQ_WARPS_PER_BLOCK * 32, KBBQ_P3R_MIN_BLOCKS)
pass3_repair_kernel(BatchView b, BloomView trusted, int k, const uint32_t* __restrict__ tflags, uint32_t* __restrict__ err_out,
                    uint32_t* __restrict__ mod_out, const uint32_t* __restrict__ repair_list,
                    const unsigned int* __restrict__ repair_count, unsigned int* __restrict__ work_counter) {
    __shared__ ErrScratch scratch[KBBQ_WARPS_PER_BLOCK];
    __shared__ BloomView filter;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    ErrScratch& sc = scratch[wib];
    if (threadIdx.x == 0) filter = trusted;
    __syncthreads();
    DevQuery q{&filter};
    const uint32_t n_list = *repair_count;
    for (;;) {
        // reads differ a lot in cost (a single repair to ~1000 probes): pull, don't stride
        uint32_t idx = 0;
        if (lane == 0) idx = atomicAdd(work_counter, 1u);
        idx = __shfl_sync(KBBQ_FULL, idx, 0);
        if (idx >= n_list) break;
        const uint32_t r = repair_list[idx];
        const int len = (int)b.len[r];
        const uint32_t wo = b.word_off[r];
        const int nw = (len + 31) >> 5;
        const uint8_t* qual = b.qual + (size_t)wo * 32;
        if (lane <= nw) {
            sc.P_w[lane] = lane < nw ? b.bases[wo + lane] : 0ull;
            sc.P_nm[lane] = lane < nw ? b.nmask[wo + lane] : 0u;
            sc.E[lane] = 0;
            sc.T_orig[lane] = 32 * lane < len - k + 1 ? tflags[wo + lane] : 0u;  // the scan wrote one word per 32 k-mers
            // low-quality flags for the overcorrection half weight (bits past the read's end cleared)
            const int left = len - 32 * lane;
            sc.QL[lane] = lane < nw ? b.qmask[wo + lane].y & (left >= 32 ? ~0u : ((1u << left) - 1u)) : 0u;
        }
        __syncwarp();
        const bool has_n = __any_sync(KBBQ_FULL, lane < nw && sc.P_nm[lane] != 0u);
        ErrCaller<DevQuery> ec(q, k, qual, sc, has_n);
        const ErrResult res = ec.run(len, true);
        __syncwarp();
        if (lane < nw) err_out[wo + lane] = sc.E[lane];
        if (lane == 0) mod_out[r] = res.mod_idx >= 0 ? (uint32_t)(4 * res.mod_idx + res.mod_base) : 0xFFFFFFFFu;
        __syncwarp();
    }
}

// ---- pass 4: CReadData::recalibrate (readutils.cc:572-595) as a fused table lookup ---------------------------------
// The flat traversal (above): one thread per 16 bases, one 16-byte load and one 16-byte store of qualities each, two
// table lookups per base.  SMEM: the lookup tables live in shared memory (the launch checks that they fit): cy as it is,
// and `dib` = base + di folded into one int16 table of 17 entries per row (entry 16 = no dinucleotide term); larger
// models read the three global tables through the L1.
struct P4Tables {
    const int8_t* cy;
    const int16_t* dib;
};
// bytes of shared memory the SMEM variant needs
__host__ __device__ inline size_t p4_smem_bytes(int n_rg, int max_cycle) {
    const size_t cy = ((size_t)n_rg * KBBQ_NQ * 2 * max_cycle + 15) & ~(size_t)15;
    return cy + (size_t)n_rg * KBBQ_NQ * 17 * 2;
}
template <bool SMEM>
__global__ void __launch_bounds__(256)
pass4_apply_kernel(BatchView b, FlatMap fm, ModelView mv, int n_rg, uint8_t* __restrict__ out, unsigned long long first_half,
                   unsigned long long end_half) {
    extern __shared__ __align__(16) unsigned char p4_smem[];
    P4Tables st{nullptr, nullptr};
    if (SMEM) {
        const size_t n_cy = (size_t)n_rg * KBBQ_NQ * 2 * mv.max_cycle, cy_bytes = (n_cy + 15) & ~(size_t)15;
        int8_t* cy = reinterpret_cast<int8_t*>(p4_smem);
        int16_t* dib = reinterpret_cast<int16_t*>(p4_smem + cy_bytes);
        for (size_t i = threadIdx.x; i < n_cy; i += blockDim.x) cy[i] = mv.cy[i];
        for (int i = threadIdx.x; i < n_rg * KBBQ_NQ * 17; i += blockDim.x) {
            const int row = i / 17, dd = i - row * 17;
            dib[i] = (int16_t)(mv.base[row] + (dd < 16 ? mv.di[row * 16 + dd] : 0));
        }
        __syncthreads();
        st.cy = cy;
        st.dib = dib;
    }
    const int C = mv.max_cycle;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long h = first_half + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; h < end_half; h += stride) {
        HalfWord hw;
        if (!load_half_word(b, fm, h, end_half, hw)) continue;  // past the read's end: the output stays zero like the input
        const uint32_t qw[4] = {hw.q.x, hw.q.y, hw.q.z, hw.q.w};
        uint32_t res[4];
        const int n = hw.len - hw.i0 < 16 ? hw.len - hw.i0 : 16;
        const uint32_t cy0 = (hw.rgrow * 2u + (uint32_t)hw.sec) * (uint32_t)C + (uint32_t)hw.i0;  // cy index of quality 0 at the first base
        const uint32_t di0 = hw.rgrow * 17u, di0g = hw.rgrow * 16u;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t packed = 0;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = 4 * w + jj;
                const uint32_t qi = (qw[w] >> (8 * jj)) & 255u;
                uint32_t nv = qi;
                if (j < n && qi >= 6 && qi <= KBBQ_MAXQ) {  // minqual: lower qualities pass through unchanged (so does anything the tables cannot index)
                    const bool has_di = ((hw.nflags >> j) & 3u) == 0;  // this base and the one before it are ACGT (and i > 0)
                    const uint32_t x = (uint32_t)(hw.codes >> (2 * j)) & 15u;  // previous base in the low pair
                    const uint32_t dn = ((x & 3u) << 2) | (x >> 2);
                    int v;
                    if (SMEM) {
                        v = st.cy[cy0 + qi * 2u * (uint32_t)C + j] + st.dib[di0 + qi * 17u + (has_di ? dn : 16u)];
                    } else {
                        v = mv.base[hw.rgrow + qi] + mv.cy[cy0 + qi * 2u * (uint32_t)C + j];
                        if (has_di) v += mv.di[di0g + qi * 16u + dn];
                    }
                    nv = (uint32_t)(v < 0 ? 0 : (v > KBBQ_MAXQ ? KBBQ_MAXQ : v));
                }
                packed |= nv << (8 * jj);
            }
            res[w] = packed;
        }
        *reinterpret_cast<uint4*>(out + (size_t)hw.gw * 32 + hw.o) = make_uint4(res[0], res[1], res[2], res[3]);
    }
}

// ---- applying the staged items (INS_STAGED): four lanes per item, region after region --------------------------------
// Every block first forms the running totals of the regions' fill counts (shared memory, n_buckets + 1 entries); the
// grid then strides over the DENSE sequence of filled positions -- region 0's items, then region 1's, ... -- so that at
// any moment all SMs update the same 16 MiB slice of the table, whatever share of each region is filled.
__global__ void __launch_bounds__(256) bloom_apply_staged_kernel(BloomView f, StageSink sink) {
    extern __shared__ unsigned long long fill_end[];  // [n_buckets + 1]: fill_end[b] = items in regions 0 .. b-1
    __shared__ unsigned long long warp_tot[8];
    {   // exclusive scan of the (clamped) cursors: each thread owns a contiguous run of regions
        const uint32_t per = (sink.n_buckets + blockDim.x - 1) / blockDim.x;
        const uint32_t lo = threadIdx.x * per, hi = lo + per < sink.n_buckets ? lo + per : sink.n_buckets;
        unsigned long long mine = 0;
        for (uint32_t bkt = lo; bkt < hi; ++bkt) {
            const unsigned long long c = __ldg(sink.cursor + bkt);
            mine += c < sink.cap ? c : sink.cap;  // reservations past the end of a region were refused
        }
        unsigned long long incl = mine;
        const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long y = __shfl_up_sync(KBBQ_FULL, incl, d);
            if (lane >= d) incl += y;
        }
        if (lane == 31) warp_tot[wib] = incl;
        __syncthreads();
        unsigned long long before = incl - mine;
        for (int w = 0; w < wib; ++w) before += warp_tot[w];
        for (uint32_t bkt = lo; bkt < hi; ++bkt) {
            fill_end[bkt] = before;
            const unsigned long long c = __ldg(sink.cursor + bkt);
            before += c < sink.cap ? c : sink.cap;
        }
        if (hi == sink.n_buckets && lo < hi) fill_end[sink.n_buckets] = before;
        if (sink.n_buckets == 0 && threadIdx.x == 0) fill_end[0] = 0;
        __syncthreads();
    }
    const unsigned long long total = fill_end[sink.n_buckets];
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = ((unsigned long long)gridDim.x * blockDim.x) >> 2;
    const int sub = (int)(threadIdx.x & 3);
    uint32_t bkt = 0;
    for (unsigned long long i = t >> 2; i < total; i += stride) {
        while (i >= fill_end[bkt + 1]) ++bkt;  // positions only grow: the region index is found by walking on
        const unsigned long long it = __ldg(sink.items + (unsigned long long)bkt * sink.cap + (i - fill_end[bkt]));
        if (it == KBBQ_STAGE_SENTINEL) continue;
        ulonglong2 p = __ldg(reinterpret_cast<const ulonglong2*>(f.patterns + (it & 0xFFFFull) * 8) + sub);
        unsigned long long* B = f.words + (it >> 16) * 8 + 2 * sub;
        const ulonglong2 have = __ldcg(reinterpret_cast<const ulonglong2*>(B));
        p.x &= ~have.x;
        p.y &= ~have.y;
        if (p.x) atomicOr(B, p.x);
        if (p.y) atomicOr(B + 1, p.y);
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

// ---- --fixed mode (kbbq.cc:367-378): a base is an error when it differs from the same base of the pre-corrected read ----
// gather the even bits of a 64-bit word into 32 bits
__device__ __forceinline__ uint32_t even_bits(unsigned long long x) {
    x &= 0x5555555555555555ull;
    x = (x | (x >> 1)) & 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
    x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
    x = (x | (x >> 16)) & 0x00000000FFFFFFFFull;
    return (uint32_t)x;
}
// One warp per read, one lane per 32-base word.  Read r of `a` is compared with read r of `f` over the bases both have
// (the reference's std::transform runs over read.seq and reads past the end of a shorter fixed read: undefined there,
// error-free here); packed form compares letters up to case and up to which non-ACGT letter a base is.
__global__ void __launch_bounds__(256) fixed_diff_kernel(BatchView a, BatchView f, uint32_t n_first, uint32_t* __restrict__ err_out) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < a.n_reads; r += nwarps) {
        const int la = (int)a.len[r];
        const uint32_t wa = a.word_off[r];
        const int lf = r < n_first && r < f.n_reads ? (int)f.len[r] : 0;
        const uint32_t wf = lf ? f.word_off[r] : 0u;
        const int both = la < lf ? la : lf;
        for (int w = lane; 32 * w < la; w += 32) {
            uint32_t m = 0;
            const int left = both - 32 * w;
            if (left > 0) {
                const unsigned long long x = a.bases[wa + w] ^ f.bases[wf + w];
                const uint32_t na = a.nmask[wa + w], nf = f.nmask[wf + w];
                m = ((even_bits(x | (x >> 1)) & ~na & ~nf) | (na ^ nf)) & (left >= 32 ? ~0u : (1u << left) - 1u);
            }
            err_out[wa + w] = m;
        }
    }
}

// ---- coverage estimation (kbbq.cc:229-236): total sequence length of a batch ----------------------------------------
__global__ void __launch_bounds__(256) total_bases_kernel(const uint32_t* __restrict__ len, uint32_t n, unsigned long long* __restrict__ out) {
    unsigned long long s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += len[i];
#pragma unroll
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(KBBQ_FULL, s, d);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// ---- order-independent digest of a word array: {sum, xor} of mix(word ^ mix(index)) -- equality checks across ranks/runs
__device__ __forceinline__ unsigned long long digest_mix(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void __launch_bounds__(256) digest_kernel(const unsigned long long* __restrict__ w, unsigned long long n,
                                                     unsigned long long* __restrict__ out) {
    unsigned long long s = 0, x = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long v = __ldg(w + i);
        if (v) {  // zero words contribute nothing: an empty table digests to {0, 0} whatever its size
            const unsigned long long h = digest_mix(v ^ digest_mix(i));
            s += h;
            x ^= h;
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        s += __shfl_xor_sync(KBBQ_FULL, s, d);
        x ^= __shfl_xor_sync(KBBQ_FULL, x, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, s);
        atomicXor(out + 1, x);
    }
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

// ---- the whole multi-GPU filter merge in one kernel over peer-mapped memory (NVLink / NVSwitch) -----------------------
// Bitwise OR has no NCCL reduction.  Every rank owns one slice of the bit array: it reads that slice from every
// peer's filter (128-bit loads through the switch), ORs them with its own and writes the result back into EVERY
// rank's filter -- reduce-scatter and all-gather fused, 2(N-1)/N of the array per GPU on the wire, no staging
// buffer.  Slices are disjoint, so the only synchronisation is a barrier before (all inserts done) and after.
#define KBBQ_MAX_PEERS 16
struct PeerPtrs {
    ulonglong2* p[KBBQ_MAX_PEERS];  // rotated: p[0] = this rank's filter words, p[s] = rank (rank+s) % world's
};
template <int WORLD>  // 0 = any world size (pointer array indexed at run time)
__global__ void __launch_bounds__(256) or_allreduce_p2p_kernel(PeerPtrs pp, int rank, int world_rt, unsigned long long n2) {
    const int world = WORLD ? WORLD : world_rt;
    const unsigned long long per = (n2 + world - 1) / world;
    const unsigned long long lo = per * rank, hi = lo + per < n2 ? lo + per : n2;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
        ulonglong2 v = __ldcg(pp.p[0] + i);
        if (WORLD) {
            ulonglong2 x[WORLD ? WORLD : 1];
#pragma unroll
            for (int s = 1; s < WORLD; ++s) x[s] = __ldcg(pp.p[s] + i);  // all peer loads in flight together
#pragma unroll
            for (int s = 1; s < WORLD; ++s) { v.x |= x[s].x; v.y |= x[s].y; }
#pragma unroll
            for (int s = 0; s < WORLD; ++s) __stcg(pp.p[s] + i, v);
        } else {
            for (int s = 1; s < world; ++s) {
                const ulonglong2 y = __ldcg(pp.p[s] + i);
                v.x |= y.x; v.y |= y.y;
            }
            for (int s = 0; s < world; ++s) __stcg(pp.p[s] + i, v);
        }
    }
}

}  // namespace kbbq
