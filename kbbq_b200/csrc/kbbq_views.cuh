// kbbq_views.cuh -- what the two device translation units share: the views of a batch and of the tables, the launch
// constants, and the one-lane filter probe.  (The long-read repair kernel lives in a translation unit of its own,
// kbbq_long.cu: with it in the same module the common repair kernel came out 1.3 % larger and 2.3 % slower.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kbbq_core.cuh"

namespace kbbq {

struct BatchView {
    const unsigned long long* bases;
    const uint32_t* nmask;
    const uint8_t* qual;
    const uint2* qmask;  // per word: .x = bases with qual <= 2, .y = bases with qual <= 6 (kbbq_batch::qmask)
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
#define KBBQ_HP_INVALID 0xFFFFFFFFu  // (hash, pattern) parked in shared memory: pattern slot of a k-mer that holds a non-ACGT base
#define KBBQ_HP_KNOWN 0xFFFFFFFEu    // ... of a k-mer known to be present (no probe)
#define KBBQ_FULL 0xffffffffu

// ---- filter probes ------------------------------------------------------------------------------------------------
// 32 bytes in one LDG.E.256 (sm_100): a 64-byte block or pattern is two instructions
__device__ __forceinline__ void ldg256(const unsigned long long* p, unsigned long long& a, unsigned long long& b,
                                       unsigned long long& c, unsigned long long& d) {
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
// Bring the line of a filter block into the L2 without holding a register for it.  A warp that probes one 32-k-mer
// round at a time has 32 blocks in flight and spends its life waiting for DRAM (the scan moved 4.0 TB/s where the same
// probes, unburdened, move 5.1); prefetching the blocks of ALL of a read's k-mers first puts ~120 in flight per warp
// and turns the rounds that follow into L2 hits.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
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

}  // namespace kbbq
