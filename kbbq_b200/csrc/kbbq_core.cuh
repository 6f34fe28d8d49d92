// kbbq_core.cuh -- primitives shared by every kernel of the recalibration hot path.
//
// Data model: a read is a run of 32-base words.  Word w of a read holds bases
// [32w, 32w+32): base i sits at bits [2*(i%32), 2*(i%32)+2) of `bases[w]` (A=0,C=1,G=2,T=3) and bit (i%32) of
// `nmask[w]` is set when the base is not A/C/G/T (the reference resets its rolling k-mer there, bloom.hh:350-358).
// With this little-endian layout a k-mer read straight out of the words IS the reference's reverse-strand word
// x[1] complemented (bloom.hh:354), and the forward word x[0] is its 2-bit-group reversal (bloom.hh:353):
// no rolling state is needed, so every k-mer of a read can be produced by an independent lane.
//
// Everything here is __host__ __device__ so that the bit-twiddling can be unit-tested on a box without a GPU
// (tests/emu); the product only ever calls it from kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KBBQ_HD __host__ __device__ __forceinline__
#else
#define KBBQ_HD inline
#endif

// Loops in the per-read logic are scalar bookkeeping executed once per read by a whole warp: unrolling them buys
// nothing and multiplies the code size of a kernel that lives or dies by the instruction cache.
#if defined(__CUDACC__) && !defined(KBBQ_ALLOW_UNROLL)
#define KBBQ_NOUNROLL _Pragma("unroll 1")
#else
#define KBBQ_NOUNROLL
#endif

#define KBBQ_MAX_K 32          // KBBQ_MAX_KMER, bloom.hh:15
#define KBBQ_MAXQ 93           // KBBQ_MAXQ, covariateutils.hh:3
#define KBBQ_NQ 94
#define KBBQ_NUM_PATTERNS 65536  // pattern_blocked_bf::num_patterns, bloom.hh:185
#define KBBQ_MAX_READ_LEN 512  // longest read the STATIC per-warp shared-memory scratch holds; batches with longer reads
                               // (up to KBBQ_ABI_MAX_READ_LEN) take the LONG instantiations and kbbq_long.cu
#define KBBQ_MAXW (KBBQ_MAX_READ_LEN / 32)

namespace kbbq {

// Device-side view of one pattern-blocked Bloom filter (bloom.hh:181-332).
struct BloomView {
    unsigned long long* words;           // n_blocks * 8 little-endian u64 (64-byte blocks)
    const unsigned long long* patterns;  // 65536 * 8
    unsigned long long* inserted;        // insert-call counter (counts duplicates, bloom.hh:266)
    uint64_t n_blocks;
    uint64_t fastmod_m;                  // ceil(2^64 / n_blocks), for h % n_blocks without a divide
    uint32_t salt0, salt1;               // salt_[0] picks the block, salt_[1] the pattern (bloom.hh:257-258)
};

// bloom_filter::hash_ap for an 8-byte key: one iteration of bloom_filter.hpp:556-565 on the two LE halves.
KBBQ_HD uint32_t hash_ap8(uint64_t key, uint32_t h) {
    const uint32_t i1 = (uint32_t)key, i2 = (uint32_t)(key >> 32);
    h ^= (h << 7) ^ (i1 * (h >> 3)) ^ (~((h << 11) + (i2 ^ (h >> 5))));
    return h;
}

KBBQ_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// h % n for 32-bit h and n < 2^32 (Lemire, "Faster remainder by direct computation", 2019): exact.
KBBQ_HD uint32_t fastmod_u32(uint32_t h, uint64_t m, uint64_t n) { return (uint32_t)mulhi64(m * (uint64_t)h, n); }

// block index: get_block(), bloom.hh:99-106; pattern index: get_pattern(), bloom.hh:250-253
KBBQ_HD uint64_t bloom_block(const BloomView& b, uint64_t key) {
    const uint32_t h = hash_ap8(key, b.salt0);
    return b.n_blocks <= 0xFFFFFFFFull ? (uint64_t)fastmod_u32(h, b.fastmod_m, b.n_blocks) : (uint64_t)h;
}
// the same from an already computed hash_ap8(key, salt0)
KBBQ_HD uint64_t bloom_block_of_hash(const BloomView& b, uint32_t h) {
    return b.n_blocks <= 0xFFFFFFFFull ? (uint64_t)fastmod_u32(h, b.fastmod_m, b.n_blocks) : (uint64_t)h;
}
KBBQ_HD uint32_t bloom_pattern(const BloomView& b, uint64_t key) { return hash_ap8(key, b.salt1) & (KBBQ_NUM_PATTERNS - 1); }

// reverse the 32 two-bit groups of a word
KBBQ_HD uint64_t rev2(uint64_t v) {
#if defined(__CUDA_ARCH__)
    v = __brevll(v);
#else
    v = ((v >> 1) & 0x5555555555555555ull) | ((v & 0x5555555555555555ull) << 1);
    v = ((v >> 2) & 0x3333333333333333ull) | ((v & 0x3333333333333333ull) << 2);
    v = ((v >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((v & 0x0F0F0F0F0F0F0F0Full) << 4);
    v = __builtin_bswap64(v);
#endif
    // a full bit reversal also swapped the two bits inside each group: swap them back
    return ((v >> 1) & 0x5555555555555555ull) | ((v & 0x5555555555555555ull) << 1);
}

KBBQ_HD uint64_t kmask(int k) { return k < 32 ? ((1ull << (2 * k)) - 1) : ~0ull; }  // bloom.hh:347

// canonical k-mer = min(x[0], x[1]) (bloom.hh:360) from the little-endian k-mer word v (2k significant bits)
KBBQ_HD uint64_t canonical(uint64_t v, int k) {
    const uint64_t x1 = ~v & kmask(k);
    const uint64_t x0 = rev2(v) >> (64 - 2 * k);
    return x0 < x1 ? x0 : x1;
}

// bases [start, start+k) of a packed read as a little-endian k-mer word; needs 0 <= start, start+k <= 32*words
template <class W>
KBBQ_HD uint64_t extract_le(const W* w, int start, int k) {
    const int wi = start >> 5, o = start & 31;
    uint64_t v = (uint64_t)w[wi] >> (2 * o);
    if (o + k > 32) v |= (uint64_t)w[wi + 1] << (64 - 2 * o);
    return v & kmask(k);
}
// N flags of bases [start, start+k) as the low k bits of a word
KBBQ_HD uint32_t extract_n(const uint32_t* nm, int start, int k) {
    const int wi = start >> 5, o = start & 31;
    uint32_t m = nm[wi] >> o;
    if (o + k > 32) m |= nm[wi + 1] << (32 - o);
    return k < 32 ? (m & ((1u << k) - 1)) : m;
}

// A read (or working copy of one) in packed form.
struct SeqRef {
    unsigned long long* w;
    uint32_t* nm;
    int len;
};
KBBQ_HD int seq_code(const SeqRef& s, int i) { return (int)((s.w[i >> 5] >> (2 * (i & 31))) & 3); }
KBBQ_HD bool seq_is_n(const SeqRef& s, int i) { return (s.nm[i >> 5] >> (i & 31)) & 1; }
// `seq[i] == c` for c in "ACGT": true only for a real base equal to c
KBBQ_HD bool seq_eq(const SeqRef& s, int i, int c) { return !seq_is_n(s, i) && seq_code(s, i) == c; }
KBBQ_HD void seq_set(const SeqRef& s, int i, int c) {
    const int o = 2 * (i & 31);
    s.w[i >> 5] = (s.w[i >> 5] & ~(3ull << o)) | ((unsigned long long)c << o);
    s.nm[i >> 5] &= ~(1u << (i & 31));
}

// Canonical key of the k-mer starting at j, with base p (if inside the k-mer) replaced by code c (c < 0: none).
// Returns false when the k-mer holds a non-ACGT base, i.e. Kmer::valid() would be false (bloom.hh:368).
// NM_OPTIONAL: s.nm may be nullptr, meaning the caller knows the sequence holds no N (pass 3 skips the masks then).
template <bool NM_OPTIONAL = false, class S>
KBBQ_HD bool kmer_key(const S& s, int j, int k, int p, int c, uint64_t& key) {
    uint64_t v = extract_le(s.w, j, k);
    uint32_t n = (!NM_OPTIONAL || s.nm) ? extract_n(s.nm, j, k) : 0u;
    if (c >= 0 && p >= j && p < j + k) {
        const int o = p - j;
        v = (v & ~(3ull << (2 * o))) | ((uint64_t)c << (2 * o));
        n &= ~(1u << o);
    }
    key = canonical(v, k);
    return n == 0;
}

// ---- multi-word bit masks (bit i of word i/32) -------------------------------------------------------------------
KBBQ_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
KBBQ_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
KBBQ_HD bool bit_get(const uint32_t* m, int i) { return (m[i >> 5] >> (i & 31)) & 1; }
KBBQ_HD void bit_set(uint32_t* m, int i) { m[i >> 5] |= 1u << (i & 31); }
KBBQ_HD void bit_clr(uint32_t* m, int i) { m[i >> 5] &= ~(1u << (i & 31)); }
// smallest index >= from whose bit is set (clear), or n when there is none
KBBQ_HD int next_one(const uint32_t* m, int n, int from) {
    KBBQ_NOUNROLL
    while (from < n) {
        const uint32_t x = m[from >> 5] >> (from & 31);
        if (x) { const int r = from + ctz32(x); return r < n ? r : n; }
        from = (from | 31) + 1;
    }
    return n;
}
KBBQ_HD int next_zero(const uint32_t* m, int n, int from) {
    KBBQ_NOUNROLL
    while (from < n) {
        const uint32_t x = ~m[from >> 5] >> (from & 31);
        if (x) { const int r = from + ctz32(x); return r < n ? r : n; }
        from = (from | 31) + 1;
    }
    return n;
}
// largest index <= from whose bit is clear, or -1 when there is none (from >= 0)
KBBQ_HD int prev_zero(const uint32_t* m, int from) {
    KBBQ_NOUNROLL
    while (from >= 0) {
        const int o = from & 31;
        const uint32_t x = ~m[from >> 5] << (31 - o);  // bit `from` moved to the top
        if (x) {
#if defined(__CUDA_ARCH__)
            return from - __clz((int)x);
#else
            return from - __builtin_clz(x);
#endif
        }
        from = (from & ~31) - 1;
    }
    return -1;
}
// number of set bits in [lo, hi) for windows of at most 32 bits (every k-mer window is); reads m[(lo >> 5) + 1]
KBBQ_HD int count_window32(const uint32_t* m, int lo, int hi) {
    const uint64_t v = (((uint64_t)m[(lo >> 5) + 1] << 32) | m[lo >> 5]) >> (lo & 31);
    const uint32_t x = (uint32_t)v & (hi - lo >= 32 ? ~0u : ((1u << (hi - lo)) - 1u));
    return popc32(x);
}
// number of set bits in [lo, hi)
KBBQ_HD int count_range(const uint32_t* m, int lo, int hi) {
    int c = 0;
    KBBQ_NOUNROLL
    while (lo < hi) {
        const int o = lo & 31;
        int take = 32 - o;
        if (take > hi - lo) take = hi - lo;
        uint32_t x = m[lo >> 5] >> o;
        if (take < 32) x &= (1u << take) - 1;
        c += popc32(x);
        lo += take;
    }
    return c;
}

}  // namespace kbbq
