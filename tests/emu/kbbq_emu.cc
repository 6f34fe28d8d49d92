// tests/emu/kbbq_emu.cc -- TEST TOOL, never shipped and never loaded by kbbq_b200/.
//
// Compiles the product's __host__ __device__ headers (kbbq_core.cuh, kbbq_errcall.cuh) with plain g++ and drives
// them with a serial stand-in for the warp query, so that the bit-level primitives and the warp-uniform error-calling
// state machine can be checked against the oracle on a box that has no GPU.  The kernels proper are only ever
// exercised by the `-m gpu` tests.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../kbbq_b200/csrc/kbbq_core.cuh"
#include "../../kbbq_b200/csrc/kbbq_errcall.cuh"

using namespace kbbq;

namespace {
bool contains_host(const BloomView& f, uint64_t key) {
    const uint64_t blk = bloom_block(f, key);
    const uint32_t pat = bloom_pattern(f, key);
    for (int w = 0; w < 8; ++w) {
        const uint64_t p = f.patterns[(uint64_t)pat * 8 + w];
        if ((f.words[blk * 8 + w] & p) != p) return false;
    }
    return true;
}
void insert_host(const BloomView& f, uint64_t key) {
    const uint64_t blk = bloom_block(f, key);
    const uint32_t pat = bloom_pattern(f, key);
    for (int w = 0; w < 8; ++w) f.words[blk * 8 + w] |= f.patterns[(uint64_t)pat * 8 + w];
    ++*f.inserted;
}
struct HostQuery {
    BloomView f;
    uint64_t nq = 0;
    template <class F>
    uint32_t query(int count, F fn) {
        uint32_t m = 0;
        for (int t = 0; t < count; ++t) {
            uint64_t key;
            if (fn(t, key)) {
                ++nq;
                if (contains_host(f, key)) m |= 1u << t;
            }
        }
        return m;
    }
    template <class F>
    uint32_t map32(F fn) {
        uint32_t m = 0;
        for (int t = 0; t < 32; ++t)
            if (fn(t)) m |= 1u << t;
        return m;
    }
};
// ErrCaller holds its query object by value (on the device it is a single pointer): hand it a reference wrapper
struct HostQueryRef {
    HostQuery* h;
    template <class F>
    uint32_t query(int count, F fn) { return h->query(count, fn); }
    template <class F>
    uint32_t map32(F fn) { return h->map32(fn); }
    template <class F>
    void par(int n, F fn) {  // the device runs fn(i) for i in [0,n) on different lanes; any order must do
        for (int i = n - 1; i >= 0; --i) fn(i);
    }
};
BloomView make_view(uint64_t* words, const uint64_t* patterns, uint64_t* inserted, uint64_t n_blocks, uint32_t s0, uint32_t s1) {
    BloomView v;
    v.words = (unsigned long long*)words;
    v.patterns = (const unsigned long long*)patterns;
    v.inserted = (unsigned long long*)inserted;
    v.n_blocks = n_blocks;
    v.fastmod_m = n_blocks > 1 ? (~0ull / n_blocks) + 1 : 0;
    v.salt0 = s0;
    v.salt1 = s1;
    return v;
}
struct CSeq {
    const unsigned long long* w;
    const uint32_t* nm;
    int len;
};
}  // namespace

extern "C" {

uint64_t emu_canonical_at(const uint64_t* bases, const uint32_t* nmask, int len, int j, int k, int p, int c, int* valid) {
    CSeq s{(const unsigned long long*)bases, nmask, len};
    uint64_t key = 0;
    *valid = kmer_key(s, j, k, p, c, key);
    return key;
}
uint32_t emu_fastmod(uint32_t h, uint64_t n) { return fastmod_u32(h, n > 1 ? (~0ull / n) + 1 : 0, n); }
uint32_t emu_hash(uint64_t key, uint32_t salt) { return hash_ap8(key, salt); }
uint64_t emu_rev2(uint64_t v) { return rev2(v); }

// pass 2 on the host with the kernel's formulation (presence mask -> popcount windows -> error mask -> inserts)
void emu_pass2(uint32_t n, const uint64_t* bases, const uint32_t* nmask, const uint8_t* qual, const uint32_t* word_off,
               const uint32_t* len, int k, const int* thr, uint64_t* s_words, const uint64_t* s_pat, uint64_t s_blocks,
               uint32_t s_s0, uint32_t s_s1, uint64_t* t_words, const uint64_t* t_pat, uint64_t t_blocks, uint32_t t_s0,
               uint32_t t_s1, uint64_t* t_inserted, uint32_t* err_bits) {
    uint64_t dummy = 0;
    BloomView sv = make_view(s_words, s_pat, &dummy, s_blocks, s_s0, s_s1);
    BloomView tv = make_view(t_words, t_pat, t_inserted, t_blocks, t_s0, t_s1);
    for (uint32_t r = 0; r < n; ++r) {
        const int L = (int)len[r], nk = L - k + 1;
        if (nk <= 0) continue;
        const uint32_t wo = word_off[r];
        CSeq s{(const unsigned long long*)bases + wo, nmask + wo, L};
        const uint8_t* q = qual + (size_t)wo * 32;
        std::vector<uint32_t> Pm_v((size_t)(L + 31) / 32 + 2, 0), Em_v((size_t)(L + 31) / 32 + 2, 0);
        uint32_t *Pm = Pm_v.data(), *Em = Em_v.data();
        for (int j = 0; j < nk; ++j) {
            uint64_t key;
            if (kmer_key(s, j, k, 0, -1, key) && contains_host(sv, key)) bit_set(Pm, j);
        }
        for (int i = 0; i < L; ++i) {
            const int lo = i - k + 1 > 0 ? i - k + 1 : 0;
            const int hi = (i < nk - 1 ? i : nk - 1) + 1;
            if (count_window32(Pm, lo, hi) <= thr[hi - lo] || q[i] <= 2) bit_set(Em, i);
        }
        for (int w = 0; w * 32 < L; ++w) err_bits[wo + w] = Em[w];
        for (int j = 0; j < nk; ++j) {
            uint64_t key;
            if (count_window32(Em, j, j + k) == 0 && kmer_key(s, j, k, 0, -1, key)) insert_host(tv, key);
        }
    }
}

// pass 3's error calling on the host through the very ErrCaller the kernel instantiates
uint64_t emu_get_errors(uint32_t n, const uint64_t* bases, const uint32_t* nmask, const uint8_t* qual,
                        const uint32_t* word_off, const uint32_t* len, int k, uint64_t* t_words, const uint64_t* t_pat,
                        uint64_t t_blocks, uint32_t t_s0, uint32_t t_s1, uint32_t* err_bits, int32_t* mod_idx,
                        int32_t* mod_base) {
    uint64_t dummy = 0;
    HostQuery q{make_view(t_words, t_pat, &dummy, t_blocks, t_s0, t_s1)};
    for (uint32_t r = 0; r < n; ++r) {
        const int L = (int)len[r];
        const uint32_t wo = word_off[r];
        const int nw = (L + 31) >> 5;
        bool has_n = false;  // as the kernel: key building skips the N masks of reads without N
        for (int w = 0; w < nw; ++w) has_n = has_n || nmask[wo + w] != 0;
        auto fill = [&](auto& sc) {
            for (int w = 0; w < nw; ++w) { sc.P_w[w] = bases[wo + w]; sc.P_nm[w] = nmask[wo + w]; }
            for (int i = 0; i < L; ++i)
                if (qual[(size_t)wo * 32 + i] <= 6) bit_set(sc.QL, i);
        };
        ErrResult res;
        if (L <= KBBQ_MAX_READ_LEN) {  // the statically sized scratch of pass3_repair_kernel
            ErrScratch sc;
            std::memset(&sc, 0, sizeof sc);
            fill(sc);
            ErrCaller<HostQueryRef> ec(HostQueryRef{&q}, k, qual + (size_t)wo * 32, sc, has_n);
            res = ec.run(L);
            for (int w = 0; w < nw; ++w) err_bits[wo + w] = sc.E[w];
        } else {  // ... and the one cut from a buffer sized for the read, as pass3_repair_long_kernel does
            std::vector<unsigned long long> buf(ErrScratchDyn::bytes(nw) / 8 + 1, 0);
            ErrScratchDyn sc;
            sc.carve(buf.data(), nw);
            fill(sc);
            ErrCaller<HostQueryRef, ErrScratchDyn> ec(HostQueryRef{&q}, k, qual + (size_t)wo * 32, sc, has_n);
            res = ec.run(L);
            for (int w = 0; w < nw; ++w) err_bits[wo + w] = sc.E[w];
        }
        mod_idx[r] = res.mod_idx;
        mod_base[r] = res.mod_base;
    }
    return q.nq;
}

}  // extern "C"
