// kbbq_errcall.cuh -- per-read error calling against the trusted filter (pass 3), one warp per read.
//
// This is CReadData::get_errors (readutils.cc:238-570) and its helpers (bloom.cc:83-305, readutils.cc:195-235)
// re-designed for a warp: the read lives in shared memory as 2-bit words, every question the reference asks of the
// filter one k-mer at a time is asked for up to 32 k-mers at once (one lane per k-mer, one ballot), and all the
// bookkeeping (anchors, streaks, windows, error flags) is bit-mask arithmetic that every lane carries out identically,
// so the control flow is warp-uniform and the only shared-memory writes are idempotent.
//
// `Q` supplies the two warp-wide primitives:
//   template<class F> uint32_t query(int count, F f)
//       evaluates f(t, key) -> bool valid for t in [0,count), count <= 32, and returns the mask of t for which the
//       k-mer is valid AND present in the trusted filter (Bloom::query, bloom.hh:398);
//   template<class F> uint32_t map32(F f)
//       returns the mask of t in [0,32) for which f(t) is true;
//   template<class F> void par(int n, F f)
//       runs f(i) for every i in [0,n), in any order (one lane each on the device): the per-word loops.
// On the device those are one lane per t and a __ballot_sync; tests/emu provides serial stand-ins so the logic can be
// exercised without a GPU.
//
// Probe economy (none of it changes a result):
//   * the initial scan's flags are kept (T): a k-mer the running call has not modified is never queried again --
//     the reference re-queries the whole original read for its "adjust" test (readutils.cc:435-454) and walks the
//     rest of the read for every candidate base (bloom.cc:147-172);
//   * candidate bases are pre-tested on their first k-mer in ONE query; only survivors get the other k-mers that
//     contain the repaired base queried.
//
// Code size matters here (the kernel runs from the instruction cache or not at all): the one heavy routine,
// kmers_mask, is out of line, the two sides of the anchor share one loop, and the three get_errors invocations of a
// read (whole read, bad prefix, bad suffix) share one call site.
//
// Conventions: positions are int; NPOS (-1) stands for std::string::npos where the reference uses it as a flag.
// Deliberate reproduction of reference behaviour is marked "ref quirk".
#pragma once
#include "kbbq_core.cuh"

#if defined(__CUDACC__)
#define KBBQ_HD_NOINLINE __host__ __device__ __noinline__
#else
#define KBBQ_HD_NOINLINE __attribute__((noinline))
#endif

namespace kbbq {

#define KBBQ_NPOS (-1)

// Per-warp scratch (shared memory on the device).
struct ErrScratch {
    unsigned long long P_w[KBBQ_MAXW + 1];  // the read as `this->seq` of the top-level call: original + right-side fixes
    unsigned long long A_w[KBBQ_MAXW + 1];  // `this->seq` of a sub-read call (readutils.cc:551,560)
    unsigned long long C_w[KBBQ_MAXW + 1];  // `revcomped` (readutils.cc:351-353)
    uint32_t P_nm[KBBQ_MAXW + 1];
    uint32_t A_nm[KBBQ_MAXW + 1];
    uint32_t C_nm[KBBQ_MAXW + 1];
    uint32_t O_nm[KBBQ_MAXW + 1];   // N flags of `original_seq` of the running call
    uint32_t T_orig[KBBQ_MAXW + 1]; // trusted flag of every k-mer of original_seq (bit j = k-mer starting at j)
    uint32_t T_cur[KBBQ_MAXW + 1];  // same for seq after correct_one, when that ran
    uint32_t E[KBBQ_MAXW + 1];      // this->errors of the top-level read
    uint32_t QL[KBBQ_MAXW + 1];     // bit i: qual[i] <= 6 (the overcorrection half weight), top-level coordinates
    uint32_t F[KBBQ_MAXW + 1];      // overcorrection check: flagged bases whose original base was ACGT ...
    uint32_t FQ[KBBQ_MAXW + 1];     // ... and those of them with qual <= 6
};

// The same scratch with its arrays somewhere else: reads longer than KBBQ_MAX_READ_LEN get theirs cut from dynamic shared
// memory sized for the batch's longest read (pass3_repair_long_kernel); every array holds maxw + 1 words.
struct ErrScratchDyn {
    unsigned long long *P_w, *A_w, *C_w;
    uint32_t *P_nm, *A_nm, *C_nm, *O_nm, *T_orig, *T_cur, *E, *QL, *F, *FQ;
    static KBBQ_HD size_t bytes(int maxw) { return (size_t)(maxw + 1) * (3 * 8 + 10 * 4); }
    KBBQ_HD void carve(void* base, int maxw) {
        const size_t n = (size_t)maxw + 1;
        unsigned long long* w = static_cast<unsigned long long*>(base);
        P_w = w; A_w = w + n; C_w = w + 2 * n;
        uint32_t* u = reinterpret_cast<uint32_t*>(w + 3 * n);
        P_nm = u; A_nm = u + n; C_nm = u + 2 * n; O_nm = u + 3 * n; T_orig = u + 4 * n; T_cur = u + 5 * n;
        E = u + 6 * n; QL = u + 7 * n; F = u + 8 * n; FQ = u + 9 * n;
    }
};

struct ErrResult {
    int mod_idx;   // ref quirk: the all-trusted early return after correct_one leaves the corrected base in seq
    int mod_base;  // (readutils.cc:263-265 vs 568); -1 when seq ends up equal to the input
};

template <class Q, class SC = ErrScratch>
struct ErrCaller {
    Q q;  // held by value: on the device it is one pointer, and nothing here may force the caller's state into memory
    const int k;
    const uint8_t* qual;  // qualities of the top-level read (only correct_one's tie-break reads them)
    SC& sc;
    const uint32_t* T = nullptr;  // cached trusted flags of the running call's sequence (see cached_run)
    bool T_mirrored = false;
    const bool has_n;  // false: the read holds no non-ACGT base (nor does anything derived from it): key building skips the N masks
    KBBQ_HD ErrCaller(Q q_, int k_, const uint8_t* qual_, SC& sc_, bool has_n_ = true)
        : q(q_), k(k_), qual(qual_), sc(sc_), has_n(has_n_) {}

    // THE probe routine.  Lane t asks about the k-mer of s starting at j0 + t*stride with base p replaced:
    //   stride == 1: consecutive k-mers, replacement base c (c < 0: no replacement);
    //   stride == 0: one k-mer, lane t tries base (cpack >> 2t) & 3  -- the candidate pre-test.
    // Out of line, and static with scalar arguments only: a noinline MEMBER would pin `*this` (memo, anchors, flags) in
    // local memory for the whole call tree.
    static KBBQ_HD_NOINLINE uint32_t kmers_mask_s(Q q, unsigned long long* w, uint32_t* nm, int kk, int j0, int stride,
                                                  int count, int p, int c, uint32_t cpack) {
        const SeqRef s{w, nm, 0};
        return q.query(count, [&](int t, uint64_t& key) {
            return kmer_key<true>(s, j0 + t * stride, kk, p, stride ? c : (int)((cpack >> (2 * t)) & 3u), key);
        });
    }
    KBBQ_HD uint32_t kmers_mask(const SeqRef& s, int j0, int stride, int count, int p, int c, uint32_t cpack) {
        return kmers_mask_s(q, s.w, has_n ? s.nm : nullptr, k, j0, stride, count, p, c, cpack);
    }
    // number of consecutive trusted k-mers starting at j0, looking no further than j1 (exclusive)
    KBBQ_HD int trusted_run(const SeqRef& s, int j0, int j1, int p, int c) {
        int j = j0;
        KBBQ_NOUNROLL
        while (j < j1) {
            const int count = j1 - j < 32 ? j1 - j : 32;
            const uint32_t m = kmers_mask(s, j, 1, count, p, c, 0);
            const uint32_t full = count < 32 ? ((1u << count) - 1) : ~0u;
            if (m != full) return j - j0 + ctz32(~m);
            j += count;
        }
        return j1 - j0;
    }
    // fill T[0 .. nk) with the trusted flag of every k-mer of s
    KBBQ_HD void scan(const SeqRef& s, uint32_t* Tout) {
        const int nk = s.len - k + 1;
        KBBQ_NOUNROLL
        for (int j = 0; j < nk; j += 32) Tout[j >> 5] = kmers_mask(s, j, 1, nk - j < 32 ? nk - j : 32, 0, -1, 0);
    }

    // bloom::find_longest_trusted_seq (bloom.cc:96-128) on the k-mer flags: first longest streak wins
    KBBQ_HD void longest_trusted(const uint32_t* Tm, int len, int& a0, int& a1) {
        const int nk = len - k + 1;
        int best = 0, bs = 0, be = 0, pos = 0;
        KBBQ_NOUNROLL
        while (true) {
            const int s = next_one(Tm, nk, pos);
            if (s >= nk) break;
            const int e = next_zero(Tm, nk, s);
            if (e - s > best) { best = e - s; bs = s; be = e; }
            pos = e;
        }
        if (best == 0) { a0 = KBBQ_NPOS; a1 = KBBQ_NPOS; return; }
        a0 = bs;
        a1 = be == nk ? KBBQ_NPOS : be - 1 + k - 1;
    }

    // Trusted flags of k-mers the running call has NOT modified since its scan, looked up instead of re-queried.
    // `T` holds one bit per k-mer of the scanned sequence in forward coordinates; when the working sequence is the
    // reverse complement, k-mer j of it is k-mer nk-1-j of the forward read (canonical k-mers are strand symmetric).
    // Returns the number of consecutive trusted k-mers of the working sequence starting at index j.
    KBBQ_HD int cached_run(int j, int len) const {
        const int nk = len - k + 1;
        if (j >= nk) return 0;
        if (!T_mirrored) return next_zero(T, nk, j) - j;
        const int f = nk - 1 - j;
        return f - prev_zero(T, f);
    }

    // bases of "ACGT" (or "TGCA") that differ from s[p], packed 2 bits each in test order
    KBBQ_HD uint32_t candidates(const SeqRef& s, int p, bool reverse_order, int& ncand) {
        // all four bases in test order (bloom.cc:140-141), two bits each: ACGT = 0xE4, TGCA = 0x1B; a real base at p is
        // cut out of the sequence by closing the gap above it (an N keeps all four)
        const uint32_t order = reverse_order ? 0x1Bu : 0xE4u;
        if (seq_is_n(s, p)) {
            ncand = 4;
            return order;
        }
        const int code = seq_code(s, p);
        const int at = reverse_order ? 3 - code : code;  // where the read's own base sits in the test order
        const uint32_t low = order & ((1u << (2 * at)) - 1u);
        ncand = 3;
        return low | ((order >> (2 * at + 2)) << (2 * at));
    }

    // What the pre-test and the walks over the k-mers containing position p found, remembered between
    // adjust_right_anchor's first phase and the find_longest_fix that immediately follows it at the same position on
    // the same, still unmodified, sequence (they ask the filter exactly the same questions).
    struct Memo {
        int p;            // position, -1 = empty
        bool mirrored;    // which working sequence
        uint32_t first;   // bit c: first k-mer with base c trusted
        uint32_t run_in;  // per base c, 8 bits at 8c: trusted streak over the k-mers containing p (1 .. k), 0 = not walked yet
    };
    // four small counters in one register (indexed arrays would live in local memory once loops are not unrolled)
    static KBBQ_HD int get16(uint64_t v, int i) { return (int)((v >> (16 * i)) & 0xFFFFu); }
    static KBBQ_HD uint64_t set16(uint64_t v, int i, int x) { return (v & ~(0xFFFFull << (16 * i))) | ((uint64_t)(unsigned)x << (16 * i)); }
    static KBBQ_HD int get8(uint32_t v, int i) { return (int)((v >> (8 * i)) & 0xFFu); }
    static KBBQ_HD uint32_t set8(uint32_t v, int i, int x) { return (v & ~(0xFFu << (8 * i))) | ((uint32_t)x << (8 * i)); }
    Memo memo;

    KBBQ_HD void memo_reset() { memo.p = -1; }
    // first-k-mer flags of all candidate bases at p (bit c of the result), through the memo
    KBBQ_HD uint32_t first_flags(const SeqRef& s, int p, uint32_t cpack, int ncand) {
        if (memo.p == p && memo.mirrored == T_mirrored) return memo.first;
        const uint32_t m0 = kmers_mask(s, p - k + 1, 0, ncand, p, 0, cpack);
        memo.p = p;
        memo.mirrored = T_mirrored;
        memo.run_in = 0;
        // bit t of m0 (test order) -> bit c of first (base code); m0 has no bit at or above ncand
        memo.first = ((m0 & 1u) << (cpack & 3u)) | (((m0 >> 1) & 1u) << ((cpack >> 2) & 3u)) | (((m0 >> 2) & 1u) << ((cpack >> 4) & 3u)) |
                     (((m0 >> 3) & 1u) << ((cpack >> 6) & 3u));
        return memo.first;
    }
    // trusted streak over the k-mers that contain p (there are nin of them) with base c at p, first one known trusted
    KBBQ_HD int run_containing(const SeqRef& s, int p, int nin, int c) {
        int r = get8(memo.run_in, c);
        if (r == 0) {
            r = 1 + trusted_run(s, p - k + 2, p - k + 1 + nin, p, c);
            memo.run_in = set8(memo.run_in, c, r);
        }
        return r;
    }

    // bloom::find_longest_fix (bloom.cc:130-188) for the substring of s starting at `start`; the base under repair is
    // p = start + k - 1.  Returns the number of best candidates; fix0 = the first of them, fixlen = best_i.
    // The walk past the read end with guessed bases (bloom.cc:83-94,149) only ever matters as a tie-break between
    // candidates that BOTH keep every in-read k-mer trusted: a lone such candidate is the unique best whatever its
    // extension, and the caller only uses its fixlen as "reaches the read end".  It is skipped in that case.
    KBBQ_HD int longest_fix(const SeqRef& s, int start, bool reverse_order, int& fix0, int& fixlen, bool& multiple) {
        const int p = start + k - 1;
        const int sublen = s.len - start;
        const int nsub = sublen - k + 1;                          // k-mers of the substring that lie inside the read
        const int i_stop = 2 * k - 1 > sublen ? 2 * k - 1 : sublen;  // bloom.cc:147
        const int nin = nsub < k ? nsub : k;                      // ... of which these contain p
        int nbest = 0, best_i = 0, first_ok = 0, ncand, nreach = 0;
        uint64_t runs = 0;  // per candidate t, 16 bits at 16t
        fix0 = 0;
        const uint32_t cpack = candidates(s, p, reverse_order, ncand);
        const uint32_t first = first_flags(s, p, cpack, ncand);
        KBBQ_NOUNROLL
        for (int t = 0; t < ncand; ++t) {
            const int c = (int)((cpack >> (2 * t)) & 3u);
            int run = 0;
            if ((first >> c) & 1u) {
                ++first_ok;  // `single`/`multiple` bookkeeping, bloom.cc:163-166
                run = run_containing(s, p, nin, c);
                if (run == nin && nsub > nin) run += cached_run(start + nin, s.len);
            }
            runs = set16(runs, t, run);
            nreach += run == nsub;
        }
        const bool extend = sublen < i_stop && nreach >= 2;
        KBBQ_NOUNROLL
        for (int t = 0; t < ncand; ++t) {
            const int c = (int)((cpack >> (2 * t)) & 3u);
            const int run = get16(runs, t);
            int i = k - 1 + run;
            if (extend && run == nsub) {
                uint64_t v = extract_le(s.w, s.len - k, k);
                if (p >= s.len - k) {
                    const int o = p - (s.len - k);
                    v = (v & ~(3ull << (2 * o))) | ((uint64_t)c << (2 * o));
                }
                const int kk = k;
                KBBQ_NOUNROLL
                for (i = sublen; i < i_stop; ++i) {
                    const uint64_t shifted = v >> 2;
                    const uint32_t m = q.query(4, [&](int tt, uint64_t& key) {
                        const int cc = reverse_order ? 3 - tt : tt;
                        key = canonical(shifted | ((uint64_t)cc << (2 * (kk - 1))), kk);
                        return true;
                    });
                    if (m == 0) break;  // no trusted next k-mer
                    const int tt = ctz32(m);
                    v = shifted | ((uint64_t)(reverse_order ? 3 - tt : tt) << (2 * (k - 1)));
                }
            }
            if (i > best_i) { nbest = 1; fix0 = c; best_i = i; }
            else if (i == best_i) { ++nbest; }
        }
        fixlen = best_i;
        multiple = first_ok >= 2;
        return nbest;
    }

    // bloom::adjust_right_anchor (bloom.cc:208-277)
    KBBQ_HD int adjust_right_anchor(int anchor, const SeqRef& s, bool& multiple) {
        multiple = false;
        const int kk = k;
        {   // does a single substitution at anchor+1 repair everything it touches?  then the anchor stands
            const int m = anchor + 1;
            int cnt = s.len - m;  // queries possible before the read ends: i = 0 .. min(k, len-1-m)
            if (cnt > k + 1) cnt = k + 1;
            const int nin = cnt < k ? cnt : k;  // k-mers that contain m; the (k+1)-th, if any, starts past m
            int ncand;
            const uint32_t cpack = candidates(s, m, false, ncand);
            const uint32_t first = first_flags(s, m, cpack, ncand);
            KBBQ_NOUNROLL
            for (int t = 0; t < ncand; ++t) {
                const int c = (int)((cpack >> (2 * t)) & 3u);
                if (!((first >> c) & 1u)) continue;
                int run = run_containing(s, m, nin, c);
                if (run == nin && cnt > nin) run += cached_run(m + 1, s.len) >= 1 ? 1 : 0;
                if (run == cnt) return anchor;
            }
        }
        // otherwise try pulling the anchor back by up to k/2: first k-mers of every (position, base) pair in bulk
        // positions i = k/2-1 .. 0 while anchor > i+k-1 (bloom.cc:228): the test only gets easier as i falls, so the walk
        // covers all k/2 positions or stops at once
        const int ni = anchor > k / 2 - 1 + k - 1 ? k / 2 : 0;
        const int npairs = 4 * ni;
        const int need = k / 2 + 2;  // k-mers ending at m .. m+k/2+1 must all be trusted
        const int jmax = s.len - k + 1;
        KBBQ_NOUNROLL
        for (int c0 = 0; c0 < npairs; c0 += 32) {
            const int count = npairs - c0 < 32 ? npairs - c0 : 32;
            uint32_t hits = q.query(count, [&](int t, uint64_t& key) {
                const int i = kk / 2 - 1 - ((c0 + t) >> 2), c = (c0 + t) & 3;
                const int m = anchor - i;
                if (seq_eq(s, m, c)) return false;
                return kmer_key(s, m - kk + 1, kk, m, c, key);
            });
            KBBQ_NOUNROLL
            while (hits) {
                const int t = ctz32(hits);
                hits &= hits - 1;
                const int i = k / 2 - 1 - ((c0 + t) >> 2), c = (c0 + t) & 3;
                const int m = anchor - i;
                multiple = true;
                const int j1 = m - k + 1 + need;
                if (j1 <= jmax && 1 + trusted_run(s, m - k + 2, j1, m, c) == need) return m - 1;
            }
        }
        return anchor;
    }

    // CReadData::correct_one (readutils.cc:195-235) + bloom::biggest_consecutive_trusted_block (bloom.cc:279-305)
    // Static for the same reason as kmers_mask_s; returns 4*position + base, or KBBQ_NPOS.
    static KBBQ_HD_NOINLINE int correct_one_s(Q q, unsigned long long* w, uint32_t* nm, int len, int k, const uint8_t* qual) {
        const SeqRef s{w, nm, len};
        int best_len = 0, best_base = 0, best_pos = KBBQ_NPOS;
        const int ncand = 4 * len;  // candidate id = 4*i + c; identical bases are skipped like `continue`
        const int kk = k;
        KBBQ_NOUNROLL
        for (int c0 = 0; c0 < ncand; c0 += 32) {
            const int count = ncand - c0 < 32 ? ncand - c0 : 32;
            // the "magic k-mer" pre-test (readutils.cc:207-213) for 32 (position, base) candidates at once
            uint32_t hits = q.query(count, [&](int t, uint64_t& key) {
                const int i = (c0 + t) >> 2, c = (c0 + t) & 3;
                if (seq_eq(s, i, c)) return false;
                int ms = 0;
                if (i > kk / 2 - 1) { ms = i - kk / 2 + 1; if (ms > len - kk) ms = len - kk; }
                return kmer_key(s, ms, kk, i, c, key);
            });
            KBBQ_NOUNROLL
            while (hits) {
                const int t = ctz32(hits);
                hits &= hits - 1;
                const int i = (c0 + t) >> 2, c = (c0 + t) & 3;
                const int start = i > k - 1 ? i - k + 1 : 0;
                int nwin = len - k + 1 - start;  // k-mers of substr(start, 2k-1)
                if (nwin > k) nwin = k;
                uint32_t m = kmers_mask_s(q, w, nm, k, start, 1, nwin, i, c, 0);
                // the walk stops once k - (#misses) < best_len (bloom.cc:295): cut the mask at that miss
                if (best_len >= 1) {
                    const int stop_after = k - best_len + 1;  // this many misses end the walk
                    uint32_t z = ~m & (nwin < 32 ? ((1u << nwin) - 1) : ~0u);
                    int seen = 0;
                    KBBQ_NOUNROLL
                    while (z) {
                        const int b = ctz32(z);
                        z &= z - 1;
                        if (++seen == stop_after) { m &= (1u << b) - 1; break; }
                    }
                }
                int n_in = 0;
                KBBQ_NOUNROLL
                while (m) {  // longest streak of ones
                    m &= m << 1;
                    ++n_in;
                }
                if (n_in > best_len) { best_base = c; best_pos = i; best_len = n_in; }
                else if (n_in == best_len && qual[i] < qual[best_pos]) { best_base = c; best_pos = i; }
            }
        }
        if (best_len > 0) seq_set(s, best_pos, best_base);
        return best_pos == KBBQ_NPOS ? KBBQ_NPOS : 4 * best_pos + best_base;
    }
    KBBQ_HD int correct_one(const SeqRef& s, int qbase, int& out_base) {
        const int r = correct_one_s(q, s.w, s.nm, s.len, k, qual + qbase);
        out_base = r & 3;
        return r == KBBQ_NPOS ? KBBQ_NPOS : r >> 2;
    }

    // reverse complement of src into dst (readutils.cc:351-353): N stays N
    KBBQ_HD void revcomp(const SeqRef& src, const SeqRef& dst) {
        const int len = src.len;
        const int nw = (len + 31) >> 5;
        q.par(nw + 1, [&](int wd) {
            if (wd == nw) {
                dst.w[nw] = 0;
                dst.nm[nw] = 0;
                return;
            }
            // dst bases [32wd, 32wd+32) = complement of src bases len-1-32wd down to len-32-32wd
            const int hi = len - 1 - 32 * wd;
            int lo = hi - 31, cnt = 32;
            if (lo < 0) { cnt = hi + 1; lo = 0; }
            uint64_t v = extract_le(src.w, lo, cnt);          // src[lo..hi], src[lo] in the low bits
            uint32_t n = extract_n(src.nm, lo, cnt);
            v = rev2(v) >> (64 - 2 * cnt);                    // now src[hi] in the low bits
            uint32_t nr = 0;
#if defined(__CUDA_ARCH__)
            nr = __brev(n) >> (32 - cnt);
#else
            for (int b = 0; b < cnt; ++b) nr |= ((n >> b) & 1u) << (cnt - 1 - b);
#endif
            dst.w[wd] = ~v & kmask(cnt);
            dst.nm[wd] = nr;
        });
    }

    // One invocation of get_errors on `cur` (length cur.len), whose base 0 is base `base` of the top-level read.
    // Error flags live in sc.E at offset `base`.  Returns bad_prefix / bad_suffix for the caller's recursion test.
    // `prescanned`: sc.T_orig already holds the trusted flag of every k-mer of `cur` (the scan kernel made them).
    KBBQ_HD void call(const SeqRef& cur, int base, bool top, int& bad_prefix, int& bad_suffix, ErrResult& res, bool prescanned = false) {
        const int len = cur.len;
        bad_prefix = 0;
        bad_suffix = KBBQ_NPOS;
        if (len < k) return;  // the reference is undefined here (bloom.cc:30 / readutils.cc:207 underflow)
        const int nw = (len + 31) >> 5;
        q.par(nw + 1, [&](int w) { sc.O_nm[w] = cur.nm[w]; });
        bool multiple = false;
        memo_reset();
        if (!prescanned) scan(cur, sc.T_orig);
        T = sc.T_orig;
        T_mirrored = false;
        int a0, a1;
        longest_trusted(sc.T_orig, len, a0, a1);
        if (a0 == KBBQ_NPOS) {  // no trusted k-mer in the read (readutils.cc:250-262)
            multiple = true;
            int fixed_base;
            const int idx = correct_one(cur, base, fixed_base);
            if (idx == KBBQ_NPOS) return;
            scan(cur, sc.T_cur);
            T = sc.T_cur;
            longest_trusted(sc.T_cur, len, a0, a1);
            bit_set(sc.E, base + idx);
            if (a0 == 0 && a1 == KBBQ_NPOS) {  // ref quirk: returns with the corrected base still in seq
                if (top) { res.mod_idx = idx; res.mod_base = fixed_base; }
                return;
            }
        } else if (a0 == 0 && a1 == KBBQ_NPOS) {
            return;  // every k-mer trusted (readutils.cc:263-265)
        }
        const int anchor_len = (a1 == KBBQ_NPOS ? len - 1 : a1) + 1 - a0;
        bool corrected = false;
        // ---- the two sides of the anchor: right on `cur` (readutils.cc:271-346), then left on its reverse complement
        //      (348-422).  In the reverse complement's own coordinates the left walk is the same loop: the base under
        //      repair sits at j, the next one at start + fixlen; only the flag index (len-1-j) is mirrored.
        SeqRef rc{sc.C_w, sc.C_nm, len};
        KBBQ_NOUNROLL
        for (int side = 0; side < 2; ++side) {
            int anchor;
            if (side == 0) {
                if (a1 == KBBQ_NPOS) continue;
                anchor = a1;
            } else {
                if (a0 == 0) continue;
                revcomp(cur, rc);  // of the read WITH its right-side fixes
                T_mirrored = true;
                anchor = len - a0 - 1;
            }
            const SeqRef s = side ? rc : cur;  // by value: a reference to one of two structs would pin both in local memory
            if (anchor_len - k + 1 >= k) {  // at least k trusted k-mers: the anchor's tail may be false positives
                bool cm;
                anchor = adjust_right_anchor(anchor, s, cm);  // ref quirk: ACGT order on both sides (readutils.cc:374)
                multiple = multiple || cm;
            }
            int bad = KBBQ_NPOS;
            KBBQ_NOUNROLL
            for (int j = anchor + 1; j < len;) {
                const int start = j - k + 1;
                int fix0, fixlen;
                bool cm;
                const int nfix = longest_fix(s, start, side != 0, fix0, fixlen, cm);
                multiple = multiple || cm;
                const int next_untrusted = start + fixlen;
                if (next_untrusted > j) {
                    if (nfix > 1) {
                        multiple = true;
                        const int largest = j + k - 1 < len - 1 ? j + k - 1 : len - 1;
                        if (next_untrusted <= largest || largest - j + 1 < k) { bad = j; break; }
                    } else {
                        seq_set(s, j, fix0);  // ref quirk: left-side fixes land in the reverse complement only
                        memo_reset();
                        bit_set(sc.E, base + (side ? len - 1 - j : j));
                    }
                    corrected = true;
                    j = next_untrusted;
                } else {
                    bad = j;
                    break;
                }
            }
            if (bad != KBBQ_NPOS) {
                if (side == 0) bad_suffix = bad;
                else bad_prefix = len - 1 - bad;
            }
        }
        // ---- overcorrection check (readutils.cc:429-546)
        if (corrected) {
            bool adjust = true;
            {   // no flagged base may lie under a streak of k-mers that were trusted in the ORIGINAL read;
                // ref quirk: the streak that runs to the end of the read is never examined (455-465)
                const int nk = len - k + 1;
                int pos = 0;
                KBBQ_NOUNROLL
                while (adjust) {
                    const int s = next_one(sc.T_orig, nk, pos);
                    if (s >= nk) break;
                    const int e = next_zero(sc.T_orig, nk, s);
                    if (e >= nk) break;
                    if (count_range(sc.E, base + s, base + e - 1 + k) != 0) adjust = false;
                    pos = e;
                }
            }
            adjust = adjust && !multiple;
            // sliding 20-base window over the flagged bases whose ORIGINAL base was ACGT: weight 1, or 1/2 when
            // qual <= minqual(6); flagged position i is overcorrected when the window ending at i weighs more than
            // the threshold (4, or 5 in the interior when `adjust`).  Counted in half units, 32 positions per step.
            const int ocwindow = 20, base_threshold = 4;
            uint32_t* OC = sc.T_cur;  // T_cur is dead from here on: reuse it for the overcorrected-index list
            bool any_oc = false;
            // F = flagged bases that count (original base ACGT), FQ = those of them with the half weight, both in the
            // call's own coordinates: a window's weight is two popcounts instead of a 20-step loop per base
            uint32_t* F = sc.F;
            uint32_t* FQ = sc.FQ;
            q.par(nw + 1, [&](int w) {
                uint32_t e = 0, eq = 0;
                if (w < nw) {
                    const int cnt = len - 32 * w < 32 ? len - 32 * w : 32;
                    e = extract_n(sc.E, base + 32 * w, cnt) & ~sc.O_nm[w];
                    eq = e & extract_n(sc.QL, base + 32 * w, cnt);
                }
                F[w] = e;
                FQ[w] = eq;
            });
            KBBQ_NOUNROLL
            for (int w = 0; w < nw; ++w) {
                uint32_t m = 0;
                if (extract_n(sc.E, base + 32 * w, len - 32 * w < 32 ? len - 32 * w : 32)) {  // only flagged bases can be overcorrected
                    m = q.map32([&](int t) {
                        const int i = 32 * w + t;
                        if (i >= len || !bit_get(sc.E, base + i)) return false;
                        const int lo = i - ocwindow + 1 > 0 ? i - ocwindow + 1 : 0;
                        const int occ2 = 2 * count_window32(F, lo, i + 1) - count_window32(FQ, lo, i + 1);
                        const int threshold = adjust && i >= ocwindow && i + ocwindow - 1 < len ? base_threshold + 1 : base_threshold;
                        return occ2 > 2 * threshold;
                    });
                }
                OC[w] = m;
                any_oc = any_oc || m != 0;
            }
            OC[nw] = 0;
            if (any_oc) {  // un-flag everything within k of an overcorrected base, transitively (521-545)
                KBBQ_NOUNROLL
                for (int oc = next_one(OC, len, 0); oc < len; oc = next_one(OC, len, oc + 1)) {
                    if (!bit_get(sc.E, base + oc)) continue;
                    int start = oc - k + 1;
                    start = start >= 0 ? start : 0;
                    int end = oc + k;
                    end = end < len ? end : len;
                    KBBQ_NOUNROLL
                    for (int i = start; i < end; ++i) {
                        if (bit_get(sc.E, base + i)) {
                            bit_clr(sc.E, base + i);
                            if (i + k > end) end = i + k < len ? i + k : len;
                            if (i - k < start) {
                                i = i - k + 1 >= 0 ? i - k : -1;
                                start = i;
                            }
                        }
                    }
                }
            }
        }
    }

    // Top level: get_errors(first_call=true) including the one-level recursion on a long bad prefix / suffix
    // (readutils.cc:547-563).  sc.P_* must hold the read, sc.QL its low-quality flags, sc.E must be zero.
    // On return sc.E holds read.errors.
    // `prescanned`: see call(); applies to the whole read (job 0) only -- the sub-reads of the recursion are cut from
    // the partly corrected read and scanned here.
    KBBQ_HD ErrResult run(int len, bool prescanned = false) {
        ErrResult res;
        res.mod_idx = -1;
        res.mod_base = -1;
        if (len < k) return res;
        SeqRef P{sc.P_w, sc.P_nm, len};
        SeqRef A{sc.A_w, sc.A_nm, 0};
        int top_prefix = 0, top_suffix = KBBQ_NPOS;
        KBBQ_NOUNROLL
        for (int job = 0; job < 3; ++job) {  // 0: the read; 1: substr(0, bad_prefix+1); 2: substr(bad_suffix)
            int from = 0;
            if (job == 1) {
                if (!(top_prefix > 0 && (top_prefix >= len / 2 || top_prefix >= 2 * k))) continue;
                A.len = top_prefix + 1;
            } else if (job == 2) {
                if (!(top_suffix != KBBQ_NPOS && top_suffix < len && (len - top_suffix > len / 2 || len - top_suffix > 2 * k))) continue;
                from = top_suffix;
                A.len = len - top_suffix;
            }
            if (job) {  // the sub-read is cut from the partly corrected read and inherits its flags (readutils.cc:597-608)
                const int nw = (A.len + 31) >> 5;
                KBBQ_NOUNROLL
                for (int w = 0; w < nw; ++w) {
                    const int cnt = A.len - 32 * w < 32 ? A.len - 32 * w : 32;
                    A.w[w] = extract_le(P.w, from + 32 * w, cnt);
                    A.nm[w] = extract_n(P.nm, from + 32 * w, cnt);
                }
                A.w[nw] = 0;
                A.nm[nw] = 0;
            }
            int bp, bs;
            ErrResult one;  // by value, like the sequence: references picked at run time would pin their targets in local memory
            one.mod_idx = -1;
            one.mod_base = -1;
            const SeqRef cur = job ? A : P;
            call(cur, from, job == 0, bp, bs, one, prescanned && job == 0);
            if (job == 0) { top_prefix = bp; top_suffix = bs; res = one; }
        }
        return res;
    }
};

}  // namespace kbbq
