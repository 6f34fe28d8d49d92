// kbbq_errcall.cuh -- per-read error calling against the trusted filter (pass 3), one warp per read.
//
// This is CReadData::get_errors (readutils.cc:238-570) and its helpers (bloom.cc:83-305, readutils.cc:195-235)
// re-designed for a warp: the read lives in shared memory as 2-bit words, every question the reference asks of the
// filter one k-mer at a time is asked for up to 32 k-mers at once through `Q::query` (one lane per k-mer, one
// ballot), and all the bookkeeping (anchors, streaks, windows, error flags) is bit-mask arithmetic that every lane
// carries out identically, so the control flow is warp-uniform and the only shared-memory writes are idempotent.
//
// `Q` supplies   template<class F> uint32_t query(int count, F f)
// which evaluates f(t, key) -> bool valid for t in [0,count), count <= 32, and returns the mask of t for which the
// k-mer is valid AND present in the trusted filter (Bloom::query, bloom.hh:398).  On the device that is one lane per
// t and a __ballot_sync; tests/emu provides a serial stand-in so the logic can be exercised without a GPU.
//
// Conventions: positions are int; NPOS (-1) stands for std::string::npos where the reference uses it as a flag.
// Deliberate reproduction of reference behaviour is marked "ref quirk".
#pragma once
#include "kbbq_core.cuh"

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
};

struct ErrResult {
    int mod_idx;   // ref quirk: the all-trusted early return after correct_one leaves the corrected base in seq
    int mod_base;  // (readutils.cc:263-265 vs 568); -1 when seq ends up equal to the input
};

template <class Q>
struct ErrCaller {
    Q& q;
    const int k;
    const uint8_t* qual;  // qualities of the top-level read
    ErrScratch& sc;
    KBBQ_HD ErrCaller(Q& q_, int k_, const uint8_t* qual_, ErrScratch& sc_) : q(q_), k(k_), qual(qual_), sc(sc_) {}

    // mask of trusted k-mers among those starting at j0 .. j0+count-1 (count <= 32) of s with base p replaced by c
    KBBQ_HD uint32_t trusted_mask(const SeqRef& s, int j0, int count, int p, int c) {
        const int kk = k;
        return q.query(count, [&](int t, uint64_t& key) { return kmer_key(s, j0 + t, kk, p, c, key); });
    }
    // number of consecutive trusted k-mers starting at j0, looking no further than j1 (exclusive)
    KBBQ_HD int trusted_run(const SeqRef& s, int j0, int j1, int p, int c) {
        int j = j0;
        while (j < j1) {
            const int count = j1 - j < 32 ? j1 - j : 32;
            const uint32_t m = trusted_mask(s, j, count, p, c);
            const uint32_t full = count < 32 ? ((1u << count) - 1) : ~0u;
            if (m != full) return j - j0 + ctz32(~m);
            j += count;
        }
        return j1 - j0;
    }
    // fill T[0 .. nk) with the trusted flag of every k-mer of s
    KBBQ_HD void scan(const SeqRef& s, uint32_t* T) {
        const int nk = s.len - k + 1;
        for (int j = 0; j < nk; j += 32) {
            const int count = nk - j < 32 ? nk - j : 32;
            T[j >> 5] = trusted_mask(s, j, count, 0, -1);
        }
    }

    // bloom::find_longest_trusted_seq (bloom.cc:96-128) on the k-mer flags: first longest streak wins
    KBBQ_HD void longest_trusted(const uint32_t* T, int len, int& a0, int& a1) {
        const int nk = len - k + 1;
        int best = 0, bs = 0, be = 0, pos = 0;
        while (true) {
            const int s = next_one(T, nk, pos);
            if (s >= nk) break;
            const int e = next_zero(T, nk, s);
            if (e - s > best) { best = e - s; bs = s; be = e; }
            pos = e;
        }
        if (best == 0) { a0 = KBBQ_NPOS; a1 = KBBQ_NPOS; return; }
        a0 = bs;
        a1 = be == nk ? KBBQ_NPOS : be - 1 + k - 1;
    }

    // bloom::find_longest_fix (bloom.cc:130-188) for the substring of s starting at `start`; the base under repair is
    // p = start + k - 1.  Returns the number of best candidates; fix0 = the first of them, fixlen = best_i.
    KBBQ_HD int longest_fix(const SeqRef& s, int start, bool reverse_order, int& fix0, int& fixlen, bool& multiple) {
        const int p = start + k - 1;
        const int sublen = s.len - start;
        const int nsub = sublen - k + 1;                          // k-mers of the substring that lie inside the read
        const int i_stop = 2 * k - 1 > sublen ? 2 * k - 1 : sublen;  // bloom.cc:147
        int nbest = 0, best_i = 0, first_ok = 0;
        fix0 = 0;
        for (int ci = 0; ci < 4; ++ci) {
            const int c = reverse_order ? 3 - ci : ci;  // ACGT or TGCA (bloom.cc:140-141)
            if (seq_eq(s, p, c)) continue;
            int run = trusted_run(s, start, start + nsub, p, c);
            int i = k - 1 + run;
            if (run >= 1) ++first_ok;  // `single`/`multiple` bookkeeping, bloom.cc:163-166
            if (run == nsub && sublen < i_stop) {
                // walked off the read end with every k-mer trusted: extend with guessed bases (bloom.cc:83-94,149)
                uint64_t v = extract_le(s.w, s.len - k, k);
                if (p >= s.len - k) {
                    const int o = p - (s.len - k);
                    v = (v & ~(3ull << (2 * o))) | ((uint64_t)c << (2 * o));
                }
                const int kk = k;
                for (i = sublen; i < i_stop; ++i) {
                    const uint64_t shifted = v >> 2;
                    const uint32_t m = q.query(4, [&](int t, uint64_t& key) {
                        const int cc = reverse_order ? 3 - t : t;
                        key = canonical(shifted | ((uint64_t)cc << (2 * (kk - 1))), kk);
                        return true;
                    });
                    if (m == 0) break;  // no trusted next k-mer
                    const int t = ctz32(m);
                    v = shifted | ((uint64_t)(reverse_order ? 3 - t : t) << (2 * (k - 1)));
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
        int m = anchor + 1;
        {   // does a single substitution at anchor+1 repair everything it touches?  then the anchor stands
            int cnt = s.len - m;  // queries possible before the read ends: i = 0 .. min(k, len-1-m)
            if (cnt > k + 1) cnt = k + 1;
            for (int c = 0; c < 4; ++c) {
                if (seq_eq(s, m, c)) continue;
                if (trusted_run(s, m - k + 1, m - k + 1 + cnt, m, c) == cnt) return anchor;
            }
        }
        // otherwise try pulling the anchor back by up to k/2
        for (int i = k / 2 - 1; i >= 0 && anchor > i + k - 1; --i) {
            m = anchor - i;
            const int need = k / 2 + 2;  // k-mers ending at m .. m+k/2+1 must all be trusted
            for (int c = 0; c < 4; ++c) {
                if (seq_eq(s, m, c)) continue;
                int j1 = m - k + 1 + need;
                const int jmax = s.len - k + 1;
                const bool fits = j1 <= jmax;
                if (!fits) j1 = jmax;
                const int run = trusted_run(s, m - k + 1, j1, m, c);
                if (run >= 1) multiple = true;
                if (fits && run == need) return m - 1;
            }
        }
        return anchor;
    }

    // CReadData::correct_one (readutils.cc:195-235) + bloom::biggest_consecutive_trusted_block (bloom.cc:279-305)
    KBBQ_HD int correct_one(const SeqRef& s, int qbase, int& out_base) {
        const int len = s.len;
        int best_len = 0, best_base = 0, best_pos = KBBQ_NPOS;
        const int ncand = 4 * len;  // candidate id = 4*i + c; identical bases are skipped like `continue`
        const int kk = k;
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
            while (hits) {
                const int t = ctz32(hits);
                hits &= hits - 1;
                const int i = (c0 + t) >> 2, c = (c0 + t) & 3;
                const int start = i > k - 1 ? i - k + 1 : 0;
                int nwin = len - k + 1 - start;  // k-mers of substr(start, 2k-1)
                if (nwin > k) nwin = k;
                uint32_t m = trusted_mask(s, start, nwin, i, c);
                // the walk stops once k - (#misses) < best_len (bloom.cc:295): cut the mask at that miss
                if (best_len >= 1) {
                    const int stop_after = k - best_len + 1;  // this many misses end the walk
                    uint32_t z = ~m & (nwin < 32 ? ((1u << nwin) - 1) : ~0u);
                    int seen = 0;
                    while (z) {
                        const int b = ctz32(z);
                        z &= z - 1;
                        if (++seen == stop_after) { m &= (1u << b) - 1; break; }
                    }
                }
                int n_in = 0;
                while (m) {  // longest streak of ones
                    m &= m << 1;
                    ++n_in;
                }
                if (n_in > best_len) { best_base = c; best_pos = i; best_len = n_in; }
                else if (n_in == best_len && qual[qbase + i] < qual[qbase + best_pos]) { best_base = c; best_pos = i; }
            }
        }
        out_base = best_base;
        if (best_len > 0) seq_set(s, best_pos, best_base);
        return best_pos;
    }

    // reverse complement of src into dst (readutils.cc:351-353): N stays N
    KBBQ_HD void revcomp(const SeqRef& src, const SeqRef& dst) {
        const int len = src.len;
        const int nw = (len + 31) >> 5;
        for (int wd = 0; wd < nw; ++wd) {
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
        }
        dst.w[nw] = 0;
        dst.nm[nw] = 0;
    }

    // One invocation of get_errors on `cur` (length cur.len), whose base 0 is base `base` of the top-level read.
    // Error flags live in sc.E at offset `base`.  Returns bad_prefix / bad_suffix for the caller's recursion test.
    KBBQ_HD void call(const SeqRef& cur, int base, bool top, int& bad_prefix, int& bad_suffix, ErrResult& res) {
        const int len = cur.len;
        bad_prefix = 0;
        bad_suffix = KBBQ_NPOS;
        if (len < k) return;  // the reference is undefined here (bloom.cc:30 / readutils.cc:207 underflow)
        const int nw = (len + 31) >> 5;
        for (int w = 0; w <= nw; ++w) sc.O_nm[w] = cur.nm[w];
        bool multiple = false;
        scan(cur, sc.T_orig);
        int a0, a1;
        longest_trusted(sc.T_orig, len, a0, a1);
        if (a0 == KBBQ_NPOS) {  // no trusted k-mer in the read (readutils.cc:250-262)
            multiple = true;
            int fixed_base;
            const int idx = correct_one(cur, base, fixed_base);
            if (idx == KBBQ_NPOS) return;
            scan(cur, sc.T_cur);
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
        // ---- right of the anchor (readutils.cc:271-346)
        if (a1 != KBBQ_NPOS) {
            if (anchor_len - k + 1 >= k) {
                bool cm;
                a1 = adjust_right_anchor(a1, cur, cm);
                multiple = multiple || cm;
            }
            for (int i = a1 + 1; i < len;) {
                const int start = i - k + 1;
                int fix0, fixlen;
                bool cm;
                const int nfix = longest_fix(cur, start, false, fix0, fixlen, cm);
                multiple = multiple || cm;
                const int next_untrusted = start + fixlen;
                if (next_untrusted > i) {
                    if (nfix > 1) {
                        multiple = true;
                        const int largest = i + k - 1 < len - 1 ? i + k - 1 : len - 1;
                        if (next_untrusted <= largest || largest - i + 1 < k) { bad_suffix = i; break; }
                    } else {
                        seq_set(cur, i, fix0);
                        bit_set(sc.E, base + i);
                    }
                    corrected = true;
                    i += fixlen - k + 1;
                } else {
                    bad_suffix = i;
                    break;
                }
            }
        }
        // ---- left of the anchor, on the reverse complement (readutils.cc:348-422)
        if (a0 != 0) {
            SeqRef rc{sc.C_w, sc.C_nm, len};
            revcomp(cur, rc);
            if (anchor_len - k + 1 >= k) {
                bool cm;
                const int left_adjust = adjust_right_anchor(len - a0 - 1, rc, cm);  // ref quirk: ACGT order here
                a0 = len - left_adjust - 1;
                multiple = multiple || cm;
            }
            for (int i = a0 - 1; i >= 0;) {
                const int j = len - i - 1;
                const int start = j - k + 1;
                int fix0, fixlen;
                bool cm;
                const int nfix = longest_fix(rc, start, true, fix0, fixlen, cm);
                multiple = multiple || cm;
                const int next_untrusted = start + fixlen;
                if (next_untrusted > j) {
                    if (nfix > 1) {
                        multiple = true;
                        const int largest = j + k - 1 < len - 1 ? j + k - 1 : len - 1;
                        if (next_untrusted <= largest || largest - j + 1 < k) { bad_prefix = i; break; }
                    } else {
                        seq_set(rc, j, fix0);  // ref quirk: left-side fixes never reach this->seq
                        bit_set(sc.E, base + i);
                    }
                    corrected = true;
                    i -= next_untrusted - j;
                } else {
                    bad_prefix = i;
                    break;
                }
            }
        }
        // ---- overcorrection check (readutils.cc:429-546)
        if (corrected) {
            bool adjust = true;
            {   // no flagged base may lie under a streak of k-mers that were trusted in the ORIGINAL read;
                // ref quirk: the streak that runs to the end of the read is never examined (455-465)
                const int nk = len - k + 1;
                int pos = 0;
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
            // sliding 20-base window, weight 1 per flagged base, 1/2 when qual <= minqual(6); counted in half units
            const int ocwindow = 20, base_threshold = 4;
            int occ2 = 0;
            bool any_oc = false;
            uint32_t* OC = sc.T_cur;  // T_cur is dead from here on: reuse it for the overcorrected-index list
            for (int w = 0; w <= nw; ++w) OC[w] = 0;
            for (int i = 0; i < len; ++i) {
                if (bit_get(sc.E, base + i) && !bit_get(sc.O_nm, i)) occ2 += qual[base + i] <= 6 ? 1 : 2;
                if (i >= ocwindow && bit_get(sc.E, base + i - ocwindow) && !bit_get(sc.O_nm, i - ocwindow))
                    occ2 -= qual[base + i - ocwindow] <= 6 ? 1 : 2;
                const int threshold = adjust && i >= ocwindow && i + ocwindow - 1 < len ? base_threshold + 1 : base_threshold;
                if (occ2 > 2 * threshold && bit_get(sc.E, base + i)) { bit_set(OC, i); any_oc = true; }
            }
            if (any_oc) {  // un-flag everything within k of an overcorrected base, transitively (521-545)
                for (int oc = next_one(OC, len, 0); oc < len; oc = next_one(OC, len, oc + 1)) {
                    if (!bit_get(sc.E, base + oc)) continue;
                    int start = oc - k + 1;
                    start = start >= 0 ? start : 0;
                    int end = oc + k;
                    end = end < len ? end : len;
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
    // (readutils.cc:547-563).  `P` must hold the read; sc.E must be zero.  On return sc.E holds read.errors.
    KBBQ_HD ErrResult run(int len) {
        ErrResult res;
        res.mod_idx = -1;
        res.mod_base = -1;
        SeqRef P{sc.P_w, sc.P_nm, len};
        int bad_prefix, bad_suffix;
        call(P, 0, true, bad_prefix, bad_suffix, res);
        if (len < k) return res;
        if (bad_prefix > 0 && (bad_prefix >= len / 2 || bad_prefix >= 2 * k)) {
            const int sl = bad_prefix + 1;  // substr(0, bad_prefix+1) of the partly corrected read
            SeqRef A{sc.A_w, sc.A_nm, sl};
            const int nw = (sl + 31) >> 5;
            for (int w = 0; w < nw; ++w) { A.w[w] = P.w[w]; A.nm[w] = P.nm[w]; }
            if (sl & 31) {
                A.w[nw - 1] &= kmask(sl & 31);
                A.nm[nw - 1] &= (1u << (sl & 31)) - 1;
            }
            A.w[nw] = 0;
            A.nm[nw] = 0;
            int bp, bs;
            ErrResult dummy;
            call(A, 0, false, bp, bs, dummy);
        }
        if (bad_suffix != KBBQ_NPOS && bad_suffix < len && (len - bad_suffix > len / 2 || len - bad_suffix > 2 * k)) {
            const int sl = len - bad_suffix;  // substr(bad_suffix)
            SeqRef A{sc.A_w, sc.A_nm, sl};
            const int nw = (sl + 31) >> 5;
            for (int w = 0; w < nw; ++w) {
                const int s0 = bad_suffix + 32 * w;
                const int cnt = sl - 32 * w < 32 ? sl - 32 * w : 32;
                A.w[w] = extract_le(P.w, s0, cnt);
                A.nm[w] = extract_n(P.nm, s0, cnt);
            }
            A.w[nw] = 0;
            A.nm[nw] = 0;
            int bp, bs;
            ErrResult dummy;
            call(A, bad_suffix, false, bp, bs, dummy);
        }
        return res;
    }
};

}  // namespace kbbq
