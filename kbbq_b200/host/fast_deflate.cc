// fast_deflate.cc -- see fast_deflate.hh.  Written from RFC 1951; no code from zlib, libdeflate or htslib.
#include "fast_deflate.hh"

#include <algorithm>
#include <cstring>

namespace kbbq {
namespace fdeflate {
namespace {

const int kNumLitLen = 286, kNumDist = 30, kNumClen = 19, kEob = 256;
const int kMaxMatch = 258, kWindow = 32768, kHashBits = 13;

inline uint32_t load32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline uint64_t load64(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }

// ---- RFC 1951 §3.2.5: length and distance symbols -------------------------------------------------------------------
struct Tables {
    uint8_t len_sym[256];     // match length - 3 -> symbol - 257
    uint8_t len_extra[29];
    uint16_t len_base[29];
    uint8_t dist_sym[512];    // zlib-style two-level lookup: d-1 < 256 ? [d-1] : [256 + ((d-1) >> 7)]
    uint8_t dist_extra[30];
    uint16_t dist_base[30];
    Tables() {
        int len = 3;
        for (int s = 0; s < 28; ++s) {
            len_extra[s] = (uint8_t)(s < 8 ? 0 : (s - 4) / 4);
            len_base[s] = (uint16_t)len;
            for (int i = 0; i < (1 << len_extra[s]); ++i, ++len)
                if (len <= 257) len_sym[len - 3] = (uint8_t)s;
        }
        len_extra[28] = 0;  // 258 has a symbol of its own
        len_base[28] = 258;
        len_sym[255] = 28;
        int d = 1;
        for (int s = 0; s < 30; ++s) {
            dist_extra[s] = (uint8_t)(s < 4 ? 0 : (s - 2) / 2);
            dist_base[s] = (uint16_t)d;
            for (int i = 0; i < (1 << dist_extra[s]); ++i, ++d) {
                if (d - 1 < 256) dist_sym[d - 1] = (uint8_t)s;
                else if (((d - 1) & 127) == 0) dist_sym[256 + ((d - 1) >> 7)] = (uint8_t)s;
            }
        }
    }
    int dsym(int dist) const { return dist <= 256 ? dist_sym[dist - 1] : dist_sym[256 + ((dist - 1) >> 7)]; }
};
const Tables T;

// ---- length-limited Huffman code lengths ----------------------------------------------------------------------------
// Plain Huffman depths by the two-queue construction over the sorted used symbols, then the textbook repair of the
// length counts (move the too-deep leaves up to the limit and pay for it by pushing shallower leaves down until the
// Kraft sum is 1 again), lengths handed out in order of frequency.
void code_lengths(const uint32_t* freq, int n, int limit, uint8_t* len) {
    int order[kNumLitLen];
    int m = 0;
    for (int s = 0; s < n; ++s) {
        len[s] = 0;
        if (freq[s]) order[m++] = s;
    }
    if (m == 0) return;
    if (m == 1) {  // one bit for it and one for a symbol that never occurs: inflaters want a complete code
        len[order[0]] = 1;
        len[order[0] == 0 ? 1 : 0] = 1;
        return;
    }
    std::sort(order, order + m, [&](int a, int b) { return freq[a] != freq[b] ? freq[a] < freq[b] : a < b; });
    uint64_t w[2 * kNumLitLen];
    int parent[2 * kNumLitLen];
    for (int i = 0; i < m; ++i) w[i] = freq[order[i]];
    int leaf = 0, inner = m, made = m;
    while (made < 2 * m - 1) {
        int pick[2];
        for (int t = 0; t < 2; ++t) {
            if (leaf < m && (inner >= made || w[leaf] <= w[inner])) pick[t] = leaf++;
            else pick[t] = inner++;
        }
        w[made] = w[pick[0]] + w[pick[1]];
        parent[pick[0]] = parent[pick[1]] = made;
        ++made;
    }
    int count[64] = {0};
    int depth[2 * kNumLitLen];
    depth[2 * m - 2] = 0;
    for (int i = 2 * m - 3; i >= 0; --i) {
        depth[i] = depth[parent[i]] + 1;
        if (i < m) ++count[depth[i] < 63 ? depth[i] : 63];
    }
    for (int d = limit + 1; d < 64; ++d) { count[limit] += count[d]; count[d] = 0; }
    uint64_t kraft = 0;
    for (int d = 1; d <= limit; ++d) kraft += (uint64_t)count[d] << (limit - d);
    while (kraft > (1ull << limit)) {  // one leaf at the limit pairs up with a leaf pushed down from above it
        --count[limit];
        for (int d = limit - 1; d >= 1; --d)
            if (count[d]) { --count[d]; count[d + 1] += 2; break; }
        --kraft;
    }
    int i = 0;  // rarest symbols first: they take the longest codes
    for (int d = limit; d >= 1; --d)
        for (int c = 0; c < count[d]; ++c) len[order[i++]] = (uint8_t)d;
}

inline uint32_t reverse_bits(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}
// canonical codes (RFC 1951 §3.2.2), stored bit-reversed: DEFLATE sends Huffman codes most significant bit first
void canonical_codes(const uint8_t* len, int n, int limit, uint16_t* code) {
    uint32_t count[17] = {0}, next[17] = {0};
    for (int s = 0; s < n; ++s) ++count[len[s]];
    count[0] = 0;
    for (int d = 1; d <= limit; ++d) next[d] = (next[d - 1] + count[d - 1]) << 1;
    for (int s = 0; s < n; ++s) code[s] = len[s] ? (uint16_t)reverse_bits(next[len[s]]++, len[s]) : 0;
}

// ---- bit output: least significant bit first, flushed a byte multiple at a time -------------------------------------
struct BitWriter {
    uint8_t* p;
    uint64_t acc = 0;
    int n = 0;
    explicit BitWriter(uint8_t* out) : p(out) {}
    inline void add(uint32_t v, int bits) { acc |= (uint64_t)v << n; n += bits; }
    inline void flush() {  // the caller keeps n below 64 between flushes; eight bytes of slack past the stream's end
        std::memcpy(p, &acc, 8);
        p += n >> 3;
        acc >>= n & ~7;
        n &= 7;
    }
    uint8_t* finish() {
        flush();
        if (n) { *p++ = (uint8_t)acc; }
        return p;
    }
};

struct Seq {  // a run of literals (taken from the input itself) followed by one match (len 0: none, the block's tail)
    uint16_t lits, len, dist;
};

inline void count_literals(const uint8_t* p, size_t n, uint32_t (*h)[256]) {
    size_t i = 0;
    for (; i + 4 <= n; i += 4) {  // four tables: runs of one symbol would otherwise serialise on one counter
        ++h[0][p[i]];
        ++h[1][p[i + 1]];
        ++h[2][p[i + 2]];
        ++h[3][p[i + 3]];
    }
    for (; i < n; ++i) ++h[0][p[i]];
}

size_t stored(const uint8_t* in, size_t n, uint8_t* out) {
    out[0] = 1;  // BFINAL = 1, BTYPE = 00, padding
    out[1] = (uint8_t)(n & 0xff);
    out[2] = (uint8_t)(n >> 8);
    out[3] = (uint8_t)(~n & 0xff);
    out[4] = (uint8_t)((~n >> 8) & 0xff);
    std::memcpy(out + 5, in, n);
    return n + 5;
}

}  // namespace

size_t compress(const uint8_t* in, size_t n, uint8_t* out, size_t cap, int min_match) {
    if (n > kMaxInput || cap < bound(n)) return 0;
    if (n == 0) {  // fixed-Huffman block holding only the end-of-block symbol
        out[0] = 3;
        out[1] = 0;
        return 2;
    }
    // ---- pass 1: sequences and symbol statistics
    static thread_local Seq seqs[kMaxInput / 4 + 2];
    static thread_local uint16_t head[1 << kHashBits];
    uint32_t lit_hist[4][256];
    uint32_t lfreq[kNumLitLen] = {0}, dfreq[kNumDist] = {0};
    std::memset(lit_hist, 0, sizeof lit_hist);
    size_t nseq = 0, lit_start = 0;
    if (min_match < 8) min_match = 8;
    if (min_match <= kMaxMatch && n >= 32) {
        // Candidates are found by hashing EIGHT bytes: only matches worth keeping ever reach the verification, however
        // repetitive the text (quality lines are full of 4-byte repeats).  Misses lengthen the stride up to 8, so that
        // stretches of incompressible bases are probed sparsely; a match found late is grown backwards to its start.
        std::memset(head, 0, sizeof head);
        const size_t last = n - 8;  // last position hashed (keeps every 8-byte load inside the input)
        size_t i = 0, misses = 0;
        while (i <= last) {
            const uint64_t v = load64(in + i);
            const uint32_t h = (uint32_t)((v * 0x9E3779B185EBCA87ull) >> (64 - kHashBits));
            size_t cand = head[h];
            head[h] = (uint16_t)i;
            if (cand < i && i - cand <= (size_t)kWindow && load64(in + cand) == v) {
                size_t at = i;
                while (at > lit_start && cand > 0 && in[at - 1] == in[cand - 1]) { --at; --cand; }
                const size_t room = std::min<size_t>(kMaxMatch, n - at);
                size_t len = i + 8 - at;
                if (len > room) len = room;
                while (len + 8 <= room) {
                    const uint64_t x = load64(in + cand + len) ^ load64(in + at + len);
                    if (x) { len += (size_t)(__builtin_ctzll(x) >> 3); goto extended; }
                    len += 8;
                }
                while (len < room && in[cand + len] == in[at + len]) ++len;
            extended:
                if (len >= (size_t)min_match) {
                    count_literals(in + lit_start, at - lit_start, lit_hist);
                    seqs[nseq++] = Seq{(uint16_t)(at - lit_start), (uint16_t)len, (uint16_t)(at - cand)};
                    ++lfreq[257 + T.len_sym[len - 3]];
                    ++dfreq[T.dsym((int)(at - cand))];
                    i = at + len;
                    lit_start = i;
                    misses = 0;
                    continue;
                }
            }
            i += 1 + std::min<size_t>(misses++ >> 4, 7);
        }
    }
    count_literals(in + lit_start, n - lit_start, lit_hist);
    seqs[nseq++] = Seq{(uint16_t)(n - lit_start), 0, 0};
    for (int s = 0; s < 256; ++s) lfreq[s] = lit_hist[0][s] + lit_hist[1][s] + lit_hist[2][s] + lit_hist[3][s];
    lfreq[kEob] = 1;

    // ---- the block's codes
    uint8_t llen[kNumLitLen + kNumDist];
    uint8_t* dlen = llen + kNumLitLen;
    uint16_t lcode[kNumLitLen], dcode[kNumDist];
    code_lengths(lfreq, kNumLitLen, 15, llen);
    code_lengths(dfreq, kNumDist, 15, dlen);
    int nlit = kNumLitLen, ndist = kNumDist;
    while (nlit > 257 && llen[nlit - 1] == 0) --nlit;
    while (ndist > 1 && dlen[ndist - 1] == 0) --ndist;
    if (ndist == 1 && dlen[0] == 0) dlen[0] = 1;  // no match at all: one unused distance code of one bit
    canonical_codes(llen, nlit, 15, lcode);
    canonical_codes(dlen, ndist, 15, dcode);
    // the two length arrays as one sequence, run-length coded with symbols 16 / 17 / 18 (RFC 1951 §3.2.7)
    uint8_t lens[kNumLitLen + kNumDist];
    std::memcpy(lens, llen, nlit);
    std::memcpy(lens + nlit, dlen, ndist);
    const int nlens = nlit + ndist;
    struct Rle { uint8_t sym, extra; };
    Rle rle[kNumLitLen + kNumDist];
    int nrle = 0;
    uint32_t cfreq[kNumClen] = {0};
    for (int i = 0; i < nlens;) {
        int run = 1;
        while (i + run < nlens && lens[i + run] == lens[i]) ++run;
        const int v = lens[i];
        i += run;
        if (v == 0) {
            while (run >= 11) { const int r = std::min(run, 138); rle[nrle++] = Rle{18, (uint8_t)(r - 11)}; ++cfreq[18]; run -= r; }
            if (run >= 3) { rle[nrle++] = Rle{17, (uint8_t)(run - 3)}; ++cfreq[17]; run = 0; }
        } else {
            rle[nrle++] = Rle{(uint8_t)v, 0};
            ++cfreq[v];
            --run;
            while (run >= 3) { const int r = std::min(run, 6); rle[nrle++] = Rle{16, (uint8_t)(r - 3)}; ++cfreq[16]; run -= r; }
        }
        for (; run > 0; --run) { rle[nrle++] = Rle{(uint8_t)v, 0}; ++cfreq[v]; }
    }
    uint8_t clen[kNumClen];
    uint16_t ccode[kNumClen];
    code_lengths(cfreq, kNumClen, 7, clen);
    canonical_codes(clen, kNumClen, 7, ccode);
    static const uint8_t clen_order[kNumClen] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int nclen = kNumClen;
    while (nclen > 4 && clen[clen_order[nclen - 1]] == 0) --nclen;

    // ---- does it pay?  exact size of the dynamic block in bits
    uint64_t bits = 3 + 5 + 5 + 4 + 3 * (uint64_t)nclen;
    for (int i = 0; i < nrle; ++i) bits += clen[rle[i].sym] + (rle[i].sym == 16 ? 2 : rle[i].sym == 17 ? 3 : rle[i].sym == 18 ? 7 : 0);
    for (int s = 0; s < nlit; ++s) bits += (uint64_t)lfreq[s] * llen[s] + (s > 256 ? (uint64_t)lfreq[s] * T.len_extra[s - 257] : 0);
    for (int s = 0; s < ndist; ++s) bits += (uint64_t)dfreq[s] * (dlen[s] + T.dist_extra[s]);
    if ((bits + 7) / 8 >= n + 5) return stored(in, n, out);

    // ---- pass 2: emit
    BitWriter bw(out);
    bw.add(1, 1);  // BFINAL
    bw.add(2, 2);  // BTYPE = dynamic Huffman
    bw.add((uint32_t)(nlit - 257), 5);
    bw.add((uint32_t)(ndist - 1), 5);
    bw.add((uint32_t)(nclen - 4), 4);
    bw.flush();
    for (int i = 0; i < nclen; ++i) {
        bw.add(clen[clen_order[i]], 3);
        if ((i & 7) == 7) bw.flush();
    }
    bw.flush();
    for (int i = 0; i < nrle; ++i) {
        const int s = rle[i].sym;
        bw.add(ccode[s], clen[s]);
        if (s >= 16) bw.add(rle[i].extra, s == 16 ? 2 : s == 17 ? 3 : 7);
        bw.flush();
    }
    uint32_t lit_pack[256];  // code | length << 16
    for (int s = 0; s < 256; ++s) lit_pack[s] = lcode[s] | ((uint32_t)llen[s] << 16);
    const uint8_t* p = in;
    for (size_t q = 0; q < nseq; ++q) {
        size_t lits = seqs[q].lits;
        for (; lits >= 3; lits -= 3, p += 3) {  // three codes of at most 15 bits on top of at most 7 pending bits
            const uint32_t a = lit_pack[p[0]], b = lit_pack[p[1]], c = lit_pack[p[2]];
            bw.add(a & 0xffffu, (int)(a >> 16));
            bw.add(b & 0xffffu, (int)(b >> 16));
            bw.add(c & 0xffffu, (int)(c >> 16));
            bw.flush();
        }
        for (; lits; --lits, ++p) {
            const uint32_t a = lit_pack[*p];
            bw.add(a & 0xffffu, (int)(a >> 16));
        }
        bw.flush();
        const int len = seqs[q].len;
        if (len) {  // 15 + 5 + 15 + 13 bits at most
            const int ls = T.len_sym[len - 3], dist = seqs[q].dist, ds = T.dsym(dist);
            bw.add(lcode[257 + ls], llen[257 + ls]);
            bw.add((uint32_t)(len - T.len_base[ls]), T.len_extra[ls]);
            bw.add(dcode[ds], dlen[ds]);
            bw.add((uint32_t)(dist - T.dist_base[ds]), T.dist_extra[ds]);
            bw.flush();
            p += len;
        }
    }
    bw.add(lcode[kEob], llen[kEob]);
    return (size_t)(bw.finish() - out);
}

}  // namespace fdeflate
}  // namespace kbbq
