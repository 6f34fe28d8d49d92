// kbbq_synth.cuh -- synthetic reads of the shape SURVEY.md §8(d) / Appendix C describes, as a pure function of
// (seed, read index, base index): the same code fills a packed batch straight in HBM (synth_reads_kernel, one thread per
// 32-base word) and ASCII arrays on the host (kbbq_synth_host, for the oracle), so a benchmark at config-3/4 sizes
// needs neither a multi-GB input file nor minutes of host-side generation, and any slice of it can be handed to the
// CPU reference.  MEASUREMENT TOOLING: no pass of the hot path depends on it.
//
//   genome   i.i.d. uniform ACGT, base p = bits of mix(genome_seed + p/32)
//   read r   start uniform in [0, G-L], strand uniform (reverse strand = reverse complement), fixed length L,
//            second = r & 1, read group = (r / 2) % n_rg
//   quality  per base from {2, 11, 25, 37, 40} with p = {.01, .04, .10, .35, .50}
//   errors   substitution with probability err_scale * 10^(-q/10), alternative base uniform over the other three
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KBBQ_SYNTH_HD __host__ __device__ __forceinline__
#else
#define KBBQ_SYNTH_HD inline
#endif

namespace kbbq {

struct SynthSpec {
    uint64_t seed;         // read stream
    uint64_t genome_seed;  // genome (shared by all shards of a run)
    uint64_t genome_len;
    uint32_t read_len;
    uint32_t n_rg;
    uint32_t q_cum[5];     // cumulative class thresholds on a 32-bit uniform draw
    uint32_t err_thr[5];   // per class: error iff a second 32-bit uniform draw is below this
};

KBBQ_SYNTH_HD uint64_t synth_mix(uint64_t x) {  // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
KBBQ_SYNTH_HD uint64_t synth_mulhi(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
KBBQ_SYNTH_HD int synth_genome_base(const SynthSpec& s, uint64_t p) {
    return (int)((synth_mix(s.genome_seed * 0xD1342543DE82EF95ull + (p >> 5)) >> (2 * (p & 31))) & 3u);
}
struct SynthRead {
    uint64_t start, key;
    int reverse;
};
KBBQ_SYNTH_HD SynthRead synth_read(const SynthSpec& s, uint64_t r) {
    const uint64_t h = synth_mix(s.seed * 0x9FB21C651E98DF25ull + 2 * r);
    SynthRead rd;
    rd.start = synth_mulhi(h, s.genome_len - s.read_len + 1);
    rd.key = synth_mix(s.seed * 0x9FB21C651E98DF25ull + 2 * r + 1);
    rd.reverse = (int)(rd.key >> 63);
    return rd;
}
// base code (0-3) and phred quality of base i of read rd
KBBQ_SYNTH_HD void synth_base(const SynthSpec& s, const SynthRead& rd, uint32_t i, int& code, int& q) {
    const int g = rd.reverse ? 3 - synth_genome_base(s, rd.start + (s.read_len - 1 - i)) : synth_genome_base(s, rd.start + i);
    const uint64_t h = synth_mix(rd.key + ((uint64_t)i << 1));
    const uint32_t u = (uint32_t)h, e = (uint32_t)(h >> 32);
    const int cls = (u >= s.q_cum[0]) + (u >= s.q_cum[1]) + (u >= s.q_cum[2]) + (u >= s.q_cum[3]);
    q = cls == 0 ? 2 : cls == 1 ? 11 : cls == 2 ? 25 : cls == 3 ? 37 : 40;
    code = g;
    if (e < s.err_thr[cls]) code = (g + 1 + (int)(((h >> 8) & 0xFFFFu) * 3u >> 16)) & 3;
}

}  // namespace kbbq
