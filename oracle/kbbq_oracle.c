/*
 * oracle/kbbq_oracle.c -- TEST INFRASTRUCTURE ONLY (see kbbq_oracle.h).
 *
 * CPU restatement of the reference's hot path in plain C.  Every function
 * cites the reference lines it follows (paths relative to /root/reference).
 * The code favours literal fidelity over speed: size_t wrap-around idioms
 * (std::string::npos), int/size_t mixed comparisons and evaluation order are
 * kept as in the reference, because the outputs are compared bit for bit.
 *
 * Third-party arithmetic restated here (not under /root/reference):
 *  - libstdc++ 13 std::bernoulli_distribution / generate_canonical<double,53>,
 *    std::uniform_int_distribution<int> (Lemire "nearly divisionless", the
 *    128-bit path taken for a 64-bit URBG) and std::shuffle (two swap
 *    positions per draw) -- /usr/include/c++/13/bits/{random.tcc,
 *    uniform_int_dist.h,stl_algo.h};
 *  - htslib's seq_nt16_table / seq_nt16_int composition (only ACGT, acgt and
 *    '0'..'3' map below 4).
 *  - glibc libm long double functions are CALLED (same libm as the reference).
 */
#define _GNU_SOURCE
#include "kbbq_oracle.h"

#include <errno.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NPOS ((size_t)-1)

/* ---- a1: nucleotide code = seq_nt16_int[seq_nt16_table[ch]] (bloom.hh:351) ---- */
int ko_nt4(unsigned char ch) {
    switch (ch) {
    case 'A': case 'a': case '0': return 0;
    case 'C': case 'c': case '1': return 1;
    case 'G': case 'g': case '2': return 2;
    case 'T': case 't': case '3': return 3;
    default: return 4;
    }
}

/* seq_nt16_str[seq_nt16_table['0' + 3 - nt4(c)]] (readutils.cc:353): complement, N for anything else */
static char complement_char(char c) {
    int v = ko_nt4((unsigned char)c);
    return v < 4 ? "ACGT"[3 - v] : 'N';
}

/* ---- a1: bloom::Kmer (bloom.hh:339-379) ---- */
typedef struct {
    size_t s;
    int k;
    uint64_t x[2];
    uint64_t mask;
    uint64_t shift;
} kmer_t;

static void kmer_init(kmer_t *km, int k) {
    km->k = k;
    km->s = 0;
    km->mask = k < 32 ? (1ULL << (k * 2)) - 1 : (uint64_t)-1; /* bloom.hh:347 */
    km->shift = (uint64_t)(k - 1) * 2;
    km->x[0] = km->x[1] = 0;
}
static void kmer_reset(kmer_t *km) { km->s = 0; km->x[0] = km->x[1] = 0; }
static size_t kmer_push(kmer_t *km, char ch) { /* bloom.hh:350-358 */
    int c = ko_nt4((unsigned char)ch);
    if (c < 4) {
        km->x[0] = (km->x[0] << 2 | (uint64_t)c) & km->mask;
        km->x[1] = km->x[1] >> 2 | (uint64_t)(3 - c) << km->shift;
        ++km->s;
    } else {
        kmer_reset(km);
    }
    return km->s;
}
static uint64_t kmer_get(const kmer_t *km) { return km->x[0] < km->x[1] ? km->x[0] : km->x[1]; } /* bloom.hh:360 */
static int kmer_valid(const kmer_t *km) { return km->s >= (size_t)km->k; }                        /* bloom.hh:368 */

void ko_kmer_of(int k, const char *s, size_t n, uint64_t out[5]) {
    kmer_t km;
    kmer_init(&km, k);
    for (size_t i = 0; i < n; ++i) kmer_push(&km, s[i]);
    out[0] = km.x[0]; out[1] = km.x[1]; out[2] = kmer_get(&km); out[3] = km.s; out[4] = (uint64_t)kmer_valid(&km);
}

/* ---- a2: bloom_filter::hash_ap on an 8-byte key (bloom_filter.hpp:551-565) ---- */
uint32_t ko_hash_ap8(uint64_t key, uint32_t hash) {
    uint32_t i1 = (uint32_t)key, i2 = (uint32_t)(key >> 32); /* little-endian halves */
    hash ^= (hash << 7) ^ i1 * (hash >> 3) ^ (~((hash << 11) + (i2 ^ (hash >> 5))));
    return hash;
}

/* ---- a5: minion::Random = xoshiro256** (minion.hpp:103-143), SeedSeq32 (305-353), Seed (355-375) ---- */
static uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
uint64_t ko_rng_next(uint64_t s[4]) {
    uint64_t result = rotl64(s[1] * 5, 7) * 9;
    uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return result;
}
static uint64_t splitmix64(uint64_t *state) { /* minion.hpp:292-299 */
    *state += UINT64_C(0x9e3779b97f4a7c15);
    uint64_t z = *state;
    z = (z ^ (z >> 30)) * UINT64_C(0xbf58476d1ce4e5b9);
    z = (z ^ (z >> 27)) * UINT64_C(0x94d049bb133111eb);
    return z ^ (z >> 31);
}
void ko_rng_seed(uint32_t seed, uint64_t st[4]) {
    /* Random::Seed(uint32) -> SeedSeq32 ss(1, s): a vector holding ONE element equal to s */
    uint32_t values[8] = {0x9272B87Fu, 0xD9F64D09u, 0x6640D56Cu, 0x8CDA60ACu,
                          0xDEED25EDu, 0x8495FC63u, 0xAEA86A02u, 0x9F129AB9u};
    for (int i = 0; i < 8; ++i) { /* SeedSeq32::Generate, minion.hpp:320-335 */
        uint64_t s = (uint64_t)values[i];
        uint64_t sum = splitmix64(&s);
        sum += splitmix64(&s) * (uint64_t)seed;
        sum += splitmix64(&s) * 1;
        values[i] = (uint32_t)(sum >> 32);
    }
    uint64_t seeds[4];
    memcpy(seeds, values, sizeof seeds);
    /* InitState, minion.hpp:121-136 */
    st[0] = UINT64_C(0x5FAF84EE2AA04CFF) + seeds[0];
    st[1] = UINT64_C(0xB3A2EF3524D89987) + seeds[1];
    st[2] = UINT64_C(0x5A82B68EF098F79D) + seeds[2];
    st[3] = UINT64_C(0x5D7AA03298486D6E) + seeds[3];
    if (st[0] == 0 && st[1] == 0 && st[2] == 0 && st[3] == 0) st[1] = UINT64_C(0x1615CA18E55EE70C);
    for (int i = 0; i < 256; ++i) ko_rng_next(st);
}
void ko_rng_outputs(uint32_t seed, uint64_t skip, uint64_t n, uint64_t state_after_seed[4], uint64_t *out) {
    uint64_t st[4];
    ko_rng_seed(seed, st);
    memcpy(state_after_seed, st, sizeof st);
    for (uint64_t i = 0; i < skip; ++i) ko_rng_next(st);
    for (uint64_t i = 0; i < n; ++i) out[i] = ko_rng_next(st);
}

/* libstdc++ bernoulli_distribution(p)(urng) with a 64-bit URBG:
 * generate_canonical<double,53> draws ONE value: double(x) / 2^64, clamped below 1; sampled iff < p */
static int bernoulli_draw(uint64_t st[4], double p) {
    double r = (double)ko_rng_next(st) / 18446744073709551616.0;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r < p;
}
void ko_bernoulli(uint32_t seed, double p, uint64_t n, uint8_t *out) {
    uint64_t st[4];
    ko_rng_seed(seed, st);
    for (uint64_t i = 0; i < n; ++i) out[i] = (uint8_t)bernoulli_draw(st, p);
}

/* libstdc++ uniform_int_distribution on a 64-bit URBG: Lemire's method with 128-bit products */
static uint64_t uniform_below(uint64_t st[4], uint64_t range) {
    unsigned __int128 product = (unsigned __int128)ko_rng_next(st) * range;
    uint64_t low = (uint64_t)product;
    if (low < range) {
        uint64_t threshold = -range % range;
        while (low < threshold) {
            product = (unsigned __int128)ko_rng_next(st) * range;
            low = (uint64_t)product;
        }
    }
    return (uint64_t)(product >> 64);
}

/* ---- a19: bloom_parameters::compute_optimal_parameters (bloom_filter.hpp:108-160) ---- */
static const uint32_t predef_salt[128] = { /* data table of bloom_filter.hpp:476-509 (A. Partow, MIT) */
    0xAAAAAAAA, 0x55555555, 0x33333333, 0xCCCCCCCC, 0x66666666, 0x99999999, 0xB5B5B5B5, 0x4B4B4B4B,
    0xAA55AA55, 0x55335533, 0x33CC33CC, 0xCC66CC66, 0x66996699, 0x99B599B5, 0xB54BB54B, 0x4BAA4BAA,
    0xAA33AA33, 0x55CC55CC, 0x33663366, 0xCC99CC99, 0x66B566B5, 0x994B994B, 0xB5AAB5AA, 0xAAAAAA33,
    0x555555CC, 0x33333366, 0xCCCCCC99, 0x666666B5, 0x9999994B, 0xB5B5B5AA, 0xFFFFFFFF, 0xFFFF0000,
    0xB823D5EB, 0xC1191CDF, 0xF623AEB3, 0xDB58499F, 0xC8D42E70, 0xB173F616, 0xA91A5967, 0xDA427D63,
    0xB1E8A2EA, 0xF6C0D155, 0x4909FEA3, 0xA68CC6A7, 0xC395E782, 0xA26057EB, 0x0CD5DA28, 0x467C5492,
    0xF15E6982, 0x61C6FAD3, 0x9615E352, 0x6E9E355A, 0x689B563E, 0x0C9831A8, 0x6753C18B, 0xA622689B,
    0x8CA63C47, 0x42CC2884, 0x8E89919B, 0x6EDBD7D3, 0x15B6796C, 0x1D6FDFE4, 0x63FF9092, 0xE7401432,
    0xEFFE9412, 0xAEAEDF79, 0x9F245A31, 0x83C136FC, 0xC3DA4A8C, 0xA5112C8C, 0x5271F491, 0x9A948DAB,
    0xCEE59A8D, 0xB5F525AB, 0x59D13217, 0x24E7C331, 0x697C2103, 0x84B0A460, 0x86156DA9, 0xAEF2AC68,
    0x23243DA5, 0x3F649643, 0x5FA495A8, 0x67710DF8, 0x9A6C499E, 0xDCFB0227, 0x46A43433, 0x1832B07A,
    0xC46AFF3C, 0xB9C8FFF0, 0xC9500467, 0x34431BDF, 0xB652432B, 0xE367F12B, 0x427F4C1B, 0x224C006E,
    0x2E7E5A89, 0x96F99AA5, 0x0BEB452A, 0x2FD87C39, 0x74B2E1FB, 0x222EFD24, 0xF357F60C, 0x440FCB1E,
    0x8BBE030F, 0x6704DC29, 0x1144D12F, 0x948B1355, 0x6D8FD7E9, 0x1C11A014, 0xADD1592F, 0xFB3C712E,
    0xFC77642F, 0xF9C4CE8C, 0x31312FB9, 0x08B0DD79, 0x318FA6E7, 0xC040D23D, 0xC0589AA7, 0x0CA5C075,
    0xF874B172, 0x0CF914D5, 0x784D3280, 0x4E8CFEBC, 0xC569F575, 0xCDB2A091, 0x2CC016B4, 0x5C5F4421};

ko_bloom *ko_bloom_new(uint64_t projected, double fpr, uint64_t seed) {
    if (projected == 0 || fpr < 0.0 || isinf(fpr) || seed == 0 || seed == UINT64_C(0xFFFFFFFFFFFFFFFF))
        return NULL; /* bloom_parameters::operator!, bloom_filter.hpp:59-70; Bloom ctor throws (bloom.cc:18-21) */
    double min_m = INFINITY, min_k = 0.0, k = 1.0;
    while (k < 1000.0) {
        const double numerator = (-k * (double)projected);
        const double denominator = log(1.0 - pow(fpr, 1.0 / k));
        const double curr_m = numerator / denominator;
        if (curr_m < min_m) { min_m = curr_m; min_k = k; }
        k += 1.0;
    }
    ko_bloom *b = (ko_bloom *)calloc(1, sizeof *b);
    b->optimal_k = (uint32_t)min_k;
    b->optimal_table_size = (uint64_t)min_m;
    b->optimal_table_size += (b->optimal_table_size % 8) != 0 ? 8 - (b->optimal_table_size % 8) : 0;
    if (b->optimal_k < 1) b->optimal_k = 1;
    if (b->optimal_table_size < 1) b->optimal_table_size = 1;
    /* blocked_bloom_filter ctor, bloom.hh:36-56 */
    b->random_seed = seed * UINT64_C(0xA5A5A5A5) + 1;
    b->n_salts = b->optimal_k > 2u ? b->optimal_k : 2u;
    b->table_bits = b->optimal_table_size;
    b->table_bits += (b->table_bits % 512) != 0 ? 512 - (b->table_bits % 512) : 0;
    b->n_blocks = b->table_bits / 512;
    if (b->n_salts > 128) { free(b); return NULL; } /* rand()-based salts (bloom_filter.hpp:530-548): not restated */
    /* generate_unique_salt, bloom_filter.hpp:511-528: mixed IN PLACE in index order */
    for (uint32_t i = 0; i < b->n_salts; ++i) b->salts[i] = predef_salt[i];
    for (uint32_t i = 0; i < b->n_salts; ++i)
        b->salts[i] = b->salts[i] * b->salts[(i + 3) % b->n_salts] + (uint32_t)b->random_seed;
    if (posix_memalign((void **)&b->words, 64, b->table_bits / 8)) { free(b); return NULL; }
    memset(b->words, 0, b->table_bits / 8);
    if (posix_memalign((void **)&b->patterns, 64, (size_t)KO_NUM_PATTERNS * 64)) { free(b->words); free(b); return NULL; }
    memset(b->patterns, 0, (size_t)KO_NUM_PATTERNS * 64);
    /* pattern_blocked_bf ctor, bloom.hh:189-231 */
    uint64_t st[4];
    ko_rng_seed((uint32_t)b->random_seed, st); /* rng.Seed(random_seed_) binds to Seed(uint32_t) */
    size_t bits[512];
    for (size_t i = 0; i < 512; ++i) bits[i] = i;
    { /* std::shuffle(begin, end, rng): 512 elements, URBG range >= 512^2 -> paired draws */
        size_t i = 1, t;
        size_t pos = (size_t)uniform_below(st, 2); /* even count: first swap from distribution{0,1} */
        t = bits[i]; bits[i] = bits[pos]; bits[pos] = t; ++i;
        while (i != 512) {
            uint64_t swap_range = (uint64_t)i + 1;
            uint64_t x = uniform_below(st, swap_range * (swap_range + 1));
            size_t p0 = (size_t)(x / (swap_range + 1)), p1 = (size_t)(x % (swap_range + 1));
            t = bits[i]; bits[i] = bits[p0]; bits[p0] = t; ++i;
            t = bits[i]; bits[i] = bits[p1]; bits[p1] = t; ++i;
        }
    }
    for (size_t i = 0; i < KO_NUM_PATTERNS; ++i) {
        for (uint32_t j = 0; j < b->n_salts; ++j) { /* partial Fisher-Yates continuing from the previous pattern */
            size_t sampled = (size_t)uniform_below(st, 512 - j) + j; /* d(rng, param_type{j, 511}) */
            size_t t = bits[j]; bits[j] = bits[sampled]; bits[sampled] = t;
        }
        for (uint32_t j = 0; j < b->n_salts; ++j) {
            size_t bit = bits[j];
            /* get_vector_unit (bloom.hh:110-113): vec=(bit/8)/32, unit=(bit/8)%4 -- unit is the BYTE index mod 4 */
            size_t vec = (bit / 8) / 32, unit = (bit / 8) % 4;
            b->patterns[i * 8 + vec * 4 + unit] |= UINT64_C(1) << (bit % 64);
        }
    }
    return b;
}

void ko_bloom_free(ko_bloom *b) {
    if (!b) return;
    free(b->words);
    free(b->patterns);
    free(b);
}

/* a3: pattern_blocked_bf::insert (bloom.hh:255-267) */
void ko_bloom_insert(ko_bloom *b, uint64_t key) {
    uint64_t block = ko_hash_ap8(key, b->salts[0]) % b->n_blocks;
    uint64_t pat = ko_hash_ap8(key, b->salts[1]) & (KO_NUM_PATTERNS - 1);
    for (int w = 0; w < 8; ++w) b->words[block * 8 + w] |= b->patterns[pat * 8 + w];
    ++b->inserted;
}
/* a4: pattern_blocked_bf::contains (bloom.hh:276-292) */
int ko_bloom_contains(const ko_bloom *b, uint64_t key) {
    uint64_t block = ko_hash_ap8(key, b->salts[0]) % b->n_blocks;
    uint64_t pat = ko_hash_ap8(key, b->salts[1]) & (KO_NUM_PATTERNS - 1);
    for (int w = 0; w < 8; ++w) {
        uint64_t p = b->patterns[pat * 8 + w];
        if ((b->words[block * 8 + w] & p) != p) return 0;
    }
    return 1;
}
static void bloom_insert_kmer(ko_bloom *b, const kmer_t *km) { if (kmer_valid(km)) ko_bloom_insert(b, kmer_get(km)); } /* bloom.hh:395 */
static int bloom_query_kmer(const ko_bloom *b, const kmer_t *km) { return kmer_valid(km) && ko_bloom_contains(b, kmer_get(km)); } /* bloom.hh:398 */

/* a19: pattern_blocked_bf::effective_fpp (bloom.hh:318-330) */
double ko_bloom_fprate(const ko_bloom *b) {
    long double c = (long double)(b->table_bits / b->inserted); /* INTEGER division, bloom.hh:319 */
    long double lambda = 512 / c;
    long double fpp = 0;
    const unsigned long nsalt = b->n_salts;
    for (int i = 0; i < 3 * lambda; ++i) {
        long double p_block = powl(lambda, i) * expl(-lambda) / tgammal(i + 1);
        long double p_collision = 1.0l - powl(1.0l - 1.0l / (KO_NUM_PATTERNS), i);
        long double fpr_inner = powl(1.0l - expl(-1.0l * nsalt * i / 512), 1.0l * nsalt);
        fpr_inner = p_collision + (1.0l - p_collision) * fpr_inner;
        fpp += p_block * fpr_inner;
    }
    return (double)fpp;
}

/* a19: bloom::calculate_phit (bloom.cc:190-195).  The unqualified pow() there resolves to ::pow(double,double). */
long double ko_phit(const ko_bloom *sampled, long double alpha) {
    long double fpr = ko_bloom_fprate(sampled);
    double exponent = alpha < 0.1 ? 0.2 / alpha : 2;
    long double pa = 1 - pow((double)(1 - alpha), exponent);
    return pa + fpr - fpr * pa;
}

/* a19: covariateutils::log_binom_pmf / log_binom_cdf / calculate_thresholds (covariateutils.hh:54-85) */
static long double log_binom_pmf(unsigned long long k, unsigned long long n, long double p) {
    /* std::lgamma(integer) is the DOUBLE overload */
    long double coefficient = (lgamma((double)(n + 1)) - (lgamma((double)(k + 1)) + lgamma((double)(n - k + 1))));
    return coefficient + (long double)k * logl(p) + (long double)(n - k) * log1pl(-p);
}
void ko_thresholds(int k, long double p, int *out) {
    const long double quartile = .995l;
    long double *cdf = (long double *)malloc(sizeof(long double) * (size_t)(k + 1));
    out[0] = 0;
    for (int i = 1; i <= k; ++i) {
        out[i] = 0;
        cdf[0] = log_binom_pmf(0, (unsigned long long)i, p);
        for (int j = 1; j <= i; ++j) cdf[j] = logl(expl(cdf[j - 1]) + expl(log_binom_pmf((unsigned long long)j, (unsigned long long)i, p)));
        for (int j = 0; j <= i; ++j) {
            if (cdf[j] >= logl(quartile)) { out[i] = j; break; }
        }
    }
    free(cdf);
}
void ko_thresholds_from_filter(const ko_bloom *sampled, long double alpha_num, long double alpha_den, int k, int *out) {
    long double alpha = alpha_num / alpha_den;
    ko_thresholds(k, ko_phit(sampled, alpha), out);
}

/* ---- a5/a6: KmerSubsampler::next + subsample_kmers (htsiter.cc:89-129; recalibrateutils.cc:7-13) ---- */
uint64_t ko_subsample(uint64_t n, const char *seq, const uint64_t *off, int k, double alpha, uint32_t seed,
                      ko_bloom *sampled) {
    return ko_subsample_skip(n, seq, off, k, alpha, seed, 0, sampled);
}

/* the same over a shard of the input whose first k-mer slot has ordinal `skip` in the file-wide stream */
uint64_t ko_subsample_skip(uint64_t n, const char *seq, const uint64_t *off, int k, double alpha, uint32_t seed,
                           uint64_t skip, ko_bloom *sampled) {
    uint64_t st[4], slots = 0;
    ko_rng_seed(seed, st);
    for (uint64_t i = 0; i < skip; ++i) ko_rng_next(st);
    kmer_t km;
    kmer_init(&km, k);
    for (uint64_t r = 0; r < n; ++r) {
        size_t len = (size_t)(off[r + 1] - off[r]);
        if (len == 0) break; /* empty sequence = EOF sentinel (htsiter.cc:95-97) */
        const char *s = seq + off[r];
        kmer_reset(&km);
        for (size_t i = 0; i < len; ++i) {
            kmer_push(&km, s[i]);
            if (i >= (size_t)(k - 1)) { /* one slot, one draw, valid or not (htsiter.cc:100-105,120) */
                ++slots;
                if (bernoulli_draw(st, alpha) && kmer_valid(&km)) bloom_insert_kmer(sampled, &km);
            }
        }
    }
    return slots;
}

/* ---- a7: bloom::overlapping_kmers_in_bf (bloom.cc:28-67); a8: infer_read_errors (readutils.cc:173-193) ---- */
static void infer_read_errors(const char *s, const uint8_t *q, size_t len, int k, const ko_bloom *b,
                              const int *thresholds, uint8_t *errors) {
    kmer_t km;
    kmer_init(&km, k);
    size_t nk = len - (size_t)k + 1; /* reads shorter than k are out of contract (bloom.cc:30 underflows) */
    uint8_t *present = (uint8_t *)calloc(nk ? nk : 1, 1);
    for (size_t i = 0; i < len; ++i) {
        kmer_push(&km, s[i]);
        if (i >= (size_t)(k - 1)) present[i - (size_t)k + 1] = kmer_valid(&km) ? (uint8_t)bloom_query_kmer(b, &km) : 0;
    }
    size_t incount = 0, outcount = 0;
    for (size_t i = 0; i < len; ++i) {
        if (i < len - (size_t)k + 1) { if (present[i]) ++incount; else ++outcount; }
        if (i >= (size_t)k) { if (present[i - (size_t)k]) --incount; else --outcount; }
        size_t in = incount, possible = incount + outcount;
        errors[i] = (in <= (size_t)thresholds[possible] || q[i] <= 2 /* INFER_ERROR_BAD_QUAL */);
    }
    free(present);
}
void ko_infer_errors(uint64_t n, const char *seq, const uint64_t *off, const uint8_t *qual, int k,
                     const ko_bloom *sampled, const int *thresholds, uint8_t *errors) {
    for (uint64_t r = 0; r < n; ++r)
        infer_read_errors(seq + off[r], qual + off[r], (size_t)(off[r + 1] - off[r]), k, sampled, thresholds, errors + off[r]);
}

/* ---- a9: find_trusted_kmers (recalibrateutils.cc:15-40) ---- */
void ko_find_trusted(uint64_t n, const char *seq, const uint64_t *off, const uint8_t *qual, int k,
                     const ko_bloom *sampled, const int *thresholds, ko_bloom *trusted) {
    kmer_t km;
    kmer_init(&km, k);
    for (uint64_t r = 0; r < n; ++r) {
        size_t len = (size_t)(off[r + 1] - off[r]);
        const char *s = seq + off[r];
        uint8_t *errors = (uint8_t *)calloc(len ? len : 1, 1);
        infer_read_errors(s, qual + off[r], len, k, sampled, thresholds, errors);
        int n_trusted = 0;
        kmer_reset(&km);
        for (int i = 0; i < (int)len; ++i) {
            kmer_push(&km, s[i]);
            if (!errors[i]) ++n_trusted;
            if (i >= k && !errors[i - k]) --n_trusted;
            if (kmer_valid(&km) && n_trusted == k) bloom_insert_kmer(trusted, &km);
        }
        free(errors);
    }
}

/* ---- a13 helper: get_next_trusted_char (bloom.cc:83-94) ---- */
static char get_next_trusted_char(const kmer_t *km, const ko_bloom *t, int reverse_test_order) {
    const char *order = !reverse_test_order ? "ACGT" : "TGCA";
    for (int j = 0; j < 4; ++j) {
        kmer_t extra = *km;
        kmer_push(&extra, order[j]);
        if (bloom_query_kmer(t, &extra)) return order[j];
    }
    return 0;
}

/* ---- a10: find_longest_trusted_seq (bloom.cc:96-128) ---- */
static void find_longest_trusted_seq(const char *s, size_t len, const ko_bloom *b, int k, size_t anchor[2]) {
    kmer_t km;
    kmer_init(&km, k);
    size_t anchor_start = NPOS, anchor_end = NPOS, anchor_best = 0, anchor_current = 0;
    for (size_t i = 0; i < len; ++i) {
        if (kmer_push(&km, s[i]) >= (size_t)k) {
            if (bloom_query_kmer(b, &km)) {
                anchor_current++;
            } else {
                if (anchor_current > anchor_best) {
                    anchor_best = anchor_current;
                    anchor_end = i - 1;
                    anchor_start = i + 1 - (size_t)k - anchor_current;
                }
                anchor_current = 0;
            }
        } else if (anchor_current != 0) {
            if (anchor_current > anchor_best) {
                anchor_best = anchor_current;
                anchor_end = i - 1;
                anchor_start = i + 1 - (size_t)k - anchor_current;
            }
            anchor_current = 0;
        }
    }
    if (anchor_current > anchor_best) {
        anchor_best = anchor_current;
        anchor_end = NPOS;
        anchor_start = len + 1 - (size_t)k - anchor_current;
    }
    anchor[0] = anchor_start;
    anchor[1] = anchor_end;
}

/* ---- a13: find_longest_fix (bloom.cc:130-188).  fix[] receives up to 3 candidate chars; returns count. ---- */
static int find_longest_fix(const char *sub, size_t sublen, const ko_bloom *t, int k, int reverse_test_order,
                            char fix[4], size_t *fixlen, int *multiple_out) {
    char *seq = (char *)malloc(sublen + 1);
    memcpy(seq, sub, sublen);
    kmer_t km;
    kmer_init(&km, k);
    int nbest = 0;
    size_t best_i = 0;
    int single = 0, multiple = 0;
    char unfixed_char = seq[k - 1];
    const char *order = !reverse_test_order ? "ACGT" : "TGCA";
    for (int ci = 0; ci < 4; ++ci) {
        char c = order[ci];
        if (c == unfixed_char) continue;
        seq[k - 1] = c;
        kmer_reset(&km);
        size_t i;
        size_t i_stop = (size_t)2 * (size_t)k - 1 > sublen ? (size_t)2 * (size_t)k - 1 : sublen;
        for (i = 0; i < i_stop; ++i) {
            char nch = i < sublen ? seq[i] : get_next_trusted_char(&km, t, reverse_test_order);
            if (nch == 0) break;
            kmer_push(&km, nch);
            if (i >= (size_t)(k - 1)) {
                if (kmer_valid(&km)) {
                    if (!bloom_query_kmer(t, &km)) {
                        break;
                    } else if (i == (size_t)(k - 1)) {
                        if (single) multiple = 1;
                        single = 1;
                    }
                } else {
                    break;
                }
            }
        }
        if (i > best_i) {
            nbest = 0;
            fix[nbest++] = c;
            best_i = i;
        } else if (i == best_i) {
            fix[nbest++] = c;
        }
    }
    free(seq);
    *fixlen = best_i;
    *multiple_out = multiple;
    return nbest;
}

/* ---- a12: adjust_right_anchor (bloom.cc:208-277) ---- */
static size_t adjust_right_anchor(size_t anchor, const char *seq, size_t len, const ko_bloom *t, int k, int *multiple_out) {
    kmer_t km, nk;
    kmer_init(&km, k);
    int multiple = 0;
    size_t modified_idx = anchor + 1;
    static const char acgt[4] = {'A', 'C', 'G', 'T'};
    for (size_t i = modified_idx - (size_t)k + 1; i < modified_idx; ++i) kmer_push(&km, seq[i]);
    for (int ci = 0; ci < 4; ++ci) {
        char c = acgt[ci];
        if (seq[modified_idx] == c) continue;
        nk = km;
        kmer_push(&nk, c);
        for (size_t i = 0; i <= (size_t)k; ++i) {
            if (!bloom_query_kmer(t, &nk)) break;
            if (modified_idx + i == len - 1 || i == (size_t)k) {
                *multiple_out = multiple;
                return anchor;
            } else {
                kmer_push(&nk, seq[modified_idx + i + 1]);
            }
        }
    }
    for (int i = k / 2 - 1; i >= 0 && anchor > (size_t)(i + k - 1); --i) {
        kmer_reset(&km);
        modified_idx = anchor - (size_t)i;
        for (size_t j = modified_idx - (size_t)k + 1; j < modified_idx; ++j) kmer_push(&km, seq[j]);
        for (int ci = 0; ci < 4; ++ci) {
            char c = acgt[ci];
            if (seq[modified_idx] == c) continue;
            nk = km;
            kmer_push(&nk, c);
            if (kmer_valid(&nk) && bloom_query_kmer(t, &nk)) {
                multiple = 1;
                for (size_t j = 0; kmer_valid(&nk) && bloom_query_kmer(t, &nk) && modified_idx + 1 + j < len &&
                                   j <= (size_t)(k / 2); ++j) {
                    kmer_push(&nk, seq[modified_idx + 1 + j]);
                    if (j == (size_t)(k / 2) && kmer_valid(&nk) && bloom_query_kmer(t, &nk)) {
                        *multiple_out = multiple;
                        return modified_idx - 1;
                    }
                }
            }
        }
    }
    *multiple_out = multiple;
    return anchor;
}

/* ---- a11: biggest_consecutive_trusted_block (bloom.cc:279-305) ---- */
static int biggest_consecutive_trusted_block(const char *seq, size_t len, const ko_bloom *t, int k, int current_len) {
    kmer_t km;
    kmer_init(&km, k);
    int in = 0, out = 0, best = 0;
    for (size_t i = 0; i < len; ++i) {
        kmer_push(&km, seq[i]);
        if (i >= (size_t)(k - 1)) {
            if (bloom_query_kmer(t, &km)) {
                ++in;
            } else {
                if (in > best) best = in;
                in = 0;
                ++out;
                if (k - out < current_len) break;
            }
        }
    }
    if (in > best) best = in;
    return best;
}

/* ---- a11: CReadData::correct_one (readutils.cc:195-235) ---- */
static size_t correct_one(char *seq, const uint8_t *qual, size_t len, const ko_bloom *t, int k) {
    int best_fix_len = 0;
    char best_fix_base = 0;
    size_t best_fix_pos = NPOS;
    static const char acgt[4] = {'A', 'C', 'G', 'T'};
    char *orig = (char *)malloc(len + 1);
    for (size_t i = 0; i < len; ++i) {
        memcpy(orig, seq, len);
        for (int ci = 0; ci < 4; ++ci) {
            char c = acgt[ci];
            if (seq[i] == c) continue;
            orig[i] = c;
            size_t start = i > (size_t)(k - 1) ? i - (size_t)k + 1 : 0;
            size_t a = i - (size_t)(k / 2) + 1, b = len - (size_t)k;
            size_t magic_start = i > (size_t)(k / 2 - 1) ? (a < b ? a : b) : 0;
            kmer_t magic;
            kmer_init(&magic, k);
            for (size_t j = magic_start; j <= magic_start + (size_t)k - 1; ++j) kmer_push(&magic, orig[j]);
            if (bloom_query_kmer(t, &magic)) {
                size_t sublen = (size_t)(2 * k - 1);
                if (start + sublen > len) sublen = len - start; /* std::string::substr clamps */
                int n_in = biggest_consecutive_trusted_block(orig + start, sublen, t, k, best_fix_len);
                if (n_in > best_fix_len) {
                    best_fix_base = c;
                    best_fix_pos = i;
                    best_fix_len = n_in;
                } else if (n_in == best_fix_len && qual[i] < qual[best_fix_pos]) {
                    best_fix_base = c;
                    best_fix_pos = i;
                }
            }
        }
    }
    free(orig);
    if (best_fix_len > 0) seq[best_fix_pos] = best_fix_base;
    return best_fix_pos;
}

/* ---- a14: CReadData::get_errors (readutils.cc:238-570).  seq/errors are modified in place; on the normal
 *      path seq is restored to the original (568), on the two early returns it is not (254, 264). ---- */
static void get_errors(char *seq, const uint8_t *qual, uint8_t *errors, size_t len, const ko_bloom *trusted, int k,
                       int minqual, int first_call) {
    char *original_seq = (char *)malloc(len + 1);
    memcpy(original_seq, seq, len);
    size_t bad_prefix = 0;
    size_t bad_suffix = NPOS;
    int multiple = 0;
    size_t anchor[2];
    find_longest_trusted_seq(seq, len, trusted, k, anchor);
    if (anchor[0] == NPOS) {
        multiple = 1;
        size_t corrected_idx = correct_one(seq, qual, len, trusted, k);
        if (corrected_idx == NPOS) {
            free(original_seq);
            return;
        } else {
            find_longest_trusted_seq(seq, len, trusted, k, anchor);
            errors[corrected_idx] = 1;
        }
    }
    if (anchor[0] == 0 && anchor[1] == NPOS) {
        free(original_seq);
        return;
    }
    size_t anchor_len = (anchor[1] < len - 1 ? anchor[1] : len - 1) + 1 - anchor[0];
    int corrected = 0;
    /* right side (271-346) */
    if (anchor[1] != NPOS) {
        if (anchor_len - (size_t)k + 1 >= (size_t)k) {
            int current_multiple;
            anchor[1] = adjust_right_anchor(anchor[1], seq, len, trusted, k, &current_multiple);
            multiple = multiple || current_multiple;
        }
        for (size_t i = anchor[1] + 1; i < len;) {
            size_t start = i - (size_t)k + 1;
            char fix[4];
            size_t fixlen;
            int current_multiple;
            int nfix = find_longest_fix(seq + start, len - start, trusted, k, 0, fix, &fixlen, &current_multiple);
            multiple = multiple || current_multiple;
            size_t next_untrusted_idx = start + fixlen;
            if (next_untrusted_idx > i) {
                if (nfix > 1) {
                    multiple = 1;
                    size_t largest_possible_idx = i + (size_t)k - 1 < len - 1 ? i + (size_t)k - 1 : len - 1;
                    if (next_untrusted_idx <= largest_possible_idx || largest_possible_idx - i + 1 < (size_t)k) {
                        bad_suffix = i;
                        break;
                    }
                } else {
                    seq[i] = fix[0];
                    errors[i] = 1;
                }
                corrected = 1;
                i += fixlen - (size_t)k + 1;
            } else {
                bad_suffix = i;
                break;
            }
        }
    }
    /* left side (348-422) */
    if (anchor[0] != 0) {
        char *revcomped = (char *)malloc(len + 1);
        for (size_t i = 0; i < len; ++i) revcomped[i] = complement_char(seq[len - 1 - i]);
        if (anchor_len - (size_t)k + 1 >= (size_t)k) {
            int current_multiple;
            size_t left_adjust = adjust_right_anchor(len - anchor[0] - 1, revcomped, len, trusted, k, &current_multiple);
            anchor[0] = len - left_adjust - 1;
            multiple = multiple || current_multiple;
        }
        for (int i = (int)(anchor[0] - 1); i >= 0;) {
            int j = (int)(len - (size_t)i - 1);
            size_t start = (size_t)(j - k + 1);
            char fix[4];
            size_t fixlen;
            int current_multiple;
            int nfix = find_longest_fix(revcomped + start, len - start, trusted, k, 1, fix, &fixlen, &current_multiple);
            multiple = multiple || current_multiple;
            size_t next_untrusted_idx = start + fixlen;
            if (next_untrusted_idx > (size_t)j) {
                if (nfix > 1) {
                    multiple = 1;
                    size_t largest_possible_idx = (size_t)j + (size_t)k - 1 < len - 1 ? (size_t)j + (size_t)k - 1 : len - 1;
                    if (next_untrusted_idx <= largest_possible_idx || largest_possible_idx - (size_t)j + 1 < (size_t)k) {
                        bad_prefix = (size_t)i;
                        break;
                    }
                } else {
                    revcomped[j] = fix[0];
                    errors[i] = 1;
                }
                corrected = 1;
                i = (int)((size_t)i - (next_untrusted_idx - (size_t)j));
            } else {
                bad_prefix = (size_t)i;
                break;
            }
        }
        free(revcomped);
    }
    /* overcorrection check (429-546) */
    if (corrected) {
        int adjust = 1;
        kmer_t km;
        kmer_init(&km, k);
        size_t trusted_start = NPOS, trusted_end = NPOS;
        for (size_t i = 0; i < len && adjust == 1; ++i) {
            kmer_push(&km, original_seq[i]);
            if (kmer_valid(&km) && bloom_query_kmer(trusted, &km)) {
                size_t cand = i - (size_t)k + 1;
                trusted_start = trusted_start < cand ? trusted_start : cand;
                trusted_end = i;
            } else {
                if (i > trusted_end) {
                    for (size_t j = trusted_start; j <= trusted_end; ++j) {
                        if (errors[j]) { adjust = 0; break; }
                    }
                    trusted_start = NPOS;
                    trusted_end = NPOS;
                }
            }
        }
        adjust = adjust && !multiple;
        int ocwindow = 20;
        int base_threshold = 4;
        int threshold = base_threshold;
        double occount = 0;
        int *overcorrected_idx = (int *)malloc(sizeof(int) * (len ? len : 1));
        size_t n_oc = 0;
        for (int i = 0; (size_t)i < len; ++i) {
            if (errors[i] && ko_nt4((unsigned char)original_seq[i]) < 4) {
                if (qual[i] <= minqual) occount += 0.5; else ++occount;
            }
            if (i >= ocwindow && errors[i - ocwindow] && ko_nt4((unsigned char)original_seq[i - ocwindow]) < 4) {
                if (qual[i - ocwindow] <= minqual) occount -= 0.5; else --occount;
            }
            threshold = adjust && i >= ocwindow && (size_t)(i + ocwindow - 1) < len ? base_threshold + 1 : base_threshold;
            if (occount > threshold && errors[i]) overcorrected_idx[n_oc++] = i;
        }
        for (size_t oi = 0; oi < n_oc; ++oi) {
            int oc_idx = overcorrected_idx[oi];
            if (errors[oc_idx]) {
                int start = oc_idx - k + 1;
                start = start >= 0 ? start : 0;
                int end = oc_idx + k;
                end = (size_t)end < len ? end : (int)len;
                for (int i = start; i < end; ++i) {
                    if (errors[i]) {
                        errors[i] = 0;
                        if (i + k > end) end = (size_t)(i + k) < len ? i + k : (int)len;
                        if (i - k < start) {
                            i = i - k + 1 >= 0 ? i - k : -1;
                            start = i;
                        }
                    }
                }
            }
        }
        free(overcorrected_idx);
    }
    /* one level of recursion on a long bad prefix / suffix (547-563) */
    if (first_call && bad_prefix > 0 && (bad_prefix >= len / 2 || bad_prefix >= (size_t)(2 * k))) {
        size_t sl = bad_prefix + 1; /* substr(0, bad_prefix+1): current seq, qual, errors (597-608) */
        char *sseq = (char *)malloc(sl + 1);
        uint8_t *serr = (uint8_t *)malloc(sl);
        memcpy(sseq, seq, sl);
        memcpy(serr, errors, sl);
        get_errors(sseq, qual, serr, sl, trusted, k, minqual, 0);
        memcpy(errors, serr, sl);
        free(sseq);
        free(serr);
    }
    if (first_call && bad_suffix < NPOS && bad_suffix < len &&
        (len - bad_suffix > len / 2 || len - bad_suffix > (size_t)(2 * k))) {
        size_t sl = len - bad_suffix;
        char *sseq = (char *)malloc(sl + 1);
        uint8_t *serr = (uint8_t *)malloc(sl);
        memcpy(sseq, seq + bad_suffix, sl);
        memcpy(serr, errors + bad_suffix, sl);
        get_errors(sseq, qual + bad_suffix, serr, sl, trusted, k, minqual, 0);
        memcpy(errors + bad_suffix, serr, sl);
        free(sseq);
        free(serr);
    }
    memcpy(seq, original_seq, len); /* this->seq = original_seq (568) */
    free(original_seq);
}

void ko_get_errors(uint64_t n, const char *seq, const uint64_t *off, const uint8_t *qual, int k,
                   const ko_bloom *trusted, uint8_t *errors, char *seq_after) {
    for (uint64_t r = 0; r < n; ++r) {
        size_t len = (size_t)(off[r + 1] - off[r]);
        char *s = (char *)malloc(len + 1);
        memcpy(s, seq + off[r], len);
        memset(errors + off[r], 0, len);
        get_errors(s, qual + off[r], errors + off[r], len, trusted, k, 6, 1);
        if (seq_after) memcpy(seq_after + off[r], s, len);
        free(s);
    }
}

/* ---- a15: C{RG,Q,Cycle,Dinuc}Covariate::consume_read (covariateutils.cc:30-42,65-76,102-116,147-164) ---- */
void ko_consume(uint64_t n, const char *seq, const uint64_t *off, const uint8_t *qual, const int32_t *rg,
                const uint8_t *second, const uint8_t *errors, int R, int maxlen, uint64_t *rgc, uint64_t *qc,
                uint64_t *cyc, uint64_t *dic) {
    (void)R;
    for (uint64_t r = 0; r < n; ++r) {
        size_t len = (size_t)(off[r + 1] - off[r]);
        const char *s = seq + off[r];
        const uint8_t *q = qual + off[r], *e = errors + off[r];
        size_t g = (size_t)(rg ? rg[r] : 0);
        int sec = second ? second[r] != 0 : 0;
        uint64_t nerr = 0;
        for (size_t i = 0; i < len; ++i) nerr += e[i] != 0;
        rgc[g * 2] += nerr;     /* skips are never set (covariateutils.cc:194-197) */
        rgc[g * 2 + 1] += len;
        for (size_t i = 0; i < len; ++i) {
            size_t qi = (g * 94 + q[i]) * 2;
            qc[qi] += e[i] != 0;
            qc[qi + 1] += 1;
            size_t ci = (((g * 94 + q[i]) * 2 + (size_t)sec) * (size_t)maxlen + i) * 2;
            cyc[ci] += e[i] != 0;
            cyc[ci + 1] += 1;
            if (i >= 1 && ko_nt4((unsigned char)s[i]) < 4 && ko_nt4((unsigned char)s[i - 1]) < 4 && q[i] >= 6) {
                int dinuc = 15 & ((ko_nt4((unsigned char)s[i - 1]) << 2) | ko_nt4((unsigned char)s[i]));
                size_t di = ((g * 94 + q[i]) * 16 + (size_t)dinuc) * 2;
                dic[di] += e[i] != 0;
                dic[di + 1] += 1;
            }
        }
    }
}

/* ---- a16: NormalPrior (covariateutils.cc:7-19), delta_q x4, get_dqs (204-230), q_to_p/p_to_q (recalibrateutils.hh:36-37) ---- */
static long double q_to_p(int q) { return powl(10.0l, -((long double)q / 10.0l)); }
static int p_to_q(long double p) { return p > 0 ? (int)(-10 * log10l(p)) : 42; }
static long double normal_prior(int j) {
    errno = 0;
    long double v = logl(.9l * expl(-(powl(((long double)j / .5l), 2.0l)) / 2.0l));
    if (errno != 0) v = -LDBL_MAX; /* numeric_limits<long double>::lowest() */
    return v;
}
static int map_delta(uint64_t err, uint64_t tot, int prior) {
    int map_q = 0;
    long double best_posterior = -LDBL_MAX;
    for (int possible = 0; possible < KO_MAXQ + 1; ++possible) {
        int diff = abs(prior - possible);
        long double prior_prob = normal_prior(diff);
        long double p = q_to_p(possible);
        long double loglike = log_binom_pmf(err + 1, tot + 2, p);
        long double posterior = prior_prob + loglike;
        if (posterior > best_posterior) { map_q = possible; best_posterior = posterior; }
    }
    return map_q - prior;
}
void ko_get_dqs(int R, int maxlen, const uint64_t *rgc, const uint64_t *qc, const uint64_t *cyc, const uint64_t *dic,
                int32_t *meanq, int32_t *rgdq, int32_t *qdq, int32_t *cydq, int32_t *didq) {
    /* vector extents as the reference grows them: q extent = 1 + largest q seen; cycle extent = 1 + largest
     * cycle seen for (rg,q,second); dinuc extent is 16 or 0; dicov's q extent = 1 + largest eligible q seen. */
    for (int r = 0; r < R; ++r) {
        int qext = 0, dqext = 0;
        for (int q = 0; q < 94; ++q) {
            if (qc[((size_t)r * 94 + q) * 2 + 1]) qext = q + 1;
            for (int n = 0; n < 16; ++n) if (dic[(((size_t)r * 94 + q) * 16 + n) * 2 + 1]) dqext = q + 1;
        }
        long double expected_errors = 0;
        for (int q = 0; q < qext; ++q) expected_errors += (q_to_p(q) * qc[((size_t)r * 94 + q) * 2 + 1]);
        meanq[r] = p_to_q(expected_errors / rgc[(size_t)r * 2 + 1]);
        rgdq[r] = map_delta(rgc[(size_t)r * 2], rgc[(size_t)r * 2 + 1], meanq[r]);
        int rgprior = meanq[r] + rgdq[r];
        for (int q = 0; q < qext; ++q) {
            size_t qi = (size_t)r * 94 + q;
            qdq[qi] = map_delta(qc[qi * 2], qc[qi * 2 + 1], rgprior);
            int qprior = rgprior + qdq[qi];
            for (int s = 0; s < 2; ++s) {
                int cext = 0;
                for (int c = 0; c < maxlen; ++c) if (cyc[((qi * 2 + s) * (size_t)maxlen + c) * 2 + 1]) cext = c + 1;
                for (int c = 0; c < cext; ++c) {
                    size_t ci = (qi * 2 + s) * (size_t)maxlen + c;
                    cydq[ci] = map_delta(cyc[ci * 2], cyc[ci * 2 + 1], qprior);
                }
            }
            if (q < dqext) {
                int any = 0;
                for (int n = 0; n < 16; ++n) if (dic[(qi * 16 + n) * 2 + 1]) any = 1;
                if (any)
                    for (int n = 0; n < 16; ++n) didq[qi * 16 + n] = map_delta(dic[(qi * 16 + n) * 2], dic[(qi * 16 + n) * 2 + 1], qprior);
            }
        }
    }
}

/* ---- a17: CReadData::recalibrate (readutils.cc:572-595) ---- */
void ko_recalibrate(uint64_t n, const char *seq, const uint64_t *off, const uint8_t *qual, const int32_t *rg,
                    const uint8_t *second, int R, int maxlen, const int32_t *meanq, const int32_t *rgdq,
                    const int32_t *qdq, const int32_t *cydq, const int32_t *didq, uint8_t *out) {
    (void)R;
    for (uint64_t r = 0; r < n; ++r) {
        size_t len = (size_t)(off[r + 1] - off[r]);
        const char *s = seq + off[r];
        const uint8_t *qq = qual + off[r];
        size_t g = (size_t)(rg ? rg[r] : 0);
        int sec = second ? second[r] != 0 : 0;
        for (size_t i = 0; i < len; ++i) {
            uint8_t q = qq[i];
            int v = q;
            if (q >= 6) {
                v = meanq[g] + rgdq[g] + qdq[g * 94 + q] + cydq[((g * 94 + q) * 2 + (size_t)sec) * (size_t)maxlen + i];
                if (i > 0) {
                    int first = ko_nt4((unsigned char)s[i - 1]), sc = ko_nt4((unsigned char)s[i]);
                    if (first < 4 && sc < 4) v += didq[(g * 94 + q) * 16 + (15 & ((first << 2) | sc))];
                }
            }
            out[off[r] + i] = (uint8_t)(v < 0 ? 0 : KO_MAXQ < v ? KO_MAXQ : v);
        }
    }
}
