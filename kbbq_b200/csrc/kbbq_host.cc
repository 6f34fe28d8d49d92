// kbbq_host.cc -- see kbbq_host.h.  Compiled with -ffp-contract=off: results must not depend on FMA availability.
#include "kbbq_host.h"

#include <cstdio>
#include <string>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <numeric>
#include <random>
#include <thread>

namespace kbbq {
namespace host {

int nt4(unsigned char ch) {
    switch (ch) {
    case 'A': case 'a': case '0': return 0;
    case 'C': case 'c': case '1': return 1;
    case 'G': case 'g': case '2': return 2;
    case 'T': case 't': case '3': return 3;
    default: return 4;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// xoshiro256** with minion's seeding
// ------------------------------------------------------------------------------------------------------------------
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

uint64_t Xoshiro256ss::next() {  // minion.hpp:103-118
    const uint64_t out = rotl(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return out;
}

static inline uint64_t splitmix64(uint64_t& x) {  // minion.hpp:292-299
    x += 0x9e3779b97f4a7c15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

void Xoshiro256ss::seed(uint32_t seed) {
    // Random::Seed(uint32_t) builds SeedSeq32(1, seed) -- ONE stored value -- and hashes eight fixed words with it
    // (minion.hpp:355-375, 320-335): a multilinear hash whose coefficients come from splitmix64.
    uint32_t v[8] = {0x9272B87Fu, 0xD9F64D09u, 0x6640D56Cu, 0x8CDA60ACu, 0xDEED25EDu, 0x8495FC63u, 0xAEA86A02u, 0x9F129AB9u};
    for (uint32_t& x : v) {
        uint64_t st = x;
        uint64_t sum = splitmix64(st);
        sum += splitmix64(st) * seed;
        sum += splitmix64(st);  // the trailing "* 1" term that keeps a zero seed word from vanishing
        x = (uint32_t)(sum >> 32);
    }
    uint64_t w[4];
    std::memcpy(w, v, sizeof w);
    // InitState (minion.hpp:121-136): add to well-mixed constants, never all-zero, burn 256 outputs
    s[0] = 0x5FAF84EE2AA04CFFull + w[0];
    s[1] = 0xB3A2EF3524D89987ull + w[1];
    s[2] = 0x5A82B68EF098F79Dull + w[2];
    s[3] = 0x5D7AA03298486D6Eull + w[3];
    if ((s[0] | s[1] | s[2] | s[3]) == 0) s[1] = 0x1615CA18E55EE70Cull;
    for (int i = 0; i < 256; ++i) next();
}

uint32_t default_seed(uint64_t time_now, uint32_t pid) {
    // create_seed_seq: {time lo, time hi, pid, 0xC8F978DB, 0x0B32F62E} (the __cpluscplus typo in minion.hpp:385 disables
    // the chrono / random_device branch); GenerateOne hashes the single word 0xFD57D105 with it
    const uint32_t seq[5] = {(uint32_t)time_now, (uint32_t)(time_now >> 32), pid, 0xC8F978DBu, 0x0B32F62Eu};
    uint64_t st = 0xFD57D105u;
    uint64_t sum = splitmix64(st);
    for (uint32_t u : seq) sum += splitmix64(st) * u;
    sum += splitmix64(st);
    return (uint32_t)(sum >> 32);
}

static void matvec(const uint64_t* cols, const uint64_t in[4], uint64_t out[4]) {
    uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int w = 0; w < 4; ++w) {
        uint64_t bits = in[w];
        while (bits) {
            const int c = w * 64 + __builtin_ctzll(bits);
            bits &= bits - 1;
            const uint64_t* col = cols + (size_t)c * 4;
            a0 ^= col[0]; a1 ^= col[1]; a2 ^= col[2]; a3 ^= col[3];
        }
    }
    out[0] = a0; out[1] = a1; out[2] = a2; out[3] = a3;
}

const JumpTables& jump_tables() {
    static JumpTables jt;
    static std::once_flag once;
    std::call_once(once, [] {
        jt.pow2.resize((size_t)64 * 1024);
        uint64_t* m0 = jt.pow2.data();
        for (int c = 0; c < 256; ++c) {  // one step applied to each basis vector
            Xoshiro256ss e;
            e.s[0] = e.s[1] = e.s[2] = e.s[3] = 0;
            e.s[c >> 6] = 1ull << (c & 63);
            e.next();
            std::memcpy(m0 + (size_t)c * 4, e.s, 32);
        }
        for (int e = 1; e < 64; ++e) {  // square
            const uint64_t* prev = jt.pow2.data() + (size_t)(e - 1) * 1024;
            uint64_t* cur = jt.pow2.data() + (size_t)e * 1024;
            for (int c = 0; c < 256; ++c) matvec(prev, prev + (size_t)c * 4, cur + (size_t)c * 4);
        }
    });
    return jt;
}

void rng_advance(uint64_t state[4], uint64_t steps) {
    const JumpTables& jt = jump_tables();
    for (int e = 0; steps; ++e, steps >>= 1) {
        if (steps & 1) {
            uint64_t out[4];
            matvec(jt.matrix(e), state, out);
            std::memcpy(state, out, 32);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Bloom geometry
// ------------------------------------------------------------------------------------------------------------------
static const uint32_t kPredefSalt[128] = {  // data table of bloom_filter.hpp:476-509 (Open Bloom Filter, A. Partow, MIT)
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

bool bloom_geometry(uint64_t projected, double fpr, uint64_t seed, BloomGeom& g) {
    // bloom_parameters::operator! (bloom_filter.hpp:59-70) with the default min/max members
    if (projected == 0 || fpr < 0.0 || std::numeric_limits<double>::infinity() == std::abs(fpr) || seed == 0 ||
        seed == 0xFFFFFFFFFFFFFFFFull)
        return false;
    g.projected = projected;
    g.fpr = fpr;
    g.seed = seed;
    // compute_optimal_parameters (bloom_filter.hpp:108-160): scan k = 1..999 for the smallest table
    double best_m = std::numeric_limits<double>::infinity(), best_k = 0.0;
    for (double kk = 1.0; kk < 1000.0; kk += 1.0) {
        const double numerator = (-kk * projected);
        const double denominator = std::log(1.0 - std::pow(fpr, 1.0 / kk));
        const double m = numerator / denominator;
        if (m < best_m) { best_m = m; best_k = kk; }
    }
    g.optimal_k = std::max(static_cast<unsigned int>(best_k), 1u);
    g.optimal_table_size = static_cast<unsigned long long>(best_m);
    g.optimal_table_size += (g.optimal_table_size % 8) ? 8 - (g.optimal_table_size % 8) : 0;
    g.optimal_table_size = std::max<uint64_t>(g.optimal_table_size, 1);
    // blocked_bloom_filter ctor (bloom.hh:36-45)
    g.random_seed = seed * 0xA5A5A5A5ull + 1;
    g.n_salts = std::max(g.optimal_k, 2u);
    g.table_bits = g.optimal_table_size;
    g.table_bits += (g.table_bits % 512) ? 512 - (g.table_bits % 512) : 0;
    g.n_blocks = g.table_bits / 512;
    if (g.n_salts > 128) return false;  // would need the rand()-based salts (bloom_filter.hpp:530-548); fpr < 2^-128
    // generate_unique_salt (bloom_filter.hpp:511-528): each salt mixed in place with the one three ahead (wrapping)
    std::copy(kPredefSalt, kPredefSalt + g.n_salts, g.salts);
    for (uint32_t i = 0; i < g.n_salts; ++i) g.salts[i] = g.salts[i] * g.salts[(i + 3) % g.n_salts] + (uint32_t)g.random_seed;
    return true;
}

void bloom_patterns(const BloomGeom& g, uint64_t* out) {
    // pattern_blocked_bf ctor (bloom.hh:202-230).  The libstdc++ facilities are CALLED, as the reference calls them,
    // with the same engine: the draw sequence of std::shuffle / uniform_int_distribution is implementation-defined.
    std::memset(out, 0, (size_t)65536 * 64);
    Xoshiro256ss rng;
    rng.seed((uint32_t)g.random_seed);  // rng.Seed(random_seed_) binds to the uint32_t overload
    std::uniform_int_distribution<> d(0, 511);
    std::vector<size_t> bits(512);
    std::iota(bits.begin(), bits.end(), 0);
    std::shuffle(bits.begin(), bits.end(), rng);
    for (size_t i = 0; i < 65536; ++i) {
        for (int j = 0; j < (int)g.n_salts; ++j) {  // partial Fisher-Yates continuing from the previous pattern
            const size_t pick = d(rng, std::uniform_int_distribution<>::param_type{j, 511});
            std::swap(bits[j], bits[pick]);
        }
        for (size_t j = 0; j < g.n_salts; ++j) {
            const size_t b = bits[j];
            // get_vector_unit (bloom.hh:110-113): vector = byte/32, unit = byte%4 (sic), bit within the u64 = b%64
            const size_t byte = b / 8;
            out[i * 8 + (byte / 32) * 4 + (byte % 4)] |= 1ull << (b % 64);
        }
    }
}

double effective_fpp(uint64_t table_bits, uint64_t inserted, uint32_t n_salts) {
    if (inserted == 0) return std::numeric_limits<double>::quiet_NaN();  // the reference divides by zero here
    const size_t salts = n_salts, block_size = 512, num_patterns = 65536;
    long double c = table_bits / inserted;  // integer division (bloom.hh:319)
    long double lambda = block_size / c;
    long double fpp = 0;
    for (int i = 0; i < 3 * lambda; ++i) {
        long double p_block = std::pow(lambda, i) * std::exp(-lambda) / std::tgammal(i + 1);
        long double p_collision = 1.0l - std::pow(1.0l - 1.0l / (num_patterns), i);
        long double fpr_inner = std::pow(1.0l - std::exp(-1.0l * salts * i / block_size), 1.0l * salts);
        fpr_inner = p_collision + (1.0l - p_collision) * fpr_inner;
        fpp += p_block * fpr_inner;
    }
    return fpp;
}

long double calculate_phit(double fpr_d, long double alpha) {
    long double fpr = fpr_d;
    double exponent = alpha < 0.1 ? 0.2 / alpha : 2;
    // the reference's unqualified pow() (bloom.cc:193) is ::pow(double, double)
    long double pa = 1 - ::pow(static_cast<double>(1 - alpha), exponent);
    return pa + fpr - fpr * pa;
}

static inline long double log_binom_pmf(unsigned long long k, unsigned long long n, long double p) {
    long double coefficient = (std::lgamma(n + 1) - (std::lgamma(k + 1) + std::lgamma(n - k + 1)));  // double lgamma
    return coefficient + (long double)k * std::log(p) + (long double)(n - k) * std::log1p(-p);
}

void calculate_thresholds(int k, long double p, int* out) {
    const long double quartile = .995l;
    std::vector<long double> cdf((size_t)k + 1);
    out[0] = 0;
    for (int i = 1; i <= k; ++i) {
        out[i] = 0;
        cdf[0] = log_binom_pmf(0, i, p);
        for (int j = 1; j <= i; ++j) cdf[j] = std::log(std::exp(cdf[j - 1]) + std::exp(log_binom_pmf(j, i, p)));
        for (int j = 0; j <= i; ++j)
            if (cdf[j] >= std::log(quartile)) { out[i] = j; break; }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// model fit
// ------------------------------------------------------------------------------------------------------------------
namespace {
struct FitConsts {
    long double prior[94];   // NormalPrior::get_normal_prior(diff) (covariateutils.cc:7-19)
    long double logp[94];    // log(q_to_p(q))
    long double log1mp[94];  // log1p(-q_to_p(q))
    long double p[94];
    FitConsts() {
        for (int i = 0; i < 94; ++i) {
            errno = 0;
            long double v = std::log(.9l * std::exp(-(std::pow(((long double)i / .5l), 2.0l)) / 2.0l));
            if (errno != 0) v = std::numeric_limits<long double>::lowest();
            prior[i] = v;
            p[i] = std::pow(10.0l, -((long double)i / 10.0l));  // q_to_p, recalibrateutils.hh:36
            logp[i] = std::log(p[i]);
            log1mp[i] = std::log1p(-p[i]);
        }
        errno = 0;
    }
};
const FitConsts& fit_consts() {
    static FitConsts fc;
    return fc;
}

// the MAP search shared by the four delta_q functions (covariateutils.cc:44-63 etc.)
int map_delta(const FitConsts& fc, unsigned long long err, unsigned long long tot, int prior) {
    const unsigned long long k = err + 1, n = tot + 2;
    const long double coefficient = (std::lgamma(n + 1) - (std::lgamma(k + 1) + std::lgamma(n - k + 1)));
    const long double kl = (long double)k, nkl = (long double)(n - k);
    int map_q = 0;
    long double best = std::numeric_limits<long double>::lowest();
    for (int possible = 0; possible < 94; ++possible) {
        const int diff = std::abs(prior - possible);
        const long double prior_prob = diff < 94 ? fc.prior[diff] : std::numeric_limits<long double>::lowest();
        const long double loglike = coefficient + kl * fc.logp[possible] + nkl * fc.log1mp[possible];
        const long double posterior = prior_prob + loglike;
        if (posterior > best) { map_q = possible; best = posterior; }
    }
    return map_q - prior;
}
}  // namespace

void fit_model(int R, int L, const uint64_t* rgc, const uint64_t* qc, const uint64_t* cyc, const uint64_t* dic,
               int32_t* meanq, int32_t* rgdq, int32_t* qdq, int32_t* cydq, int32_t* didq, int n_threads) {
    const FitConsts& fc = fit_consts();
    struct Cell { const uint64_t* cnt; int32_t* out; int prior; };
    std::vector<Cell> cells;
    for (int r = 0; r < R; ++r) {
        // vector extents as the reference grows them lazily: a q row exists up to the largest q seen
        int qext = 0, dqext = 0;
        for (int q = 0; q < 94; ++q) {
            if (qc[((size_t)r * 94 + q) * 2 + 1]) qext = q + 1;
            for (int n = 0; n < 16; ++n)
                if (dic[(((size_t)r * 94 + q) * 16 + n) * 2 + 1]) dqext = q + 1;
        }
        long double expected_errors = 0;
        for (int q = 0; q < qext; ++q) expected_errors += (fc.p[q] * qc[((size_t)r * 94 + q) * 2 + 1]);
        const long double pm = expected_errors / rgc[(size_t)r * 2 + 1];
        meanq[r] = pm > 0 ? (int)(-10 * std::log10(pm)) : 42;  // p_to_q, recalibrateutils.hh:37
        rgdq[r] = map_delta(fc, rgc[(size_t)r * 2], rgc[(size_t)r * 2 + 1], meanq[r]);
        const int rgprior = meanq[r] + rgdq[r];
        for (int q = 0; q < qext; ++q) {
            const size_t qi = (size_t)r * 94 + q;
            qdq[qi] = map_delta(fc, qc[qi * 2], qc[qi * 2 + 1], rgprior);
            const int qprior = rgprior + qdq[qi];
            for (int s = 0; s < 2; ++s) {
                int cext = 0;
                for (int c = 0; c < L; ++c)
                    if (cyc[((qi * 2 + s) * (size_t)L + c) * 2 + 1]) cext = c + 1;
                for (int c = 0; c < cext; ++c) {
                    const size_t ci = (qi * 2 + s) * (size_t)L + c;
                    cells.push_back({cyc + ci * 2, cydq + ci, qprior});
                }
            }
            if (q < dqext) {
                bool any = false;
                for (int n = 0; n < 16; ++n) any = any || dic[(qi * 16 + n) * 2 + 1];
                if (any)
                    for (int n = 0; n < 16; ++n) cells.push_back({dic + (qi * 16 + n) * 2, didq + qi * 16 + n, qprior});
            }
        }
    }
    // cycle and dinucleotide cells are independent given their prior: spread them over host threads
    auto work = [&](size_t lo, size_t hi) {
        int cached_prior = 1 << 30, cached_val = 0;  // cells never observed share one answer per prior
        for (size_t i = lo; i < hi; ++i) {
            const Cell& c = cells[i];
            if (c.cnt[0] == 0 && c.cnt[1] == 0) {
                if (c.prior != cached_prior) { cached_prior = c.prior; cached_val = map_delta(fc, 0, 0, c.prior); }
                *c.out = cached_val;
            } else {
                *c.out = map_delta(fc, c.cnt[0], c.cnt[1], c.prior);
            }
        }
    };
    if (n_threads <= 1 || cells.size() < 256) {
        work(0, cells.size());
    } else {
        std::vector<std::thread> th;
        const size_t per = (cells.size() + n_threads - 1) / n_threads;
        for (int t = 0; t < n_threads; ++t) {
            const size_t lo = std::min(cells.size(), (size_t)t * per), hi = std::min(cells.size(), lo + per);
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto& t : th) t.join();
    }
}

// ---- GATK-style report -------------------------------------------------------------------------------------------------
namespace {
struct ReportRow {
    std::vector<std::string> cell;
};
void write_table(FILE* fh, const char* name, const char* fmt_line, const std::vector<std::string>& head, const std::vector<ReportRow>& rows) {
    std::vector<size_t> w(head.size());
    for (size_t c = 0; c < head.size(); ++c) w[c] = head[c].size();
    for (const ReportRow& r : rows)
        for (size_t c = 0; c < head.size(); ++c) w[c] = std::max(w[c], r.cell[c].size());
    fprintf(fh, "#:GATKTable:%zu:%zu:%s:;\n#:GATKTable:%s:\n", head.size(), rows.size(), fmt_line, name);
    for (size_t c = 0; c < head.size(); ++c) fprintf(fh, "%-*s%s", (int)w[c], head[c].c_str(), c + 1 < head.size() ? "  " : "\n");
    for (const ReportRow& r : rows)
        for (size_t c = 0; c < head.size(); ++c) {
            // strings left aligned, numbers right aligned, like GATK's reports
            const bool text = c < 2 || head[c] == "CovariateValue" || head[c] == "CovariateName" || head[c] == "EventType";
            fprintf(fh, text ? "%-*s%s" : "%*s%s", (int)w[c], r.cell[c].c_str(), c + 1 < head.size() ? "  " : "\n");
        }
    fputc('\n', fh);
}
std::string fixed(double v, int digits) {
    char buf[64];
    snprintf(buf, sizeof buf, "%.*f", digits, v);
    return buf;
}
// GATK's empirical quality: -10 log10((errors + 1) / (observations + 2)), capped at 93
double empirical_q(uint64_t errors, uint64_t obs) {
    const double q = -10.0 * std::log10(((double)errors + 1.0) / ((double)obs + 2.0));
    return q > 93.0 ? 93.0 : q;
}
}  // namespace

bool write_gatk_report(const char* path, int R, int C, const uint64_t* rgc, const uint64_t* qc, const uint64_t* cyc, const uint64_t* dic,
                       const char* const* rg_names) {
    FILE* fh = fopen(path, "w");
    if (!fh) return false;
    auto rgname = [&](int r) { return rg_names && rg_names[r] ? std::string(rg_names[r]) : "rg" + std::to_string(r); };
    fprintf(fh, "#:GATKReport.v1.1:3\n");
    {
        std::vector<ReportRow> rows;
        for (int r = 0; r < R; ++r) {
            const uint64_t err = rgc[2 * r], obs = rgc[2 * r + 1];
            if (!obs) continue;
            long double expected = 0;  // errors the reported qualities predict: sum over q of n_q * 10^(-q/10)
            for (int q = 0; q < 94; ++q) expected += (long double)qc[((size_t)r * 94 + q) * 2 + 1] * std::pow(10.0l, -(long double)q / 10.0l);
            const double est = expected > 0 ? (double)(-10.0l * std::log10(expected / (long double)obs)) : 93.0;
            rows.push_back({{rgname(r), "M", fixed(empirical_q(err, obs), 4), fixed(est, 4), std::to_string(obs), fixed((double)err, 2)}});
        }
        write_table(fh, "RecalTable0", "%s:%s:%.4f:%.4f:%d:%.2f",
                    {"ReadGroup", "EventType", "EmpiricalQuality", "EstimatedQReported", "Observations", "Errors"}, rows);
    }
    {
        std::vector<ReportRow> rows;
        for (int r = 0; r < R; ++r)
            for (int q = 0; q < 94; ++q) {
                const uint64_t err = qc[((size_t)r * 94 + q) * 2], obs = qc[((size_t)r * 94 + q) * 2 + 1];
                if (!obs) continue;
                rows.push_back({{rgname(r), std::to_string(q), "M", fixed(empirical_q(err, obs), 4), std::to_string(obs), fixed((double)err, 2)}});
            }
        write_table(fh, "RecalTable1", "%s:%s:%s:%.4f:%d:%.2f",
                    {"ReadGroup", "QualityScore", "EventType", "EmpiricalQuality", "Observations", "Errors"}, rows);
    }
    {
        std::vector<ReportRow> rows;
        for (int r = 0; r < R; ++r)
            for (int q = 0; q < 94; ++q) {
                for (int d = 0; d < 16; ++d) {  // dinucleotide context: previous base, this base
                    const uint64_t* c = dic + (((size_t)r * 94 + q) * 16 + d) * 2;
                    if (!c[1]) continue;
                    const std::string ctx{"ACGT"[d >> 2], "ACGT"[d & 3]};
                    rows.push_back({{rgname(r), std::to_string(q), ctx, "Context", "M", fixed(empirical_q(c[0], c[1]), 4), std::to_string(c[1]), fixed((double)c[0], 2)}});
                }
                for (int sec = 0; sec < 2; ++sec)  // cycle: 1-based, negative for the second read of a pair (GATK's convention)
                    for (int i = 0; i < C; ++i) {
                        const uint64_t* c = cyc + ((((size_t)r * 94 + q) * 2 + sec) * C + i) * 2;
                        if (!c[1]) continue;
                        rows.push_back({{rgname(r), std::to_string(q), std::to_string(sec ? -(i + 1) : i + 1), "Cycle", "M", fixed(empirical_q(c[0], c[1]), 4),
                                         std::to_string(c[1]), fixed((double)c[0], 2)}});
                    }
            }
        write_table(fh, "RecalTable2", "%s:%s:%s:%s:%s:%.4f:%d:%.2f",
                    {"ReadGroup", "QualityScore", "CovariateValue", "CovariateName", "EventType", "EmpiricalQuality", "Observations", "Errors"}, rows);
    }
    return fclose(fh) == 0;
}

}  // namespace host
}  // namespace kbbq
