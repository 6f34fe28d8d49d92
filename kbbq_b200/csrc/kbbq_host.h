// kbbq_host.h -- host-only pieces of the hot path.
//
// These are the few things the reference computes once per run, between passes, in double / x87 long double
// arithmetic or through libstdc++ distributions: filter geometry, salts and pattern table, the sampling RNG's
// seeding, fpr -> p -> thresholds, and the hierarchical model fit.  They feed the kernels (thresholds, geometry,
// lookup tables) and therefore must agree with the reference bit for bit; they stay on the host, in C++, calling
// the same libm / libstdc++ entry points the reference calls.
#pragma once
#include <stdint.h>
#include <vector>

namespace kbbq {
namespace host {

// ---- pattern-blocked Bloom filter geometry (bloom.cc:12-24, bloom.hh:36-56, bloom_filter.hpp:108-160,467-549) ----
struct BloomGeom {
    uint64_t projected;
    double fpr;
    uint64_t seed;
    uint32_t optimal_k;
    uint64_t optimal_table_size;
    uint64_t random_seed;  // blocked_bloom_filter::random_seed_ = seed * 0xA5A5A5A5 + 1
    uint64_t table_bits;   // padded to a multiple of 512
    uint64_t n_blocks;
    uint32_t n_salts;
    uint32_t salts[128];
};
// false when the reference would throw std::invalid_argument (bloom.cc:18-21)
bool bloom_geometry(uint64_t projected, double fpr, uint64_t seed, BloomGeom& g);
// the 65536 x 512-bit pattern table (bloom.hh:189-231); out holds 65536*8 words
void bloom_patterns(const BloomGeom& g, uint64_t* out);
// pattern_blocked_bf::effective_fpp (bloom.hh:318-330)
double effective_fpp(uint64_t table_bits, uint64_t inserted, uint32_t n_salts);
// bloom::calculate_phit (bloom.cc:190-195) and covariateutils::calculate_thresholds (covariateutils.hh:71-85)
long double calculate_phit(double fpr, long double alpha);
void calculate_thresholds(int k, long double p, int* out);

// ---- minion::Random = xoshiro256** (minion.hpp:72-143, 305-375) ----
struct Xoshiro256ss {
    typedef uint64_t result_type;
    uint64_t s[4];
    static constexpr result_type min() { return 0; }
    static constexpr result_type max() { return UINT64_MAX; }
    result_type operator()() { return next(); }
    uint64_t next();
    void seed(uint32_t seed);  // Random::Seed(uint32_t)
};
// minion::create_seed_seq().GenerateOne(): time, pid and two fixed words hashed by SeedSeq32 (minion.hpp:337-408)
uint32_t default_seed(uint64_t time_now, uint32_t pid);
// The state transition is linear over GF(2).  jump_pow2[e] is the matrix of 2^e steps, stored as 256 columns of
// 4 words (column c = image of state bit c).  Built once per process.
struct JumpTables {
    std::vector<uint64_t> pow2;  // 64 * 256 * 4 words
    const uint64_t* matrix(int e) const { return pow2.data() + (size_t)e * 1024; }
};
const JumpTables& jump_tables();
void rng_advance(uint64_t state[4], uint64_t steps);  // O(popcount(steps)) matrix-vector products

// ---- hierarchical model fit (covariateutils.cc:5-19,44-63,78-100,118-145,166-191,204-230) ----
// dense count tables in, dense dq tables out; layouts as in include/kbbq_b200.h
void fit_model(int R, int max_cycle, const uint64_t* rgc, const uint64_t* qc, const uint64_t* cyc, const uint64_t* dic,
               int32_t* meanq, int32_t* rgdq, int32_t* qdq, int32_t* cydq, int32_t* didq, int n_threads);

// ---- GATK-style recalibration report of the count tables (the intent of the reference's unused gatkreport.hh) ----
// Writes "#:GATKReport.v1.1:3": RecalTable0 (per read group), RecalTable1 (per read group x quality), RecalTable2
// (cycle and dinucleotide context rows) with Observations / Errors / EmpiricalQuality columns; only rows that were
// observed.  rg_names may be null (rg<id> is used).  Returns false when the file cannot be written.
bool write_gatk_report(const char* path, int R, int max_cycle, const uint64_t* rgc, const uint64_t* qc, const uint64_t* cyc,
                       const uint64_t* dic, const char* const* rg_names);

// ---- packing ASCII reads into 2-bit words (the host feeder's inner loop) ----
// seq_nt16_int[seq_nt16_table[ch]] (bloom.hh:351): 0..3 for ACGTacgt0123, 4 otherwise
int nt4(unsigned char ch);

}  // namespace host
}  // namespace kbbq
