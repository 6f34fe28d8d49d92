/*
 * oracle/ref_harness.cc -- TEST INFRASTRUCTURE ONLY.
 *
 * A thin C-callable harness around the UNMODIFIED reference sources
 * (/root/reference/src/kbbq/{bloom,covariateutils,readutils,recalibrateutils,
 * htsiter}.cc).  It is compiled by oracle/Makefile into
 * oracle/_ref/libkbbq_ref.so and loaded with ctypes by tests/ and by
 * bench.py's cpu_baseline / --impl reference legs.  Nothing under kbbq_b200/
 * links or loads it.
 *
 * The harness feeds the reference's own pass drivers (recalibrateutils.hh:29-44)
 * from an in-memory read set through a subclass of the reference's abstract
 * htsiter::HTSFile (htsiter.hh:40-49), calling them exactly as main() does
 * (kbbq.cc:264-457), and dumps every intermediate: Bloom words, pattern table,
 * salts, per-read error flags, covariate tables, dq tables, new qualities.
 */
#include "htsiter.hh"
#include "bloom.hh"
#include "covariateutils.hh"
#include "recalibrateutils.hh"
#include "readutils.hh"

#include <chrono>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct ReadSet {
    size_t n = 0;
    const char* seq = nullptr;        // concatenated, no terminators
    const uint64_t* off = nullptr;    // n+1 base offsets
    const uint8_t* qual = nullptr;    // phred values (not +33), same offsets
    const int32_t* rg = nullptr;      // dense read-group id per read
    const uint8_t* second = nullptr;  // 0/1 per read
};

// In-memory HTSFile: what BamFile/FastqFile (htsiter.cc:5-86) are to a file.
class MemFile : public htsiter::HTSFile {
public:
    const ReadSet* rs;
    size_t cur;       // index of the *next* read
    uint8_t* out_q;   // destination for recalibrate(); may be null
    std::vector<std::string>* rgnames;
    MemFile(const ReadSet* r, std::vector<std::string>* names) : rs(r), cur(0), out_q(nullptr), rgnames(names) {}
    int next() override {
        if (cur >= rs->n) return -1;
        ++cur;
        return 0;
    }
    std::string next_str() override {
        if (next() < 0) return "";
        size_t i = cur - 1;
        return std::string(rs->seq + rs->off[i], rs->off[i + 1] - rs->off[i]);
    }
    readutils::CReadData get() override {
        size_t i = cur - 1;
        size_t len = rs->off[i + 1] - rs->off[i];
        readutils::CReadData r;
        r.seq.assign(rs->seq + rs->off[i], len);
        r.qual.assign(rs->qual + rs->off[i], rs->qual + rs->off[i] + len);
        r.skips.assign(len, false);
        r.errors.assign(len, false);
        r.name = "r" + std::to_string(i);
        r.rg = (*rgnames)[rs->rg ? rs->rg[i] : 0];
        r.second = rs->second ? rs->second[i] != 0 : false;
        // same registration the reference's constructors do (readutils.cc:54-58,100-103)
        if (readutils::CReadData::rg_to_pu.count(r.rg) == 0) {
            readutils::CReadData::rg_to_int[r.rg] = readutils::CReadData::rg_to_int.size();
            readutils::CReadData::rg_to_pu[r.rg] = r.rg;
        }
        return r;
    }
    void recalibrate(const std::vector<uint8_t>& q) override {
        if (out_q) std::memcpy(out_q + rs->off[cur - 1], q.data(), q.size());
    }
    int open_out(std::string) override { return 0; }
    int write() override { return 0; }
};

struct Harness {
    int k;
    ReadSet rs;
    std::vector<std::string> rgnames;
    covariateutils::CCovariateData data;
    covariateutils::dq_t dqs;
};

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

void* ref_new(int k, int n_rg) {
    Harness* h = new Harness();
    h->k = k;
    readutils::CReadData::rg_to_int.clear();
    readutils::CReadData::rg_to_pu.clear();
    for (int i = 0; i < n_rg; ++i) {
        std::string name = n_rg == 1 ? std::string("") : "rg" + std::to_string(i);
        h->rgnames.push_back(name);
        // pre-register in id order so that dense id == reference rg_to_int id
        int id = (int)readutils::CReadData::rg_to_int.size();
        readutils::CReadData::rg_to_int[name] = id;
        readutils::CReadData::rg_to_pu[name] = name;
    }
    return h;
}
void ref_free(void* hp) { delete static_cast<Harness*>(hp); }

void ref_set_reads(void* hp, uint64_t n, const char* seq, const uint64_t* off, const uint8_t* qual,
                   const int32_t* rg, const uint8_t* second) {
    Harness* h = static_cast<Harness*>(hp);
    h->rs.n = n; h->rs.seq = seq; h->rs.off = off; h->rs.qual = qual; h->rs.rg = rg; h->rs.second = second;
}

/* ---- Bloom (bloom.hh:385-402, bloom.cc:12-24) ---------------------------- */
void* ref_bloom_new(unsigned long long projected, double fpr) { return new bloom::Bloom(projected, fpr); }
void* ref_bloom_new_seed(unsigned long long projected, double fpr, unsigned long long seed) {
    return new bloom::Bloom(projected, fpr, seed);
}
void ref_bloom_free(void* b) { delete static_cast<bloom::Bloom*>(b); }
/* info[0]=table bits, [1]=blocks, [2]=#salts, [3]=random_seed_, [4]=inserted, [5]=optimal k, [6]=optimal table size */
void ref_bloom_info(void* bp, uint64_t* info, uint32_t* salts) {
    bloom::Bloom* b = static_cast<bloom::Bloom*>(bp);
    info[0] = b->bloom.size();
    info[1] = b->bloom.num_blocks();
    info[2] = b->bloom.salt_.size();
    info[3] = b->bloom.random_seed_;
    info[4] = b->bloom.element_count();
    info[5] = b->params.optimal_parameters.number_of_hashes;
    info[6] = b->params.optimal_parameters.table_size;
    for (size_t i = 0; i < b->bloom.salt_.size(); ++i) salts[i] = b->bloom.salt_[i];
}
const uint64_t* ref_bloom_words(void* bp) {
    return reinterpret_cast<const uint64_t*>(static_cast<bloom::Bloom*>(bp)->bloom.bit_table_.get());
}
const uint64_t* ref_bloom_patterns(void* bp) {
    return reinterpret_cast<const uint64_t*>(static_cast<bloom::Bloom*>(bp)->bloom.patterns.get());
}
void ref_bloom_insert_u64(void* bp, uint64_t key) { static_cast<bloom::Bloom*>(bp)->bloom.insert(key); }
int ref_bloom_contains_u64(void* bp, uint64_t key) { return static_cast<bloom::Bloom*>(bp)->bloom.contains(key); }
void ref_bloom_hashes_u64(void* bp, uint64_t key, uint32_t* out) {
    bloom::Bloom* b = static_cast<bloom::Bloom*>(bp);
    out[0] = b->bloom.block_hash(key);
    out[1] = b->bloom.pattern_hash(key);
}
double ref_bloom_fprate(void* bp) { return static_cast<bloom::Bloom*>(bp)->fprate(); }

/* Kmer (bloom.hh:339-379): push a string, report fwd/rc/get/size */
void ref_kmer(int k, const char* s, uint64_t n, uint64_t* out) {
    bloom::Kmer km(k);
    for (uint64_t i = 0; i < n; ++i) km.push_back(s[i]);
    out[0] = km.x[0]; out[1] = km.x[1]; out[2] = km.get(); out[3] = km.size(); out[4] = km.valid();
}

/* minion::Random (minion.hpp:355-375): state after Seed(s), then n outputs */
void ref_rng(uint32_t seed, uint64_t skip, uint64_t n, uint64_t* state4, uint64_t* out) {
    minion::Random r;
    r.Seed(seed);
    for (int i = 0; i < 4; ++i) state4[i] = r.state()[i];
    r.Discard(skip);
    for (uint64_t i = 0; i < n; ++i) out[i] = r.bits();
}
/* bernoulli draws exactly as KmerSubsampler::next does (htsiter.cc:120) */
void ref_bernoulli(uint32_t seed, double p, uint64_t n, uint8_t* out) {
    minion::Random r;
    r.Seed(seed);
    std::bernoulli_distribution d(p);
    for (uint64_t i = 0; i < n; ++i) out[i] = d(r);
}

/* ---- pass 1 (kbbq.cc:277-280; recalibrateutils.cc:7-13) ------------------- */
double ref_subsample(void* hp, void* sampled, double alpha, uint32_t seed, uint64_t* total_kmers) {
    Harness* h = static_cast<Harness*>(hp);
    MemFile f(&h->rs, &h->rgnames);
    double t0 = now_s();
    htsiter::KmerSubsampler sub(&f, h->k, alpha, seed);
    recalibrateutils::subsample_kmers(sub, *static_cast<bloom::Bloom*>(sampled));
    double t1 = now_s();
    if (total_kmers) *total_kmers = sub.total_kmers;
    return t1 - t0;
}

/* ---- host maths between passes (kbbq.cc:304-313) -------------------------- */
double ref_phit(void* sampled, long double alpha) {
    return (double)bloom::calculate_phit(*static_cast<bloom::Bloom*>(sampled), alpha);
}
/* thresholds from the sampled filter exactly as main() chains them; alpha passed as long double num/den */
void ref_thresholds_from_filter(void* sampled, long double alpha_num, long double alpha_den, int k, int* out) {
    long double alpha = alpha_num / alpha_den;
    long double p = bloom::calculate_phit(*static_cast<bloom::Bloom*>(sampled), alpha);
    std::vector<int> t = covariateutils::calculate_thresholds(k, p);
    for (int i = 0; i <= k; ++i) out[i] = t[i];
}
void ref_thresholds(int k, double p, int* out) {
    std::vector<int> t = covariateutils::calculate_thresholds(k, (long double)p);
    for (int i = 0; i <= k; ++i) out[i] = t[i];
}

/* ---- pass 2 (recalibrateutils.cc:15-40) ------------------------------------ */
double ref_find_trusted(void* hp, void* trusted, void* sampled, const int* thresholds) {
    Harness* h = static_cast<Harness*>(hp);
    MemFile f(&h->rs, &h->rgnames);
    std::vector<int> th(thresholds, thresholds + h->k + 1);
    double t0 = now_s();
    recalibrateutils::find_trusted_kmers(&f, *static_cast<bloom::Bloom*>(trusted),
                                         *static_cast<bloom::Bloom*>(sampled), th, h->k);
    return now_s() - t0;
}
/* per-read flags of infer_read_errors (readutils.cc:173-193) */
void ref_infer_errors(void* hp, void* sampled, const int* thresholds, uint8_t* out) {
    Harness* h = static_cast<Harness*>(hp);
    MemFile f(&h->rs, &h->rgnames);
    std::vector<int> th(thresholds, thresholds + h->k + 1);
    while (f.next() >= 0) {
        readutils::CReadData r = f.get();
        r.infer_read_errors(*static_cast<bloom::Bloom*>(sampled), th, h->k);
        size_t o = h->rs.off[f.cur - 1];
        for (size_t i = 0; i < r.errors.size(); ++i) out[o + i] = r.errors[i];
    }
}

/* ---- pass 3 (recalibrateutils.cc:42-89) ------------------------------------- */
/* per-read flags of get_errors (readutils.cc:238-570); seq_after, if non-null, receives read.seq after the call */
double ref_get_errors(void* hp, void* trusted, uint8_t* out, char* seq_after) {
    Harness* h = static_cast<Harness*>(hp);
    MemFile f(&h->rs, &h->rgnames);
    double t0 = now_s();
    while (f.next() >= 0) {
        readutils::CReadData r = f.get();
        r.get_errors(*static_cast<bloom::Bloom*>(trusted), h->k, 6);
        size_t o = h->rs.off[f.cur - 1];
        for (size_t i = 0; i < r.errors.size(); ++i) out[o + i] = r.errors[i];
        if (seq_after) std::memcpy(seq_after + o, r.seq.data(), r.seq.size());
    }
    return now_s() - t0;
}
double ref_covariates(void* hp, void* trusted) {
    Harness* h = static_cast<Harness*>(hp);
    MemFile f(&h->rs, &h->rgnames);
    double t0 = now_s();
    h->data = recalibrateutils::get_covariatedata(&f, *static_cast<bloom::Bloom*>(trusted), h->k);
    return now_s() - t0;
}
/* --fixed style (kbbq.cc:367-378): errors supplied, then consume_read */
double ref_consume_with_errors(void* hp, const uint8_t* errors) {
    Harness* h = static_cast<Harness*>(hp);
    MemFile f(&h->rs, &h->rgnames);
    h->data = covariateutils::CCovariateData();
    double t0 = now_s();
    while (f.next() >= 0) {
        readutils::CReadData r = f.get();
        size_t o = h->rs.off[f.cur - 1];
        for (size_t i = 0; i < r.errors.size(); ++i) r.errors[i] = errors[o + i] != 0;
        h->data.consume_read(r);
    }
    return now_s() - t0;
}
/* dense dumps: rg[R][2], q[R][94][2], cy[R][94][2][maxlen][2], di[R][94][16][2]; returns #rg */
int ref_dump_covariates(void* hp, int R, int maxlen, uint64_t* rgc, uint64_t* qc, uint64_t* cyc, uint64_t* dic) {
    Harness* h = static_cast<Harness*>(hp);
    covariateutils::CCovariateData& d = h->data;
    for (size_t r = 0; r < d.rgcov.size() && (int)r < R; ++r) { rgc[2 * r] = d.rgcov[r][0]; rgc[2 * r + 1] = d.rgcov[r][1]; }
    for (size_t r = 0; r < d.qcov.size() && (int)r < R; ++r)
        for (size_t q = 0; q < d.qcov[r].size() && q < 94; ++q) {
            qc[(r * 94 + q) * 2] = d.qcov[r][q][0];
            qc[(r * 94 + q) * 2 + 1] = d.qcov[r][q][1];
        }
    for (size_t r = 0; r < d.cycov.size() && (int)r < R; ++r)
        for (size_t q = 0; q < d.cycov[r].size() && q < 94; ++q)
            for (int s = 0; s < 2; ++s)
                for (size_t c = 0; c < d.cycov[r][q][s].size() && (int)c < maxlen; ++c) {
                    size_t idx = (((r * 94 + q) * 2 + s) * maxlen + c) * 2;
                    cyc[idx] = d.cycov[r][q][s][c][0];
                    cyc[idx + 1] = d.cycov[r][q][s][c][1];
                }
    for (size_t r = 0; r < d.dicov.size() && (int)r < R; ++r)
        for (size_t q = 0; q < d.dicov[r].size() && q < 94; ++q)
            for (size_t n = 0; n < d.dicov[r][q].size() && n < 16; ++n) {
                size_t idx = ((r * 94 + q) * 16 + n) * 2;
                dic[idx] = d.dicov[r][q][n][0];
                dic[idx + 1] = d.dicov[r][q][n][1];
            }
    return (int)d.rgcov.size();
}
/* load dense tables back into the reference's containers (for model-fit-only checks) */
void ref_load_covariates(void* hp, int R, int maxlen, const uint64_t* rgc, const uint64_t* qc, const uint64_t* cyc,
                         const uint64_t* dic) {
    Harness* h = static_cast<Harness*>(hp);
    covariateutils::CCovariateData d;
    d.rgcov.resize(R);
    d.qcov.resize(R);
    d.cycov.resize(R);
    d.dicov.resize(R);
    for (int r = 0; r < R; ++r) {
        d.rgcov[r] = {rgc[2 * r], rgc[2 * r + 1]};
        int qmax = -1;
        for (int q = 0; q < 94; ++q) if (qc[(r * 94 + q) * 2 + 1]) qmax = q;
        d.qcov[r].resize(qmax + 1);
        d.cycov[r].resize(qmax + 1);
        int dqmax = -1;
        for (int q = 0; q < 94; ++q)
            for (int n = 0; n < 16; ++n) if (dic[((r * 94 + q) * 16 + n) * 2 + 1]) dqmax = q;
        d.dicov[r].resize(dqmax + 1);
        for (int q = 0; q <= qmax; ++q) {
            d.qcov[r][q] = {qc[(r * 94 + q) * 2], qc[(r * 94 + q) * 2 + 1]};
            for (int s = 0; s < 2; ++s) {
                int cmax = -1;
                for (int c = 0; c < maxlen; ++c) if (cyc[((((size_t)r * 94 + q) * 2 + s) * maxlen + c) * 2 + 1]) cmax = c;
                d.cycov[r][q][s].resize(cmax + 1);
                for (int c = 0; c <= cmax; ++c) {
                    size_t idx = ((((size_t)r * 94 + q) * 2 + s) * maxlen + c) * 2;
                    d.cycov[r][q][s][c] = {cyc[idx], cyc[idx + 1]};
                }
            }
        }
        for (int q = 0; q <= dqmax; ++q) {
            bool any = false;
            for (int n = 0; n < 16; ++n) if (dic[((r * 94 + q) * 16 + n) * 2 + 1]) any = true;
            if (!any) continue;
            d.dicov[r][q].resize(16);
            for (int n = 0; n < 16; ++n) d.dicov[r][q][n] = {dic[((r * 94 + q) * 16 + n) * 2], dic[((r * 94 + q) * 16 + n) * 2 + 1]};
        }
    }
    h->data = d;
}

/* ---- model fit (covariateutils.cc:204-230) ---------------------------------- */
double ref_get_dqs(void* hp) {
    Harness* h = static_cast<Harness*>(hp);
    double t0 = now_s();
    h->dqs = h->data.get_dqs();
    return now_s() - t0;
}
/* dense int32 dumps: meanq[R], rgdq[R], qdq[R][94], cydq[R][94][2][maxlen], didq[R][94][16] */
int ref_dump_dqs(void* hp, int R, int maxlen, int32_t* meanq, int32_t* rgdq, int32_t* qdq, int32_t* cydq, int32_t* didq) {
    Harness* h = static_cast<Harness*>(hp);
    covariateutils::dq_t& d = h->dqs;
    for (size_t r = 0; r < d.meanq.size() && (int)r < R; ++r) { meanq[r] = d.meanq[r]; rgdq[r] = d.rgdq[r]; }
    for (size_t r = 0; r < d.qscoredq.size() && (int)r < R; ++r)
        for (size_t q = 0; q < d.qscoredq[r].size() && q < 94; ++q) qdq[r * 94 + q] = d.qscoredq[r][q];
    for (size_t r = 0; r < d.cycledq.size() && (int)r < R; ++r)
        for (size_t q = 0; q < d.cycledq[r].size() && q < 94; ++q)
            for (int s = 0; s < 2; ++s)
                for (size_t c = 0; c < d.cycledq[r][q][s].size() && (int)c < maxlen; ++c)
                    cydq[((r * 94 + q) * 2 + s) * maxlen + c] = d.cycledq[r][q][s][c];
    for (size_t r = 0; r < d.dinucdq.size() && (int)r < R; ++r)
        for (size_t q = 0; q < d.dinucdq[r].size() && q < 94; ++q)
            for (size_t n = 0; n < d.dinucdq[r][q].size() && n < 16; ++n) didq[(r * 94 + q) * 16 + n] = d.dinucdq[r][q][n];
    return (int)d.meanq.size();
}

/* ---- pass 4 (recalibrateutils.cc:91-105; readutils.cc:572-595) --------------- */
double ref_recalibrate(void* hp, uint8_t* out_q) {
    Harness* h = static_cast<Harness*>(hp);
    MemFile f(&h->rs, &h->rgnames);
    f.out_q = out_q;
    double t0 = now_s();
    recalibrateutils::recalibrate_and_write(&f, h->dqs, "-");
    return now_s() - t0;
}

}  // extern "C"
