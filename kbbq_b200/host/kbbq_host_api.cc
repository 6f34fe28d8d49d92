// kbbq_host_api.cc -- see kbbq_host_api.hh.
#include "kbbq_host_api.hh"

#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstring>
#include <iostream>
#include <thread>

namespace kbbq {

void throw_last(int rc) {
    const std::string msg = kbbq_last_error();
    if (rc == KBBQ_ENOMEM) throw std::bad_alloc();
    if (rc == KBBQ_EINVAL || rc == KBBQ_ERANGE) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
#define CK(call)                      \
    do {                              \
        int rc__ = (call);            \
        if (rc__ != KBBQ_OK) kbbq::throw_last(rc__); \
    } while (0)

static Session* g_current = nullptr;
Session& Session::current() {
    if (!g_current) throw std::runtime_error("kbbq::Session not initialised (no GPU context; this build has no CPU path)");
    return *g_current;
}
void Session::set_current(Session* s) { g_current = s; }

Session::Session(int k, int n_rg, int max_cycle, int device) : k_(k) {
    CK(kbbq_ctx_create(device, k, n_rg, max_cycle, &ctx_));
    if (!g_current) g_current = this;
}
Session::~Session() {
    for (kbbq_dbatch* b : batches_) kbbq_batch_free(ctx_, b);
    kbbq_ctx_destroy(ctx_);
    if (g_current == this) g_current = nullptr;
}

void Session::ensure_resident(const ReadStore& rs) {
    if (resident_ == &rs) return;
    if (resident_) throw std::runtime_error("a Session holds one input at a time");
    const size_t n = rs.n();
    const size_t per_batch = 1u << 21;  // 2 M reads per batch keeps word offsets well inside 32 bits
    const int nt = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    for (size_t lo = 0; lo < n || (n == 0 && lo == 0); lo += per_batch) {
        const size_t hi = std::min(n, lo + per_batch);
        const size_t m = hi - lo;
        std::vector<uint64_t> off(m + 1);
        for (size_t i = 0; i <= m; ++i) off[i] = rs.off[lo + i] - rs.off[lo];
        uint64_t nw = 0;
        uint32_t ml = 0;
        CK(kbbq_pack_size(m, off.data(), &nw, &ml));
        std::vector<uint64_t> bases(std::max<uint64_t>(nw, 1));
        std::vector<uint32_t> nmask(std::max<uint64_t>(nw, 1)), word_off(m + 1), len(std::max<size_t>(m, 1));
        std::vector<uint8_t> qual(std::max<uint64_t>(nw, 1) * 32);
        CK(kbbq_pack_reads(m, rs.seq.data() + rs.off[lo], off.data(), rs.qual.data() + rs.off[lo], bases.data(), nmask.data(),
                           qual.data(), word_off.data(), len.data(), nt));
        kbbq_batch hb;
        hb.n_reads = (uint32_t)m; hb.max_len = ml; hb.n_words = nw;
        hb.bases = bases.data(); hb.nmask = nmask.data(); hb.qual = qual.data(); hb.word_off = word_off.data();
        hb.len = len.data(); hb.rg = rs.rg.data() + lo; hb.second = rs.second.data() + lo;
        kbbq_dbatch* db = nullptr;
        CK(kbbq_batch_upload(ctx_, &hb, &db));
        batches_.push_back(db);
        ranges_.emplace_back(lo, hi);
        word_off_.push_back(std::move(word_off));
        n_words_.push_back(nw);
        if (n == 0) break;
    }
    resident_ = &rs;
}

}  // namespace kbbq

using kbbq::Session;

// ---------------------------------------------------------------------------------------------------------------------
namespace bloom {

static int nt4(unsigned char ch) {
    switch (ch) {
    case 'A': case 'a': case '0': return 0;
    case 'C': case 'c': case '1': return 1;
    case 'G': case 'g': case '2': return 2;
    case 'T': case 't': case '3': return 3;
    default: return 4;
    }
}
size_t Kmer::push_back(char ch) {
    const int c = nt4((unsigned char)ch);
    if (c < 4) {
        x[0] = (x[0] << 2 | (uint64_t)c) & mask;
        x[1] = x[1] >> 2 | (uint64_t)(3 - c) << shift;
        ++s;
    } else {
        reset();
    }
    return s;
}

Bloom::Bloom(unsigned long long projected, double fpr, unsigned long long seed) {
    Session& ses = Session::current();
    slot_ = ses.next_filter_slot();
    if (slot_ > 1) throw std::invalid_argument("this build holds two filters per session (subsampled, trusted)");
    CK(kbbq_bloom_create(ses.ctx(), slot_, projected, fpr, seed));  // KBBQ_EINVAL -> std::invalid_argument (bloom.cc:18-21)
    CK(kbbq_bloom_info(ses.ctx(), slot_, &params));
}
void Bloom::insert(const Kmer& kmer) {
    if (!kmer.valid()) return;  // bloom.hh:395
    const uint64_t key = kmer.get();
    CK(kbbq_bloom_insert_keys(Session::current().ctx(), slot_, &key, 1));
}
bool Bloom::query(const Kmer& kmer) const {
    if (!kmer.valid()) return false;  // bloom.hh:398
    const uint64_t key = kmer.get();
    uint8_t hit = 0;
    CK(kbbq_bloom_query_keys(Session::current().ctx(), slot_, &key, 1, &hit));
    return hit != 0;
}
double Bloom::fprate() const {
    double f = 0;
    CK(kbbq_bloom_fprate(Session::current().ctx(), slot_, &f));
    return f;
}
unsigned long long Bloom::inserted_elements() const {
    uint64_t v = 0;
    CK(kbbq_bloom_inserted(Session::current().ctx(), slot_, &v));
    return v;
}

}  // namespace bloom

// ---------------------------------------------------------------------------------------------------------------------
namespace covariateutils {

std::vector<int> calculate_thresholds(unsigned long long k, long double p, long double) {
    std::vector<int> t(k + 1, 0);
    CK(kbbq_thresholds_from_p((int)k, p, t.data()));
    return t;
}

dq_t CCovariateData::get_dqs() {
    const int R = n_rg, C = max_cycle;
    std::vector<int32_t> m(R), g(R), q((size_t)R * 94), cy((size_t)R * 94 * 2 * C), di((size_t)R * 94 * 16);
    CK(kbbq_fit_model_host(R, C, d_rgc.data(), d_qc.data(), d_cyc.data(), d_dic.data(), m.data(), g.data(), q.data(), cy.data(), di.data()));
    dq_t dq;
    dq.meanq.assign(m.begin(), m.end());
    dq.rgdq.assign(g.begin(), g.end());
    dq.qscoredq.resize(R);
    dq.cycledq.resize(R);
    dq.dinucdq.resize(R);
    for (int r = 0; r < R; ++r) {  // vector extents as the reference grows them
        dq.qscoredq[r].resize(qcov[r].size());
        dq.cycledq[r].resize(cycov[r].size());
        dq.dinucdq[r].resize(dicov[r].size());
        for (size_t qq = 0; qq < qcov[r].size(); ++qq) {
            dq.qscoredq[r][qq] = q[(size_t)r * 94 + qq];
            for (int s = 0; s < 2; ++s) {
                dq.cycledq[r][qq][s].resize(cycov[r][qq][s].size());
                for (size_t c = 0; c < cycov[r][qq][s].size(); ++c) dq.cycledq[r][qq][s][c] = cy[(((size_t)r * 94 + qq) * 2 + s) * C + c];
            }
        }
        for (size_t qq = 0; qq < dicov[r].size(); ++qq) {
            dq.dinucdq[r][qq].resize(dicov[r][qq].size());
            for (size_t n = 0; n < dicov[r][qq].size(); ++n) dq.dinucdq[r][qq][n] = di[((size_t)r * 94 + qq) * 16 + n];
        }
    }
    return dq;
}

}  // namespace covariateutils

namespace bloom {
long double calculate_phit(const Bloom& bf, long double alpha) {
    long double p = 0;
    CK(kbbq_phit(bf.fprate(), alpha, &p));
    return p;
}
}  // namespace bloom

// ---------------------------------------------------------------------------------------------------------------------
namespace readutils {
std::unordered_map<std::string, std::string> CReadData::rg_to_pu{};
std::unordered_map<std::string, int> CReadData::rg_to_int{};
}  // namespace readutils

namespace htsiter {

// read-group / second-of-pair parsing of CReadData's FASTQ constructor (readutils.cc:64-104), same string operations
static void parse_name(const std::string& full, std::string& rg, bool& second) {
    const std::string delim = "_";
    std::string fullname(full);
    size_t current_pos = fullname.find(delim);
    std::string first_name = fullname.substr(0, current_pos);
    rg = "";
    while (rg == "" && current_pos != std::string::npos) {
        fullname = fullname.substr(current_pos + 1);
        current_pos = fullname.find(delim);
        if (fullname.substr(0, 3) == "RG:") {
            size_t last_colon = fullname.find_last_of(":", current_pos);
            rg = fullname.substr(last_colon + 1, current_pos);  // sic: a position used as a count
        }
    }
    std::string tail = first_name.substr(first_name.length() - 2);  // throws std::out_of_range for 1-char names, as the reference
    second = (tail == "/2");
}

static std::shared_ptr<kbbq::ReadStore> load_fastq(const std::string& filename) {
    gzFile gz = filename == "-" ? gzdopen(0, "rb") : gzopen(filename.c_str(), "rb");
    if (!gz) throw std::runtime_error("Error opening file " + filename);
    gzbuffer(gz, 1 << 22);
    std::vector<char> text;
    {
        std::vector<char> buf(1 << 24);
        int got;
        while ((got = gzread(gz, buf.data(), (unsigned)buf.size())) > 0) text.insert(text.end(), buf.begin(), buf.begin() + got);
        gzclose(gz);
    }
    auto rs = std::make_shared<kbbq::ReadStore>();
    rs->path = filename;
    rs->name_off.push_back(0);
    rs->comment_off.push_back(0);
    rs->off.push_back(0);
    readutils::CReadData::rg_to_int.clear();
    readutils::CReadData::rg_to_pu.clear();
    const char* p = text.data();
    const char* end = p + text.size();
    auto line_end = [&](const char* s) { const char* e = (const char*)memchr(s, '\n', end - s); return e ? e : end; };
    while (p < end) {
        while (p < end && *p != '@' && *p != '>') ++p;  // kseq: skip to the next header
        if (p >= end) break;
        const char* le = line_end(p);
        const char* he = le;
        if (he > p && he[-1] == '\r') --he;
        const char* ne = p + 1;
        while (ne < he && *ne != ' ' && *ne != '\t') ++ne;
        rs->names.insert(rs->names.end(), p + 1, ne);
        if (ne < he) rs->comments.insert(rs->comments.end(), ne + 1, he);
        p = le < end ? le + 1 : end;
        // sequence lines until a line starting with '+', '@' or '>'
        while (p < end && *p != '+' && *p != '@' && *p != '>') {
            le = line_end(p);
            const char* se = le;
            if (se > p && se[-1] == '\r') --se;
            rs->seq.insert(rs->seq.end(), p, se);
            p = le < end ? le + 1 : end;
        }
        const size_t slen = rs->seq.size() - rs->off.back();
        size_t qlen = 0;
        if (p < end && *p == '+') {
            le = line_end(p);
            p = le < end ? le + 1 : end;
            while (p < end && qlen < slen) {
                le = line_end(p);
                const char* qe = le;
                if (qe > p && qe[-1] == '\r') --qe;
                for (const char* c = p; c < qe; ++c) rs->qual.push_back((uint8_t)(*c - 33));
                qlen += qe - p;
                p = le < end ? le + 1 : end;
            }
        }
        if (qlen != slen) throw std::runtime_error("truncated quality string in " + filename);  // kseq_read == -2
        rs->name_off.push_back(rs->names.size());
        rs->comment_off.push_back(rs->comments.size());
        rs->off.push_back(rs->seq.size());
        rs->max_len = std::max<uint32_t>(rs->max_len, (uint32_t)slen);
        std::string name(rs->names.data() + rs->name_off[rs->name_off.size() - 2], rs->names.data() + rs->names.size());
        std::string rg;
        bool second;
        parse_name(name, rg, second);
        auto& r2i = readutils::CReadData::rg_to_int;
        if (readutils::CReadData::rg_to_pu.count(rg) == 0) {  // ids in order of first appearance, starting at 0
            const int id = (int)r2i.size();
            r2i[rg] = id;
            readutils::CReadData::rg_to_pu[rg] = rg;
        }
        rs->rg.push_back((uint16_t)r2i[rg]);
        rs->second.push_back(second ? 1 : 0);
    }
    return rs;
}

FastqFile::FastqFile(std::string filename, int nthreads) : store_(load_fastq(filename)), nthreads_(nthreads) {}
FastqFile::~FastqFile() {
    if (out_) {
        flush_pending();
        FILE* f = (FILE*)out_;
        if (f == stdout) fflush(f); else fclose(f);
    }
}
int FastqFile::next() {
    if (cur_ >= store_->n()) return -1;
    ++cur_;
    return (int)(store_->off[cur_] - store_->off[cur_ - 1]);
}
std::string FastqFile::next_str() {
    if (next() < 0) return "";
    return std::string(store_->seq.data() + store_->off[cur_ - 1], store_->seq.data() + store_->off[cur_]);
}
readutils::CReadData FastqFile::get() {
    const size_t i = cur_ - 1;
    readutils::CReadData r;
    r.seq.assign(store_->seq.data() + store_->off[i], store_->seq.data() + store_->off[i + 1]);
    r.qual.assign(store_->qual.data() + store_->off[i], store_->qual.data() + store_->off[i + 1]);
    r.skips.assign(r.seq.size(), false);
    r.errors.assign(r.seq.size(), false);
    std::string full(store_->names.data() + store_->name_off[i], store_->names.data() + store_->name_off[i + 1]);
    r.name = full.substr(0, full.find('_'));
    if (r.name.size() >= 2 && (r.name.substr(r.name.size() - 2) == "/1" || r.name.substr(r.name.size() - 2) == "/2"))
        r.name = r.name.substr(0, r.name.size() - 2);
    for (auto& kv : readutils::CReadData::rg_to_int)
        if (kv.second == store_->rg[i]) r.rg = kv.first;
    r.second = store_->second[i];
    return r;
}
void FastqFile::recalibrate(const std::vector<uint8_t>& qual) { cur_qual_ = qual; }
int FastqFile::open_out(std::string filename) {
    out_ = filename == "-" ? (void*)stdout : (void*)fopen(filename.c_str(), "wb");
    return out_ == nullptr ? -1 : 0;
}
static void format_record(const kbbq::ReadStore& rs, size_t i, const uint8_t* q, std::string& s) {
    // "@" + name + "\n" + seq + "\n+" + comment + "\n" + qual + "\n": the comment moves to the '+' line (htsiter.cc:84)
    s.push_back('@');
    s.append(rs.names.data() + rs.name_off[i], rs.names.data() + rs.name_off[i + 1]);
    s.push_back('\n');
    s.append(rs.seq.data() + rs.off[i], rs.seq.data() + rs.off[i + 1]);
    s.append("\n+");
    s.append(rs.comments.data() + rs.comment_off[i], rs.comments.data() + rs.comment_off[i + 1]);
    s.push_back('\n');
    const size_t L = rs.off[i + 1] - rs.off[i];
    for (size_t j = 0; j < L; ++j) s.push_back((char)(q[j] + 33));
    s.push_back('\n');
}
// The output is a sequence of gzip members (what BGZF is, minus the block index): members can be compressed by
// independent threads and concatenated in input order; zlib/gzip/htslib read the result as one stream.
static bool deflate_member(const std::string& in, std::string& out) {
    z_stream zs;
    std::memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out.resize(deflateBound(&zs, in.size()) + 64);
    zs.next_in = (Bytef*)in.data();
    zs.avail_in = (uInt)in.size();
    zs.next_out = (Bytef*)out.data();
    zs.avail_out = (uInt)out.size();
    const int rc = deflate(&zs, Z_FINISH);
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return rc == Z_STREAM_END;
}

int FastqFile::flush_pending() {
    if (!out_ || pending_.empty()) return 0;
    std::string z;
    if (!deflate_member(pending_, z)) return -1;
    pending_.clear();
    return fwrite(z.data(), 1, z.size(), (FILE*)out_) == z.size() ? 0 : -1;
}

int FastqFile::write() {
    if (!out_) return -1;
    const size_t before = pending_.size();
    format_record(*store_, cur_ - 1, cur_qual_.data(), pending_);
    const int n = (int)(pending_.size() - before);
    if (pending_.size() > (8u << 20) && flush_pending() < 0) return -1;
    return n;
}

int FastqFile::write_all(const kbbq::Session& ses, const std::vector<std::vector<uint8_t>>& newq) {
    if (!out_ || flush_pending() < 0) return -1;
    const kbbq::ReadStore& rs = *store_;
    const size_t chunk = 1u << 15;
    const int nt = (int)std::max(1u, nthreads_ > 0 ? (unsigned)nthreads_ : std::min(16u, std::thread::hardware_concurrency()));
    struct Job { size_t b, lo, hi; };
    std::vector<Job> jobs;
    for (size_t b = 0; b < ses.batch_ranges().size(); ++b)
        for (size_t lo = ses.batch_ranges()[b].first; lo < ses.batch_ranges()[b].second; lo += chunk)
            jobs.push_back({b, lo, std::min(ses.batch_ranges()[b].second, lo + chunk)});
    const size_t wave = (size_t)nt * 4;
    std::vector<std::string> outs(wave);
    for (size_t j0 = 0; j0 < jobs.size(); j0 += wave) {
        const size_t j1 = std::min(jobs.size(), j0 + wave);
        std::vector<std::thread> th;
        std::vector<int> ok(nt, 1);
        for (int t = 0; t < nt; ++t)
            th.emplace_back([&, t] {
                std::string text;
                for (size_t j = j0 + t; j < j1; j += nt) {
                    const Job& jb = jobs[j];
                    text.clear();
                    const size_t first = ses.batch_ranges()[jb.b].first;
                    for (size_t i = jb.lo; i < jb.hi; ++i)
                        format_record(rs, i, newq[jb.b].data() + (size_t)ses.word_off(jb.b)[i - first] * 32, text);
                    if (!deflate_member(text, outs[j - j0])) ok[t] = 0;
                }
            });
        for (auto& t : th) t.join();
        for (int t = 0; t < nt; ++t)
            if (!ok[t]) return -1;
        for (size_t j = j0; j < j1; ++j)
            if (fwrite(outs[j - j0].data(), 1, outs[j - j0].size(), (FILE*)out_) != outs[j - j0].size()) return -1;
    }
    cur_ = rs.n();
    return 0;
}

KmerSubsampler::KmerSubsampler(HTSFile* file_, int k_, double alpha_, uint64_t seed_) : file(file_), k(k_), alpha(alpha_), seed(seed_) {
    std::cerr << "p: " << alpha << std::endl;  // htsiter.hh:143
}

}  // namespace htsiter

// ---------------------------------------------------------------------------------------------------------------------
namespace recalibrateutils {

void subsample_kmers(htsiter::KmerSubsampler& s, bloom::Bloom& sampled) {
    Session& ses = Session::current();
    ses.ensure_resident(s.file->store());
    if (sampled.slot() != KBBQ_FILTER_SAMPLED) throw std::invalid_argument("the first Bloom of a session is the subsampled one");
    // Random::Seed takes a uint32 (the uint64 seed is truncated by overload resolution, htsiter.hh:143)
    CK(kbbq_sampler_seed(ses.ctx(), (uint32_t)s.seed, s.alpha));
    for (kbbq_dbatch* b : ses.batches()) {
        CK(kbbq_pass1_subsample(ses.ctx(), b));
        s.total_kmers += kbbq_batch_slots(b);
    }
    s.not_eof = false;
}

void find_trusted_kmers(htsiter::HTSFile* file, bloom::Bloom& trusted, const bloom::Bloom& sampled, std::vector<int> thresholds, int k) {
    Session& ses = Session::current();
    ses.ensure_resident(file->store());
    if (k != ses.k() || (int)thresholds.size() < k + 1) throw std::invalid_argument("k / thresholds do not match the session");
    (void)trusted; (void)sampled;
    CK(kbbq_set_thresholds(ses.ctx(), thresholds.data()));
    for (kbbq_dbatch* b : ses.batches()) CK(kbbq_pass2_find_trusted(ses.ctx(), b, nullptr));
}

static void densify(covariateutils::CCovariateData& d, int R, int C) {
    // grow the reference's lazily sized vectors to the extents its consume_read would have produced
    using namespace covariateutils;
    d.rgcov.assign(R, covariate_t{0, 0});
    d.qcov.assign(R, CCovariate());
    d.cycov.assign(R, std::vector<std::array<CCovariate, 2>>());
    d.dicov.assign(R, std::vector<CCovariate>());
    for (int r = 0; r < R; ++r) {
        d.rgcov[r] = {d.d_rgc[2 * r], d.d_rgc[2 * r + 1]};
        int qext = 0, dqext = 0;
        for (int q = 0; q < 94; ++q) {
            if (d.d_qc[((size_t)r * 94 + q) * 2 + 1]) qext = q + 1;
            for (int n = 0; n < 16; ++n)
                if (d.d_dic[(((size_t)r * 94 + q) * 16 + n) * 2 + 1]) dqext = q + 1;
        }
        d.qcov[r].resize(qext);
        d.cycov[r].resize(qext);
        d.dicov[r].resize(dqext);
        for (int q = 0; q < qext; ++q) {
            const size_t qi = (size_t)r * 94 + q;
            d.qcov[r][q] = {d.d_qc[qi * 2], d.d_qc[qi * 2 + 1]};
            for (int s = 0; s < 2; ++s) {
                int cext = 0;
                for (int c = 0; c < C; ++c)
                    if (d.d_cyc[((qi * 2 + s) * C + c) * 2 + 1]) cext = c + 1;
                d.cycov[r][q][s].resize(cext);
                for (int c = 0; c < cext; ++c) d.cycov[r][q][s][c] = {d.d_cyc[((qi * 2 + s) * C + c) * 2], d.d_cyc[((qi * 2 + s) * C + c) * 2 + 1]};
            }
            if (q < dqext) {
                bool any = false;
                for (int n = 0; n < 16; ++n) any = any || d.d_dic[(qi * 16 + n) * 2 + 1];
                if (any) {
                    d.dicov[r][q].resize(16);
                    for (int n = 0; n < 16; ++n) d.dicov[r][q][n] = {d.d_dic[(qi * 16 + n) * 2], d.d_dic[(qi * 16 + n) * 2 + 1]};
                }
            }
        }
    }
}

covariateutils::CCovariateData download_covariates(Session& ses, int R, int C) {
    covariateutils::CCovariateData d;
    d.n_rg = R;
    d.max_cycle = C;
    d.d_rgc.resize((size_t)R * 2);
    d.d_qc.resize((size_t)R * 94 * 2);
    d.d_cyc.resize((size_t)R * 94 * 2 * C * 2);
    d.d_dic.resize((size_t)R * 94 * 16 * 2);
    CK(kbbq_covariates_download(ses.ctx(), d.d_rgc.data(), d.d_qc.data(), d.d_cyc.data(), d.d_dic.data()));
    densify(d, R, C);
    return d;
}

covariateutils::CCovariateData get_covariatedata(htsiter::HTSFile* file, const bloom::Bloom& trusted, int k) {
    Session& ses = Session::current();
    ses.ensure_resident(file->store());
    if (k != ses.k()) throw std::invalid_argument("k does not match the session");
    (void)trusted;
    CK(kbbq_covariates_reset(ses.ctx()));
    for (kbbq_dbatch* b : ses.batches()) CK(kbbq_pass3_covariates(ses.ctx(), b, nullptr));
    const int R = (int)std::max<size_t>(1, readutils::CReadData::rg_to_int.size());
    return download_covariates(ses, R, (int)std::max<uint32_t>(1, file->store().max_len));
}

void recalibrate_and_write(htsiter::HTSFile* in, const covariateutils::dq_t& dqs, std::string outfn) {
    Session& ses = Session::current();
    ses.ensure_resident(in->store());
    const int R = (int)dqs.meanq.size(), C = (int)std::max<uint32_t>(1, in->store().max_len);
    std::vector<int32_t> m(R), g(R), q((size_t)R * 94, 0), cy((size_t)R * 94 * 2 * C, 0), di((size_t)R * 94 * 16, 0);
    for (int r = 0; r < R; ++r) {
        m[r] = dqs.meanq[r];
        g[r] = dqs.rgdq[r];
        for (size_t qq = 0; qq < dqs.qscoredq[r].size(); ++qq) q[(size_t)r * 94 + qq] = dqs.qscoredq[r][qq];
        for (size_t qq = 0; qq < dqs.cycledq[r].size(); ++qq)
            for (int s = 0; s < 2; ++s)
                for (size_t c = 0; c < dqs.cycledq[r][qq][s].size(); ++c) cy[(((size_t)r * 94 + qq) * 2 + s) * C + c] = dqs.cycledq[r][qq][s][c];
        for (size_t qq = 0; qq < dqs.dinucdq[r].size(); ++qq)
            for (size_t n = 0; n < dqs.dinucdq[r][qq].size(); ++n) di[((size_t)r * 94 + qq) * 16 + n] = dqs.dinucdq[r][qq][n];
    }
    CK(kbbq_set_model(ses.ctx(), m.data(), g.data(), q.data(), cy.data(), di.data()));
    if (in->open_out(outfn) < 0) return;  // error!! TODO (recalibrateutils.cc:92-95)
    std::vector<std::vector<uint8_t>> newq(ses.batches().size());
    for (size_t b = 0; b < ses.batches().size(); ++b) {
        newq[b].resize(std::max<uint64_t>(ses.n_words(b), 1) * 32);
        CK(kbbq_pass4_recalibrate(ses.ctx(), ses.batches()[b], newq[b].data()));
    }
    // FASTQ fast path: format + compress in parallel, members written in input order
    if (auto* fq = dynamic_cast<htsiter::FastqFile*>(in)) {
        fq->write_all(ses, newq);
        return;
    }
    // generic path, one record at a time like the reference's loop (recalibrateutils.cc:96-104)
    size_t b = 0, idx = 0;
    const kbbq::ReadStore& rs = in->store();
    while (in->next() >= 0) {
        while (idx >= ses.batch_ranges()[b].second) ++b;
        const size_t L = rs.off[idx + 1] - rs.off[idx];
        const uint8_t* src = newq[b].data() + (size_t)ses.word_off(b)[idx - ses.batch_ranges()[b].first] * 32;
        in->recalibrate(std::vector<uint8_t>(src, src + L));
        ++idx;
        if (in->write() < 0) return;  // error! TODO (recalibrateutils.cc:101-103)
    }
}

}  // namespace recalibrateutils
