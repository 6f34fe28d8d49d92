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

FastqFile::FastqFile(std::string filename, int nthreads) : StoreFile(load_fastq(filename), nthreads) {}

StoreFile::~StoreFile() {
    if (out_) {
        flush_pending(true);
        FILE* f = (FILE*)out_;
        if (f == stdout) fflush(f); else fclose(f);
    }
}
int StoreFile::next() {
    if (cur_ >= store_->n()) return -1;
    ++cur_;
    return (int)(store_->off[cur_] - store_->off[cur_ - 1]);
}
std::string StoreFile::next_str() {
    if (next() < 0) return "";
    return std::string(store_->seq.data() + store_->off[cur_ - 1], store_->seq.data() + store_->off[cur_]);
}
readutils::CReadData StoreFile::get() {
    const size_t i = cur_ - 1;
    readutils::CReadData r;
    r.seq.assign(store_->seq.data() + store_->off[i], store_->seq.data() + store_->off[i + 1]);
    r.qual.assign(store_->qual.data() + store_->off[i], store_->qual.data() + store_->off[i + 1]);
    r.skips.assign(r.seq.size(), false);
    r.errors.assign(r.seq.size(), false);
    r.name.assign(store_->names.data() + store_->name_off[i], store_->names.data() + store_->name_off[i + 1]);
    if (store_->bam_off.empty()) {  // FASTQ: first '_' field without its /1 or /2 (readutils.cc:74-97); BAM: the qname
        r.name = r.name.substr(0, r.name.find('_'));
        if (r.name.size() >= 2 && (r.name.substr(r.name.size() - 2) == "/1" || r.name.substr(r.name.size() - 2) == "/2"))
            r.name = r.name.substr(0, r.name.size() - 2);
    }
    for (auto& kv : readutils::CReadData::rg_to_int)
        if (kv.second == store_->rg[i]) r.rg = kv.first;
    r.second = store_->second[i];
    return r;
}
int StoreFile::open_out(std::string filename) {
    out_ = filename == "-" ? (void*)stdout : (void*)fopen(filename.c_str(), "wb");
    if (!out_) return -1;
    format_header(pending_);
    return 0;
}

// ---- BGZF: gzip members of <= 64 KiB with the "BC" extra field, closed by the empty EOF block --------------------
static const size_t kBgzfInput = 0xff00;
static bool bgzf_block(const char* in, size_t n, std::string& out) {
    z_stream zs;
    std::memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    const size_t start = out.size();
    out.resize(start + 18 + deflateBound(&zs, n) + 8);
    unsigned char* o = (unsigned char*)&out[start];
    static const unsigned char head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    std::memcpy(o, head, 16);
    zs.next_in = (Bytef*)in;
    zs.avail_in = (uInt)n;
    zs.next_out = o + 18;
    zs.avail_out = (uInt)(out.size() - start - 18 - 8);
    const int rc = deflate(&zs, Z_FINISH);
    const size_t clen = zs.total_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return false;
    const size_t bsize = 18 + clen + 8;
    o[16] = (unsigned char)((bsize - 1) & 0xff);
    o[17] = (unsigned char)((bsize - 1) >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)in, (uInt)n), isize = (uint32_t)n;
    std::memcpy(o + 18 + clen, &crc, 4);
    std::memcpy(o + 18 + clen + 4, &isize, 4);
    out.resize(start + bsize);
    return true;
}
static bool bgzf_compress(const std::string& in, std::string& out) {
    for (size_t p = 0; p < in.size(); p += kBgzfInput)
        if (!bgzf_block(in.data() + p, std::min(kBgzfInput, in.size() - p), out)) return false;
    return true;
}

int StoreFile::flush_pending(bool final) {
    if (!out_) return 0;
    std::string z;
    if (!pending_.empty() && !bgzf_compress(pending_, z)) return -1;
    pending_.clear();
    if (final) {
        static const unsigned char eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        z.append((const char*)eof, 28);
    }
    return fwrite(z.data(), 1, z.size(), (FILE*)out_) == z.size() ? 0 : -1;
}

int StoreFile::write() {
    if (!out_) return -1;
    const size_t before = pending_.size();
    format_record(cur_ - 1, cur_qual_.data(), pending_);
    const int n = (int)(pending_.size() - before);
    if (pending_.size() > (8u << 20) && flush_pending(false) < 0) return -1;
    return n;
}

int StoreFile::write_all(const kbbq::Session& ses, const std::vector<std::vector<uint8_t>>& newq) {
    if (!out_ || flush_pending(false) < 0) return -1;
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
                    outs[j - j0].clear();
                    const size_t first = ses.batch_ranges()[jb.b].first;
                    for (size_t i = jb.lo; i < jb.hi; ++i)
                        format_record(i, newq[jb.b].data() + (size_t)ses.word_off(jb.b)[i - first] * 32, text);
                    if (!bgzf_compress(text, outs[j - j0])) ok[t] = 0;
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

void FastqFile::format_record(size_t i, const uint8_t* q, std::string& s) const {
    // "@" + name + "\n" + seq + "\n+" + comment + "\n" + qual + "\n": the comment moves to the '+' line (htsiter.cc:84)
    const kbbq::ReadStore& rs = *store_;
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

// ---- BAM ---------------------------------------------------------------------------------------------------------------
static std::vector<uint8_t> gunzip_all(const std::string& filename) {
    gzFile gz = filename == "-" ? gzdopen(0, "rb") : gzopen(filename.c_str(), "rb");
    if (!gz) throw std::runtime_error("Error opening file " + filename);
    gzbuffer(gz, 1 << 22);
    std::vector<uint8_t> data;
    std::vector<uint8_t> buf(1 << 24);
    int got;
    while ((got = gzread(gz, buf.data(), (unsigned)buf.size())) > 0) data.insert(data.end(), buf.begin(), buf.begin() + got);
    gzclose(gz);
    return data;
}

bool is_bam_file(const std::string& filename) {
    gzFile gz = gzopen(filename.c_str(), "rb");
    if (!gz) throw std::runtime_error("Error opening file " + filename);
    unsigned char head[4] = {0, 0, 0, 0};
    const int n = gzread(gz, head, 4);
    gzclose(gz);
    if (n >= 4 && std::memcmp(head, "BAM\1", 4) == 0) return true;
    if (n >= 1 && head[0] == '@') return false;
    throw std::runtime_error("Error: File format must be bam, cram, or fastq.");
}

// tag value of a Z-typed aux field, or nullptr; `*pos` (optional) receives the offset of the tag's first byte
static const uint8_t* aux_find(const uint8_t* aux, const uint8_t* end, const char tag[2], const uint8_t** field_begin, const uint8_t** field_end) {
    const uint8_t* s = aux;
    while (end - s >= 3) {
        const uint8_t* f = s;
        const uint8_t type = s[2];
        const uint8_t* v = s + 3;
        size_t sz = 0;
        switch (type) {
        case 'A': case 'c': case 'C': sz = 1; break;
        case 's': case 'S': sz = 2; break;
        case 'i': case 'I': case 'f': sz = 4; break;
        case 'd': sz = 8; break;
        case 'Z': case 'H': { const uint8_t* z = v; while (z < end && *z) ++z; sz = (size_t)(z - v) + 1; break; }
        case 'B': {
            if (end - v < 5) return nullptr;
            uint32_t cnt; std::memcpy(&cnt, v + 1, 4);
            const size_t es = (v[0] == 'c' || v[0] == 'C') ? 1 : (v[0] == 's' || v[0] == 'S') ? 2 : 4;
            sz = 5 + es * cnt;
            break;
        }
        default: return nullptr;
        }
        if ((size_t)(end - v) < sz) return nullptr;
        if (f[0] == (uint8_t)tag[0] && f[1] == (uint8_t)tag[1]) {
            if (field_begin) *field_begin = f;
            if (field_end) *field_end = v + sz;
            return type == 'Z' ? v : nullptr;
        }
        s = v + sz;
    }
    return nullptr;
}

struct BamRec {
    uint8_t l_read_name; uint16_t n_cigar, flag; int32_t l_seq;
    const uint8_t *name, *seq, *qual, *aux, *end;
};
static BamRec bam_parse(const uint8_t* rec, size_t len) {
    BamRec r;
    r.l_read_name = rec[8];
    std::memcpy(&r.n_cigar, rec + 12, 2);
    std::memcpy(&r.flag, rec + 14, 2);
    std::memcpy(&r.l_seq, rec + 16, 4);
    r.name = rec + 32;
    r.seq = r.name + r.l_read_name + 4 * (size_t)r.n_cigar;
    r.qual = r.seq + (r.l_seq + 1) / 2;
    r.aux = r.qual + r.l_seq;
    r.end = rec + len;
    return r;
}

static std::shared_ptr<kbbq::ReadStore> load_bam(const std::string& filename, bool use_oq) {
    std::vector<uint8_t> d = gunzip_all(filename);
    auto rs = std::make_shared<kbbq::ReadStore>();
    rs->path = filename;
    if (d.size() < 12 || std::memcmp(d.data(), "BAM\1", 4) != 0) throw std::runtime_error("Error opening file " + filename);
    size_t p = 4;
    int32_t l_text, n_ref;
    std::memcpy(&l_text, &d[p], 4); p += 4 + (size_t)l_text;
    std::memcpy(&n_ref, &d[p], 4); p += 4;
    for (int i = 0; i < n_ref; ++i) {
        int32_t l_name; uint32_t l_ref;
        std::memcpy(&l_name, &d[p], 4); p += 4 + (size_t)l_name;
        std::memcpy(&l_ref, &d[p], 4); p += 4;
        rs->ref_len.push_back(l_ref);
    }
    rs->bam_header.assign(d.begin(), d.begin() + p);
    rs->name_off.push_back(0); rs->comment_off.push_back(0); rs->off.push_back(0); rs->bam_off.push_back(0);
    readutils::CReadData::rg_to_int.clear();
    readutils::CReadData::rg_to_pu.clear();
    static const char nt16[] = "=ACMGRSVTWYHKDBN";
    static const int nt16_int[16] = {4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4};
    while (p + 4 <= d.size()) {
        int32_t block_size;
        std::memcpy(&block_size, &d[p], 4);
        p += 4;
        if (block_size < 32 || p + (size_t)block_size > d.size()) throw std::runtime_error("truncated BAM record in " + filename);
        const uint8_t* rec = &d[p];
        const BamRec r = bam_parse(rec, (size_t)block_size);
        const bool rev = r.flag & 16;
        const std::string name((const char*)r.name);
        // bam_seq_str (readutils.hh:30-42): forward orientation; a reverse-strand record is complemented, then reversed
        const size_t L = (size_t)r.l_seq, s0 = rs->seq.size();
        rs->seq.resize(s0 + L);
        for (size_t i = 0; i < L; ++i) {
            const int nib = (r.seq[i >> 1] >> ((~i & 1) << 2)) & 0xf;
            char c = nt16[nib];
            if (rev) c = nt16_int[nib] < 4 ? "ACGT"[3 - nt16_int[nib]] : 'N';
            rs->seq[s0 + (rev ? L - 1 - i : i)] = c;
        }
        rs->qual.resize(s0 + L);
        if (use_oq) {
            const uint8_t* oq = aux_find(r.aux, r.end, "OQ", nullptr, nullptr);
            if (!oq) {
                std::cerr << "Error: --use-oq was specified but unable to read OQ tag on read " << name << std::endl;
                throw std::invalid_argument("Unable to read OQ tag.");
            }
            for (size_t i = 0; i < L; ++i) rs->qual[s0 + (rev ? L - 1 - i : i)] = (uint8_t)(oq[i] - 33);
        } else {
            for (size_t i = 0; i < L; ++i) rs->qual[s0 + (rev ? L - 1 - i : i)] = r.qual[i];
        }
        const uint8_t* rgz = aux_find(r.aux, r.end, "RG", nullptr, nullptr);
        if (!rgz) {
            std::cerr << "Error: Unable to read RG tag on read " << name << std::endl;
            throw std::invalid_argument("Unable to read RG tag.");
        }
        const std::string rg((const char*)rgz);
        auto& r2i = readutils::CReadData::rg_to_int;
        if (readutils::CReadData::rg_to_pu.count(rg) == 0) {
            const int id = (int)r2i.size();
            r2i[rg] = id;
            readutils::CReadData::rg_to_pu[rg] = rg;
        }
        rs->rg.push_back((uint16_t)r2i[rg]);
        rs->second.push_back((r.flag & 0x80) ? 1 : 0);
        rs->names.insert(rs->names.end(), name.begin(), name.end());
        rs->name_off.push_back(rs->names.size());
        rs->comment_off.push_back(0);
        rs->off.push_back(rs->seq.size());
        rs->max_len = std::max<uint32_t>(rs->max_len, (uint32_t)L);
        rs->bam_blob.insert(rs->bam_blob.end(), rec, rec + block_size);
        rs->bam_off.push_back(rs->bam_blob.size());
        p += (size_t)block_size;
    }
    return rs;
}

BamFile::BamFile(std::string filename, int nthreads, bool use_oq, bool set_oq)
    : StoreFile(load_bam(filename, use_oq), nthreads), use_oq_(use_oq), set_oq_(set_oq) {}

void BamFile::format_header(std::string& out) const { out.append((const char*)store_->bam_header.data(), store_->bam_header.size()); }

void BamFile::format_record(size_t i, const uint8_t* newq, std::string& out) const {
    // BamFile::recalibrate + write (htsiter.cc:11-35,45): optional OQ tag from the record's current qualities, then the
    // new qualities, reversed back for a reverse-strand record
    const kbbq::ReadStore& rs = *store_;
    const uint8_t* rec = rs.bam_blob.data() + rs.bam_off[i];
    const size_t len = rs.bam_off[i + 1] - rs.bam_off[i];
    const BamRec r = bam_parse(rec, len);
    const size_t L = (size_t)r.l_seq;
    std::string body((const char*)rec, len);
    if (set_oq_) {
        std::string oq(L, '!');
        for (size_t j = 0; j < L; ++j) oq[j] = (char)(r.qual[j] + 33);
        const uint8_t *fb = nullptr, *fe = nullptr;
        std::string field = std::string("OQZ") + oq + '\0';
        if (aux_find(r.aux, r.end, "OQ", &fb, &fe) || fb) body.replace((size_t)(fb - rec), (size_t)(fe - fb), field);  // bam_aux_update_str: in place
        else body.append(field);
    }
    uint8_t* q = (uint8_t*)&body[(size_t)(r.qual - rec)];
    const bool rev = r.flag & 16;
    for (size_t j = 0; j < L; ++j) q[j] = newq[rev ? L - 1 - j : j];
    const int32_t block_size = (int32_t)body.size();
    out.append((const char*)&block_size, 4);
    out.append(body);
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
    // fast path: format + compress in parallel, BGZF blocks written in input order
    if (auto* fq = dynamic_cast<htsiter::StoreFile*>(in)) {
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
