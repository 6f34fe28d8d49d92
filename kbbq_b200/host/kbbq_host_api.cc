// kbbq_host_api.cc -- see kbbq_host_api.hh.
#include "kbbq_host_api.hh"
#include "fast_deflate.hh"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <chrono>
#include <cstring>
#include <iostream>
#include <thread>

namespace kbbq {

void* big_alloc(size_t bytes) {
    const size_t huge = (size_t)2 << 20;
    void* p = nullptr;
    if (bytes >= 2 * huge) {
        const size_t rounded = (bytes + huge - 1) & ~(huge - 1);
        if (posix_memalign(&p, huge, rounded) != 0) p = nullptr;
        if (p) madvise(p, rounded, MADV_HUGEPAGE);
    } else {
        p = std::malloc(std::max<size_t>(bytes, 1));
    }
    if (!p) throw std::bad_alloc();
    return p;
}

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

template <class T>
void Pinned<T>::resize(size_t n) {
    release();
    void* p = nullptr;
    CK(kbbq_host_alloc(std::max<size_t>(n, 1) * sizeof(T), &p));
    p_ = static_cast<T*>(p);
    n_ = n;
}
template <class T>
void Pinned<T>::release() {
    if (p_) kbbq_host_free(p_);
    p_ = nullptr;
    n_ = 0;
}
template class Pinned<uint64_t>;
template class Pinned<uint32_t>;
template class Pinned<uint8_t>;

Session::Session(int k, int n_rg, int max_cycle, int device) : Session(k, n_rg, max_cycle, std::vector<int>{device}) {}

Session::Session(int k, int n_rg, int max_cycle, const std::vector<int>& devices, bool background) : devices_(devices), k_(k) {
    if (devices.empty()) throw std::invalid_argument("a Session needs at least one device");
    streaming_.assign(devices.size(), false);
    auto create = [this, k, n_rg, max_cycle] {
        try {
            for (int d : devices_) {
                kbbq_ctx* c = nullptr;
                CK(kbbq_ctx_create(d, k, n_rg, max_cycle, &c));
                ctx_.push_back(c);
            }
        } catch (...) {
            for (kbbq_ctx* c : ctx_) kbbq_ctx_destroy(c);
            ctx_.clear();
            throw;
        }
    };
    if (background) {
        ready_ = std::async(std::launch::async, create).share();
    } else {
        create();
        std::promise<void> done;
        done.set_value();
        ready_ = done.get_future().share();
    }
    if (!g_current) g_current = this;
}
void Session::wait_contexts() const {
    std::shared_future<void> f = ready_;  // a copy per caller: get() on one object from several threads is not allowed
    f.get();                              // rethrows what the creation threw, every time
}
Session::~Session() { release(); }
void Session::release() {
    try { wait_contexts(); } catch (...) {}
    const bool timing = getenv("KBBQ_HOST_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!timing) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::cerr << "[host] session teardown: " << what << " " << std::chrono::duration<double, std::milli>(t1 - t0).count() << " ms" << std::endl;
        t0 = t1;
    };
    for (SessionBatch& b : batches_)
        if (b.dev) kbbq_batch_free(ctx_[b.gpu], b.dev);
    mark("batches freed");
    for (kbbq_ctx* c : ctx_) kbbq_ctx_destroy(c);
    ctx_.clear();
    mark("contexts destroyed");
    batches_.clear();
    mark("host copies released");
    if (g_current == this) g_current = nullptr;
}

// Pack reads [sb.lo, sb.hi) of the store into the kbbq_batch layout: into the batch's own pinned members (kept for
// re-streaming; needs a context), into its ordinary pre_* members (prepack), or into thread-local scratch that lives
// until the next call on this thread.
void Session::pack(const ReadStore& rs, SessionBatch& sb, PackTo to, kbbq_batch& hb) {
    const size_t m = sb.hi - sb.lo;
    const int nt = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency() / (unsigned)std::max<size_t>(1, devices_.size())));
    std::vector<uint64_t> off(m + 1);
    for (size_t i = 0; i <= m; ++i) off[i] = rs.off[sb.lo + i] - rs.off[sb.lo];
    uint64_t nw = 0;
    uint32_t ml = 0;
    CK(kbbq_pack_size(m, off.data(), &nw, &ml));
    sb.n_words = nw;
    sb.max_len = ml;
    sb.word_off.resize(m + 1);
    // kbbq_pack_reads writes every word and byte (padding included): no need to zero hundreds of MB first
    thread_local Column<uint64_t> t_bases;
    thread_local Column<uint32_t> t_nmask, t_len;
    thread_local Column<uint8_t> t_qual;
    uint64_t* bases;
    uint32_t *nmask, *len;
    uint8_t* qual;
    const size_t nw1 = std::max<uint64_t>(nw, 1);
    if (to == kPinned) {
        sb.bases.resize(nw1); sb.nmask.resize(nw1); sb.qual.resize(nw1 * 32); sb.len.resize(std::max<size_t>(m, 1)); sb.qmask.resize(nw1 * 2);
        bases = sb.bases.data(); nmask = sb.nmask.data(); qual = sb.qual.data(); len = sb.len.data();
        sb.has_host_copy = true;
    } else if (to == kOwned) {
        sb.pre_bases.resize(nw1); sb.pre_nmask.resize(nw1); sb.pre_qual.resize(nw1 * 32); sb.pre_len.resize(std::max<size_t>(m, 1));
        bases = sb.pre_bases.data(); nmask = sb.pre_nmask.data(); qual = sb.pre_qual.data(); len = sb.pre_len.data();
        sb.prepacked = true;
    } else {
        t_bases.resize(nw1); t_nmask.resize(nw1); t_qual.resize(nw1 * 32); t_len.resize(std::max<size_t>(m, 1));
        bases = t_bases.data(); nmask = t_nmask.data(); qual = t_qual.data(); len = t_len.data();
    }
    if (nw == 0) { bases[0] = 0; nmask[0] = 0; std::memset(qual, 0, 32); }
    CK(kbbq_pack_reads(m, rs.seq.data() + rs.off[sb.lo], off.data(), rs.qual.data() + rs.off[sb.lo], bases, nmask, qual, sb.word_off.data(), len, nt));
    uint64_t slots = 0;
    for (size_t i = 0; i < m; ++i) slots += len[i] >= (uint32_t)k_ ? len[i] - (uint32_t)k_ + 1 : 0;
    sb.slots = slots;
    hb.n_reads = (uint32_t)m; hb.max_len = ml; hb.n_words = nw;
    hb.bases = bases; hb.nmask = nmask; hb.qual = qual; hb.word_off = sb.word_off.data();
    hb.len = len; hb.rg = rs.rg.data() + sb.lo; hb.second = rs.second.data() + sb.lo;
    hb.qmask = nullptr;  // resident batches: derived on the device (the upload is synchronous anyway)
    if (to == kPinned) {  // streamed batches bring the masks, so that pass 2 never waits for more than its own upload
        CK(kbbq_quality_masks(nw, qual, sb.qmask.data(), nt));
        hb.qmask = sb.qmask.data();
    }
}

static void host_batch_of(const ReadStore& rs, const SessionBatch& sb, kbbq_batch& hb) {
    hb.n_reads = (uint32_t)(sb.hi - sb.lo); hb.max_len = sb.max_len; hb.n_words = sb.n_words;
    hb.bases = sb.bases.data(); hb.nmask = sb.nmask.data(); hb.qual = sb.qual.data(); hb.word_off = sb.word_off.data();
    hb.len = sb.len.data(); hb.rg = rs.rg.data() + sb.lo; hb.second = rs.second.data() + sb.lo; hb.qmask = sb.qmask.data();
}

void Session::cut_batches(const ReadStore& rs) {
    if (cut_for_ == &rs) return;
    if (cut_for_) throw std::runtime_error("a Session holds one input at a time");
    const size_t n = rs.n(), G = devices_.size();
    // at most 2 M reads per batch keeps word offsets well inside 32 bits; several GPUs each get a share of a small input
    size_t per_batch = 1u << 21;
    if (G > 1) per_batch = std::max<size_t>(1, std::min(per_batch, (n + G - 1) / G));
    for (size_t lo = 0; lo < n || (n == 0 && lo == 0); lo += per_batch) {
        SessionBatch sb;
        sb.lo = lo;
        sb.hi = std::min(n, lo + per_batch);
        sb.gpu = (int)(batches_.size() % G);
        batches_.push_back(std::move(sb));
        if (n == 0) break;
    }
    cut_for_ = &rs;
}

void Session::prepack(const ReadStore& rs) {
    if (resident_ == &rs) return;
    cut_batches(rs);
    const size_t G = devices_.size();
    std::vector<std::exception_ptr> err(G);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; ++g)
        th.emplace_back([&, g] {
            try {
                for (SessionBatch& sb : batches_) {
                    if ((size_t)sb.gpu != g) continue;
                    kbbq_batch hb;
                    if (!sb.prepacked) pack(rs, sb, kOwned, hb);
                    break;  // the first batch of this GPU only: the others reuse one scratch buffer as they are uploaded
                }
            } catch (...) {
                err[g] = std::current_exception();
            }
        });
    for (auto& t : th) t.join();
    for (auto& e : err)
        if (e) std::rethrow_exception(e);
}

void Session::ensure_resident(const ReadStore& rs) {
    if (resident_ == &rs) return;
    if (resident_) throw std::runtime_error("a Session holds one input at a time");
    cut_batches(rs);
    wait_contexts();
    const size_t G = ctx_.size();
    // what a GPU's share would take resident (inputs, flags, new qualities), against what it has left beside the filters
    // and the staging arrays of one batch's inserts
    std::vector<uint64_t> need(G, 0), biggest(G, 0);
    for (const SessionBatch& sb : batches_) {
        const uint64_t bases = rs.off[sb.hi] - rs.off[sb.lo], reads = sb.hi - sb.lo;
        const uint64_t words = bases / 32 + reads;  // upper bound
        need[sb.gpu] += words * 96 + reads * 32;
        biggest[sb.gpu] = std::max(biggest[sb.gpu], bases * 16 + (words * 96 + reads * 32));
    }
    for (size_t g = 0; g < G; ++g) {
        uint64_t budget = 0;
        if (const char* e = getenv("KBBQ_RESIDENT_BYTES")) {
            budget = std::strtoull(e, nullptr, 10);
        } else {
            uint64_t free_b = 0, total_b = 0;
            CK(kbbq_device_memory(devices_[g], &free_b, &total_b));
            const uint64_t reserve = biggest[g] + (2ull << 30);
            budget = free_b > reserve ? free_b - reserve : 0;
        }
        streaming_[g] = need[g] > budget;
    }
    // pack (and upload the resident ones), one host thread per GPU
    std::vector<std::exception_ptr> err(G);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; ++g)
        th.emplace_back([&, g] {
            try {
                for (SessionBatch& sb : batches_) {
                    if ((size_t)sb.gpu != g) continue;
                    kbbq_batch hb;
                    if (sb.prepacked && !streaming_[g]) {  // packed while the contexts were being created
                        hb.n_reads = (uint32_t)(sb.hi - sb.lo); hb.max_len = sb.max_len; hb.n_words = sb.n_words;
                        hb.bases = sb.pre_bases.data(); hb.nmask = sb.pre_nmask.data(); hb.qual = sb.pre_qual.data();
                        hb.word_off = sb.word_off.data(); hb.len = sb.pre_len.data();
                        hb.rg = rs.rg.data() + sb.lo; hb.second = rs.second.data() + sb.lo; hb.qmask = nullptr;
                    } else {
                        pack(rs, sb, streaming_[g] ? kPinned : kScratch, hb);
                    }
                    if (!streaming_[g]) CK(kbbq_batch_upload(ctx_[g], &hb, &sb.dev));
                    if (sb.prepacked) {
                        sb.pre_bases = Column<uint64_t>(); sb.pre_nmask = Column<uint32_t>(); sb.pre_len = Column<uint32_t>(); sb.pre_qual = Column<uint8_t>();
                        sb.prepacked = false;
                    }
                }
            } catch (...) {
                err[g] = std::current_exception();
            }
        });
    for (auto& t : th) t.join();
    for (auto& e : err)
        if (e) std::rethrow_exception(e);
    uint64_t first = 0;  // the one serial sampling stream (htsiter.cc:113-129): a batch starts where the batches before it end
    for (SessionBatch& sb : batches_) {
        sb.first_slot = first;
        first += sb.slots;
    }
    resident_ = &rs;
}

void Session::for_each_batch(const std::function<void(kbbq_ctx*, kbbq_dbatch*, SessionBatch&)>& fn) {
    if (!resident_) throw std::runtime_error("no input is resident");
    const size_t G = ctx_.size();
    std::vector<std::exception_ptr> err(G);
    auto work = [&](size_t g) {
        try {
            std::vector<SessionBatch*> mine;
            for (SessionBatch& sb : batches_)
                if ((size_t)sb.gpu == g) mine.push_back(&sb);
            if (!streaming_[g]) {
                for (SessionBatch* sb : mine) fn(ctx_[g], sb->dev, *sb);
                return;
            }
            // streamed: batch i+1 travels (copy stream) while the kernels of batch i run; two batches in device memory
            auto upload = [&](SessionBatch* sb) {
                kbbq_batch hb;
                host_batch_of(*resident_, *sb, hb);
                kbbq_dbatch* d = nullptr;
                CK(kbbq_batch_upload_stream(ctx_[g], &hb, &d));
                return d;
            };
            kbbq_dbatch* next = mine.empty() ? nullptr : upload(mine[0]);
            for (size_t i = 0; i < mine.size(); ++i) {
                kbbq_dbatch* cur = next;
                next = nullptr;
                try {
                    if (i + 1 < mine.size()) next = upload(mine[i + 1]);
                    fn(ctx_[g], cur, *mine[i]);
                } catch (...) {
                    kbbq_sync(ctx_[g]);
                    kbbq_batch_free(ctx_[g], cur);
                    if (next) kbbq_batch_free(ctx_[g], next);
                    throw;
                }
                kbbq_batch_free(ctx_[g], cur);  // stream ordered: after the kernels that read it
            }
            CK(kbbq_sync(ctx_[g]));
        } catch (...) {
            err[g] = std::current_exception();
        }
    };
    if (G == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; ++g) th.emplace_back(work, g);
        for (auto& t : th) t.join();
    }
    for (auto& e : err)
        if (e) std::rethrow_exception(e);
}

void Session::merge_filter(int which) {
    wait_contexts();
    if (ctx_.size() > 1) CK(kbbq_bloom_or_allreduce_local(ctx_.data(), (int)ctx_.size(), which));
}

kbbq_dbatch* Session::upload_range(const ReadStore& rs, size_t lo, size_t hi, int g) {
    SessionBatch sb;
    sb.lo = lo;
    sb.hi = hi;
    sb.gpu = g;
    kbbq_batch hb;
    pack(rs, sb, kScratch, hb);
    kbbq_dbatch* d = nullptr;
    CK(kbbq_batch_upload(ctx(g), &hb, &d));
    return d;
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
    for (kbbq_ctx* c : ses.contexts())  // every GPU holds a full-size filter; partial results are merged after the pass
        CK(kbbq_bloom_create(c, slot_, projected, fpr, seed));  // KBBQ_EINVAL -> std::invalid_argument (bloom.cc:18-21)
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

// ---- reading the input ------------------------------------------------------------------------------------------------
// The whole decompressed input is brought into memory once (the reference re-reads the file for every pass).  BGZF
// input (bgzip'd FASTQ, every BAM) is a sequence of independent gzip members of <= 64 KiB: the members are located by
// their BSIZE fields and inflated by a pool of threads.  Plain gzip is one serial zlib stream; plain text is read.
// KBBQ_HOST_TIMING=1: phase times of the host side on stderr
struct PhaseTimer {
    const bool on = getenv("KBBQ_HOST_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::cerr << "[host] " << what << " " << std::chrono::duration<double, std::milli>(t1 - t0).count() << " ms" << std::endl;
        t0 = t1;
    }
};
static int pool_size(int nthreads) {
    return (int)std::max(1u, nthreads > 0 ? (unsigned)nthreads : std::min(16u, std::thread::hardware_concurrency()));
}
template <class F>
static void parallel_for(size_t n, int nt, F f) {  // f(begin, end, thread index) over contiguous ranges
    nt = (int)std::min<size_t>((size_t)std::max(nt, 1), std::max<size_t>(n, 1));
    if (nt <= 1) { f((size_t)0, n, 0); return; }
    // an exception leaving a std::thread is std::terminate: carry the first one back to the caller instead
    std::vector<std::exception_ptr> err((size_t)nt);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([=, &err] {
            try {
                f(n * t / nt, n * (t + 1) / nt, t);
            } catch (...) {
                err[(size_t)t] = std::current_exception();
            }
        });
    for (auto& t : th) t.join();
    for (auto& e : err)
        if (e) std::rethrow_exception(e);
}

}  // namespace htsiter (reopened below)

kbbq::Bytes& kbbq::Bytes::operator=(Bytes&& o) noexcept {
    if (this != &o) {
        release();
        mapped_ = o.mapped_;
        heap_ = std::move(o.heap_);
        vec_ = std::move(o.vec_);
        n_ = o.n_;
        p_ = mapped_ ? o.p_ : !vec_.empty() ? vec_.data() : heap_.data();
        o.p_ = nullptr;
        o.n_ = 0;
        o.mapped_ = false;
    }
    return *this;
}
void kbbq::Bytes::release() {
    if (mapped_ && p_) munmap(const_cast<uint8_t*>(p_), n_);
    p_ = nullptr;
    n_ = 0;
    mapped_ = false;
}
bool kbbq::Bytes::map_file(const std::string& path, Bytes& out) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size <= 0) { close(fd); return false; }
    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return false;
    madvise(m, (size_t)st.st_size, MADV_WILLNEED);
    out.release();
    out.p_ = (const uint8_t*)m;
    out.n_ = (size_t)st.st_size;
    out.mapped_ = true;
    return true;
}
kbbq::Bytes kbbq::Bytes::uninitialised(size_t n) {
    Bytes b;
    b.heap_.resize(n);
    b.p_ = b.heap_.data();
    b.n_ = n;
    return b;
}
kbbq::Bytes kbbq::Bytes::from(std::vector<uint8_t>&& v) {
    Bytes b;
    b.vec_ = std::move(v);
    b.p_ = b.vec_.data();
    b.n_ = b.vec_.size();
    return b;
}

namespace htsiter {

struct BgzfBlock { size_t cdata, clen, out_off; uint32_t isize; };
// member boundaries of a BGZF file, or false when some member is not a BGZF block (then zlib reads it serially)
static bool bgzf_index(const kbbq::Bytes& raw, std::vector<BgzfBlock>& blocks, size_t& total) {
    size_t o = 0;
    total = 0;
    while (o < raw.size()) {
        if (raw.size() - o < 18 || raw[o] != 0x1f || raw[o + 1] != 0x8b || raw[o + 2] != 8 || !(raw[o + 3] & 4)) return false;
        const size_t xlen = raw[o + 10] | (raw[o + 11] << 8);
        size_t bsize = 0;
        for (size_t x = o + 12; x + 4 <= o + 12 + xlen && x + 4 <= raw.size();) {  // extra subfields: SI1 SI2 SLEN data
            const size_t slen = raw[x + 2] | (raw[x + 3] << 8);
            if (raw[x] == 'B' && raw[x + 1] == 'C' && slen == 2 && x + 6 <= raw.size()) bsize = (size_t)(raw[x + 4] | (raw[x + 5] << 8)) + 1;
            x += 4 + slen;
        }
        if (bsize < 12 + xlen + 8 || o + bsize > raw.size() || (raw[o + 3] & ~4)) return false;  // other gzip flags: not BGZF
        BgzfBlock b;
        b.cdata = o + 12 + xlen;
        b.clen = bsize - 12 - xlen - 8;
        std::memcpy(&b.isize, &raw[o + bsize - 4], 4);
        b.out_off = total;
        total += b.isize;
        blocks.push_back(b);
        o += bsize;
    }
    return !blocks.empty();
}

static kbbq::Bytes read_all_decompressed(const std::string& filename, int nthreads) {
    kbbq::Bytes raw;
    if (filename != "-" && kbbq::Bytes::map_file(filename, raw)) {
        if (raw.size() < 2 || raw[0] != 0x1f || raw[1] != 0x8b) return raw;  // not compressed: parsed where it lies
        std::vector<BgzfBlock> blocks;
        size_t total = 0;
        if (bgzf_index(raw, blocks, total)) {
            kbbq::Bytes data = kbbq::Bytes::uninitialised(total);  // every byte is written by the pool below
            uint8_t* out = data.writable();
            std::vector<char> bad((size_t)pool_size(nthreads), 0);
            parallel_for(blocks.size(), pool_size(nthreads), [&](size_t lo, size_t hi, int t) {
                z_stream zs;
                for (size_t i = lo; i < hi; ++i) {
                    const BgzfBlock& b = blocks[i];
                    if (b.isize == 0) continue;
                    std::memset(&zs, 0, sizeof zs);
                    if (inflateInit2(&zs, -15) != Z_OK) { bad[t] = 1; return; }
                    zs.next_in = (Bytef*)&raw[b.cdata];
                    zs.avail_in = (uInt)b.clen;
                    zs.next_out = out + b.out_off;
                    zs.avail_out = b.isize;
                    const int rc = inflate(&zs, Z_FINISH);
                    inflateEnd(&zs);
                    if (rc != Z_STREAM_END || zs.avail_out != 0) { bad[t] = 1; return; }
                }
            });
            for (char c : bad)
                if (c) throw std::runtime_error("corrupt BGZF block in " + filename);
            return data;
        }
    }
    // plain gzip, or something that cannot be mapped (a pipe, standard input): one serial stream through zlib, which
    // passes uncompressed bytes through unchanged
    gzFile gz = filename == "-" ? gzdopen(0, "rb") : gzopen(filename.c_str(), "rb");
    if (!gz) throw std::runtime_error("Error opening file " + filename);
    gzbuffer(gz, 1 << 22);
    std::vector<uint8_t> data, buf(1 << 24);
    int got;
    while ((got = gzread(gz, buf.data(), (unsigned)buf.size())) > 0) data.insert(data.end(), buf.begin(), buf.begin() + got);
    gzclose(gz);
    return kbbq::Bytes::from(std::move(data));
}

// Read groups are numbered in order of first appearance in the file (readutils.cc:99-102).  Parsers working on
// ranges of the file keep a local table each; merging the tables range by range gives the file's order.
struct RgTable {
    std::vector<std::string> names;                 // local id -> name, in order of first appearance in the range
    std::unordered_map<std::string, uint16_t> ids;
    uint16_t id(const std::string& rg) {
        auto it = ids.find(rg);
        if (it != ids.end()) return it->second;
        if (names.size() >= 65535) throw std::runtime_error("more than 65535 read groups");
        const uint16_t v = (uint16_t)names.size();
        names.push_back(rg);
        ids.emplace(rg, v);
        return v;
    }
};
static std::vector<std::vector<uint16_t>> merge_rg_tables(const std::vector<RgTable>& parts) {
    readutils::CReadData::rg_to_int.clear();
    readutils::CReadData::rg_to_pu.clear();
    auto& r2i = readutils::CReadData::rg_to_int;
    std::vector<std::vector<uint16_t>> remap(parts.size());
    for (size_t t = 0; t < parts.size(); ++t)
        for (const std::string& rg : parts[t].names) {
            if (readutils::CReadData::rg_to_pu.count(rg) == 0) {  // ids in order of first appearance, starting at 0
                const int id = (int)r2i.size();
                r2i[rg] = id;
                readutils::CReadData::rg_to_pu[rg] = rg;
            }
            remap[t].push_back((uint16_t)r2i[rg]);
        }
    return remap;
}

struct FastqPart {  // what one thread parsed: same columns as ReadStore, offsets relative to the part
    std::vector<char> names, comments, seq;
    std::vector<uint8_t> qual;
    std::vector<uint64_t> name_end, comment_end, seq_end;
    std::vector<uint16_t> rg;
    std::vector<uint8_t> second;
    RgTable rgs;
    uint32_t max_len = 0;
    bool strict = true;  // false: something other than 4-line records was seen; the serial kseq-style parser takes over
    size_t stopped_at = 0;
};

// kseq_read semantics (kseq.hh) on [p, end): header '@' or '>', name up to the first blank, rest = comment, sequence
// lines until a line starting with '+', '@' or '>', then as many quality characters as bases.  `strict4` parses only
// plain 4-line records and gives up (strict = false) on anything else.
static void parse_fastq_range(const char* base, size_t begin, size_t end_, bool strict4, const std::string& filename, FastqPart& out) {
    const char* p = base + begin;
    const char* end = base + end_;
    auto line_end = [&](const char* s) { const char* e = (const char*)memchr(s, '\n', end - s); return e ? e : end; };
    {   // a 4-line record is name + 2 x bases + a few separators: reserve once instead of growing by doubling
        const size_t bytes = end_ - begin;
        out.seq.reserve(bytes / 2 + 64);
        out.qual.reserve(bytes / 2 + 64);
        out.names.reserve(bytes / 8 + 64);
        const size_t recs = bytes / 64 + 16;
        out.name_end.reserve(recs); out.comment_end.reserve(recs); out.seq_end.reserve(recs); out.rg.reserve(recs); out.second.reserve(recs);
    }
    while (p < end) {
        if (strict4) {
            if (*p != '@') { out.strict = false; return; }
        } else {
            while (p < end && *p != '@' && *p != '>') ++p;  // kseq: skip to the next header
            if (p >= end) break;
        }
        const char* rec_start = p;
        const char* le = line_end(p);
        const char* he = le;
        if (he > p && he[-1] == '\r') --he;
        const char* ne = p + 1;
        while (ne < he && *ne != ' ' && *ne != '\t') ++ne;
        const size_t name0 = out.names.size(), comment0 = out.comments.size(), seq0 = out.seq.size();
        out.names.insert(out.names.end(), p + 1, ne);
        if (ne < he) out.comments.insert(out.comments.end(), ne + 1, he);
        p = le < end ? le + 1 : end;
        // sequence lines until a line starting with '+', '@' or '>'
        int seq_lines = 0;
        while (p < end && *p != '+' && *p != '@' && *p != '>') {
            le = line_end(p);
            const char* se = le;
            if (se > p && se[-1] == '\r') --se;
            out.seq.insert(out.seq.end(), p, se);
            p = le < end ? le + 1 : end;
            ++seq_lines;
        }
        const size_t slen = out.seq.size() - seq0;
        size_t qlen = 0;
        int qual_lines = 0;
        if (p < end && *p == '+') {
            le = line_end(p);
            p = le < end ? le + 1 : end;
            const size_t q0 = out.qual.size();
            out.qual.resize(q0 + slen);
            while (p < end && qlen < slen) {
                le = line_end(p);
                const char* qe = le;
                if (qe > p && qe[-1] == '\r') --qe;
                const size_t take = std::min<size_t>((size_t)(qe - p), slen - qlen);
                for (size_t c = 0; c < take; ++c) out.qual[q0 + qlen + c] = (uint8_t)(p[c] - 33);
                if ((size_t)(qe - p) > take) qlen = slen + 1;  // longer than the sequence: kseq_read == -2
                else qlen += take;
                p = le < end ? le + 1 : end;
                ++qual_lines;
            }
        }
        if (strict4 && (seq_lines != 1 || qual_lines != 1 || slen == 0 || qlen != slen)) {
            out.names.resize(name0); out.comments.resize(comment0); out.seq.resize(seq0); out.qual.resize(seq0);
            out.strict = false;
            out.stopped_at = (size_t)(rec_start - base);
            return;
        }
        if (qlen != slen) throw std::runtime_error("truncated quality string in " + filename);  // kseq_read == -2
        out.name_end.push_back(out.names.size());
        out.comment_end.push_back(out.comments.size());
        out.seq_end.push_back(out.seq.size());
        out.max_len = std::max<uint32_t>(out.max_len, (uint32_t)slen);
        std::string name(out.names.data() + name0, out.names.data() + out.names.size());
        std::string rg;
        bool second;
        parse_name(name, rg, second);
        out.rg.push_back(out.rgs.id(rg));
        out.second.push_back(second ? 1 : 0);
    }
    out.stopped_at = end_;
}

// start of the first 4-line record at or after `from`: a line starting with '@' whose third line starts with '+' and
// whose second and fourth lines have equal lengths (a quality line may start with '@', a sequence line never with '+')
static size_t next_record_start(const char* base, size_t size, size_t from) {
    const char* end = base + size;
    const char* p = base + from;
    if (from > 0) {  // move to a line start
        const char* nl = (const char*)memchr(p - 1, '\n', end - (p - 1));
        if (!nl) return size;
        p = nl + 1;
    }
    auto line_end = [&](const char* s) { const char* e = (const char*)memchr(s, '\n', end - s); return e ? e : end; };
    for (int tries = 0; p < end && tries < 8; ++tries) {
        if (*p == '@') {
            const char* l1 = line_end(p);
            const char* s0 = l1 < end ? l1 + 1 : end;
            const char* l2 = line_end(s0);
            const char* pl = l2 < end ? l2 + 1 : end;
            if (pl < end && *pl == '+' && s0 < end && *s0 != '@' && *s0 != '+' && *s0 != '>') {
                const char* l3 = line_end(pl);
                const char* q0 = l3 < end ? l3 + 1 : end;
                const char* l4 = line_end(q0);
                if ((l2 - s0) == (l4 - q0) && l2 > s0) return (size_t)(p - base);
            }
        }
        const char* le = line_end(p);
        p = le < end ? le + 1 : end;
    }
    return size;  // no plausible record start near: the caller falls back to the serial parser
}

static std::shared_ptr<kbbq::ReadStore> load_fastq(const std::string& filename, int nthreads) {
    PhaseTimer pt_;
    const kbbq::Bytes text = read_all_decompressed(filename, nthreads);
    pt_.mark("fastq: read + decompress");
    const char* base = (const char*)text.data();
    const size_t size = text.size();
    const int nt = (int)std::min<size_t>((size_t)pool_size(nthreads), std::max<size_t>(1, size >> 20));  // >= 1 MiB per thread
    std::vector<FastqPart> parts;
    bool ok = nt > 1;
    if (ok) {
        std::vector<size_t> cut(nt + 1, size);
        cut[0] = 0;
        for (int t = 1; t < nt && ok; ++t) {
            cut[t] = next_record_start(base, size, size * t / nt);
            if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
        }
        parts.resize(nt);
        if (pt_.on) { std::cerr << "[host] fastq cuts:"; for (size_t c : cut) std::cerr << " " << c; std::cerr << std::endl; }
        parallel_for((size_t)nt, nt, [&](size_t lo, size_t hi, int) {
            for (size_t t = lo; t < hi; ++t) parse_fastq_range(base, cut[t], cut[t + 1], true, filename, parts[t]);
        });
        for (const FastqPart& pt : parts) ok = ok && pt.strict;
        if (!(size && base[0] == '@')) ok = false;
    }
    if (!ok) {  // multi-line records, FASTA-style headers, junk between records ...: one thread, kseq semantics
        parts.assign(1, FastqPart());
        parse_fastq_range(base, 0, size, false, filename, parts[0]);
    }
    pt_.mark(ok ? "fastq: parse (parallel, 4-line records)" : "fastq: parse (serial, kseq semantics)");
    // stitch the parts together
    auto rs = std::make_shared<kbbq::ReadStore>();
    rs->path = filename;
    std::vector<RgTable> tables;
    for (FastqPart& pt : parts) tables.push_back(std::move(pt.rgs));
    const std::vector<std::vector<uint16_t>> remap = merge_rg_tables(tables);
    std::vector<size_t> r0(parts.size() + 1, 0), n0(parts.size() + 1, 0), c0(parts.size() + 1, 0), s0(parts.size() + 1, 0);
    for (size_t t = 0; t < parts.size(); ++t) {
        r0[t + 1] = r0[t] + parts[t].seq_end.size();
        n0[t + 1] = n0[t] + parts[t].names.size();
        c0[t + 1] = c0[t] + parts[t].comments.size();
        s0[t + 1] = s0[t] + parts[t].seq.size();
        rs->max_len = std::max(rs->max_len, parts[t].max_len);
    }
    const size_t n = r0.back();
    rs->names.resize(n0.back()); rs->comments.resize(c0.back()); rs->seq.resize(s0.back()); rs->qual.resize(s0.back());
    rs->name_off.assign(n + 1, 0); rs->comment_off.assign(n + 1, 0); rs->off.assign(n + 1, 0);
    rs->rg.resize(n); rs->second.resize(n);
    parallel_for(parts.size(), (int)parts.size(), [&](size_t lo, size_t hi, int) {
        for (size_t t = lo; t < hi; ++t) {
            const FastqPart& pt = parts[t];
            if (!pt.names.empty()) std::memcpy(&rs->names[n0[t]], pt.names.data(), pt.names.size());
            if (!pt.comments.empty()) std::memcpy(&rs->comments[c0[t]], pt.comments.data(), pt.comments.size());
            if (!pt.seq.empty()) std::memcpy(&rs->seq[s0[t]], pt.seq.data(), pt.seq.size());
            if (!pt.qual.empty()) std::memcpy(&rs->qual[s0[t]], pt.qual.data(), pt.qual.size());
            for (size_t i = 0; i < pt.seq_end.size(); ++i) {
                rs->name_off[r0[t] + i + 1] = n0[t] + pt.name_end[i];
                rs->comment_off[r0[t] + i + 1] = c0[t] + pt.comment_end[i];
                rs->off[r0[t] + i + 1] = s0[t] + pt.seq_end[i];
                rs->rg[r0[t] + i] = remap[t][pt.rg[i]];
                rs->second[r0[t] + i] = pt.second[i];
            }
        }
    });
    pt_.mark("fastq: stitch");
    return rs;
}

FastqFile::FastqFile(std::string filename, int nthreads) : StoreFile(load_fastq(filename, nthreads), nthreads) {}

StoreFile::~StoreFile() { close_out(); }
int StoreFile::close_out() {
    if (!out_) return 0;
    int rc = flush_pending(true);
    FILE* f = (FILE*)out_;
    out_ = nullptr;
    if (f == stdout) { if (fflush(f) != 0) rc = -1; }
    else if (fclose(f) != 0) rc = -1;
    return rc;
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
    if (store_->bam_header.empty()) {  // FASTQ: first '_' field without its /1 or /2 (readutils.cc:74-97); BAM: the qname
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
// The deflate stream inside a block: this repo's own encoder (fast_deflate.hh: long matches + one dynamic Huffman code
// per block, ~3x zlib's level 1 at its ratio) unless KBBQ_BGZF_LEVEL asks for zlib: 0..9 = that zlib level (htslib's
// bgzf_open(.., "w") compresses at zlib's default, 6), "huff" = this encoder without matching.
struct BgzfMode { int zlib_level; int min_match; };  // zlib_level < 0: fast_deflate with min_match
static BgzfMode bgzf_mode() {
    static const BgzfMode mode = [] {
        const char* e = getenv("KBBQ_BGZF_LEVEL");
        if (e && *e >= '0' && *e <= '9') return BgzfMode{*e - '0', 0};
        if (e && !strcmp(e, "huff")) return BgzfMode{-1, 1 << 30};
        return BgzfMode{-1, 8};
    }();
    return mode;
}
static bool bgzf_block(const char* in, size_t n, std::string& out) {
    const BgzfMode mode = bgzf_mode();
    const size_t start = out.size();
    size_t clen = 0;
    static const unsigned char head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    if (mode.zlib_level < 0) {
        const size_t cap = kbbq::fdeflate::bound(n);
        out.resize(start + 18 + cap + 8);
        clen = kbbq::fdeflate::compress((const uint8_t*)in, n, (uint8_t*)&out[start] + 18, cap, mode.min_match);
        if (clen == 0) return false;
    } else {
        z_stream zs;
        std::memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, mode.zlib_level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
        out.resize(start + 18 + deflateBound(&zs, n) + 8);
        zs.next_in = (Bytef*)in;
        zs.avail_in = (uInt)n;
        zs.next_out = (Bytef*)&out[start] + 18;
        zs.avail_out = (uInt)(out.size() - start - 18 - 8);
        const int rc = deflate(&zs, Z_FINISH);
        clen = zs.total_out;
        deflateEnd(&zs);
        if (rc != Z_STREAM_END) return false;
    }
    unsigned char* o = (unsigned char*)&out[start];
    std::memcpy(o, head, 16);
    const size_t bsize = 18 + clen + 8;
    if (bsize > 65536) return false;
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

int StoreFile::write_all(const kbbq::Session& ses, const std::vector<kbbq::Pinned<uint8_t>>& newq) {
    if (!out_ || flush_pending(false) < 0) return -1;
    const kbbq::ReadStore& rs = *store_;
    // Jobs of 8 Ki records are formatted and compressed by a pool (dynamic hand-out: records and compression ratios
    // vary), and written by this thread in input order as they complete; at most `window` finished jobs wait.
    const size_t chunk = 1u << 13;
    const int nt = pool_size(nthreads_);
    struct Job { size_t b, lo, hi; };
    std::vector<Job> jobs;
    for (size_t b = 0; b < ses.batches().size(); ++b)
        for (size_t lo = ses.batches()[b].lo; lo < ses.batches()[b].hi; lo += chunk)
            jobs.push_back({b, lo, std::min(ses.batches()[b].hi, lo + chunk)});
    const size_t window = (size_t)nt * 8;
    std::vector<std::string> outs(jobs.size());
    std::vector<char> done(jobs.size(), 0);
    std::mutex mu;
    std::condition_variable cv_done, cv_room;
    std::atomic<size_t> next{0};
    size_t written = 0;  // guarded by mu
    bool failed = false;
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&] {
            std::string text;
            for (;;) {
                const size_t j = next.fetch_add(1);
                if (j >= jobs.size()) return;
                {   // do not run further ahead of the writer than the window
                    std::unique_lock<std::mutex> lk(mu);
                    cv_room.wait(lk, [&] { return failed || j < written + window; });
                    if (failed) return;
                }
                const Job& jb = jobs[j];
                text.clear();
                const size_t first = ses.batches()[jb.b].lo;
                for (size_t i = jb.lo; i < jb.hi; ++i)
                    format_record(i, newq[jb.b].data() + (size_t)ses.batches()[jb.b].word_off[i - first] * 32, text);
                std::string z;
                const bool ok = bgzf_compress(text, z);
                std::lock_guard<std::mutex> lk(mu);
                outs[j].swap(z);
                done[j] = 1;
                if (!ok) failed = true;
                cv_done.notify_all();
                if (!ok) cv_room.notify_all();
            }
        });
    int rc = 0;
    for (size_t j = 0; j < jobs.size(); ++j) {
        std::string z;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return failed || done[j]; });
            if (failed) { rc = -1; break; }
            z.swap(outs[j]);
        }
        if (fwrite(z.data(), 1, z.size(), (FILE*)out_) != z.size()) {
            std::lock_guard<std::mutex> lk(mu);
            failed = true;
            rc = -1;
            cv_room.notify_all();
            break;
        }
        std::lock_guard<std::mutex> lk(mu);
        written = j + 1;
        cv_room.notify_all();
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        if (rc) failed = true;
        cv_room.notify_all();
    }
    for (auto& t : th) t.join();
    if (rc) return rc;
    cur_ = rs.n();
    return 0;
}

void FastqFile::format_record(size_t i, const uint8_t* q, std::string& s) const {
    // "@" + name + "\n" + seq + "\n+" + comment + "\n" + qual + "\n": the comment moves to the '+' line (htsiter.cc:84)
    const kbbq::ReadStore& rs = *store_;
    const size_t nl = rs.name_off[i + 1] - rs.name_off[i], cl = rs.comment_off[i + 1] - rs.comment_off[i], L = rs.off[i + 1] - rs.off[i];
    const size_t at = s.size();
    s.resize(at + 1 + nl + 1 + L + 2 + cl + 1 + L + 1);
    char* o = &s[at];
    *o++ = '@';
    std::memcpy(o, rs.names.data() + rs.name_off[i], nl); o += nl;
    *o++ = '\n';
    std::memcpy(o, rs.seq.data() + rs.off[i], L); o += L;
    *o++ = '\n';
    *o++ = '+';
    std::memcpy(o, rs.comments.data() + rs.comment_off[i], cl); o += cl;
    *o++ = '\n';
    for (size_t j = 0; j < L; ++j) o[j] = (char)(q[j] + 33);
    o[L] = '\n';
}

// ---- BAM ---------------------------------------------------------------------------------------------------------------
bool is_bam_file(const std::string& filename) {
    if (filename == "-")  // the loaders read stdin, but it cannot be peeked at and rewound: the format must be known
        throw std::runtime_error("Error: reading from standard input is not supported by this build; give a file name.");
    gzFile gz = gzopen(filename.c_str(), "rb");
    if (!gz) throw std::runtime_error("Error opening file " + filename);
    unsigned char head[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int n = gzread(gz, head, 8);
    gzclose(gz);
    if (n >= 4 && std::memcmp(head, "BAM\1", 4) == 0) return true;
    if (n >= 4 && std::memcmp(head, "CRAM", 4) == 0)
        throw std::runtime_error("Error: CRAM input is not supported by this build (bam or fastq only).");
    // a SAM text file also starts with '@' (@HD / @SQ / @RG / @PG / @CO then a tab): the reference's hts_detect_format
    // calls it sam and main() rejects it (kbbq.cc:172-190); a FASTQ name may start the same way but is never followed by a tab
    if (n >= 4 && head[0] == '@' && head[3] == '\t' &&
        (!std::memcmp(head + 1, "HD", 2) || !std::memcmp(head + 1, "SQ", 2) || !std::memcmp(head + 1, "RG", 2) ||
         !std::memcmp(head + 1, "PG", 2) || !std::memcmp(head + 1, "CO", 2)))
        throw std::runtime_error("Error: File format must be bam, cram, or fastq.");
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

static std::shared_ptr<kbbq::ReadStore> load_bam(const std::string& filename, bool use_oq, int nthreads) {
    auto rs = std::make_shared<kbbq::ReadStore>();
    rs->path = filename;
    PhaseTimer pt_;
    rs->bam_raw = read_all_decompressed(filename, nthreads);
    pt_.mark("bam: read + decompress");
    const kbbq::Bytes& d = rs->bam_raw;
    if (d.size() < 12 || std::memcmp(d.data(), "BAM\1", 4) != 0) throw std::runtime_error("Error opening file " + filename);
    size_t p = 4;
    int32_t l_text, n_ref;
    std::memcpy(&l_text, &d[p], 4); p += 4 + (size_t)l_text;
    if (p + 4 > d.size()) throw std::runtime_error("truncated BAM header in " + filename);
    std::memcpy(&n_ref, &d[p], 4); p += 4;
    for (int i = 0; i < n_ref; ++i) {
        int32_t l_name; uint32_t l_ref;
        if (p + 4 > d.size()) throw std::runtime_error("truncated BAM header in " + filename);
        std::memcpy(&l_name, &d[p], 4); p += 4 + (size_t)l_name;
        if (p + 4 > d.size()) throw std::runtime_error("truncated BAM header in " + filename);
        std::memcpy(&l_ref, &d[p], 4); p += 4;
        rs->ref_len.push_back(l_ref);
    }
    rs->bam_header.assign(d.data(), d.data() + p);
    // record boundaries and column offsets: one cheap serial walk over the block_size chain
    rs->name_off.push_back(0); rs->comment_off.push_back(0); rs->off.push_back(0);
    while (p + 4 <= d.size()) {
        int32_t block_size;
        std::memcpy(&block_size, &d[p], 4);
        p += 4;
        if (block_size < 32 || p + (size_t)block_size > d.size()) throw std::runtime_error("truncated BAM record in " + filename);
        const uint8_t* rec = &d[p];
        int32_t l_seq;
        std::memcpy(&l_seq, rec + 16, 4);
        const size_t l_name = rec[8] ? strnlen((const char*)rec + 32, rec[8]) : 0;
        rs->bam_rec_off.push_back(p);
        rs->bam_rec_len.push_back((uint32_t)block_size);
        rs->name_off.push_back(rs->name_off.back() + l_name);
        rs->comment_off.push_back(0);
        rs->off.push_back(rs->off.back() + (size_t)std::max(l_seq, 0));
        rs->max_len = std::max<uint32_t>(rs->max_len, (uint32_t)std::max(l_seq, 0));
        p += (size_t)block_size;
    }
    pt_.mark("bam: record index");
    const size_t n = rs->bam_rec_off.size();
    rs->names.resize(rs->name_off.back());
    rs->seq.resize(rs->off.back());
    rs->qual.resize(rs->off.back());
    rs->rg.resize(n);
    rs->second.resize(n);
    // the records themselves, converted by a pool of threads
    const int nt = (int)std::min<size_t>((size_t)pool_size(nthreads), std::max<size_t>(1, n >> 12));
    std::vector<RgTable> tables((size_t)nt);
    std::vector<std::string> errors((size_t)nt);
    std::vector<size_t> lo_of((size_t)nt + 1, n);
    static const char nt16[] = "=ACMGRSVTWYHKDBN";
    static const int nt16_int[16] = {4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4};
    parallel_for(n, nt, [&](size_t lo, size_t hi, int t) {
        lo_of[t] = lo;
        for (size_t i = lo; i < hi; ++i) {
            const uint8_t* rec = &d[rs->bam_rec_off[i]];
            const BamRec r = bam_parse(rec, rs->bam_rec_len[i]);
            const bool rev = r.flag & 16;
            const size_t L = (size_t)std::max(r.l_seq, 0), s0 = rs->off[i];
            if (r.aux > r.end) { errors[t] = "truncated BAM record in " + filename; return; }
            std::memcpy(rs->names.data() + rs->name_off[i], r.name, rs->name_off[i + 1] - rs->name_off[i]);
            // bam_seq_str (readutils.hh:30-42): forward orientation; a reverse-strand record is complemented, then reversed
            for (size_t j = 0; j < L; ++j) {
                const int nib = (r.seq[j >> 1] >> ((~j & 1) << 2)) & 0xf;
                char c = nt16[nib];
                if (rev) c = nt16_int[nib] < 4 ? "ACGT"[3 - nt16_int[nib]] : 'N';
                rs->seq[s0 + (rev ? L - 1 - j : j)] = c;
            }
            if (use_oq) {
                const uint8_t* oq = aux_find(r.aux, r.end, "OQ", nullptr, nullptr);
                if (!oq || strnlen((const char*)oq, (size_t)(r.end - oq)) < L) {
                    errors[t] = std::string("Error: --use-oq was specified but unable to read OQ tag on read ") + std::string((const char*)r.name) +
                                "\nUnable to read OQ tag.";
                    return;
                }
                for (size_t j = 0; j < L; ++j) rs->qual[s0 + (rev ? L - 1 - j : j)] = (uint8_t)(oq[j] - 33);
            } else {
                for (size_t j = 0; j < L; ++j) rs->qual[s0 + (rev ? L - 1 - j : j)] = r.qual[j];
            }
            const uint8_t* rgz = aux_find(r.aux, r.end, "RG", nullptr, nullptr);
            if (!rgz) {
                errors[t] = std::string("Error: Unable to read RG tag on read ") + std::string((const char*)r.name) + "\nUnable to read RG tag.";
                return;
            }
            rs->rg[i] = tables[t].id(std::string((const char*)rgz));
            rs->second[i] = (r.flag & 0x80) ? 1 : 0;
        }
    });
    for (const std::string& e : errors)
        if (!e.empty()) {  // the reference prints the first line and throws std::invalid_argument with the second (readutils.cc:17-52)
            const size_t nl = e.find('\n');
            if (nl != std::string::npos) {
                std::cerr << e.substr(0, nl) << std::endl;
                throw std::invalid_argument(e.substr(nl + 1));
            }
            throw std::runtime_error(e);
        }
    pt_.mark("bam: convert records");
    const std::vector<std::vector<uint16_t>> remap = merge_rg_tables(tables);
    parallel_for((size_t)nt, nt, [&](size_t lo, size_t hi, int) {
        for (size_t t = lo; t < hi; ++t) {
            const size_t end = t + 1 < (size_t)nt ? lo_of[t + 1] : n;
            for (size_t i = lo_of[t]; i < end && i < n; ++i) rs->rg[i] = remap[t][rs->rg[i]];
        }
    });
    return rs;
}

BamFile::BamFile(std::string filename, int nthreads, bool use_oq, bool set_oq)
    : StoreFile(load_bam(filename, use_oq, nthreads), nthreads), use_oq_(use_oq), set_oq_(set_oq) {}

void BamFile::format_header(std::string& out) const { out.append((const char*)store_->bam_header.data(), store_->bam_header.size()); }

void BamFile::format_record(size_t i, const uint8_t* newq, std::string& out) const {
    // BamFile::recalibrate + write (htsiter.cc:11-35,45): optional OQ tag from the record's current qualities, then the
    // new qualities, reversed back for a reverse-strand record
    const kbbq::ReadStore& rs = *store_;
    const uint8_t* rec = rs.bam_raw.data() + rs.bam_rec_off[i];
    const size_t len = rs.bam_rec_len[i];
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
    for (kbbq_ctx* c : ses.contexts()) CK(kbbq_sampler_seed(c, (uint32_t)s.seed, s.alpha));
    ses.for_each_batch([](kbbq_ctx* c, kbbq_dbatch* d, kbbq::SessionBatch& sb) {
        CK(kbbq_sampler_seek(c, sb.first_slot));  // one serial stream over the whole file, whichever GPU holds the batch
        CK(kbbq_pass1_subsample(c, d));
    });
    for (const kbbq::SessionBatch& sb : ses.batches()) s.total_kmers += sb.slots;
    ses.merge_filter(KBBQ_FILTER_SAMPLED);
    s.not_eof = false;
}

void find_trusted_kmers(htsiter::HTSFile* file, bloom::Bloom& trusted, const bloom::Bloom& sampled, std::vector<int> thresholds, int k) {
    Session& ses = Session::current();
    ses.ensure_resident(file->store());
    if (k != ses.k() || (int)thresholds.size() < k + 1) throw std::invalid_argument("k / thresholds do not match the session");
    (void)trusted; (void)sampled;
    for (kbbq_ctx* c : ses.contexts()) CK(kbbq_set_thresholds(c, thresholds.data()));
    ses.for_each_batch([](kbbq_ctx* c, kbbq_dbatch* d, kbbq::SessionBatch&) { CK(kbbq_pass2_find_trusted(c, d, nullptr)); });
    ses.merge_filter(KBBQ_FILTER_TRUSTED);
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
    if (ses.n_gpus() > 1) {  // the SUM merge of the per-GPU partial tables (a few hundred KB each) happens here, on the host
        std::vector<uint64_t> a(d.d_rgc.size()), b(d.d_qc.size()), c(d.d_cyc.size()), e(d.d_dic.size());
        for (int g = 1; g < ses.n_gpus(); ++g) {
            CK(kbbq_covariates_download(ses.ctx(g), a.data(), b.data(), c.data(), e.data()));
            for (size_t i = 0; i < a.size(); ++i) d.d_rgc[i] += a[i];
            for (size_t i = 0; i < b.size(); ++i) d.d_qc[i] += b[i];
            for (size_t i = 0; i < c.size(); ++i) d.d_cyc[i] += c[i];
            for (size_t i = 0; i < e.size(); ++i) d.d_dic[i] += e[i];
        }
    }
    densify(d, R, C);
    return d;
}

covariateutils::CCovariateData get_covariatedata(htsiter::HTSFile* file, const bloom::Bloom& trusted, int k) {
    Session& ses = Session::current();
    ses.ensure_resident(file->store());
    if (k != ses.k()) throw std::invalid_argument("k does not match the session");
    (void)trusted;
    for (kbbq_ctx* c : ses.contexts()) CK(kbbq_covariates_reset(c));
    ses.for_each_batch([](kbbq_ctx* c, kbbq_dbatch* d, kbbq::SessionBatch&) { CK(kbbq_pass3_covariates(c, d, nullptr)); });
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
    for (kbbq_ctx* c : ses.contexts()) CK(kbbq_set_model(c, m.data(), g.data(), q.data(), cy.data(), di.data()));
    if (in->open_out(outfn) < 0) return;  // error!! TODO (recalibrateutils.cc:92-95)
    std::vector<kbbq::Pinned<uint8_t>> newq(ses.batches().size());  // pinned: the device-to-host copies overlap the lookups
    for (size_t b = 0; b < ses.batches().size(); ++b) newq[b].resize(std::max<uint64_t>(ses.batches()[b].n_words, 1) * 32);
    ses.for_each_batch([&](kbbq_ctx* c, kbbq_dbatch* d, kbbq::SessionBatch& sb) {
        CK(kbbq_pass4_recalibrate(c, d, newq[&sb - ses.batches().data()].data()));
    });
    // fast path: format + compress in parallel, BGZF blocks written in input order
    if (auto* fq = dynamic_cast<htsiter::StoreFile*>(in)) {
        if (fq->write_all(ses, newq) < 0) throw std::runtime_error("writing the output failed");
        return;
    }
    // generic path, one record at a time like the reference's loop (recalibrateutils.cc:96-104)
    size_t b = 0, idx = 0;
    const kbbq::ReadStore& rs = in->store();
    while (in->next() >= 0) {
        while (idx >= ses.batches()[b].hi) ++b;
        const size_t L = rs.off[idx + 1] - rs.off[idx];
        const uint8_t* src = newq[b].data() + (size_t)ses.batches()[b].word_off[idx - ses.batches()[b].lo] * 32;
        in->recalibrate(std::vector<uint8_t>(src, src + L));
        ++idx;
        if (in->write() < 0) return;  // error! TODO (recalibrateutils.cc:101-103)
    }
}

}  // namespace recalibrateutils
