// kbbq_host_api.hh -- C++ host mirror of the reference's library surface for the hot path, over the C ABI.
//
// Same namespaces, names, argument meaning and error behaviour as the reference's headers
// (include/kbbq/{bloom,htsiter,readutils,covariateutils,recalibrateutils}.hh), so that the reference's main()
// (src/kbbq/kbbq.cc:229-457) -- reproduced in kbbq_cli.cc -- reads the same.  What differs is where work happens:
// the input is parsed ONCE into packed batches that stay resident in HBM (the reference re-opens and re-parses the
// file for each of its five passes), and the four pass drivers launch the sm_100a kernels through
// include/kbbq_b200.h instead of looping over records.  There is no CPU fallback: without a GPU the Session
// constructor throws.
#pragma once
#include <array>
#include <cstdint>
#include <cstdlib>
#include <future>
#include <memory>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/kbbq_b200.h"

#define KBBQ_MAX_KMER 32
#define KBBQ_MAXQ 93
#define INFER_ERROR_BAD_QUAL 2

namespace kbbq {

// Uninitialised memory for the big host arrays.  From 4 MiB on it is aligned to 2 MiB and marked MADV_HUGEPAGE: with
// transparent huge pages in "madvise" mode the threads that fill hundreds of MB take 512 times fewer page faults, and
// the kernel hands the memory back at exit in as many fewer steps.  Released with std::free.
void* big_alloc(size_t bytes);

// A column of the ReadStore: like std::vector for what the store needs, but resize() leaves new bytes uninitialised --
// the loaders fill every byte from several threads, and zero-filling hundreds of MB first would be a serial pass of
// page faults in the calling thread.
template <class T>
class Column {
public:
    Column() = default;
    Column(const Column&) = delete;
    Column& operator=(const Column&) = delete;
    Column(Column&& o) noexcept : p_(std::move(o.p_)), n_(o.n_), cap_(o.cap_) { o.n_ = o.cap_ = 0; }
    Column& operator=(Column&& o) noexcept {
        p_ = std::move(o.p_);
        n_ = o.n_;
        cap_ = o.cap_;
        o.n_ = o.cap_ = 0;
        return *this;
    }
    void resize(size_t n) {
        if (n > cap_) {
            T* p = static_cast<T*>(big_alloc(n * sizeof(T)));
            if (n_) std::memcpy(p, p_.get(), n_ * sizeof(T));
            p_.reset(p);
            cap_ = n;
        }
        n_ = n;
    }
    T* data() { return p_.get(); }
    const T* data() const { return p_.get(); }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    T& operator[](size_t i) { return p_.get()[i]; }
    const T& operator[](size_t i) const { return p_.get()[i]; }

private:
    struct Free { void operator()(T* p) const { std::free(p); } };
    std::unique_ptr<T, Free> p_;
    size_t n_ = 0, cap_ = 0;
};

// The decompressed input: a read-only mapping of the file itself when it is plain (no copy, pages come straight from
// the page cache as the parser threads touch them), else heap memory filled by the inflater threads.
class Bytes {
public:
    Bytes() = default;
    Bytes(const Bytes&) = delete;
    Bytes& operator=(const Bytes&) = delete;
    Bytes(Bytes&& o) noexcept { *this = std::move(o); }
    Bytes& operator=(Bytes&& o) noexcept;
    ~Bytes() { release(); }
    static bool map_file(const std::string& path, Bytes& out);  // false: not a regular file, or the mapping failed
    static Bytes uninitialised(size_t n);
    static Bytes from(std::vector<uint8_t>&& v);
    const uint8_t* data() const { return p_; }
    uint8_t* writable() { return mapped_ ? nullptr : const_cast<uint8_t*>(p_); }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    const uint8_t& operator[](size_t i) const { return p_[i]; }

private:
    void release();
    const uint8_t* p_ = nullptr;
    size_t n_ = 0;
    bool mapped_ = false;
    Column<uint8_t> heap_;
    std::vector<uint8_t> vec_;
};

// Everything parsed from the input, kept for all passes: the reference's per-record kseq_t / CReadData, columnar.
struct ReadStore {
    std::string path;
    Column<char> names, comments, seq;  // concatenated, no terminators
    Column<uint8_t> qual;               // phred
    std::vector<uint64_t> name_off, comment_off, off;  // n+1 each
    std::vector<uint16_t> rg;
    std::vector<uint8_t> second;
    size_t n() const { return off.empty() ? 0 : off.size() - 1; }
    uint32_t max_len = 0;
    // BAM input only: the header as it was read, the whole uncompressed stream, and where each record's body (core +
    // data, without its block_size field) sits in it; reference lengths from the header
    std::vector<uint8_t> bam_header;
    Bytes bam_raw;
    std::vector<uint64_t> bam_rec_off;
    std::vector<uint32_t> bam_rec_len;
    std::vector<uint64_t> ref_len;
};

// Page-locked host array (kbbq_host_alloc): uploads from it and downloads into it are asynchronous.
template <class T>
class Pinned {
public:
    Pinned() = default;
    Pinned(const Pinned&) = delete;
    Pinned& operator=(const Pinned&) = delete;
    Pinned(Pinned&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    Pinned& operator=(Pinned&& o) noexcept { if (this != &o) { release(); p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; } return *this; }
    ~Pinned() { release(); }
    void resize(size_t n);  // contents are NOT preserved
    void release();
    T* data() { return p_; }
    const T* data() const { return p_; }
    size_t size() const { return n_; }
    T& operator[](size_t i) { return p_[i]; }
    const T& operator[](size_t i) const { return p_[i]; }

private:
    T* p_ = nullptr;
    size_t n_ = 0;
};

// One batch of the input as the Session holds it: reads [lo, hi) of the store, dealt to one GPU.
struct SessionBatch {
    size_t lo = 0, hi = 0;
    int gpu = 0;
    uint64_t n_words = 0, slots = 0, first_slot = 0;  // k-mer slots of the batch / of all batches before it (the sampling stream)
    uint32_t max_len = 0;
    std::vector<uint32_t> word_off;                   // kept for the writer
    kbbq_dbatch* dev = nullptr;                       // resident copy, or nullptr while evicted
    // packed host copy (pinned), kept only when the GPU's batches do not all fit in its memory
    Pinned<uint64_t> bases;
    Pinned<uint32_t> nmask, len, qmask;
    Pinned<uint8_t> qual;
    bool has_host_copy = false;
    // packed ahead of the upload (Session::prepack), ordinary memory: the contexts may not exist yet
    Column<uint64_t> pre_bases;
    Column<uint32_t> pre_nmask, pre_len;
    Column<uint8_t> pre_qual;
    bool prepacked = false;
};

// The GPU contexts of one run plus the batches of one input.
//
// Reads are cut into batches of at most 2 M reads and dealt round-robin to the GPUs; every GPU holds full-size filters
// and count tables, and the partial results are merged after passes 1-3 (kbbq_bloom_or_allreduce_local over peer memory;
// the count tables, a few hundred KB, are summed on the host) -- SURVEY.md section 8(e).  Batches stay resident across
// the passes while they fit; a GPU whose share does not fit keeps packed, pinned host copies instead and streams them
// through its memory once per pass, the upload of batch i+1 (kbbq_batch_upload_stream) overlapping the kernels of batch i
// -- as the reference re-reads its file once per pass (kbbq.cc:232,258,336,365,456), minus the parsing.
// KBBQ_RESIDENT_BYTES overrides the per-GPU budget (tests force streaming with it).
class Session {
public:
    Session(int k, int n_rg, int max_cycle, int device = 0);
    // background = true: the contexts are created by a side thread (CUDA's start-up takes hundreds of ms) and the
    // constructor returns at once; whatever needs a context waits for it, and reports its failure.  prepack() does not.
    Session(int k, int n_rg, int max_cycle, const std::vector<int>& devices, bool background = false);
    ~Session();
    Session(const Session&) = delete;
    void release();  // frees the batches and destroys the contexts now (what the destructor does)
    int n_gpus() const { return (int)devices_.size(); }
    kbbq_ctx* ctx(int g = 0) const { wait_contexts(); return ctx_[g]; }
    const std::vector<kbbq_ctx*>& contexts() const { wait_contexts(); return ctx_; }
    int k() const { return k_; }
    // Host-only head start for ensure_resident: cuts the store into batches and packs every GPU's first batch into
    // ordinary memory, touching no context -- what a caller does while the contexts are still being created.
    void prepack(const ReadStore& rs);
    // pack the store into batches once (uploading those that stay resident); later calls are no-ops
    void ensure_resident(const ReadStore& rs);
    std::vector<SessionBatch>& batches() { return batches_; }
    const std::vector<SessionBatch>& batches() const { return batches_; }
    bool streaming(int g) const { return streaming_[g]; }
    // Run fn(ctx, device batch, batch) for every batch, one host thread per GPU, each GPU's batches in input order;
    // streamed batches are uploaded one ahead and freed after their turn.  The first exception of any thread is rethrown.
    void for_each_batch(const std::function<void(kbbq_ctx*, kbbq_dbatch*, SessionBatch&)>& fn);
    // merge the partial filters of all GPUs (no-op on one GPU)
    void merge_filter(int which);
    // pack + upload reads [lo, hi) of ANOTHER store to GPU g (--fixed mode's second file); the caller frees the batch
    kbbq_dbatch* upload_range(const ReadStore& rs, size_t lo, size_t hi, int g);
    static Session& current();           // the session the mirror classes use
    static void set_current(Session* s);
    int next_filter_slot() { return next_slot_++; }

private:
    enum PackTo { kScratch, kPinned, kOwned };  // thread-local scratch / the batch's pinned members / its pre_* members
    void pack(const ReadStore& rs, SessionBatch& sb, PackTo to, kbbq_batch& hb);
    void cut_batches(const ReadStore& rs);
    void wait_contexts() const;
    mutable std::shared_future<void> ready_;
    const ReadStore* cut_for_ = nullptr;
    std::vector<kbbq_ctx*> ctx_;
    std::vector<int> devices_;
    std::vector<bool> streaming_;
    int k_;
    const ReadStore* resident_ = nullptr;
    std::vector<SessionBatch> batches_;
    int next_slot_ = 0;
};

[[noreturn]] void throw_last(int rc);  // maps KBBQ_E* to the reference's exception types

}  // namespace kbbq

namespace bloom {

// bloom::Kmer (bloom.hh:339-379): kept for API compatibility; the kernels never roll a k-mer
class Kmer {
    size_t s;
    int k;
    uint64_t x[2], mask, shift;

public:
    explicit Kmer(int k_) : s(0), k(k_), mask(k_ < 32 ? (1ULL << k_ * 2) - 1 : ~0ULL), shift((k_ - 1) * 2) { x[0] = x[1] = 0; }
    size_t push_back(char ch);
    uint64_t get() const { return x[0] < x[1] ? x[0] : x[1]; }
    void reset() { s = 0; x[0] = x[1] = 0; }
    size_t size() const { return s; }
    bool valid() const { return s >= (size_t)k; }
    int ksize() const { return k; }
    explicit operator bool() const { return valid(); }
};

// bloom::Bloom (bloom.hh:385-402) living in HBM
class Bloom {
public:
    Bloom(unsigned long long projected_element_count, double fpr, unsigned long long seed = KBBQ_DEFAULT_BLOOM_SEED);
    Bloom(Bloom&&) noexcept = default;
    kbbq_bloom_geom params;  // geometry the reference keeps in params / bloom
    void insert(const Kmer& kmer);
    bool query(const Kmer& kmer) const;
    double fprate() const;
    unsigned long long inserted_elements() const;
    int slot() const { return slot_; }

private:
    int slot_;
};

long double calculate_phit(const Bloom& bf, long double alpha);  // bloom.cc:190-195

}  // namespace bloom

namespace covariateutils {

typedef std::vector<int> meanq_t;
typedef std::vector<int> rgdq_t;
typedef std::vector<std::vector<int>> qscoredq_t;
typedef std::vector<std::vector<std::array<std::vector<int>, 2>>> cycledq_t;
typedef std::vector<std::vector<std::vector<int>>> dinucdq_t;
struct dq_t {  // covariateutils.hh:27-34
    meanq_t meanq;
    rgdq_t rgdq;
    qscoredq_t qscoredq;
    cycledq_t cycledq;
    dinucdq_t dinucdq;
};
typedef std::array<unsigned long long, 2> covariate_t;
typedef std::vector<covariate_t> CCovariate;

// covariateutils.hh:157-172; the tables are filled by the pass-3 kernel and downloaded
class CCovariateData {
public:
    std::vector<covariate_t> rgcov;                                      // [rg]
    std::vector<CCovariate> qcov;                                        // [rg][q]
    std::vector<std::vector<std::array<CCovariate, 2>>> cycov;           // [rg][q][second][cycle]
    std::vector<std::vector<CCovariate>> dicov;                          // [rg][q][dinuc]
    dq_t get_dqs();  // covariateutils.cc:204-230, host long double
    // dense images (layout of include/kbbq_b200.h), kept so get_dqs feeds the fit without re-densifying
    int n_rg = 0, max_cycle = 0;
    std::vector<uint64_t> d_rgc, d_qc, d_cyc, d_dic;
};

std::vector<int> calculate_thresholds(unsigned long long k, long double p, long double quartile = .995l);  // covariateutils.hh:71-85

}  // namespace covariateutils

namespace readutils {
// the slice of CReadData (readutils.hh:44-74) that callers of the pass drivers look at
class CReadData {
public:
    static std::unordered_map<std::string, std::string> rg_to_pu;
    static std::unordered_map<std::string, int> rg_to_int;
    std::string seq, name, rg;
    std::vector<uint8_t> qual;
    std::vector<bool> skips, errors;
    bool second = false;
    int get_rg_int() const { return rg_to_int[rg]; }
};
}  // namespace readutils

namespace htsiter {

// htsiter::HTSFile (htsiter.hh:40-49)
class HTSFile {
public:
    virtual ~HTSFile() {}
    virtual int next() = 0;
    virtual std::string next_str() = 0;
    virtual readutils::CReadData get() = 0;
    virtual void recalibrate(const std::vector<uint8_t>& qual) = 0;
    virtual int open_out(std::string filename) = 0;
    virtual int write() = 0;
    virtual const kbbq::ReadStore& store() const = 0;  // mirror-only: the parsed input behind the iterator
};

// What FastqFile and BamFile share in this mirror: the input parsed once into a ReadStore, a cursor for the
// reference's next()/get()/write() protocol, and a BGZF writer (blocks compressed by independent threads).
class StoreFile : public HTSFile {
public:
    ~StoreFile() override;
    int next() override;
    std::string next_str() override;
    readutils::CReadData get() override;
    void recalibrate(const std::vector<uint8_t>& qual) override { cur_qual_ = qual; }
    int open_out(std::string filename) override;
    int write() override;
    const kbbq::ReadStore& store() const override { return *store_; }
    // mirror-only: write every record with new qualities (layout of kbbq_batch.qual per batch), in input order
    int write_all(const kbbq::Session& s, const std::vector<kbbq::Pinned<uint8_t>>& newq);
    // mirror-only: finish the output now (last block, BGZF end-of-file marker, flush) instead of in the destructor;
    // 0, or -1 when the stream reports an error
    int close_out();

protected:
    StoreFile(std::shared_ptr<kbbq::ReadStore> st, int nthreads) : store_(std::move(st)), nthreads_(nthreads) {}
    virtual void format_record(size_t i, const uint8_t* newq, std::string& out) const = 0;
    virtual void format_header(std::string&) const {}
    std::shared_ptr<kbbq::ReadStore> store_;
    size_t cur_ = 0;  // index of the NEXT record
    std::vector<uint8_t> cur_qual_;
    void* out_ = nullptr;  // FILE*: BGZF blocks are written raw
    std::string pending_;  // bytes waiting to be compressed
    int nthreads_;
    int flush_pending(bool final);
};

// htsiter::FastqFile (htsiter.hh:101-126): FASTQ, plain or gzip
class FastqFile : public StoreFile {
public:
    explicit FastqFile(std::string filename, int nthreads = 0);

protected:
    void format_record(size_t i, const uint8_t* newq, std::string& out) const override;
};

// htsiter::BamFile (htsiter.hh:51-99): unaligned or aligned BAM; --use-oq reads qualities from the OQ tag,
// --set-oq stores the input qualities in it (htsiter.cc:11-35)
class BamFile : public StoreFile {
public:
    BamFile(std::string filename, int nthreads = 0, bool use_oq = false, bool set_oq = false);

protected:
    void format_record(size_t i, const uint8_t* newq, std::string& out) const override;
    void format_header(std::string& out) const override;
    bool use_oq_, set_oq_;
};

// hts_detect_format as far as main() needs it (kbbq.cc:172-190): true for BAM, false for FASTQ, throws otherwise
bool is_bam_file(const std::string& filename);

// htsiter::KmerSubsampler (htsiter.hh:128-156): carries file, k, alpha, seed to subsample_kmers
class KmerSubsampler {
public:
    HTSFile* file;
    int k;
    double alpha;
    uint64_t seed;
    size_t total_kmers = 0;
    bool not_eof = true;
    KmerSubsampler(HTSFile* file_, int k_, double alpha_, uint64_t seed_);
};

}  // namespace htsiter

namespace recalibrateutils {
// recalibrateutils.hh:29-44
void subsample_kmers(htsiter::KmerSubsampler& s, bloom::Bloom& sampled);
void find_trusted_kmers(htsiter::HTSFile* file, bloom::Bloom& trusted, const bloom::Bloom& sampled, std::vector<int> thresholds, int k);
covariateutils::CCovariateData get_covariatedata(htsiter::HTSFile* file, const bloom::Bloom& trusted, int k);
void recalibrate_and_write(htsiter::HTSFile* in, const covariateutils::dq_t& dqs, std::string outfn);
inline long double q_to_p(int q);
}  // namespace recalibrateutils
