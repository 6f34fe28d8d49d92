// kbbq_cli.cc -- the kbbq command line over the B200 hot path.
//
// Same options, same order of operations and the same stderr lines as the reference's main()
// (src/kbbq/kbbq.cc:81-459); the pass calls are the reference's own four, now backed by kernels.  Added: --seed,
// because the reference seeds from time()+getpid() (kbbq.cc:268-270) and cannot be reproduced otherwise.
// Differences a user can see: the input is read once instead of five times; FASTQ (plain/gzip) and BAM are read by
// this repo's own BGZF/BAM code (htslib is not in this image); CRAM is not supported.
#include <getopt.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <ctime>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <thread>

#include "kbbq_host_api.hh"

namespace recalibrateutils {
covariateutils::CCovariateData download_covariates(kbbq::Session& ses, int R, int C);
}

static std::ostream& put_now(std::ostream& os) {
    std::time_t t = std::time(nullptr);
    std::tm tm = *std::localtime(&t);
    return os << std::put_time(&tm, "[%F %T %Z]");
}

// KBBQ_HOST_TIMING=1: wall time of each phase on stderr (the reference only stamps whole seconds)
struct PhaseTimer {
    const bool on = getenv("KBBQ_HOST_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), start = t0;
    void mark(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::cerr << "[cli] " << what << " " << std::chrono::duration<double, std::milli>(t1 - t0).count() << " ms (at "
                  << std::chrono::duration<double, std::milli>(t1 - start).count() << ")" << std::endl;
        t0 = t1;
    }
};

static struct option long_options[] = {
    {"ksize", required_argument, 0, 'k'},   {"use-oq", no_argument, 0, 'u'},        {"set-oq", no_argument, 0, 's'},
    {"genomelen", required_argument, 0, 'g'}, {"coverage", required_argument, 0, 'c'}, {"fixed", required_argument, 0, 'f'},
    {"alpha", required_argument, 0, 'a'},   {"threads", required_argument, 0, 't'}, {"seed", required_argument, 0, 'S'},
    {"gpus", required_argument, 0, 'G'},    {"report", required_argument, 0, 'R'},
    {0, 0, 0, 0}};

// true when every base of the store is one of "ACGTN": packed batches then compare exactly like the characters do
static bool plain_alphabet(const kbbq::ReadStore& rs, unsigned nt) {
    bool ok[256] = {false};
    for (char c : std::string("ACGTN")) ok[(unsigned char)c] = true;
    const size_t total = rs.seq.size();
    std::atomic<bool> plain{true};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t] {
            bool mine = true;
            const char* p = rs.seq.data();
            for (size_t i = total * t / nt; i < total * (t + 1) / nt; ++i) mine = mine && ok[(unsigned char)p[i]];
            if (!mine) plain.store(false, std::memory_order_relaxed);
        });
    for (auto& t : th) t.join();
    return plain.load();
}

int main(int argc, char* argv[]) {
    int k = 32;
    long double alpha = 0;
    uint64_t genomelen = 0;
    unsigned int coverage = 0;
    uint32_t seed = 0;
    int nthreads = 0, ngpus = 1;
    std::string fixedinput = "", reportfn = "";
    bool use_oq = false, set_oq = false;
    int opt = 0, opt_idx = 0;
    while ((opt = getopt_long(argc, argv, "k:usg:c:f:a:t:S:G:R:", long_options, &opt_idx)) != -1) {
        switch (opt) {
        case 'k':
            k = std::stoi(std::string(optarg));
            if (k <= 0 || k > KBBQ_MAX_KMER) {
                std::cerr << put_now << "  Error: k must be <= " << KBBQ_MAX_KMER << " and > 0." << std::endl;
                return 1;  // the reference prints and carries on into undefined behaviour
            }
            break;
        case 'u': use_oq = true; break;
        case 's': set_oq = true; break;
        case 'g': genomelen = std::stoull(std::string(optarg)); break;
        case 'c': coverage = std::stoul(std::string(optarg)); break;
        case 'f': fixedinput = std::string(optarg); break;
        case 'a': alpha = std::stold(std::string(optarg)); break;
        case 't':
            nthreads = std::stoi(std::string(optarg));
            if (nthreads < 0) std::cerr << put_now << " Error: threads must be >= 0." << std::endl;
            break;
        case 'S': seed = (uint32_t)std::stoul(std::string(optarg)); break;
        case 'G':  // GPUs to shard the reads over (0 = every device of the box)
            ngpus = std::stoi(std::string(optarg));
            if (ngpus < 0) {
                std::cerr << put_now << " Error: gpus must be >= 0." << std::endl;
                return 1;
            }
            break;
        case 'R': reportfn = std::string(optarg); break;  // GATK-style report of the covariate tables
        case '?':
        default:
            std::cerr << put_now << "  Unknown argument " << (char)opt << std::endl;
            return 1;
        }
    }
    std::string filename("-");
    if (optind < argc) {
        filename = std::string(argv[optind]);
        while (++optind < argc) std::cerr << put_now << " Warning: Extra argument " << argv[optind] << " ignored." << std::endl;
    }
    long double sampler_desiredfpr = 0.01;    // Lighter uses .01
    long double trusted_desiredfpr = 0.0005;  // and .0005

    try {
        // the one parse of the input (the reference re-opens the file for every pass, kbbq.cc:232,258,336,365,456)
        bool is_bam = false;
        try {
            is_bam = htsiter::is_bam_file(filename);  // hts_detect_format (kbbq.cc:172-190)
        } catch (const std::exception& e) {
            std::cerr << put_now << " " << e.what() << std::endl;
            return 1;
        }
        auto open_file = [&](const std::string& fn) -> std::unique_ptr<htsiter::StoreFile> {  // kbbq.cc:22-32
            if (is_bam) return std::unique_ptr<htsiter::StoreFile>(new htsiter::BamFile(fn, nthreads, use_oq, set_oq));
            return std::unique_ptr<htsiter::StoreFile>(new htsiter::FastqFile(fn, nthreads));
        };
        PhaseTimer timer;
        std::thread warmup([] { kbbq_device_warmup(0); });  // CUDA context creation runs while the input is parsed
        struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{warmup};
        std::unique_ptr<htsiter::StoreFile> file = open_file(filename);
        const kbbq::ReadStore& store = file->store();
        timer.mark("input read + parsed");
        if (store.n() == 0) {
            std::cerr << put_now << " Error: total sequence length in file " << filename << " is 0. Check that the file isn't empty." << std::endl;
            return 1;
        }
        const int n_rg = (int)std::max<size_t>(1, readutils::CReadData::rg_to_int.size());
        // --gpus N with N > 0 names the devices without asking the driver (the first CUDA call blocks until the runtime is
        // up, and the packing below should run under that start-up, not behind it); the count is checked afterwards
        std::vector<int> devices;
        auto check_count = [&](int have) {
            const int want = ngpus == 0 ? have : ngpus;
            if (want > have && have > 0) {
                std::cerr << put_now << " Error: " << want << " GPUs requested, " << have << " present." << std::endl;
                return false;
            }
            return true;
        };
        if (ngpus == 0) {
            const int have = kbbq_device_count();
            if (!check_count(have)) return 1;
            for (int d = 0; d < std::max(1, have); ++d) devices.push_back(d);
        } else {
            for (int d = 0; d < ngpus; ++d) devices.push_back(d);
        }
        // the contexts come up in the background (CUDA's start-up outlasts the parsing); packing needs none of them
        kbbq::Session session(k, n_rg, (int)std::max<uint32_t>(1, store.max_len), devices, true);
        covariateutils::CCovariateData data;
        session.prepack(store);
        timer.mark("first batches packed");
        if (ngpus != 0 && !check_count(kbbq_device_count())) return 1;
        session.contexts();
        timer.mark("GPU context");

        if (fixedinput == "") {
            if (genomelen == 0) {
                if (is_bam) {  // sum of the header's reference lengths (kbbq.cc:196-216)
                    std::cerr << put_now << " Estimating genome length" << std::endl;
                    for (uint64_t l : store.ref_len) genomelen += l;
                    if (genomelen == 0) {
                        std::cerr << put_now << " Header does not contain genome information."
                                  << " Unable to estimate genome length; please provide it on the command line"
                                  << " using the --genomelen option." << std::endl;
                        return 1;
                    }
                    std::cerr << put_now << " Genome length is " << genomelen << " bp." << std::endl;
                } else {
                    std::cerr << put_now << " Error: --genomelen must be specified if input is not a bam." << std::endl;
                    return 1;  // the reference carries on into a division by zero
                }
            }
            if (alpha == 0) {
                std::cerr << put_now << " Estimating alpha." << std::endl;
                if (coverage == 0) {
                    std::cerr << put_now << " Estimating coverage." << std::endl;
                    // sum of sequence lengths (kbbq.cc:234-236): the coverage pass, over the batches the GPUs hold
                    session.ensure_resident(store);
                    std::atomic<uint64_t> seqlen_dev{0};
                    session.for_each_batch([&](kbbq_ctx* c, kbbq_dbatch* d, kbbq::SessionBatch&) {
                        uint64_t t = 0;
                        if (kbbq_batch_total_bases(c, d, &t) != KBBQ_OK) kbbq::throw_last(-3);
                        seqlen_dev += t;
                    });
                    uint64_t seqlen = seqlen_dev.load();
                    std::cerr << put_now << " Total Sequence length: " << seqlen << std::endl;
                    std::cerr << put_now << " Genome length: " << genomelen << std::endl;
                    coverage = seqlen / genomelen;
                    std::cerr << put_now << " Estimated coverage: " << coverage << std::endl;
                    if (coverage == 0) {
                        std::cerr << put_now << " Error: estimated coverage is 0." << std::endl;
                        return 1;
                    }
                }
                alpha = 7.0l / (long double)coverage;
            }
            if (coverage == 0) coverage = 7.0l / alpha;
            std::cerr << put_now << " Sampling kmers at rate " << alpha << std::endl;
            unsigned long long int approx_kmers = genomelen * coverage * alpha;
            bloom::Bloom subsampled(approx_kmers, sampler_desiredfpr);
            bloom::Bloom trusted(approx_kmers, trusted_desiredfpr);
            if (seed == 0) seed = kbbq_default_seed((uint64_t)time(nullptr), (uint32_t)getpid());
            std::cerr << put_now << " Seed: " << seed << std::endl;

            htsiter::KmerSubsampler subsampler(file.get(), k, alpha, seed);
            recalibrateutils::subsample_kmers(subsampler, subsampled);
            timer.mark("pack + upload + pass 1");
            std::cerr << put_now << " Sampled " << subsampled.inserted_elements() << " valid kmers." << std::endl;

            long double fpr = subsampled.fprate();
            std::cerr << put_now << " Approximate false positive rate: " << fpr << std::endl;
            if (fpr > .15) {
                std::cerr << put_now << " Error: false positive rate is too high. Increase genomelen parameter and try again." << std::endl;
                return 1;
            }
            long double p = bloom::calculate_phit(subsampled, alpha);
            std::vector<int> thresholds = covariateutils::calculate_thresholds(k, p);

            std::cerr << put_now << " Finding trusted kmers" << std::endl;
            recalibrateutils::find_trusted_kmers(file.get(), trusted, subsampled, thresholds, k);
            timer.mark("pass 2 (launched)");
            std::cerr << put_now << " Finding errors" << std::endl;
            data = recalibrateutils::get_covariatedata(file.get(), trusted, k);
            timer.mark("pass 3 + tables to host");
        } else {
            // --fixed: a base is an error when it differs from the same base of the pre-corrected file (kbbq.cc:367-378)
            std::cerr << put_now << " Using fixed file to find errors." << std::endl;
            auto saved_int = readutils::CReadData::rg_to_int;
            auto saved_pu = readutils::CReadData::rg_to_pu;
            std::unique_ptr<htsiter::StoreFile> fixedfile = open_file(fixedinput);
            readutils::CReadData::rg_to_int = saved_int;
            readutils::CReadData::rg_to_pu = saved_pu;
            const kbbq::ReadStore& fx = fixedfile->store();
            session.ensure_resident(store);
            const unsigned nt = std::max(1u, std::min(16u, nthreads > 0 ? (unsigned)nthreads : std::thread::hardware_concurrency()));
            // Both files in upper-case ACGTN (the normal case): the comparison runs on the device over the packed batches
            // (kbbq_fixed_errors).  Anything else -- soft-masked or IUPAC letters, which packing folds together -- is
            // compared character by character here, like the reference's std::not_equal_to<char>.
            const bool on_device = fx.max_len <= store.max_len && plain_alphabet(store, nt) && plain_alphabet(fx, nt) && !getenv("KBBQ_FIXED_ON_HOST");
            for (kbbq_ctx* c : session.contexts())
                if (kbbq_covariates_reset(c) != KBBQ_OK) kbbq::throw_last(-3);
            session.for_each_batch([&](kbbq_ctx* c, kbbq_dbatch* d, kbbq::SessionBatch& sb) {
                // like the reference's loop (kbbq.cc:370) only reads that have a partner in the fixed file are consumed
                const size_t first = sb.lo, last = std::min(sb.hi, fx.n());
                const uint32_t n_first = (uint32_t)(last > first ? last - first : 0);
                if (on_device) {
                    if (!n_first) return;
                    kbbq_dbatch* f = session.upload_range(fx, first, last, sb.gpu);
                    const int rc = kbbq_fixed_errors(c, d, f, n_first, nullptr);
                    kbbq_batch_free(c, f);
                    if (rc != KBBQ_OK) kbbq::throw_last(rc);
                    return;
                }
                // a read owns its words of `bits`, so ranges of reads go to threads.  (The reference reads past the end of
                // a fixed read that is shorter than its partner; here the bases without a partner count as error-free.)
                std::vector<uint32_t> bits(std::max<uint64_t>(sb.n_words, 1), 0);
                std::vector<std::thread> th;
                for (unsigned t = 0; t < nt; ++t)
                    th.emplace_back([&, t] {
                        const size_t span = last > first ? last - first : 0;
                        for (size_t i = first + span * t / nt; i < first + span * (t + 1) / nt; ++i) {
                            const size_t L = std::min<size_t>(store.off[i + 1] - store.off[i], fx.off[i + 1] - fx.off[i]);
                            const uint32_t wo = sb.word_off[i - first];
                            for (size_t j = 0; j < L; ++j)
                                if (store.seq[store.off[i] + j] != fx.seq[fx.off[i] + j]) bits[wo + (j >> 5)] |= 1u << (j & 31);
                        }
                    });
                for (auto& t : th) t.join();
                if (kbbq_consume_errors_n(c, d, bits.data(), n_first) != KBBQ_OK) kbbq::throw_last(-3);
            });
            data = recalibrateutils::download_covariates(session, n_rg, (int)std::max<uint32_t>(1, store.max_len));
        }

        if (reportfn != "") {  // GATK-style text report of the count tables (the reference's gatkreport.hh was written for this)
            std::vector<std::string> names(n_rg);
            for (const auto& kv : readutils::CReadData::rg_to_int)
                if (kv.second >= 0 && kv.second < n_rg) names[kv.second] = kv.first;
            std::vector<const char*> np;
            for (int r = 0; r < n_rg; ++r) np.push_back(names[r].empty() ? nullptr : names[r].c_str());
            if (kbbq_report_tables_host(reportfn.c_str(), data.n_rg, data.max_cycle, data.d_rgc.data(), data.d_qc.data(), data.d_cyc.data(),
                                        data.d_dic.data(), np.data()) != KBBQ_OK)
                std::cerr << put_now << " Warning: " << kbbq_last_error() << std::endl;
        }
        std::cerr << put_now << " Training model" << std::endl;
        covariateutils::dq_t dqs = data.get_dqs();
        timer.mark("model fit");
        std::cerr << put_now << " Recalibrating file" << std::endl;
        recalibrateutils::recalibrate_and_write(file.get(), dqs, "-");
        if (file->close_out() < 0) {
            std::cerr << put_now << " Error: writing the output failed" << std::endl;
            return 1;
        }
        timer.mark("pass 4 + format + compress + write");
        // Everything is written.  What is left is handing back gigabytes of parsed input, pinned buffers and the CUDA
        // contexts one destructor at a time, which the kernel does in one sweep at process exit anyway
        // (KBBQ_CLEAN_EXIT=1 runs the destructors, for leak checkers).
        if (const char* em = getenv("KBBQ_EXIT_MODE")) {  // measurement: what the kernel's teardown at exit is made of
            const int mode = atoi(em);
            if (mode & 1) { file.reset(); timer.mark("exit probe: parsed input released"); }
            if (mode & 2) { session.release(); timer.mark("exit probe: batches and contexts released"); }
            if (mode & 4) { for (int d : devices) kbbq_device_reset(d); timer.mark("exit probe: devices reset"); }
            std::cerr.flush();
            _exit(0);
        }
        if (!getenv("KBBQ_CLEAN_EXIT")) {
            std::cerr.flush();
            _exit(0);
        }
        file.reset();
        timer.mark("parsed input released");
    } catch (const std::exception& e) {
        std::cerr << put_now << " Error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
