/*
 * kbbq_b200.h -- C ABI of the B200-native kbbq recalibration hot path.
 *
 * The reference (adamjorr/kbbq) has no plugin/FFI layer: its operator API for this path is the C++ surface of
 * libkbbq.a that main() drives (src/kbbq/kbbq.cc:264-457).  This header is what a maintainer's replacement pass
 * drivers bind instead of the CPU loops; each entry point names the reference interface it stands in for.
 * Plain pointers and sizes only; every function returns 0 on success or a negative KBBQ_E* code, never throws.
 * Caller owns host buffers; the library owns device memory.  One host thread drives one context (one GPU).
 *
 * Table layouts (dense, row-major; R = n_rg, C = max_cycle given to kbbq_ctx_create, 94 = KBBQ_MAXQ+1):
 *   counts (uint64 {errors,total} pairs):  rg[R][2]  q[R][94][2]  cycle[R][94][2][C][2]  dinuc[R][94][16][2]
 *   model  (int32):                        meanq[R] rgdq[R] qdq[R][94] cydq[R][94][2][C] didq[R][94][16]
 * They are covariateutils::CCovariateData / dq_t (covariateutils.hh:21-34,157-172) without the lazy vector growth.
 */
#ifndef KBBQ_B200_H
#define KBBQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KBBQ_OK 0
#define KBBQ_EINVAL (-1)    /* bad argument (std::invalid_argument in the reference) */
#define KBBQ_ENOMEM (-2)    /* host or device allocation failed (std::bad_alloc) */
#define KBBQ_ECUDA (-3)     /* a CUDA call failed; see kbbq_last_error() */
#define KBBQ_ESTATE (-4)    /* call out of order (e.g. a pass before its filter exists) */
#define KBBQ_ERANGE (-5)    /* read longer than KBBQ_MAX_READ_LEN, rg >= n_rg, ... */

#define KBBQ_FILTER_SAMPLED 0 /* `subsampled` in kbbq.cc:265 */
#define KBBQ_FILTER_TRUSTED 1 /* `trusted`    in kbbq.cc:266 */
#define KBBQ_DEFAULT_BLOOM_SEED 0xA5A5A5A55A5A5A5AULL /* bloom.hh:389 */
/* longest read a batch may hold.  Reads of up to 512 bases run from statically sized shared-memory scratch; a batch with
 * a longer read is routed to slower instantiations whose scratch is sized for it at launch. */
#define KBBQ_ABI_MAX_READ_LEN 65535

typedef struct kbbq_ctx kbbq_ctx;       /* one GPU: stream, filters, count tables, model tables */
typedef struct kbbq_dbatch kbbq_dbatch; /* a read batch resident in device memory */

/* Geometry of a pattern-blocked Bloom filter: what bloom::Bloom's constructor derives (bloom.cc:12-24,
 * bloom.hh:36-45, bloom_filter.hpp:108-160 and 467-528). */
typedef struct kbbq_bloom_geom {
    uint64_t projected;
    double fpr;
    uint64_t seed;
    uint32_t optimal_k;
    uint64_t optimal_table_size;
    uint64_t random_seed;
    uint64_t table_bits;
    uint64_t n_blocks;
    uint32_t n_salts;
    uint32_t salts[128];
} kbbq_bloom_geom;

/* A batch of reads in packed form, host memory (pinned memory makes the upload asynchronous).
 * Read r occupies words [word_off[r], word_off[r+1]); base i of it is bits [2*(i%32), +2) of
 * bases[word_off[r] + i/32] (A=0 C=1 G=2 T=3), bit i%32 of nmask[...] marks a non-ACGT base (the reference's
 * seq_nt16_int >= 4, bloom.hh:351), and its quality is qual[32*word_off[r] + i] (phred, no +33).  Unused bits and
 * bytes are zero.  This replaces the per-record CReadData (readutils.hh:44-74) that the reference's pass drivers
 * pull from htsiter::HTSFile::get() (htsiter.hh:40-49). */
typedef struct kbbq_batch {
    uint32_t n_reads;
    uint32_t max_len;
    uint64_t n_words;
    const uint64_t* bases;    /* [n_words] */
    const uint32_t* nmask;    /* [n_words]; may be NULL: the batch holds no non-ACGT base */
    const uint8_t* qual;      /* [32*n_words] */
    const uint32_t* word_off; /* [n_reads+1] */
    const uint32_t* len;      /* [n_reads] */
    const uint16_t* rg;       /* [n_reads] dense read-group id, CReadData::get_rg_int() (readutils.hh:62) */
    const uint8_t* second;    /* [n_reads] CReadData::second */
    /* optional, may be NULL: [2*n_words] per word w, qmask[2w] = bit i set when qual[32w+i] <= 2 (INFER_ERROR_BAD_QUAL,
     * readutils.hh:16) and qmask[2w+1] = bit i set when qual[32w+i] <= 6 (get_errors' minqual, readutils.cc:238);
     * kbbq_quality_masks makes them.  They are all pass 2 needs of the qualities, so a batch that brings them lets pass 2
     * start before the quality array itself has arrived (kbbq_batch_upload_async); without them the library derives
     * them on the device from `qual`. */
    const uint32_t* qmask;
} kbbq_batch;

const char* kbbq_last_error(void);
int kbbq_device_count(void);
/* Initialise the CUDA runtime on `device` (primary context creation, the slow part of a fresh process).  Optional:
 * kbbq_ctx_create does it implicitly; a host may call this from a side thread while it parses its input. */
int kbbq_device_warmup(int device);

/* page-locked host memory (what kbbq_batch_upload_async / _stream and kbbq_pass4_recalibrate need for their copies to
 * be asynchronous), usable from every device of the process; and a device's free / total memory, for a host that
 * decides how many batches may stay resident */
int kbbq_host_alloc(size_t bytes, void** out);
void kbbq_host_free(void* p);
int kbbq_device_memory(int device, uint64_t* free_bytes, uint64_t* total_bytes);
/* Give back everything this process holds on `device` (cached pool memory, the primary context).  Optional; every
 * kbbq_ctx of the device must have been destroyed and every pinned array freed. */
int kbbq_device_reset(int device);

/* ---- host-only helpers (no GPU needed) ------------------------------------------------------------------------ */
/* bloom::Bloom ctor geometry; KBBQ_EINVAL where the reference throws (bloom.cc:18-21) */
int kbbq_bloom_geometry(uint64_t projected, double fpr, uint64_t seed, kbbq_bloom_geom* out);
/* pattern table of pattern_blocked_bf (bloom.hh:189-231): patterns[65536*8] */
int kbbq_bloom_patterns(const kbbq_bloom_geom* g, uint64_t* patterns);
/* Bloom::fprate() = pattern_blocked_bf::effective_fpp (bloom.hh:318-330,399) */
double kbbq_effective_fpp(uint64_t table_bits, uint64_t inserted, uint32_t n_salts);
/* bloom::calculate_phit (bloom.cc:190-195) chained into covariateutils::calculate_thresholds (covariateutils.hh:71-85),
 * exactly as kbbq.cc:304-313; thresholds[k+1] */
int kbbq_thresholds_from_fpr(double fpr, long double alpha, int k, int* thresholds);
/* the two halves of the above on their own: bloom::calculate_phit and covariateutils::calculate_thresholds */
int kbbq_phit(double fpr, long double alpha, long double* p);
int kbbq_thresholds_from_p(int k, long double p, int* thresholds);
/* minion::create_seed_seq().GenerateOne() (minion.hpp:337-408, kbbq.cc:268-270) for a given time() and getpid() */
uint32_t kbbq_default_seed(uint64_t time_now, uint32_t pid);
/* minion::Random::Seed(uint32) state (minion.hpp:355-375) advanced by `skip` draws */
int kbbq_rng_state(uint32_t seed, uint64_t skip, uint64_t state[4]);
/* CCovariateData::get_dqs (covariateutils.cc:204-230) on dense tables */
int kbbq_fit_model_host(int n_rg, int max_cycle, const uint64_t* rgc, const uint64_t* qc, const uint64_t* cyc,
                        const uint64_t* dic, int32_t* meanq, int32_t* rgdq, int32_t* qdq, int32_t* cydq, int32_t* didq);
/* Pack ASCII reads (seq[total], off[n+1], qual[total] phred) into the kbbq_batch layout.  Two-call protocol:
 * kbbq_pack_size gives n_words; the caller allocates; kbbq_pack_reads fills.  Returns KBBQ_ERANGE for a read longer
 * than KBBQ_ABI_MAX_READ_LEN. */
int kbbq_pack_size(uint64_t n, const uint64_t* off, uint64_t* n_words, uint32_t* max_len);
int kbbq_pack_reads(uint64_t n, const char* seq, const uint64_t* off, const uint8_t* qual, uint64_t* bases,
                    uint32_t* nmask, uint8_t* qual_out, uint32_t* word_off, uint32_t* len, int n_threads);
/* kbbq_batch::qmask from a packed quality array: qmask[2*n_words] from qual[32*n_words] */
int kbbq_quality_masks(uint64_t n_words, const uint8_t* qual, uint32_t* qmask, int n_threads);

/* ---- context ---------------------------------------------------------------------------------------------------- */
int kbbq_ctx_create(int device, int k, int n_rg, int max_cycle, kbbq_ctx** out);
void kbbq_ctx_destroy(kbbq_ctx* ctx);
int kbbq_sync(kbbq_ctx* ctx);
/* the CUDA stream (cudaStream_t) all of this context's work is issued on */
void* kbbq_stream(kbbq_ctx* ctx);
/* device time of the most recent pass call in ms (CUDA events on the context's stream); syncs */
int kbbq_last_pass_ms(kbbq_ctx* ctx, float* ms);
/* the kernels of the most recent pass call and their device times (events between the launches): fills up to `max`
 * entries of names[] (static strings) and ms[], returns how many, or a negative KBBQ_E* code; syncs */
int kbbq_last_pass_kernels(kbbq_ctx* ctx, int max, const char** names, float* ms);
/* kernels launched by this context so far */
uint64_t kbbq_launch_count(kbbq_ctx* ctx);

/* ---- filters: bloom::Bloom (bloom.hh:385-402) in HBM -------------------------------------------------------------- */
int kbbq_bloom_create(kbbq_ctx* ctx, int which, uint64_t projected, double fpr, uint64_t seed);
int kbbq_bloom_info(kbbq_ctx* ctx, int which, kbbq_bloom_geom* out);
int kbbq_bloom_clear(kbbq_ctx* ctx, int which);
/* Bloom::inserted_elements() */
int kbbq_bloom_inserted(kbbq_ctx* ctx, int which, uint64_t* inserted);
/* bit array as the reference holds it in bit_table_ (bloom.hh:33): words[table_bits/64]; either pointer may be NULL */
int kbbq_bloom_download(kbbq_ctx* ctx, int which, uint64_t* words, uint64_t* inserted);
int kbbq_bloom_upload(kbbq_ctx* ctx, int which, const uint64_t* words, uint64_t inserted);
/* Bloom::insert / Bloom::query on raw 64-bit keys (bloom.hh:395-398 with a valid Kmer); hits[n] receives 0/1 */
int kbbq_bloom_insert_keys(kbbq_ctx* ctx, int which, const uint64_t* keys, uint64_t n);
int kbbq_bloom_query_keys(kbbq_ctx* ctx, int which, const uint64_t* keys, uint64_t n, uint8_t* hits);
/* Bloom::fprate() from the device-side insert counter */
int kbbq_bloom_fprate(kbbq_ctx* ctx, int which, double* fpr);
/* device pointers for collectives (bit array: table_bits/8 bytes; counter: one uint64) */
void* kbbq_bloom_device_words(kbbq_ctx* ctx, int which);
void* kbbq_bloom_device_counter(kbbq_ctx* ctx, int which);
/* words |= other (other: device pointer to an equally sized bit array, e.g. a peer GPU's mapped filter) */
int kbbq_bloom_or_from(kbbq_ctx* ctx, int which, const void* other_device_words, uint64_t first_word, uint64_t n_words);

/* Order-independent 128-bit digests {sum, xor of mix(word ^ mix(index))} of a filter's bit array, of the four count
 * tables and of a batch's recalibrated qualities, computed on the device: equality checks across ranks, runs and
 * against a CPU restatement (tests) without moving tens of GB to the host.  An all-zero array digests to {0, 0}. */
int kbbq_bloom_digest(kbbq_ctx* ctx, int which, uint64_t out[2]);
int kbbq_covariates_digest(kbbq_ctx* ctx, uint64_t out[2]);
int kbbq_batch_newqual_digest(kbbq_ctx* ctx, const kbbq_dbatch* b, uint64_t out[2]);

/* Save / restore a filter (parameters, insert counter, bit array laid out as the reference's bit_table_): checkpoint
 * between passes, fixtures across runs.  kbbq_bloom_load (re)creates filter `which` from the file's parameters.
 * The reference never serialises its filters. */
int kbbq_bloom_save(kbbq_ctx* ctx, int which, const char* path);
int kbbq_bloom_load(kbbq_ctx* ctx, int which, const char* path);

/* The whole multi-GPU merge of a filter's bit array in one kernel over peer-mapped memory (one process per GPU, all
 * on one NVLink/NVSwitch box): rank r ORs slice r of every rank's array and writes the result into every rank's array.
 *   1. every rank:  kbbq_bloom_ipc_handle -> 64 bytes; all-gather them (any transport) into handles[world][64]
 *   2. every rank:  kbbq_sync, then a barrier              (all inserts of the pass are in memory everywhere)
 *   3. every rank:  kbbq_bloom_or_allreduce_p2p, kbbq_sync, then a barrier      (all slices written everywhere)
 * The insert counters are summed separately (kbbq_bloom_device_counter).  world <= 16.  The reference has no
 * counterpart: it is single-process. */
#define KBBQ_IPC_HANDLE_BYTES 64
int kbbq_bloom_ipc_handle(kbbq_ctx* ctx, int which, void* handle64);
int kbbq_bloom_or_allreduce_p2p(kbbq_ctx* ctx, int which, int rank, int world, const void* handles);

/* The merge through NVSwitch multicast (NVLS): the switch ORs the N copies of a word (`multimem.ld_reduce.or`) and
 * writes the result to all N (`multimem.st`), so 1/N of the array arrives at and leaves each GPU instead of (N-1)/N.
 * The filters must have been created as shareable virtual-memory-API memory (environment KBBQ_FILTER_VMM=1 at
 * kbbq_bloom_create; they then cannot be exported with kbbq_bloom_ipc_handle).  One process per GPU:
 *   rank 0:       kbbq_bloom_mc_create -> a POSIX file descriptor; hand it to the other ranks (SCM_RIGHTS over a Unix socket)
 *   other ranks:  kbbq_bloom_mc_import(fd)
 *   every rank:   kbbq_bloom_mc_add_device; barrier (binding blocks until every device has been added);
 *                 kbbq_bloom_mc_bind; barrier
 *   per merge:    kbbq_sync + barrier, kbbq_bloom_or_allreduce_mc, kbbq_sync + barrier, counters summed as before
 * kbbq_bloom_mc_supported: 1 when the filter is VMM memory and the device reports multicast support.  Every call fails
 * soft (KBBQ_ECUDA / KBBQ_ESTATE with a message); dist.py falls back to the all-gather merge when any rank fails. */
int kbbq_bloom_mc_supported(kbbq_ctx* ctx, int which);
int kbbq_bloom_mc_create(kbbq_ctx* ctx, int which, int world, int* fd_out);
int kbbq_bloom_mc_import(kbbq_ctx* ctx, int which, int world, int fd);
int kbbq_bloom_mc_add_device(kbbq_ctx* ctx, int which);
int kbbq_bloom_mc_bind(kbbq_ctx* ctx, int which);
int kbbq_bloom_or_allreduce_mc(kbbq_ctx* ctx, int which, int rank, int world);

/* The same merge for several contexts of ONE process (a host that drives several GPUs itself): peer access instead of
 * IPC handles, every rank's kernel launched from this one call, the insert counters summed; synchronises every
 * context's stream before and after.  KBBQ_ECUDA when two of the devices cannot map each other's memory. */
int kbbq_bloom_or_allreduce_local(kbbq_ctx* const* ctxs, int n, int which);
/* replace the context's count tables by the caller's (dense layout as above): the host-side SUM of per-GPU partial
 * tables goes back to every GPU this way (they are a few hundred KB) */
int kbbq_covariates_upload(kbbq_ctx* ctx, const uint64_t* rgc, const uint64_t* qc, const uint64_t* cyc, const uint64_t* dic);

/* ---- batches -------------------------------------------------------------------------------------------------------- */
int kbbq_batch_upload(kbbq_ctx* ctx, const kbbq_batch* host, kbbq_dbatch** out);
/* Same, but the quality masks (when the host brings them) and the quality array -- 3/4 of the bytes -- are copied on a
 * side stream behind everything else and may still be in flight on return: pass 1 reads neither and overlaps the copies,
 * pass 2 waits on the device for the masks, passes 3-4 for the qualities.  host->qmask and host->qual must stay valid
 * and unmodified until kbbq_batch_wait() or kbbq_sync() has returned (with pinned memory the copies are truly
 * asynchronous). */
int kbbq_batch_upload_async(kbbq_ctx* ctx, const kbbq_batch* host, kbbq_dbatch** out);
/* Same, but EVERY array travels on the side stream and nothing is waited for: the call returns as soon as the copies
 * are queued, the passes wait for them on the device, and the upload of the next batch overlaps the passes of this one
 * -- double buffering for inputs larger than device memory, which are streamed through the device once per pass as the
 * reference re-reads its file once per pass (kbbq.cc:232,258,336,365,456).  Every host array must be pinned and stay
 * valid and unmodified until kbbq_batch_wait (or kbbq_sync) has returned. */
int kbbq_batch_upload_stream(kbbq_ctx* ctx, const kbbq_batch* host, kbbq_dbatch** out);
/* blocks until every host array of a kbbq_batch_upload_async / _stream call has been read */
int kbbq_batch_wait(kbbq_ctx* ctx, const kbbq_dbatch* b);
void kbbq_batch_free(kbbq_ctx* ctx, kbbq_dbatch* b);
/* number of k-mer slots of the batch = sum over reads of max(0, len-k+1): the draws pass 1 consumes */
uint64_t kbbq_batch_slots(const kbbq_dbatch* b);

/* ---- passes (each is asynchronous on the context's stream unless a host output pointer is given) ------------------ */
/* htsiter::KmerSubsampler(file,k,alpha,seed) (htsiter.hh:143): seeds the sampling stream, position 0 */
int kbbq_sampler_seed(kbbq_ctx* ctx, uint32_t seed, double alpha);
/* the same from the generator's state (minion::Random::state(), minion.hpp:88): a KmerSubsampler that was constructed
 * by the reference's own code keeps no seed, only its engine */
int kbbq_sampler_set_state(kbbq_ctx* ctx, const uint64_t state[4], double alpha);
/* position of the sampling stream in draws; set it when batches are dealt to several GPUs */
int kbbq_sampler_seek(kbbq_ctx* ctx, uint64_t slot_ordinal);
/* recalibrateutils::subsample_kmers (recalibrateutils.cc:7-13) over one batch; advances the stream by the batch's slots */
int kbbq_pass1_subsample(kbbq_ctx* ctx, const kbbq_dbatch* b);
/* thresholds[k+1] for pass 2 (kbbq.cc:313) */
int kbbq_set_thresholds(kbbq_ctx* ctx, const int* thresholds);
/* fprate -> phit -> thresholds from the sampled filter's own counter, as kbbq.cc:304-313; also installs them */
int kbbq_compute_thresholds(kbbq_ctx* ctx, long double alpha, int* thresholds_out /* may be NULL */);
/* recalibrateutils::find_trusted_kmers (recalibrateutils.cc:15-40) over one batch.  err_bits (host, may be NULL)
 * receives CReadData::errors after infer_read_errors as bit i%32 of err_bits[word_off[r] + i/32]. */
int kbbq_pass2_find_trusted(kbbq_ctx* ctx, const kbbq_dbatch* b, uint32_t* err_bits);
/* recalibrateutils::get_covariatedata (recalibrateutils.cc:42-89) over one batch: get_errors + consume_read.
 * err_bits as above (CReadData::errors after get_errors). */
int kbbq_pass3_covariates(kbbq_ctx* ctx, const kbbq_dbatch* b, uint32_t* err_bits);
/* --fixed mode (kbbq.cc:367-378): error flags supplied by the caller, then CCovariateData::consume_read */
int kbbq_consume_errors(kbbq_ctx* ctx, const kbbq_dbatch* b, const uint32_t* err_bits);
/* the same for the first n_first reads of the batch only: main()'s loop stops consuming where the shorter of the two
 * files ends (kbbq.cc:370) */
int kbbq_consume_errors_n(kbbq_ctx* ctx, const kbbq_dbatch* b, const uint32_t* err_bits, uint32_t n_first);
/* --fixed mode with both files resident (kbbq.cc:367-378): flags base i of read r of `b` when it differs from base i
 * of read r of `fixed` (std::not_equal_to<char> over read.seq; bases beyond the shorter of the two reads are error-free,
 * where the reference reads past the end), then CCovariateData::consume_read for the first n_first reads.  Packed
 * batches compare letters up to case and up to WHICH non-ACGT letter a base is: a host that must tell 'a' from 'A' or
 * 'N' from 'R' compares characters itself and calls kbbq_consume_errors_n.  err_bits (host, may be NULL) as above. */
int kbbq_fixed_errors(kbbq_ctx* ctx, const kbbq_dbatch* b, const kbbq_dbatch* fixed, uint32_t n_first, uint32_t* err_bits);
/* the coverage-estimation pass (kbbq.cc:229-236): total sequence length of a resident batch */
int kbbq_batch_total_bases(kbbq_ctx* ctx, const kbbq_dbatch* b, uint64_t* total);
int kbbq_covariates_reset(kbbq_ctx* ctx);
int kbbq_covariates_download(kbbq_ctx* ctx, uint64_t* rgc, uint64_t* qc, uint64_t* cyc, uint64_t* dic);
/* GATK-style text report of the four count tables ("#:GATKReport.v1.1:3": RecalTable0 per read group, RecalTable1 per
 * read group x quality, RecalTable2 cycle and dinucleotide-context rows; observed rows only) -- what the reference's
 * unused include/kbbq/gatkreport.hh:1-295 was written for.  rg_names[n_rg] may be NULL (rg<id>).  The _host form
 * formats dense tables laid out as above and needs no GPU. */
int kbbq_covariates_report(kbbq_ctx* ctx, const char* path, const char* const* rg_names);
int kbbq_report_tables_host(const char* path, int n_rg, int max_cycle, const uint64_t* rgc, const uint64_t* qc, const uint64_t* cyc,
                            const uint64_t* dic, const char* const* rg_names);
/* device pointer + size in uint64 words of the four count tables laid out back to back (for a sum all-reduce) */
void* kbbq_covariates_device(kbbq_ctx* ctx, uint64_t* n_words);
/* CCovariateData::get_dqs (covariateutils.cc:204-230) on the context's tables; installs the lookup tables.
 * Output pointers may be NULL. */
int kbbq_fit_model(kbbq_ctx* ctx, int32_t* meanq, int32_t* rgdq, int32_t* qdq, int32_t* cydq, int32_t* didq);
int kbbq_set_model(kbbq_ctx* ctx, const int32_t* meanq, const int32_t* rgdq, const int32_t* qdq, const int32_t* cydq,
                   const int32_t* didq);
/* recalibrateutils::recalibrate_and_write's compute half = CReadData::recalibrate (readutils.cc:572-595) over one batch.
 * qual_out (host, may be NULL) has the layout of kbbq_batch.qual. */
int kbbq_pass4_recalibrate(kbbq_ctx* ctx, const kbbq_dbatch* b, uint8_t* qual_out);
/* device pointer to the batch's recalibrated qualities (valid after kbbq_pass4_recalibrate) */
void* kbbq_batch_device_newqual(kbbq_dbatch* b);

/* ---- synthetic batches: MEASUREMENT TOOLING, not part of the reference's interface ---------------------------------- */
/* Reads of the shape SURVEY.md section 8(d) / Appendix C describes as a pure function of (seed, read index): uniform
 * genome, uniform start and strand, fixed length, qualities from {2,11,25,37,40}, substitution errors at
 * err_scale * 10^(-q/10), second = index & 1, read group = (index / 2) % n_rg.  kbbq_synth_batch fills a packed batch
 * for reads [first_read, first_read + n_reads) directly in device memory; kbbq_synth_host writes the same reads as
 * ASCII (seq[n*read_len], qual[n*read_len] phred, rg[n], second[n]) for a CPU reference; kbbq_batch_download copies a
 * resident batch's arrays back into caller-allocated host buffers laid out like the arrays of a kbbq_batch; any pointer may be NULL. */
typedef struct kbbq_synth_params {
    uint64_t seed;        /* read stream */
    uint64_t genome_seed; /* genome; shards of one run share it */
    uint64_t genome_len;
    uint32_t read_len;
    uint32_t n_rg;
    double err_scale;
} kbbq_synth_params;
int kbbq_synth_batch(kbbq_ctx* ctx, const kbbq_synth_params* p, uint64_t first_read, uint32_t n_reads, kbbq_dbatch** out);
int kbbq_synth_host(const kbbq_synth_params* p, uint64_t first_read, uint64_t n_reads, char* seq, uint8_t* qual, int32_t* rg,
                    uint8_t* second, int n_threads);
int kbbq_batch_extent(const kbbq_dbatch* b, uint32_t* n_reads, uint64_t* n_words, uint32_t* max_len);
int kbbq_batch_download(kbbq_ctx* ctx, const kbbq_dbatch* b, uint64_t* bases, uint32_t* nmask, uint8_t* qual, uint32_t* word_off,
                        uint32_t* len, uint16_t* rg, uint8_t* second, uint32_t* qmask);

#ifdef __cplusplus
}
#endif
#endif
