// kbbq_capi.cu -- implementation of include/kbbq_b200.h: contexts, device memory, kernel launches.
// No CPU fallback lives here: every pass entry point launches the sm_100a kernels or fails with KBBQ_ECUDA.
#include "../../include/kbbq_b200.h"
#include "kbbq_long.h"

#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "kbbq_host.h"
#include "kbbq_kernels.cuh"
#include "kbbq_synth.cuh"
#include "kbbq_nvls.cuh"

using namespace kbbq;

namespace {

thread_local std::string g_err;

// host threads the library may use for the model fit: 16 at most, fewer when several processes share the box
// (KBBQ_HOST_THREADS, e.g. cores / ranks: bench.py sets it)
int host_threads() {
    unsigned n = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("KBBQ_HOST_THREADS")) n = (unsigned)std::max(1, std::min(64, atoi(e)));
    return (int)n;
}

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(e__ == cudaErrorMemoryAllocation ? KBBQ_ENOMEM : KBBQ_ECUDA,                          \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                                 \
    } while (0)

struct Filter {
    bool exists = false;
    host::BloomGeom geom{};
    unsigned long long* d_words = nullptr;
    unsigned long long* d_patterns = nullptr;
    unsigned long long* d_inserted = nullptr;
    // KBBQ_FILTER_VMM=1: the bit array is virtual-memory-API memory (shareable: it can be bound to a multicast object)
    bool vmm = false;
    CUmemGenericAllocationHandle mem = 0;
    size_t vmm_bytes = 0;
    // NVSwitch multicast (kbbq_nvls.cuh): the object, what this rank did with it, and its mapping
    CUmemGenericAllocationHandle mc = 0;
    bool mc_have = false, mc_added = false, mc_bound = false;
    unsigned long long* mc_ptr = nullptr;
    int mc_world = 0;
    // peer mappings of the other ranks' bit arrays (kbbq_bloom_or_allreduce_p2p), opened once per filter
    void* peer[KBBQ_MAX_PEERS] = {};
    int peer_world = 0, peer_rank = -1;
    void close_peers() {
        for (int r = 0; r < peer_world; ++r)
            if (r != peer_rank && peer[r]) cudaIpcCloseMemHandle(peer[r]);
        std::memset(peer, 0, sizeof peer);
        peer_world = 0;
        peer_rank = -1;
    }
    BloomView view() const {
        BloomView v;
        v.words = d_words;
        v.patterns = d_patterns;
        v.inserted = d_inserted;
        v.n_blocks = geom.n_blocks;
        v.fastmod_m = geom.n_blocks > 1 ? (~0ull / geom.n_blocks) + 1 : 0;
        v.salt0 = geom.salts[0];
        v.salt1 = geom.salts[1];
        return v;
    }
};

}  // namespace

namespace {
// give back a filter's bit array and whatever multicast state hangs on it
void filter_release(int device, Filter& f) {
    const nvls::Driver& d = nvls::driver();
    f.close_peers();
    if (f.mc_ptr && d.ok) {
        d.MemUnmap((CUdeviceptr)f.mc_ptr, f.vmm_bytes);
        d.MemAddressFree((CUdeviceptr)f.mc_ptr, f.vmm_bytes);
    }
    if (f.mc_bound && d.ok) {
        CUdevice dev;
        if (d.DeviceGet(&dev, device) == CUDA_SUCCESS) d.MulticastUnbind(f.mc, dev, 0, f.vmm_bytes);
    }
    if (f.mc_have && d.ok) d.MemRelease(f.mc);
    if (f.vmm && d.ok) {
        d.MemUnmap((CUdeviceptr)f.d_words, f.vmm_bytes);
        d.MemAddressFree((CUdeviceptr)f.d_words, f.vmm_bytes);
        d.MemRelease(f.mem);
    } else {
        cudaFree(f.d_words);
    }
    cudaFree(f.d_inserted);
    f = Filter();
}
// the bit array as shareable virtual-memory-API memory; false (with a reason) when that is not to be had
bool filter_alloc_vmm(int device, size_t bytes, Filter& f, std::string& why) {
    const nvls::Driver& d = nvls::driver();
    size_t padded = 0;
    if (!nvls::padded_size(device, bytes, padded, why)) return false;
    const CUmemAllocationProp prop = nvls::alloc_prop(device);
    CUmemGenericAllocationHandle h = 0;
    CUresult r = d.MemCreate(&h, padded, &prop, 0);
    if (r != CUDA_SUCCESS) { why = nvls::cu_error("cuMemCreate", r); return false; }
    CUdeviceptr va = 0;
    r = d.MemAddressReserve(&va, padded, 0, 0, 0);
    if (r != CUDA_SUCCESS) { d.MemRelease(h); why = nvls::cu_error("cuMemAddressReserve", r); return false; }
    r = d.MemMap(va, padded, 0, h, 0);
    if (r == CUDA_SUCCESS) {
        CUmemAccessDesc acc = {};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = device;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        r = d.MemSetAccess(va, padded, &acc, 1);
        if (r != CUDA_SUCCESS) d.MemUnmap(va, padded);
    }
    if (r != CUDA_SUCCESS) {
        d.MemAddressFree(va, padded);
        d.MemRelease(h);
        why = nvls::cu_error("cuMemMap / cuMemSetAccess", r);
        return false;
    }
    f.vmm = true;
    f.mem = h;
    f.vmm_bytes = padded;
    f.d_words = reinterpret_cast<unsigned long long*>(va);
    return true;
}
}  // namespace

#define KBBQ_MAX_PASS_KERNELS 4

struct kbbq_ctx {
    int device = 0;
    int k = 32;
    int n_rg = 1;
    int max_cycle = 0;
    int sm_count = 148;
    int l2_bytes = 126 << 20;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // side stream: quality arrays travel while pass 1 runs on `stream`
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evk[KBBQ_MAX_PASS_KERNELS + 1] = {};  // marks around the kernels of the most recent pass
    cudaEvent_t evp[4] = {};                          // pass 4: piece i computed
    const char* kname[KBBQ_MAX_PASS_KERNELS] = {};
    int n_marks = 0;
    bool timed = false;
    uint64_t launches = 0;
    Filter filt[2];
    // bumped whenever a filter's bits may have been REMOVED or replaced (create, clear, upload, load): a batch remembers
    // the epoch at which it inserted k-mers, and a later pass skips the probes of k-mers it knows to be present only
    // while the epoch still matches (merges and further inserts only add bits: they keep the knowledge valid)
    uint64_t epoch[2] = {1, 1};
    // sampler
    bool sampler_seeded = false;
    uint64_t sampler_state0[4] = {0, 0, 0, 0};  // state right after Seed()
    uint64_t sampler_pos = 0;
    double alpha = 0;
    unsigned long long sample_threshold = 0;  // draw x is sampled iff take_all or x < threshold (see kbbq_sampler_seed)
    int sample_all = 0;
    unsigned long long* d_jump = nullptr;  // 26 matrices: 2^(10+b) steps
    unsigned long long* d_lane_jump = nullptr;  // 32 matrices: lane x 1024 steps
    // thresholds
    bool have_thresholds = false;
    int* d_thresholds = nullptr;
    // count tables (one allocation, four tables back to back)
    unsigned long long* d_counts = nullptr;
    size_t n_rgc = 0, n_qc = 0, n_cyc = 0, n_dic = 0;
    // model lookup tables
    bool have_model = false;
    int16_t* d_base = nullptr;
    int8_t* d_cy = nullptr;
    int8_t* d_di = nullptr;
    unsigned int* d_work = nullptr;
    // set by qual_range_kernel when an uploaded batch holds a quality above 93 (kbbq_batch is a public struct: callers may
    // fill it without kbbq_pack_reads); mapped pinned memory, reported by the next synchronising call
    int* h_badq = nullptr;
    int* d_badq = nullptr;
    bool count_attr_set = false;
    bool p4_attr_set = false;
    // both filters' pattern tables (2 x 4 MiB) in ONE allocation: every probe and insert reads one pattern, pass 2
    // reads both tables; measured on the config-3 shape, pass 2 runs in 59.4 ms with the tables adjacent against 70.2 ms
    // with two separate cudaMallocs (marking them persisting in L2 on top of that changed nothing)
    unsigned long long* d_pattern_pool = nullptr;

    CountView counts() const {
        CountView c;
        c.rgc = d_counts;
        c.qc = d_counts + n_rgc;
        c.cyc = d_counts + n_rgc + n_qc;
        c.dic = d_counts + n_rgc + n_qc + n_cyc;
        c.max_cycle = max_cycle;
        return c;
    }
    size_t n_counts() const { return n_rgc + n_qc + n_cyc + n_dic; }
};

struct kbbq_dbatch {
    uint32_t n_reads = 0;
    uint32_t max_len = 0;
    uint64_t n_words = 0;
    uint64_t n_slots = 0;
    int k = 0;
    unsigned long long* d_bases = nullptr;
    uint32_t* d_nmask = nullptr;
    uint8_t* d_qual = nullptr;
    uint2* d_qmask = nullptr;        // per word: qual <= 2 / qual <= 6 bit masks
    bool qmask_needs_quals = false;  // derived on the device from d_qual (the host brought none): ready with qual_ready
    uint32_t* d_word_off = nullptr;
    uint32_t* d_len = nullptr;
    uint16_t* d_rg = nullptr;
    uint8_t* d_second = nullptr;
    unsigned long long* d_slot_off = nullptr;
    uint32_t* d_sample_bits = nullptr;
    uint64_t sample_epoch = 0;      // ctx->epoch[SAMPLED] when pass 1 inserted this batch's sampled k-mers (0: never)
    uint32_t* d_ins_bits = nullptr;  // per k-mer (word layout): inserted into `trusted` by pass 2
    uint64_t ins_epoch = 0;         // ctx->epoch[TRUSTED] when pass 2 did
    uint32_t* d_err_bits = nullptr;
    uint32_t* d_word_read = nullptr;  // read index of every word (flat traversal), made on demand for non-uniform batches
    uint32_t uniform_wpr = 0;         // words per read when every read occupies the same number of words, else 0
    uint32_t* d_tflags = nullptr;  // pass 3: trusted flag of every k-mer (word layout), scan -> repair
    uint32_t* d_repair = nullptr;  // pass 3: indices of the reads the scan could not finish
    uint32_t* d_mod = nullptr;  // per read: the base get_errors left modified, or 0xFFFFFFFF (see covariate_count_kernel)
    uint8_t* d_newqual = nullptr;
    // pass 4 hands its output to the host in KBBQ_P4_PARTS pieces, each copied while the next is computed
    uint32_t part_read[5] = {0, 0, 0, 0, 0};   // read boundaries of the pieces
    uint64_t part_word[5] = {0, 0, 0, 0, 0};   // ... and their word offsets (from the host's word_off at upload)
    cudaEvent_t qual_ready = nullptr;  // recorded on the copy stream after d_qual's upload
    cudaEvent_t qmask_ready = nullptr; // kbbq_batch_upload_async with host masks: recorded on the copy stream after d_qmask's upload
    bool qmask_waited = true;
    bool qual_waited = true;           // false until `stream` has been made to wait for qual_ready
    bool all_on_copy = false;          // kbbq_batch_upload_stream: every array is ready with qual_ready, not before
    BatchView view() const {
        BatchView v;
        v.bases = d_bases;
        v.nmask = d_nmask;
        v.qual = d_qual;
        v.qmask = d_qmask;
        v.word_off = d_word_off;
        v.len = d_len;
        v.rg = d_rg;
        v.second = d_second;
        v.slot_off = d_slot_off;
        v.n_reads = n_reads;
        return v;
    }
};

namespace {

int use(kbbq_ctx* ctx) {
    if (!ctx) return fail(KBBQ_EINVAL, "null context");
    CU(cudaSetDevice(ctx->device));
    return KBBQ_OK;
}

// after a synchronising call: did any uploaded batch bring a quality the tables cannot index?
int check_quals(kbbq_ctx* ctx) {
    if (ctx->h_badq && *(volatile int*)ctx->h_badq)
        return fail(KBBQ_ERANGE, "quality above 93 in an uploaded batch: outside the model's table range (bases holding one were skipped)");
    return KBBQ_OK;
}

void pass_begin(kbbq_ctx* ctx) {
    cudaEventRecord(ctx->ev0, ctx->stream);
    ctx->n_marks = 0;
}
void pass_end(kbbq_ctx* ctx) {
    cudaEventRecord(ctx->ev1, ctx->stream);
    ctx->timed = true;
}
// call right before launching kernel `name`; the previous kernel's interval ends here
void mark_kernel(kbbq_ctx* ctx, const char* name) {
    if (ctx->n_marks >= KBBQ_MAX_PASS_KERNELS) return;
    cudaEventRecord(ctx->evk[ctx->n_marks], ctx->stream);
    ctx->kname[ctx->n_marks++] = name;
}
// passes that read the qualities wait (once) for their upload on the copy stream
int need_quals(kbbq_ctx* ctx, kbbq_dbatch* b) {
    if (!b->qual_waited) {
        CU(cudaStreamWaitEvent(ctx->stream, b->qual_ready, 0));
        b->qual_waited = true;
    }
    return KBBQ_OK;
}

// pass 2 reads the quality MASKS only: made by the host they travel right behind what pass 1 needs and ahead of the
// qualities; made on the device (or on a streamed batch) they are ready when the qualities are
int need_qmask(kbbq_ctx* ctx, kbbq_dbatch* b) {
    if (b->qmask_needs_quals || b->all_on_copy) return need_quals(ctx, b);
    if (!b->qmask_waited) {
        CU(cudaStreamWaitEvent(ctx->stream, b->qmask_ready, 0));
        b->qmask_waited = true;
    }
    return KBBQ_OK;
}

// one warp per read; enough blocks to fill every SM's 64 warp slots, no more than the reads need
int read_grid(const kbbq_ctx* ctx, uint32_t n_reads) {
    const long long need = ((long long)n_reads + KBBQ_WARPS_PER_BLOCK - 1) / KBBQ_WARPS_PER_BLOCK;
    const long long full = (long long)ctx->sm_count * (64 / KBBQ_WARPS_PER_BLOCK);
    return (int)std::max(1ll, std::min(need, full));
}

int check_launch(kbbq_ctx* ctx, const char* what) {
    ++ctx->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(KBBQ_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return KBBQ_OK;
}

// the read of every word, for the flat traversal of batches whose reads differ in word count (made once per batch)
int flat_map(kbbq_ctx* ctx, kbbq_dbatch* b, FlatMap& fm) {
    fm.word_read = nullptr;
    fm.wpr = b->uniform_wpr ? b->uniform_wpr : 1;
    if (b->uniform_wpr || !b->n_reads) return KBBQ_OK;
    if (!b->d_word_read) {
        CU(cudaMallocAsync(&b->d_word_read, std::max<size_t>(b->n_words, 1) * 4, ctx->stream));
        word_read_kernel<<<(b->n_reads + 255) / 256, 256, 0, ctx->stream>>>(b->d_word_off, b->n_reads, b->d_word_read);
        if (int rc = check_launch(ctx, "word_read_kernel")) return rc;
    }
    fm.word_read = b->d_word_read;
    return KBBQ_OK;
}

// covariate_count_kernel: rows of the block-private table that fit beside a second resident block, and the launch(es)
int launch_count(kbbq_ctx* ctx, kbbq_dbatch* b, const uint32_t* d_err, const uint32_t* d_mod, uint32_t n_first) {
    const int row_cells = 4 * ctx->max_cycle + 32;
    const long long fixed_words = KBBQ_HIST_DIRECT_RG * KBBQ_NQ + 2;
    const long long budget_words = (108 * 1024) / 4 - fixed_words;  // 2 blocks/SM inside the 227 KB of an SM
    int S = (int)std::max(0ll, std::min(64ll, budget_words / (row_cells + 1)));
    if (const char* e = getenv("KBBQ_HIST_S")) S = std::min(S, atoi(e));  // measurement override
    const size_t smem = ((size_t)fixed_words + (size_t)S + (size_t)S * row_cells) * 4;
    if (!ctx->count_attr_set) {
        CU(cudaFuncSetAttribute(covariate_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        ctx->count_attr_set = true;
    }
    FlatMap fm;
    if (int rc = flat_map(ctx, b, fm)) return rc;
    const unsigned long long n_half = (unsigned long long)b->n_words * 2;
    const long long need = (long long)((n_half + KBBQ_HIST_THREADS - 1) / KBBQ_HIST_THREADS);
    const int grid = (int)std::max(1ll, std::min(need, (long long)ctx->sm_count * 2));
    // a block's 32-bit cells must not overflow: a half word adds at most 16 to a cell, so at most 2^26 half words per
    // block and launch
    const unsigned long long per_launch = (unsigned long long)grid << 26;
    for (unsigned long long first = 0; first < n_half; first += per_launch) {
        const unsigned long long end = std::min(n_half, first + per_launch);
        covariate_count_kernel<<<grid, KBBQ_HIST_THREADS, smem, ctx->stream>>>(b->view(), fm, ctx->counts(), d_err, d_mod, first, end, n_first, S);
        if (int rc = check_launch(ctx, "covariate_count_kernel")) return rc;
    }
    return KBBQ_OK;
}

// d_slot_off[0..n] = exclusive prefix sum of max(0, len[r]-k+1) (the k-mer slots of the reads before r):
// slot_counts_kernel writes the n counts and a trailing 0, cub scans the n+1 items in place
int slot_prefix(kbbq_ctx* ctx, const uint32_t* d_len, uint32_t n, unsigned long long* d_slot_off, cudaStream_t st = nullptr) {
    if (!st) st = ctx->stream;
    slot_counts_kernel<<<(unsigned)(((size_t)n + 1 + 255) / 256), 256, 0, st>>>(d_len, n, ctx->k, d_slot_off);
    if (int rc = check_launch(ctx, "slot_counts_kernel")) return rc;
    size_t tmp_bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_slot_off, d_slot_off, (size_t)n + 1, st));
    void* tmp = nullptr;
    CU(cudaMallocAsync(&tmp, std::max<size_t>(tmp_bytes, 16), st));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_slot_off, d_slot_off, (size_t)n + 1, st);
    cudaFreeAsync(tmp, st);
    if (e != cudaSuccess) return fail(KBBQ_ECUDA, std::string("slot prefix: ") + cudaGetErrorString(e));
    return KBBQ_OK;
}

// How the pass-1 / pass-2 kernels insert (kbbq_kernels.cuh: INS_*): directly while the filter fits the L2; through
// bucketed staging once it does not AND the batch brings at least two items per 64-byte block of the table (sparser than
// that, a slice's lines are touched about once either way and the staging traffic buys nothing) AND the staging array
// fits a sane share of the device; read-before-OR otherwise.  KBBQ_INSERT_MODE=0|1|2 overrides (measurement, tests).
int stage_shift() {  // KBBQ_STAGE_SHIFT: tests cut a small filter into many buckets with it
    const char* e = getenv("KBBQ_STAGE_SHIFT");  // read per call: tests change it between runs of one process
    return e ? std::max(0, std::min(40, atoi(e))) : KBBQ_STAGE_BUCKET_SHIFT;
}
int insert_mode(const kbbq_ctx* ctx, int which, uint64_t n_slots) {
    const host::BloomGeom& g = ctx->filt[which].geom;
    const uint64_t n_buckets = (g.n_blocks + (1ull << stage_shift()) - 1) >> stage_shift();
    if (const char* e = getenv("KBBQ_INSERT_MODE")) return atoi(e) == INS_STAGED && n_buckets > 2048 ? INS_TEST_FIRST : atoi(e);
    if (g.table_bits / 8 <= (size_t)ctx->l2_bytes * 3 / 4) return INS_DIRECT;
    return n_slots >= 2 * g.n_blocks && n_buckets <= 2048 && n_slots * 10 <= (24ull << 30) ? INS_STAGED : INS_TEST_FIRST;
}

struct Staging {  // device buffers of one staged-insert launch; released in stream order when it goes out of scope
    kbbq_ctx* ctx = nullptr;
    unsigned long long *items = nullptr, *cursor = nullptr;
    StageSink sink{nullptr, nullptr, 0, 0, 0, KBBQ_STAGE_BUCKET_SHIFT};
    Staging() = default;
    Staging(const Staging&) = delete;
    Staging& operator=(const Staging&) = delete;
    ~Staging() {
        if (!ctx) return;
        if (items) cudaFreeAsync(items, ctx->stream);
        if (cursor) cudaFreeAsync(cursor, ctx->stream);
    }
    size_t smem() const { return (size_t)sink.n_buckets * 12; }  // per block: state (8 B) + overflow flag (4 B) per bucket
};
// one region per 16 MiB slice of filter `which`: the expected share of a batch whose every slot inserts, a quarter more,
// plus one granule per thread block (each block leaves its last granule of a region part empty)
int staging_begin(kbbq_ctx* ctx, int which, uint64_t n_slots, int grid_blocks, Staging& st) {
    st.ctx = ctx;
    const host::BloomGeom& g = ctx->filt[which].geom;
    StageSink& k = st.sink;
    k.shift = stage_shift();
    k.n_buckets = (uint32_t)((g.n_blocks + (1ull << k.shift) - 1) >> k.shift);
    k.granule = k.n_buckets > 512 ? 32 : 128;
    if (const char* e = getenv("KBBQ_STAGE_CAP")) {  // tests: a tiny region forces the overflow path
        k.cap = std::max<unsigned long long>(k.granule, strtoull(e, nullptr, 10) / k.granule * k.granule);
    } else {
        const unsigned long long want = n_slots / k.n_buckets + n_slots / k.n_buckets / 4 + (unsigned long long)(grid_blocks + 1) * k.granule;
        k.cap = (want + k.granule - 1) / k.granule * k.granule;
    }
    CU(cudaMallocAsync(&st.items, (size_t)k.n_buckets * k.cap * 8, ctx->stream));
    CU(cudaMallocAsync(&st.cursor, (size_t)k.n_buckets * 8, ctx->stream));
    CU(cudaMemsetAsync(st.cursor, 0, (size_t)k.n_buckets * 8, ctx->stream));
    k.items = st.items;
    k.cursor = st.cursor;
    return KBBQ_OK;
}
// apply the staged (block, pattern) items to filter `which`, region after region; nothing comes back to the host
int staging_apply(kbbq_ctx* ctx, int which, Staging& st) {
    mark_kernel(ctx, "bloom_apply_staged_kernel");
    bloom_apply_staged_kernel<<<ctx->sm_count * 8, 256, ((size_t)st.sink.n_buckets + 1) * 8, ctx->stream>>>(ctx->filt[which].view(), st.sink);
    return check_launch(ctx, "bloom_apply_staged_kernel");
}

int filter_ok(kbbq_ctx* ctx, int which) {
    if (which != KBBQ_FILTER_SAMPLED && which != KBBQ_FILTER_TRUSTED) return fail(KBBQ_EINVAL, "filter id must be 0 or 1");
    if (!ctx->filt[which].exists) return fail(KBBQ_ESTATE, "filter has not been created");
    return KBBQ_OK;
}

void geom_to_c(const host::BloomGeom& g, kbbq_bloom_geom* o) {
    o->projected = g.projected; o->fpr = g.fpr; o->seed = g.seed; o->optimal_k = g.optimal_k;
    o->optimal_table_size = g.optimal_table_size; o->random_seed = g.random_seed; o->table_bits = g.table_bits;
    o->n_blocks = g.n_blocks; o->n_salts = g.n_salts;
    std::memset(o->salts, 0, sizeof o->salts);
    std::memcpy(o->salts, g.salts, sizeof(uint32_t) * g.n_salts);
}
void geom_from_c(const kbbq_bloom_geom* o, host::BloomGeom& g) {
    g.projected = o->projected; g.fpr = o->fpr; g.seed = o->seed; g.optimal_k = o->optimal_k;
    g.optimal_table_size = o->optimal_table_size; g.random_seed = o->random_seed; g.table_bits = o->table_bits;
    g.n_blocks = o->n_blocks; g.n_salts = o->n_salts;
    std::memcpy(g.salts, o->salts, sizeof g.salts);
}

}  // namespace

// ---- synthetic batches (measurement tooling, kbbq_synth.cuh): one thread per 32-base word ---------------------------
namespace kbbq {
__global__ void __launch_bounds__(256) synth_reads_kernel(SynthSpec s, unsigned long long first_read, uint32_t n_reads, uint32_t wpr,
                                                          unsigned long long* __restrict__ bases, uint32_t* __restrict__ nmask,
                                                          uint8_t* __restrict__ qual, uint2* __restrict__ qmask,
                                                          uint32_t* __restrict__ word_off, uint32_t* __restrict__ len,
                                                          uint16_t* __restrict__ rg, uint8_t* __restrict__ second) {
    const unsigned long long gw = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gw >= (unsigned long long)n_reads * wpr) return;
    const uint32_t r = (uint32_t)(gw / wpr), w = (uint32_t)(gw % wpr);
    const unsigned long long gr = first_read + r;
    const SynthRead rd = synth_read(s, gr);
    unsigned long long bw = 0;
    uint32_t q4[8] = {0, 0, 0, 0, 0, 0, 0, 0}, m2 = 0, m6 = 0;
    const uint32_t cnt = s.read_len - 32 * w < 32 ? s.read_len - 32 * w : 32;
#pragma unroll
    for (uint32_t i = 0; i < 32; ++i) {
        if (i >= cnt) break;
        int code, q;
        synth_base(s, rd, 32 * w + i, code, q);
        bw |= (unsigned long long)code << (2 * i);
        q4[i >> 2] |= (uint32_t)q << (8 * (i & 3));
        m2 |= (uint32_t)(q <= 2) << i;
        m6 |= (uint32_t)(q <= 6) << i;
    }
    if (cnt < 32) {  // padding bytes hold quality 0: kbbq_quality_masks flags them too (every consumer masks by length)
        m2 |= ~0u << cnt;
        m6 |= ~0u << cnt;
    }
    bases[gw] = bw;
    nmask[gw] = 0;
    uint4* qo = reinterpret_cast<uint4*>(qual + gw * 32);
    qo[0] = make_uint4(q4[0], q4[1], q4[2], q4[3]);
    qo[1] = make_uint4(q4[4], q4[5], q4[6], q4[7]);
    qmask[gw] = make_uint2(m2, m6);
    if (w == 0) {
        word_off[r] = r * wpr;
        len[r] = s.read_len;
        rg[r] = (uint16_t)((gr >> 1) % s.n_rg);
        second[r] = (uint8_t)(gr & 1);
        if (r == n_reads - 1) word_off[n_reads] = n_reads * wpr;
    }
}
}  // namespace kbbq

namespace {
int synth_spec(const kbbq_synth_params* p, SynthSpec& s) {
    if (!p) return fail(KBBQ_EINVAL, "null synth parameters");
    if (p->read_len < 1 || p->read_len > KBBQ_ABI_MAX_READ_LEN || p->genome_len < p->read_len || p->n_rg < 1 || p->n_rg > 65535 ||
        !(p->err_scale >= 0))
        return fail(KBBQ_EINVAL, "bad synth parameters");
    s.seed = p->seed; s.genome_seed = p->genome_seed; s.genome_len = p->genome_len; s.read_len = p->read_len; s.n_rg = p->n_rg;
    const double pq[5] = {0.01, 0.04, 0.10, 0.35, 0.50};
    const int qv[5] = {2, 11, 25, 37, 40};
    double cum = 0;
    for (int i = 0; i < 5; ++i) {
        cum += pq[i];
        s.q_cum[i] = cum >= 1.0 ? 0xFFFFFFFFu : (uint32_t)(cum * 4294967296.0);
        const double pe = std::min(1.0, p->err_scale * std::pow(10.0, -qv[i] / 10.0));
        s.err_thr[i] = pe >= 1.0 ? 0xFFFFFFFFu : (uint32_t)(pe * 4294967296.0);
    }
    return KBBQ_OK;
}
}  // namespace

extern "C" {

const char* kbbq_last_error(void) { return g_err.c_str(); }

int kbbq_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int kbbq_device_warmup(int device) {
    // creates the device's primary context (hundreds of ms to seconds in a fresh process): callable from a side thread
    // while the caller parses its input
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess)
        return fail(KBBQ_ECUDA, std::string("no usable CUDA device ") + std::to_string(device) + " (this library has no CPU path)");
    return KBBQ_OK;
}

int kbbq_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(KBBQ_EINVAL, "null output");
    cudaError_t e = cudaHostAlloc(out, std::max<size_t>(bytes, 16), cudaHostAllocPortable);
    if (e != cudaSuccess) return fail(KBBQ_ENOMEM, std::string("pinned host allocation: ") + cudaGetErrorString(e));
    return KBBQ_OK;
}
void kbbq_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
int kbbq_device_reset(int device) {
    if (cudaSetDevice(device) != cudaSuccess) return fail(KBBQ_ECUDA, "no such device");
    CU(cudaDeviceReset());
    return KBBQ_OK;
}
int kbbq_device_memory(int device, uint64_t* free_bytes, uint64_t* total_bytes) {
    if (cudaSetDevice(device) != cudaSuccess) return fail(KBBQ_ECUDA, "no such device");
    size_t f = 0, t = 0;
    CU(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return KBBQ_OK;
}

// ---- host-only helpers ---------------------------------------------------------------------------------------------
int kbbq_bloom_geometry(uint64_t projected, double fpr, uint64_t seed, kbbq_bloom_geom* out) {
    if (!out) return fail(KBBQ_EINVAL, "null output");
    host::BloomGeom g;
    if (!host::bloom_geometry(projected, fpr, seed, g))
        return fail(KBBQ_EINVAL, "Error: Invalid bloom filter parameters. Adjust parameters and try again.");
    geom_to_c(g, out);
    return KBBQ_OK;
}

int kbbq_bloom_patterns(const kbbq_bloom_geom* g, uint64_t* patterns) {
    if (!g || !patterns) return fail(KBBQ_EINVAL, "null argument");
    host::BloomGeom hg;
    geom_from_c(g, hg);
    host::bloom_patterns(hg, patterns);
    return KBBQ_OK;
}

double kbbq_effective_fpp(uint64_t table_bits, uint64_t inserted, uint32_t n_salts) {
    return host::effective_fpp(table_bits, inserted, n_salts);
}

int kbbq_thresholds_from_fpr(double fpr, long double alpha, int k, int* thresholds) {
    if (!thresholds || k < 1 || k > KBBQ_MAX_K) return fail(KBBQ_EINVAL, "k must be in 1..32");
    host::calculate_thresholds(k, host::calculate_phit(fpr, alpha), thresholds);
    return KBBQ_OK;
}

int kbbq_phit(double fpr, long double alpha, long double* p) {
    if (!p) return fail(KBBQ_EINVAL, "null output");
    *p = host::calculate_phit(fpr, alpha);
    return KBBQ_OK;
}

int kbbq_thresholds_from_p(int k, long double p, int* thresholds) {
    if (!thresholds || k < 1 || k > KBBQ_MAX_K) return fail(KBBQ_EINVAL, "k must be in 1..32");
    host::calculate_thresholds(k, p, thresholds);
    return KBBQ_OK;
}

uint32_t kbbq_default_seed(uint64_t time_now, uint32_t pid) { return host::default_seed(time_now, pid); }

int kbbq_rng_state(uint32_t seed, uint64_t skip, uint64_t state[4]) {
    host::Xoshiro256ss r;
    r.seed(seed);
    host::rng_advance(r.s, skip);
    std::memcpy(state, r.s, 32);
    return KBBQ_OK;
}

int kbbq_fit_model_host(int n_rg, int max_cycle, const uint64_t* rgc, const uint64_t* qc, const uint64_t* cyc,
                        const uint64_t* dic, int32_t* meanq, int32_t* rgdq, int32_t* qdq, int32_t* cydq, int32_t* didq) {
    if (n_rg < 1 || max_cycle < 1) return fail(KBBQ_EINVAL, "bad table extents");
    const int nt = host_threads();
    host::fit_model(n_rg, max_cycle, rgc, qc, cyc, dic, meanq, rgdq, qdq, cydq, didq, nt);
    return KBBQ_OK;
}

int kbbq_pack_size(uint64_t n, const uint64_t* off, uint64_t* n_words, uint32_t* max_len) {
    uint64_t words = 0;
    uint32_t ml = 0;
    for (uint64_t r = 0; r < n; ++r) {
        const uint64_t len = off[r + 1] - off[r];
        if (len > KBBQ_ABI_MAX_READ_LEN) return fail(KBBQ_ERANGE, "read longer than KBBQ_ABI_MAX_READ_LEN");
        words += (len + 31) / 32;
        ml = std::max<uint32_t>(ml, (uint32_t)len);
    }
    if (words > 0xFFFFFFFFull) return fail(KBBQ_ERANGE, "batch too large: split it");
    *n_words = words;
    if (max_len) *max_len = ml;
    return KBBQ_OK;
}

int kbbq_pack_reads(uint64_t n, const char* seq, const uint64_t* off, const uint8_t* qual, uint64_t* bases,
                    uint32_t* nmask, uint8_t* qual_out, uint32_t* word_off, uint32_t* len, int n_threads) {
    uint64_t words = 0;
    for (uint64_t r = 0; r < n; ++r) {
        const uint64_t l = off[r + 1] - off[r];
        if (l > KBBQ_ABI_MAX_READ_LEN) return fail(KBBQ_ERANGE, "read longer than KBBQ_ABI_MAX_READ_LEN");
        word_off[r] = (uint32_t)words;
        len[r] = (uint32_t)l;
        words += (l + 31) / 32;
    }
    word_off[n] = (uint32_t)words;
    unsigned char code[256];
    for (int c = 0; c < 256; ++c) code[c] = (unsigned char)host::nt4((unsigned char)c);
    std::atomic<bool> bad_q{false};
    auto work = [&](uint64_t lo, uint64_t hi) {
        bool bad = false;
        for (uint64_t r = lo; r < hi; ++r) {
            const char* s = seq + off[r];
            const uint8_t* q = qual + off[r];
            const uint32_t l = len[r], wo = word_off[r];
            for (uint32_t w = 0; w * 32 < l; ++w) {
                uint64_t bw = 0;
                uint32_t nw = 0;
                const uint32_t cnt = std::min<uint32_t>(32, l - w * 32);
                for (uint32_t i = 0; i < cnt; ++i) {
                    const unsigned c = code[(unsigned char)s[w * 32 + i]];
                    if (c < 4) bw |= (uint64_t)c << (2 * i);
                    else nw |= 1u << i;
                    const uint8_t qq = q[w * 32 + i];
                    if (qq > KBBQ_MAXQ) bad = true;
                    qual_out[(size_t)(wo + w) * 32 + i] = qq;
                }
                for (uint32_t i = cnt; i < 32; ++i) qual_out[(size_t)(wo + w) * 32 + i] = 0;
                bases[wo + w] = bw;
                nmask[wo + w] = nw;
            }
        }
        if (bad) bad_q.store(true, std::memory_order_relaxed);
    };
    if (n_threads <= 1 || n < 4096) {
        work(0, n);
    } else {
        std::vector<std::thread> th;
        const uint64_t per = (n + n_threads - 1) / n_threads;
        for (int t = 0; t < n_threads; ++t) {
            const uint64_t lo = std::min<uint64_t>(n, (uint64_t)t * per), hi = std::min<uint64_t>(n, lo + per);
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto& t : th) t.join();
    }
    if (bad_q.load()) return fail(KBBQ_ERANGE, "quality above 93: outside the model's table range");
    return KBBQ_OK;
}

int kbbq_quality_masks(uint64_t n_words, const uint8_t* qual, uint32_t* qmask, int n_threads) {
    if (n_words && (!qual || !qmask)) return fail(KBBQ_EINVAL, "null argument");
    auto work = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t w = lo; w < hi; ++w) {
            const uint8_t* q = qual + w * 32;
            uint32_t m2 = 0, m6 = 0;
            for (int i = 0; i < 32; ++i) {
                m2 |= (uint32_t)(q[i] <= 2) << i;
                m6 |= (uint32_t)(q[i] <= 6) << i;
            }
            qmask[2 * w] = m2;
            qmask[2 * w + 1] = m6;
        }
    };
    if (n_threads <= 1 || n_words < (1u << 16)) {
        work(0, n_words);
    } else {
        std::vector<std::thread> th;
        const uint64_t per = (n_words + n_threads - 1) / n_threads;
        for (int t = 0; t < n_threads; ++t) {
            const uint64_t lo = std::min<uint64_t>(n_words, (uint64_t)t * per), hi = std::min<uint64_t>(n_words, lo + per);
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto& t : th) t.join();
    }
    return KBBQ_OK;
}

// ---- context -------------------------------------------------------------------------------------------------------
int kbbq_ctx_create(int device, int k, int n_rg, int max_cycle, kbbq_ctx** out) {
    if (!out) return fail(KBBQ_EINVAL, "null output");
    if (k < 1 || k > KBBQ_MAX_K) return fail(KBBQ_EINVAL, "Error: k must be <= 32 and > 0.");
    if (n_rg < 1 || n_rg > 65535) return fail(KBBQ_EINVAL, "n_rg out of range");
    if (max_cycle < 1 || max_cycle > KBBQ_ABI_MAX_READ_LEN) return fail(KBBQ_EINVAL, "max_cycle out of range");
    // the kernels index the model and count tables with 32-bit offsets
    if ((unsigned long long)n_rg * KBBQ_NQ * 2ull * (unsigned long long)max_cycle >= (1ull << 31))
        return fail(KBBQ_ERANGE, "n_rg x max_cycle too large for the model tables");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(KBBQ_ECUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(KBBQ_EINVAL, "no such device");
    CU(cudaSetDevice(device));
    kbbq_ctx* c = new (std::nothrow) kbbq_ctx();
    if (!c) return fail(KBBQ_ENOMEM, "out of host memory");
    c->device = device;
    c->k = k;
    c->n_rg = n_rg;
    c->max_cycle = max_cycle;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&c->l2_bytes, cudaDevAttrL2CacheSize, device);
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {   // batch buffers come from the stream-ordered pool and stay cached in it: cudaMalloc/cudaFree of hundreds of MB
        // per batch cost up to a second each on this driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    for (auto& e : c->evk) CU(cudaEventCreate(&e));
    for (auto& e : c->evp) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->n_rgc = (size_t)n_rg * 2;
    c->n_qc = (size_t)n_rg * KBBQ_NQ * 2;
    c->n_cyc = (size_t)n_rg * KBBQ_NQ * 2 * max_cycle * 2;
    c->n_dic = (size_t)n_rg * KBBQ_NQ * 16 * 2;
    CU(cudaMalloc(&c->d_counts, c->n_counts() * 8));
    CU(cudaMemsetAsync(c->d_counts, 0, c->n_counts() * 8, c->stream));
    CU(cudaMalloc(&c->d_thresholds, sizeof(int) * (KBBQ_MAX_K + 1)));
    CU(cudaMalloc(&c->d_base, sizeof(int16_t) * (size_t)n_rg * KBBQ_NQ));
    CU(cudaMalloc(&c->d_cy, (size_t)n_rg * KBBQ_NQ * 2 * max_cycle));
    CU(cudaMalloc(&c->d_di, (size_t)n_rg * KBBQ_NQ * 16));
    CU(cudaMalloc(&c->d_work, 2 * sizeof(unsigned int)));
    CU(cudaHostAlloc(&c->h_badq, sizeof(int), cudaHostAllocMapped));
    *c->h_badq = 0;
    CU(cudaHostGetDevicePointer(&c->d_badq, c->h_badq, 0));
    CU(cudaMalloc(&c->d_pattern_pool, (size_t)2 * KBBQ_NUM_PATTERNS * 64));
    *out = c;
    return KBBQ_OK;
}

void kbbq_ctx_destroy(kbbq_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& f : c->filt) filter_release(c->device, f);
    cudaFree(c->d_jump);
    cudaFree(c->d_lane_jump);
    cudaFree(c->d_thresholds);
    cudaFree(c->d_counts);
    cudaFree(c->d_base);
    cudaFree(c->d_cy);
    cudaFree(c->d_di);
    cudaFree(c->d_work);
    cudaFreeHost(c->h_badq);
    cudaFree(c->d_pattern_pool);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    for (auto& e : c->evk) cudaEventDestroy(e);
    for (auto& e : c->evp) cudaEventDestroy(e);
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->stream);
    delete c;
}

int kbbq_sync(kbbq_ctx* ctx) {
    if (int rc = use(ctx)) return rc;
    CU(cudaStreamSynchronize(ctx->copy_stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return check_quals(ctx);
}

void* kbbq_stream(kbbq_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int kbbq_last_pass_ms(kbbq_ctx* ctx, float* ms) {
    if (int rc = use(ctx)) return rc;
    if (!ctx->timed) return fail(KBBQ_ESTATE, "no pass has run");
    CU(cudaEventSynchronize(ctx->ev1));
    CU(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return KBBQ_OK;
}

int kbbq_last_pass_kernels(kbbq_ctx* ctx, int max, const char** names, float* ms) {
    if (int rc = use(ctx)) return rc < 0 ? rc : -rc;
    if (!ctx->timed) return fail(KBBQ_ESTATE, "no pass has run");
    if (cudaEventSynchronize(ctx->ev1) != cudaSuccess) return fail(KBBQ_ECUDA, "event sync failed");
    int n = 0;
    for (int i = 0; i < ctx->n_marks && n < max; ++i, ++n) {
        float t = 0;
        cudaEventElapsedTime(&t, ctx->evk[i], i + 1 < ctx->n_marks ? ctx->evk[i + 1] : ctx->ev1);
        if (names) names[n] = ctx->kname[i];
        if (ms) ms[n] = t;
    }
    return n;
}

uint64_t kbbq_launch_count(kbbq_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- filters -------------------------------------------------------------------------------------------------------
int kbbq_bloom_create(kbbq_ctx* ctx, int which, uint64_t projected, double fpr, uint64_t seed) {
    if (int rc = use(ctx)) return rc;
    if (which != 0 && which != 1) return fail(KBBQ_EINVAL, "filter id must be 0 or 1");
    host::BloomGeom g;
    if (!host::bloom_geometry(projected, fpr, seed, g))
        return fail(KBBQ_EINVAL, "Error: Invalid bloom filter parameters. Adjust parameters and try again.");
    Filter& f = ctx->filt[which];
    if (f.exists) {
        CU(cudaStreamSynchronize(ctx->stream));
        filter_release(ctx->device, f);
    }
    ++ctx->epoch[which];
    f.geom = g;
    size_t table_alloc = g.table_bits / 8;
    if (getenv("KBBQ_FILTER_VMM")) {  // shareable virtual-memory-API memory, so that the array can join a multicast object
        std::string why;
        if (filter_alloc_vmm(ctx->device, g.table_bits / 8, f, why)) table_alloc = f.vmm_bytes;
        else if (getenv("KBBQ_NVLS_VERBOSE")) fprintf(stderr, "[kbbq_b200] filter %d stays cudaMalloc memory: %s\n", which, why.c_str());
    }
    if (!f.vmm) CU(cudaMalloc(&f.d_words, g.table_bits / 8));
    if (table_alloc > g.table_bits / 8) CU(cudaMemsetAsync((char*)f.d_words + g.table_bits / 8, 0, table_alloc - g.table_bits / 8, ctx->stream));
    f.d_patterns = ctx->d_pattern_pool + (size_t)which * KBBQ_NUM_PATTERNS * 8;
    CU(cudaMalloc(&f.d_inserted, 8));
    CU(cudaMemsetAsync(f.d_words, 0, g.table_bits / 8, ctx->stream));
    CU(cudaMemsetAsync(f.d_inserted, 0, 8, ctx->stream));
    std::vector<uint64_t> pat((size_t)KBBQ_NUM_PATTERNS * 8);
    host::bloom_patterns(g, pat.data());
    CU(cudaMemcpyAsync(f.d_patterns, pat.data(), pat.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    f.exists = true;
    return KBBQ_OK;
}

int kbbq_bloom_info(kbbq_ctx* ctx, int which, kbbq_bloom_geom* out) {
    if (!ctx || !out) return fail(KBBQ_EINVAL, "null argument");
    if (int rc = filter_ok(ctx, which)) return rc;
    geom_to_c(ctx->filt[which].geom, out);
    return KBBQ_OK;
}

int kbbq_bloom_clear(kbbq_ctx* ctx, int which) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    Filter& f = ctx->filt[which];
    ++ctx->epoch[which];
    CU(cudaMemsetAsync(f.d_words, 0, f.geom.table_bits / 8, ctx->stream));
    CU(cudaMemsetAsync(f.d_inserted, 0, 8, ctx->stream));
    return KBBQ_OK;
}

int kbbq_bloom_inserted(kbbq_ctx* ctx, int which, uint64_t* inserted) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    CU(cudaMemcpyAsync(inserted, ctx->filt[which].d_inserted, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KBBQ_OK;
}

int kbbq_bloom_download(kbbq_ctx* ctx, int which, uint64_t* words, uint64_t* inserted) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    Filter& f = ctx->filt[which];
    if (words) CU(cudaMemcpyAsync(words, f.d_words, f.geom.table_bits / 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (inserted) CU(cudaMemcpyAsync(inserted, f.d_inserted, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KBBQ_OK;
}

int kbbq_bloom_upload(kbbq_ctx* ctx, int which, const uint64_t* words, uint64_t inserted) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    Filter& f = ctx->filt[which];
    ++ctx->epoch[which];
    if (words) CU(cudaMemcpyAsync(f.d_words, words, f.geom.table_bits / 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(f.d_inserted, &inserted, 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KBBQ_OK;
}

int kbbq_bloom_insert_keys(kbbq_ctx* ctx, int which, const uint64_t* keys, uint64_t n) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    if (n == 0) return KBBQ_OK;
    unsigned long long* d = nullptr;
    CU(cudaMalloc(&d, n * 8));
    CU(cudaMemcpyAsync(d, keys, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    bloom_insert_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->filt[which].view(), d, n);
    int rc = check_launch(ctx, "bloom_insert_keys_kernel");
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    return rc;
}

int kbbq_bloom_query_keys(kbbq_ctx* ctx, int which, const uint64_t* keys, uint64_t n, uint8_t* hits) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    if (n == 0) return KBBQ_OK;
    unsigned long long* d = nullptr;
    uint8_t* dh = nullptr;
    CU(cudaMalloc(&d, n * 8));
    CU(cudaMalloc(&dh, n));
    CU(cudaMemcpyAsync(d, keys, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    bloom_query_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->filt[which].view(), d, n, dh);
    int rc = check_launch(ctx, "bloom_query_keys_kernel");
    if (rc == KBBQ_OK) {
        cudaMemcpyAsync(hits, dh, n, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(KBBQ_ECUDA, "query sync failed");
    }
    cudaFree(d);
    cudaFree(dh);
    return rc;
}

int kbbq_bloom_fprate(kbbq_ctx* ctx, int which, double* fpr) {
    uint64_t ins = 0;
    if (int rc = kbbq_bloom_inserted(ctx, which, &ins)) return rc;
    const host::BloomGeom& g = ctx->filt[which].geom;
    *fpr = host::effective_fpp(g.table_bits, ins, g.n_salts);
    return KBBQ_OK;
}

static int digest_words(kbbq_ctx* ctx, const unsigned long long* d_words, uint64_t n, uint64_t out[2]) {
    unsigned long long* d_out = nullptr;
    CU(cudaMallocAsync(&d_out, 16, ctx->stream));
    CU(cudaMemsetAsync(d_out, 0, 16, ctx->stream));
    if (n) {
        const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 16);
        digest_kernel<<<grid, 256, 0, ctx->stream>>>(d_words, n, d_out);
        if (int rc = check_launch(ctx, "digest_kernel")) return rc;
    }
    CU(cudaMemcpyAsync(out, d_out, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaFreeAsync(d_out, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KBBQ_OK;
}

int kbbq_bloom_digest(kbbq_ctx* ctx, int which, uint64_t out[2]) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    if (!out) return fail(KBBQ_EINVAL, "null output");
    return digest_words(ctx, ctx->filt[which].d_words, ctx->filt[which].geom.table_bits / 64, out);
}

int kbbq_covariates_digest(kbbq_ctx* ctx, uint64_t out[2]) {
    if (int rc = use(ctx)) return rc;
    if (!out) return fail(KBBQ_EINVAL, "null output");
    return digest_words(ctx, ctx->d_counts, ctx->n_counts(), out);
}

int kbbq_batch_newqual_digest(kbbq_ctx* ctx, const kbbq_dbatch* b, uint64_t out[2]) {
    if (int rc = use(ctx)) return rc;
    if (!b || !out) return fail(KBBQ_EINVAL, "null argument");
    if (!b->d_newqual) return fail(KBBQ_ESTATE, "pass 4 has not run on this batch");
    return digest_words(ctx, reinterpret_cast<const unsigned long long*>(b->d_newqual), b->n_words * 4, out);
}

void* kbbq_bloom_device_words(kbbq_ctx* ctx, int which) {
    return (ctx && (which == 0 || which == 1) && ctx->filt[which].exists) ? ctx->filt[which].d_words : nullptr;
}
void* kbbq_bloom_device_counter(kbbq_ctx* ctx, int which) {
    return (ctx && (which == 0 || which == 1) && ctx->filt[which].exists) ? ctx->filt[which].d_inserted : nullptr;
}

int kbbq_bloom_or_from(kbbq_ctx* ctx, int which, const void* other, uint64_t first_word, uint64_t n_words) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    Filter& f = ctx->filt[which];
    if ((first_word & 1) || (n_words & 1) || first_word + n_words > f.geom.table_bits / 64)
        return fail(KBBQ_EINVAL, "word range must be even-aligned and inside the table");
    if (n_words == 0) return KBBQ_OK;
    const unsigned long long n2 = n_words / 2;
    const int grid = (int)std::min<unsigned long long>((n2 + 255) / 256, (unsigned long long)ctx->sm_count * 8);
    or_words_kernel<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<ulonglong2*>(f.d_words + first_word),
                                                   reinterpret_cast<const ulonglong2*>((const unsigned long long*)other + first_word), n2);
    return check_launch(ctx, "or_words_kernel");
}

// ---- filter (de)serialisation: checkpointing a pass, cross-run fixtures --------------------------------------------
// File: "KBBQBF1\0", the kbbq_bloom_geom the filter was created with, the insert counter, then the bit array exactly
// as the reference's bit_table_ (bloom.hh:33) lays it out.
int kbbq_bloom_save(kbbq_ctx* ctx, int which, const char* path) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    if (!path) return fail(KBBQ_EINVAL, "null path");
    Filter& f = ctx->filt[which];
    FILE* fh = fopen(path, "wb");
    if (!fh) return fail(KBBQ_EINVAL, std::string("cannot open ") + path);
    kbbq_bloom_geom g;
    geom_to_c(f.geom, &g);
    uint64_t inserted = 0;
    bool ok = fwrite("KBBQBF1", 1, 8, fh) == 8 && fwrite(&g, sizeof g, 1, fh) == 1;
    cudaError_t e = cudaMemcpyAsync(&inserted, f.d_inserted, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    ok = ok && e == cudaSuccess && fwrite(&inserted, 8, 1, fh) == 1;
    const size_t total = f.geom.table_bits / 8, chunk = (size_t)64 << 20;
    std::vector<unsigned char> buf(std::min(total, chunk));
    for (size_t off = 0; ok && off < total; off += chunk) {  // 64 MiB at a time: the human-scale filters are tens of GB
        const size_t n = std::min(chunk, total - off);
        e = cudaMemcpyAsync(buf.data(), (const unsigned char*)f.d_words + off, n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        ok = e == cudaSuccess && fwrite(buf.data(), 1, n, fh) == n;
    }
    ok = (fclose(fh) == 0) && ok;
    if (!ok) return fail(e != cudaSuccess ? KBBQ_ECUDA : KBBQ_EINVAL, std::string("writing ") + path + " failed");
    return KBBQ_OK;
}

int kbbq_bloom_load(kbbq_ctx* ctx, int which, const char* path) {
    if (int rc = use(ctx)) return rc;
    if (which != 0 && which != 1) return fail(KBBQ_EINVAL, "filter id must be 0 or 1");
    if (!path) return fail(KBBQ_EINVAL, "null path");
    FILE* fh = fopen(path, "rb");
    if (!fh) return fail(KBBQ_EINVAL, std::string("cannot open ") + path);
    char magic[8];
    kbbq_bloom_geom g;
    uint64_t inserted = 0;
    if (fread(magic, 1, 8, fh) != 8 || std::memcmp(magic, "KBBQBF1", 8) != 0 || fread(&g, sizeof g, 1, fh) != 1 ||
        fread(&inserted, 8, 1, fh) != 1) {
        fclose(fh);
        return fail(KBBQ_EINVAL, std::string(path) + " is not a kbbq_b200 filter file");
    }
    if (int rc = kbbq_bloom_create(ctx, which, g.projected, g.fpr, g.seed)) { fclose(fh); return rc; }
    Filter& f = ctx->filt[which];
    if (f.geom.table_bits != g.table_bits || f.geom.n_blocks != g.n_blocks || f.geom.n_salts != g.n_salts ||
        std::memcmp(f.geom.salts, g.salts, sizeof(uint32_t) * g.n_salts) != 0) {
        fclose(fh);
        return fail(KBBQ_EINVAL, "the file's filter geometry is not what its parameters give on this build");
    }
    const size_t total = f.geom.table_bits / 8, chunk = (size_t)64 << 20;
    std::vector<unsigned char> buf(std::min(total, chunk));
    bool ok = true;
    cudaError_t e = cudaSuccess;
    for (size_t off = 0; ok && off < total; off += chunk) {
        const size_t n = std::min(chunk, total - off);
        ok = fread(buf.data(), 1, n, fh) == n;
        if (ok) e = cudaMemcpyAsync((unsigned char*)f.d_words + off, buf.data(), n, cudaMemcpyHostToDevice, ctx->stream);
        if (ok && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        ok = ok && e == cudaSuccess;
    }
    fclose(fh);
    if (ok) e = cudaMemcpyAsync(f.d_inserted, &inserted, 8, cudaMemcpyHostToDevice, ctx->stream);
    if (ok && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (!ok || e != cudaSuccess) return fail(e != cudaSuccess ? KBBQ_ECUDA : KBBQ_EINVAL, std::string("reading ") + path + " failed");
    return KBBQ_OK;
}

int kbbq_bloom_ipc_handle(kbbq_ctx* ctx, int which, void* handle64) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    if (!handle64) return fail(KBBQ_EINVAL, "null handle buffer");
    static_assert(sizeof(cudaIpcMemHandle_t) == KBBQ_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->filt[which].d_words));
    std::memcpy(handle64, &h, sizeof h);
    return KBBQ_OK;
}

int kbbq_bloom_or_allreduce_p2p(kbbq_ctx* ctx, int which, int rank, int world, const void* handles) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    if (world < 1 || world > KBBQ_MAX_PEERS || rank < 0 || rank >= world) return fail(KBBQ_EINVAL, "bad rank/world (at most 16 peers)");
    if (world == 1) return KBBQ_OK;
    if (!handles) return fail(KBBQ_EINVAL, "null handles");
    Filter& f = ctx->filt[which];
    if (f.peer_world != world || f.peer_rank != rank) {
        f.close_peers();
        for (int r = 0; r < world; ++r) {
            if (r == rank) {
                f.peer[r] = f.d_words;
                continue;
            }
            cudaIpcMemHandle_t h;
            std::memcpy(&h, (const char*)handles + (size_t)r * sizeof h, sizeof h);
            cudaError_t e = cudaIpcOpenMemHandle(&f.peer[r], h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                f.peer_world = r;  // close what was opened so far
                f.peer_rank = rank;
                f.close_peers();
                return fail(KBBQ_ECUDA, std::string("cudaIpcOpenMemHandle (peer ") + std::to_string(r) + "): " + cudaGetErrorString(e));
            }
        }
        f.peer_world = world;
        f.peer_rank = rank;
    }
    PeerPtrs pp;
    std::memset(&pp, 0, sizeof pp);
    for (int s2 = 0; s2 < world; ++s2) pp.p[s2] = reinterpret_cast<ulonglong2*>(f.peer[(rank + s2) % world]);
    const unsigned long long n2 = f.geom.table_bits / 128;  // table_bits is a multiple of 512
    const unsigned long long per = (n2 + world - 1) / world;
    const int grid = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((per + 255) / 256, (unsigned long long)ctx->sm_count * 8));
    pass_begin(ctx);
    mark_kernel(ctx, "or_allreduce_p2p_kernel");
    switch (world) {
    case 2: or_allreduce_p2p_kernel<2><<<grid, 256, 0, ctx->stream>>>(pp, rank, world, n2); break;
    case 4: or_allreduce_p2p_kernel<4><<<grid, 256, 0, ctx->stream>>>(pp, rank, world, n2); break;
    case 8: or_allreduce_p2p_kernel<8><<<grid, 256, 0, ctx->stream>>>(pp, rank, world, n2); break;
    default: or_allreduce_p2p_kernel<0><<<grid, 256, 0, ctx->stream>>>(pp, rank, world, n2); break;
    }
    if (int rc = check_launch(ctx, "or_allreduce_p2p_kernel")) return rc;
    pass_end(ctx);
    return KBBQ_OK;
}

// ---- the merge through NVSwitch multicast (kbbq_nvls.cuh) -------------------------------------------------------------
int kbbq_bloom_mc_supported(kbbq_ctx* ctx, int which) {
    if (!ctx || (which != 0 && which != 1) || !ctx->filt[which].exists || !ctx->filt[which].vmm) return 0;
    const nvls::Driver& d = nvls::driver();
    if (!d.ok || cudaSetDevice(ctx->device) != cudaSuccess) return 0;
    CUdevice dev;
    int yes = 0;
    if (d.DeviceGet(&dev, ctx->device) != CUDA_SUCCESS || d.DeviceGetAttribute(&yes, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) return 0;
    return yes ? 1 : 0;
}

int kbbq_bloom_mc_create(kbbq_ctx* ctx, int which, int world, int* fd_out) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    if (!fd_out || world < 2) return fail(KBBQ_EINVAL, "bad argument");
    Filter& f = ctx->filt[which];
    if (!f.vmm) return fail(KBBQ_ESTATE, "the filter was not created as virtual-memory-API memory (KBBQ_FILTER_VMM=1)");
    if (f.mc_have) return fail(KBBQ_ESTATE, "the filter already has a multicast object");
    const nvls::Driver& d = nvls::driver();
    CUmulticastObjectProp prop = {};
    prop.numDevices = (unsigned)world;
    prop.size = f.vmm_bytes;
    prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUresult r = d.MulticastCreate(&f.mc, &prop);
    if (r != CUDA_SUCCESS) return fail(KBBQ_ECUDA, nvls::cu_error("cuMulticastCreate", r));
    f.mc_have = true;
    f.mc_world = world;
    int fd = -1;
    r = d.MemExportToShareableHandle(&fd, f.mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) return fail(KBBQ_ECUDA, nvls::cu_error("cuMemExportToShareableHandle", r));
    *fd_out = fd;
    return KBBQ_OK;
}

int kbbq_bloom_mc_import(kbbq_ctx* ctx, int which, int world, int fd) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    Filter& f = ctx->filt[which];
    if (!f.vmm) return fail(KBBQ_ESTATE, "the filter was not created as virtual-memory-API memory (KBBQ_FILTER_VMM=1)");
    if (f.mc_have) return fail(KBBQ_ESTATE, "the filter already has a multicast object");
    CUresult r = nvls::driver().MemImportFromShareableHandle(&f.mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (r != CUDA_SUCCESS) return fail(KBBQ_ECUDA, nvls::cu_error("cuMemImportFromShareableHandle", r));
    f.mc_have = true;
    f.mc_world = world;
    return KBBQ_OK;
}

int kbbq_bloom_mc_add_device(kbbq_ctx* ctx, int which) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    Filter& f = ctx->filt[which];
    if (!f.mc_have) return fail(KBBQ_ESTATE, "no multicast object");
    const nvls::Driver& d = nvls::driver();
    CUdevice dev;
    CUresult r = d.DeviceGet(&dev, ctx->device);
    if (r == CUDA_SUCCESS) r = d.MulticastAddDevice(f.mc, dev);
    if (r != CUDA_SUCCESS) return fail(KBBQ_ECUDA, nvls::cu_error("cuMulticastAddDevice", r));
    f.mc_added = true;
    return KBBQ_OK;
}

int kbbq_bloom_mc_bind(kbbq_ctx* ctx, int which) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    Filter& f = ctx->filt[which];
    if (!f.mc_added) return fail(KBBQ_ESTATE, "this rank's device has not been added to the multicast object");
    const nvls::Driver& d = nvls::driver();
    CUresult r = d.MulticastBindMem(f.mc, 0, f.mem, 0, f.vmm_bytes, 0);
    if (r != CUDA_SUCCESS) return fail(KBBQ_ECUDA, nvls::cu_error("cuMulticastBindMem", r));
    f.mc_bound = true;
    CUdeviceptr va = 0;
    r = d.MemAddressReserve(&va, f.vmm_bytes, 0, 0, 0);
    if (r != CUDA_SUCCESS) return fail(KBBQ_ECUDA, nvls::cu_error("cuMemAddressReserve (multicast)", r));
    r = d.MemMap(va, f.vmm_bytes, 0, f.mc, 0);
    if (r == CUDA_SUCCESS) {
        CUmemAccessDesc acc = {};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = ctx->device;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        r = d.MemSetAccess(va, f.vmm_bytes, &acc, 1);
        if (r != CUDA_SUCCESS) d.MemUnmap(va, f.vmm_bytes);
    }
    if (r != CUDA_SUCCESS) {
        d.MemAddressFree(va, f.vmm_bytes);
        return fail(KBBQ_ECUDA, nvls::cu_error("cuMemMap / cuMemSetAccess (multicast)", r));
    }
    f.mc_ptr = reinterpret_cast<unsigned long long*>(va);
    return KBBQ_OK;
}

int kbbq_bloom_or_allreduce_mc(kbbq_ctx* ctx, int which, int rank, int world) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, which)) return rc;
    Filter& f = ctx->filt[which];
    if (!f.mc_ptr || world != f.mc_world || rank < 0 || rank >= world) return fail(KBBQ_ESTATE, "the multicast mapping is not set up for this world");
    const unsigned long long n_words = f.geom.table_bits / 64;
    const unsigned long long per = (n_words + world - 1) / world;
    const int grid = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((per + 255) / 256, (unsigned long long)ctx->sm_count * 16));
    pass_begin(ctx);
    mark_kernel(ctx, "or_allreduce_mc_kernel");
    nvls::or_allreduce_mc_kernel<<<grid, 256, 0, ctx->stream>>>(f.mc_ptr, rank, world, n_words);
    if (int rc = check_launch(ctx, "or_allreduce_mc_kernel")) return rc;
    pass_end(ctx);
    return KBBQ_OK;
}

// The same merge for several contexts of ONE process (one host thread driving several GPUs, or one thread per GPU that
// meets here): peer access instead of IPC handles, the kernel of every rank launched from this call, counters summed
// on the host.  Synchronises every context's stream before and after.
int kbbq_bloom_or_allreduce_local(kbbq_ctx* const* ctxs, int n, int which) {
    if (!ctxs || n < 1 || n > KBBQ_MAX_PEERS) return fail(KBBQ_EINVAL, "bad context list (at most 16 peers)");
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i]) return fail(KBBQ_EINVAL, "null context");
        if (int rc = filter_ok(ctxs[i], which)) return rc;
        if (ctxs[i]->filt[which].geom.table_bits != ctxs[0]->filt[which].geom.table_bits) return fail(KBBQ_EINVAL, "filters differ in size");
        for (int j = 0; j < i; ++j)
            if (ctxs[j]->device == ctxs[i]->device) return fail(KBBQ_EINVAL, "two contexts on one device");
    }
    if (n == 1) return KBBQ_OK;
    for (int i = 0; i < n; ++i) {
        CU(cudaSetDevice(ctxs[i]->device));
        for (int j = 0; j < n; ++j) {
            if (i == j) continue;
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, ctxs[i]->device, ctxs[j]->device));
            if (!can) return fail(KBBQ_ECUDA, "no peer access between devices " + std::to_string(ctxs[i]->device) + " and " + std::to_string(ctxs[j]->device));
            cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return fail(KBBQ_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        }
        CU(cudaStreamSynchronize(ctxs[i]->stream));  // this rank's inserts are in memory
    }
    const unsigned long long n2 = ctxs[0]->filt[which].geom.table_bits / 128;
    const unsigned long long per = (n2 + n - 1) / n;
    for (int r = 0; r < n; ++r) {
        kbbq_ctx* ctx = ctxs[r];
        CU(cudaSetDevice(ctx->device));
        PeerPtrs pp;
        std::memset(&pp, 0, sizeof pp);
        for (int s2 = 0; s2 < n; ++s2) pp.p[s2] = reinterpret_cast<ulonglong2*>(ctxs[(r + s2) % n]->filt[which].d_words);
        const int grid = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((per + 255) / 256, (unsigned long long)ctx->sm_count * 8));
        pass_begin(ctx);
        mark_kernel(ctx, "or_allreduce_p2p_kernel");
        switch (n) {
        case 2: or_allreduce_p2p_kernel<2><<<grid, 256, 0, ctx->stream>>>(pp, r, n, n2); break;
        case 4: or_allreduce_p2p_kernel<4><<<grid, 256, 0, ctx->stream>>>(pp, r, n, n2); break;
        case 8: or_allreduce_p2p_kernel<8><<<grid, 256, 0, ctx->stream>>>(pp, r, n, n2); break;
        default: or_allreduce_p2p_kernel<0><<<grid, 256, 0, ctx->stream>>>(pp, r, n, n2); break;
        }
        if (int rc = check_launch(ctx, "or_allreduce_p2p_kernel")) return rc;
        pass_end(ctx);
    }
    unsigned long long total = 0;
    for (int r = 0; r < n; ++r) {
        CU(cudaSetDevice(ctxs[r]->device));
        CU(cudaStreamSynchronize(ctxs[r]->stream));  // every slice is written everywhere
        unsigned long long c = 0;
        CU(cudaMemcpy(&c, ctxs[r]->filt[which].d_inserted, 8, cudaMemcpyDeviceToHost));
        total += c;
    }
    for (int r = 0; r < n; ++r) {
        CU(cudaSetDevice(ctxs[r]->device));
        CU(cudaMemcpy(ctxs[r]->filt[which].d_inserted, &total, 8, cudaMemcpyHostToDevice));
    }
    return KBBQ_OK;
}

// count tables += the caller's (dense layout): the SUM merge of per-GPU partial tables when the host does it
int kbbq_covariates_upload(kbbq_ctx* ctx, const uint64_t* rgc, const uint64_t* qc, const uint64_t* cyc, const uint64_t* dic) {
    if (int rc = use(ctx)) return rc;
    if (!rgc || !qc || !cyc || !dic) return fail(KBBQ_EINVAL, "null table");
    const CountView c = ctx->counts();
    CU(cudaMemcpyAsync(c.rgc, rgc, ctx->n_rgc * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(c.qc, qc, ctx->n_qc * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(c.cyc, cyc, ctx->n_cyc * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(c.dic, dic, ctx->n_dic * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KBBQ_OK;
}

// ---- batches -------------------------------------------------------------------------------------------------------
void kbbq_batch_free(kbbq_ctx* ctx, kbbq_dbatch* b) {
    if (!b) return;
    if (ctx) cudaSetDevice(ctx->device);
    cudaStream_t st = ctx ? ctx->stream : nullptr;
    void* ptrs[] = {b->d_bases, b->d_nmask, b->d_qual, b->d_qmask, b->d_word_off, b->d_len, b->d_rg, b->d_second, b->d_slot_off,
                    b->d_sample_bits, b->d_err_bits, b->d_mod, b->d_newqual, b->d_tflags, b->d_repair, b->d_word_read, b->d_ins_bits};
    if (b->qual_ready) {
        if (ctx && !b->qual_waited) cudaStreamWaitEvent(st, b->qual_ready, 0);  // d_qual is freed in `st`'s order
        cudaEventDestroy(b->qual_ready);
    }
    if (b->qmask_ready) cudaEventDestroy(b->qmask_ready);  // recorded before qual_ready on the same stream
    for (void* p : ptrs)
        if (p) cudaFreeAsync(p, st);
    delete b;
}

// mode 0: everything has landed on return; 1: the quality array may still be in flight on the copy stream (pass 1/2 do
// not read it); 2: every array travels on the copy stream and nothing is waited for -- the upload of the next batch
// overlaps the passes of this one (double buffering for inputs that are streamed through the device pass by pass)
static int batch_upload(kbbq_ctx* ctx, const kbbq_batch* h, kbbq_dbatch** out, int mode);

int kbbq_batch_upload(kbbq_ctx* ctx, const kbbq_batch* h, kbbq_dbatch** out) { return batch_upload(ctx, h, out, 0); }
int kbbq_batch_upload_async(kbbq_ctx* ctx, const kbbq_batch* h, kbbq_dbatch** out) { return batch_upload(ctx, h, out, 1); }
int kbbq_batch_upload_stream(kbbq_ctx* ctx, const kbbq_batch* h, kbbq_dbatch** out) { return batch_upload(ctx, h, out, 2); }

int kbbq_batch_wait(kbbq_ctx* ctx, const kbbq_dbatch* b) {
    if (int rc = use(ctx)) return rc;
    if (!b) return fail(KBBQ_EINVAL, "null batch");
    if (b->qual_ready) CU(cudaEventSynchronize(b->qual_ready));
    return KBBQ_OK;
}

static int batch_upload(kbbq_ctx* ctx, const kbbq_batch* h, kbbq_dbatch** out, int mode) {
    if (int rc = use(ctx)) return rc;
    if (!h || !out) return fail(KBBQ_EINVAL, "null argument");
    if (h->max_len > KBBQ_ABI_MAX_READ_LEN) return fail(KBBQ_ERANGE, "read longer than KBBQ_ABI_MAX_READ_LEN");
    if ((int)h->max_len > ctx->max_cycle) return fail(KBBQ_ERANGE, "read longer than the context's max_cycle");
    const uint32_t n = h->n_reads;
    kbbq_dbatch* b = new (std::nothrow) kbbq_dbatch();
    if (!b) return fail(KBBQ_ENOMEM, "out of host memory");
    b->n_reads = n; b->max_len = h->max_len; b->n_words = h->n_words; b->k = ctx->k;
    for (int i = 0; i <= 4; ++i) {
        b->part_read[i] = (uint32_t)((uint64_t)n * i / 4);
        b->part_word[i] = n ? h->word_off[b->part_read[i]] : 0;
    }
    const size_t nw = std::max<uint64_t>(h->n_words, 1);
    cudaError_t e1 = cudaSuccess;
    auto bail = [&](cudaError_t e) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->copy_stream);
        kbbq_batch_free(ctx, b);
        return fail(e == cudaErrorMemoryAllocation ? KBBQ_ENOMEM : KBBQ_ECUDA, std::string("batch upload: ") + cudaGetErrorString(e));
    };
    // all device arrays come from the stream-ordered pool in `stream`'s order ...
    struct Arr { void** dst; size_t bytes; };
    const Arr arrs[] = {{(void**)&b->d_bases, (nw + 1) * 8}, {(void**)&b->d_nmask, (nw + 1) * 4}, {(void**)&b->d_qmask, nw * 8},
                        {(void**)&b->d_word_off, ((size_t)n + 1) * 4}, {(void**)&b->d_len, std::max<size_t>(n, 4) * 4},
                        {(void**)&b->d_rg, std::max<size_t>(n, 8) * 2}, {(void**)&b->d_second, std::max<size_t>(n, 16)},
                        {(void**)&b->d_slot_off, ((size_t)n + 1) * 8}, {(void**)&b->d_qual, nw * 32}};
    for (const Arr& a : arrs)
        if (e1 == cudaSuccess) e1 = cudaMallocAsync(a.dst, a.bytes, ctx->stream);
    if (e1 != cudaSuccess) return bail(e1);
    // ... and are filled in the copy stream's (mode 2: all of them; otherwise the quality array only)
    cudaEvent_t alloc_done = nullptr;
    e1 = cudaEventCreateWithFlags(&alloc_done, cudaEventDisableTiming);
    if (e1 == cudaSuccess) e1 = cudaEventRecord(alloc_done, ctx->stream);
    if (e1 == cudaSuccess) e1 = cudaStreamWaitEvent(ctx->copy_stream, alloc_done, 0);
    if (alloc_done) cudaEventDestroy(alloc_done);
    if (e1 != cudaSuccess) return bail(e1);
    cudaStream_t cs = mode == 2 ? ctx->copy_stream : ctx->stream;
#define PUT(dst, src, bytes)                                                                                   \
    do {                                                                                                       \
        if (e1 == cudaSuccess && (bytes)) {                                                                    \
            if (src) e1 = cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, cs);                  \
            else e1 = cudaMemsetAsync((dst), 0, (bytes), cs);                                                  \
        }                                                                                                      \
    } while (0)
    // one padding word after the last read keeps two-word k-mer extraction in bounds
    PUT(b->d_bases + h->n_words, (const void*)nullptr, 8);
    PUT(b->d_nmask + h->n_words, (const void*)nullptr, 4);
    PUT(b->d_bases, h->bases, h->n_words * 8);
    PUT(b->d_nmask, h->nmask, h->n_words * 4);
    const bool qmask_behind = mode == 1 && h->qmask && h->n_words;  // pass 1 does not read the masks: they follow on the copy stream
    if (h->qmask && !qmask_behind) PUT(b->d_qmask, h->qmask, (size_t)h->n_words * 8);
    PUT(b->d_word_off, h->word_off, ((size_t)n + 1) * 4);
    PUT(b->d_len, h->len, (size_t)n * 4);
    PUT(b->d_rg, h->rg, (size_t)n * 2);
    PUT(b->d_second, h->second, (size_t)n);
#undef PUT
    if (e1 != cudaSuccess) return bail(e1);
    if (n) {  // slot_off = exclusive prefix of max(0, len-k+1), n+1 entries
        if (int rc = slot_prefix(ctx, b->d_len, n, b->d_slot_off, cs)) {
            kbbq_batch_free(ctx, b);
            return rc;
        }
    } else {
        e1 = cudaMemsetAsync(b->d_slot_off, 0, 8, cs);
    }
    {   // the qualities (3/5 of the bytes) go last and on the copy stream: pass 1 does not read them and starts
        // as soon as everything above has landed
        cudaError_t e2 = e1;
        if (mode != 2 && e2 == cudaSuccess) {  // ... and BEHIND everything above: two concurrent copies would share the link,
            cudaEvent_t small_done = nullptr;  // and the bases pass 1 waits for would arrive at half speed
            e2 = cudaEventCreateWithFlags(&small_done, cudaEventDisableTiming);
            if (e2 == cudaSuccess) e2 = cudaEventRecord(small_done, ctx->stream);
            if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(ctx->copy_stream, small_done, 0);
            if (small_done) cudaEventDestroy(small_done);
        }
        if (e2 == cudaSuccess && qmask_behind) {  // the masks first (all pass 2 needs of the qualities), then the qualities
            e2 = cudaMemcpyAsync(b->d_qmask, h->qmask, (size_t)h->n_words * 8, cudaMemcpyHostToDevice, ctx->copy_stream);
            if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&b->qmask_ready, cudaEventDisableTiming);
            if (e2 == cudaSuccess) e2 = cudaEventRecord(b->qmask_ready, ctx->copy_stream);
            b->qmask_waited = false;
        }
        if (e2 == cudaSuccess && !h->n_words) e2 = cudaMemsetAsync(b->d_qual, 0, 32, ctx->copy_stream);
        if (e2 == cudaSuccess && h->n_words)
            e2 = cudaMemcpyAsync(b->d_qual, h->qual, (size_t)h->n_words * 32, cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e2 == cudaSuccess && h->n_words) {  // the count and apply kernels index tables by quality: flag anything above 93
            const unsigned long long n16 = (unsigned long long)h->n_words * 2;
            qual_range_kernel<<<(unsigned)std::min<unsigned long long>((n16 + 255) / 256, (unsigned long long)ctx->sm_count * 8), 256, 0,
                                ctx->copy_stream>>>(reinterpret_cast<const uint4*>(b->d_qual), n16, ctx->d_badq);
            e2 = cudaGetLastError();
            ++ctx->launches;
        }
        if (e2 == cudaSuccess && !h->qmask && h->n_words) {  // no masks from the host: derive them where the qualities land
            qual_masks_kernel<<<(unsigned)((h->n_words * 32 + 255) / 256), 256, 0, ctx->copy_stream>>>(b->d_qual, h->n_words, b->d_qmask);
            e2 = cudaGetLastError();
            ++ctx->launches;
            b->qmask_needs_quals = true;
        }
        if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&b->qual_ready, cudaEventDisableTiming);
        if (e2 == cudaSuccess) e2 = cudaEventRecord(b->qual_ready, ctx->copy_stream);
        if (e2 != cudaSuccess) return bail(e2);
        b->qual_waited = false;
        b->all_on_copy = mode == 2;
    }
    // validation and the slot total in tight host loops, while the copies above are in flight (the per-read prefix,
    // slot_off, is made on the device)
    {
        unsigned long long slots = 0;
        uint32_t longest = 0, biggest_rg = 0;
        const uint32_t km1 = (uint32_t)ctx->k - 1;
        bool uniform = n > 0;
        const uint32_t wpr0 = n ? (h->len[0] + 31) / 32 : 0;
        // pass 1 cannot be launched before this returns: large batches are checked by a few threads
        const int nt = (int)std::min<uint32_t>((uint32_t)std::min(host_threads(), 8), n / (1u << 18) + 1);
        struct Part { unsigned long long slots = 0; uint32_t longest = 0, biggest_rg = 0; bool uniform = true; };
        std::vector<Part> parts((size_t)nt);
        auto check = [&](int t) {
            Part pt;
            const uint32_t lo = (uint32_t)((uint64_t)n * t / nt), hi = (uint32_t)((uint64_t)n * (t + 1) / nt);
            for (uint32_t r = lo; r < hi; ++r) {
                const uint32_t l = h->len[r];
                pt.longest = std::max(pt.longest, l);
                pt.slots += l > km1 ? l - km1 : 0;
                pt.uniform = pt.uniform && (l + 31) / 32 == wpr0 && h->word_off[r] == r * wpr0;
            }
            if (h->rg)
                for (uint32_t r = lo; r < hi; ++r) pt.biggest_rg = std::max<uint32_t>(pt.biggest_rg, h->rg[r]);
            parts[(size_t)t] = pt;
        };
        {
            std::vector<std::thread> th;
            for (int t = 1; t < nt; ++t) th.emplace_back(check, t);
            check(0);
            for (auto& t : th) t.join();
        }
        for (const Part& pt : parts) {
            slots += pt.slots;
            longest = std::max(longest, pt.longest);
            biggest_rg = std::max(biggest_rg, pt.biggest_rg);
            uniform = uniform && pt.uniform;
        }
        b->uniform_wpr = uniform && wpr0 ? wpr0 : 0;
        b->n_slots = slots;
        const char* bad = longest > h->max_len ? "len[r] exceeds max_len"
                          : (n && biggest_rg >= (uint32_t)ctx->n_rg) ? "read group id >= n_rg" : nullptr;
        if (bad) {
            cudaStreamSynchronize(ctx->stream);
            cudaStreamSynchronize(ctx->copy_stream);
            kbbq_batch_free(ctx, b);
            return fail(KBBQ_ERANGE, bad);
        }
    }
    if (mode != 2) CU(cudaStreamSynchronize(ctx->stream));     // the caller may reuse everything but h->qual
    if (mode == 0) CU(cudaStreamSynchronize(ctx->copy_stream));  // the caller may reuse h->qual
    *out = b;
    return KBBQ_OK;
}

uint64_t kbbq_batch_slots(const kbbq_dbatch* b) { return b ? b->n_slots : 0; }
void* kbbq_batch_device_newqual(kbbq_dbatch* b) { return b ? b->d_newqual : nullptr; }

// ---- pass 1 --------------------------------------------------------------------------------------------------------
static int sampler_install(kbbq_ctx* ctx, const uint64_t state[4], double alpha);

int kbbq_sampler_seed(kbbq_ctx* ctx, uint32_t seed, double alpha) {
    host::Xoshiro256ss r;
    r.seed(seed);
    return sampler_install(ctx, r.s, alpha);
}

int kbbq_sampler_set_state(kbbq_ctx* ctx, const uint64_t state[4], double alpha) {
    if (!state) return fail(KBBQ_EINVAL, "null state");
    return sampler_install(ctx, state, alpha);
}

static int sampler_install(kbbq_ctx* ctx, const uint64_t state[4], double alpha) {
    if (int rc = use(ctx)) return rc;
    std::memcpy(ctx->sampler_state0, state, 32);
    ctx->sampler_pos = 0;
    ctx->alpha = alpha;
    ctx->sampler_seeded = true;
    {   // libstdc++'s bernoulli_distribution(p)(urng) on a 64-bit URBG: double(x) / 2^64, clamped below 1, < p.
        // The predicate is monotone in x: find the smallest x for which it is false.
        auto sampled = [alpha](unsigned long long x) {
            double r = (double)x / 18446744073709551616.0;
            if (r >= 1.0) r = std::nextafter(1.0, 0.0);
            return r < alpha;
        };
        if (sampled(~0ull)) {
            ctx->sample_all = 1;
            ctx->sample_threshold = ~0ull;
        } else {
            unsigned long long lo = 0, hi = ~0ull;  // sampled(hi) is false; find the first false
            while (lo < hi) {
                const unsigned long long mid = lo + (hi - lo) / 2;
                if (sampled(mid)) lo = mid + 1; else hi = mid;
            }
            ctx->sample_all = 0;
            ctx->sample_threshold = lo;
        }
    }
    if (!ctx->d_jump) {
        const host::JumpTables& jt = host::jump_tables();
        CU(cudaMalloc(&ctx->d_jump, (size_t)26 * 1024 * 8));
        CU(cudaMemcpyAsync(ctx->d_jump, jt.matrix(10), (size_t)26 * 1024 * 8, cudaMemcpyHostToDevice, ctx->stream));
        // lane_jump[l] = (1024 steps)^l: column c of matrix l is column c of matrix l-1 advanced by 1024 steps
        std::vector<uint64_t> lj((size_t)32 * 1024, 0);
        for (int c = 0; c < 256; ++c) lj[(size_t)c * 4 + (c >> 6)] = 1ull << (c & 63);  // identity
        for (int l = 1; l < 32; ++l)
            for (int c = 0; c < 256; ++c) {
                uint64_t v[4];
                std::memcpy(v, &lj[((size_t)(l - 1) * 256 + c) * 4], 32);
                host::rng_advance(v, 1024);
                std::memcpy(&lj[((size_t)l * 256 + c) * 4], v, 32);
            }
        CU(cudaMalloc(&ctx->d_lane_jump, lj.size() * 8));
        CU(cudaMemcpyAsync(ctx->d_lane_jump, lj.data(), lj.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return KBBQ_OK;
}

int kbbq_sampler_seek(kbbq_ctx* ctx, uint64_t slot_ordinal) {
    if (!ctx || !ctx->sampler_seeded) return fail(KBBQ_ESTATE, "sampler not seeded");
    ctx->sampler_pos = slot_ordinal;
    return KBBQ_OK;
}

int kbbq_pass1_subsample(kbbq_ctx* ctx, const kbbq_dbatch* cb) {
    if (int rc = use(ctx)) return rc;
    if (!ctx->sampler_seeded) return fail(KBBQ_ESTATE, "sampler not seeded");
    if (int rc = filter_ok(ctx, KBBQ_FILTER_SAMPLED)) return rc;
    kbbq_dbatch* b = const_cast<kbbq_dbatch*>(cb);
    if (!b || b->k != ctx->k) return fail(KBBQ_EINVAL, "batch was uploaded for another k");
    if (b->all_on_copy)
        if (int rc = need_quals(ctx, b)) return rc;
    pass_begin(ctx);
    if (b->n_slots) {
        const uint64_t chunks = (b->n_slots + 1023) / 1024;
        if (!b->d_sample_bits) CU(cudaMallocAsync(&b->d_sample_bits, chunks * 32 * 4, ctx->stream));
        uint64_t rs[4];
        std::memcpy(rs, ctx->sampler_state0, 32);
        host::rng_advance(rs, ctx->sampler_pos);
        mark_kernel(ctx, "rng_sample_kernel");
        rng_sample_kernel<<<(unsigned)((chunks + 127) / 128), 128, 0, ctx->stream>>>(rs[0], rs[1], rs[2], rs[3], ctx->d_jump, ctx->d_lane_jump,
                                                                                     b->n_slots, ctx->sample_threshold, ctx->sample_all, b->d_sample_bits);
        if (int rc = check_launch(ctx, "rng_sample_kernel")) return rc;
        mark_kernel(ctx, "pass1_insert_kernel");
        const int grid = read_grid(ctx, b->n_reads);
        const int mode = insert_mode(ctx, KBBQ_FILTER_SAMPLED, b->n_slots);
        Staging st;
        if (mode == INS_STAGED)
            if (int rc = staging_begin(ctx, KBBQ_FILTER_SAMPLED, b->n_slots, grid, st)) return rc;
        const StageSink sink = st.sink;
        const BloomView fv = ctx->filt[KBBQ_FILTER_SAMPLED].view();
#define P1_LAUNCH(M, L, SM) pass1_insert_kernel<M, L><<<grid, KBBQ_WARPS_PER_BLOCK * 32, SM, ctx->stream>>>(b->view(), fv, ctx->k, b->d_sample_bits, sink)
        if (b->max_len > KBBQ_MAX_READ_LEN) {  // a read longer than the kernel's position list
            if (mode == INS_STAGED) P1_LAUNCH(INS_STAGED, true, st.smem());
            else if (mode == INS_TEST_FIRST) P1_LAUNCH(INS_TEST_FIRST, true, 0);
            else P1_LAUNCH(INS_DIRECT, true, 0);
        } else {
            if (mode == INS_STAGED) P1_LAUNCH(INS_STAGED, false, st.smem());
            else if (mode == INS_TEST_FIRST) P1_LAUNCH(INS_TEST_FIRST, false, 0);
            else P1_LAUNCH(INS_DIRECT, false, 0);
        }
#undef P1_LAUNCH
        if (int rc = check_launch(ctx, "pass1_insert_kernel")) return rc;
        if (mode == INS_STAGED) {
            if (int rc = staging_apply(ctx, KBBQ_FILTER_SAMPLED, st)) return rc;
        }
    }
    pass_end(ctx);
    b->sample_epoch = b->n_slots ? ctx->epoch[KBBQ_FILTER_SAMPLED] : 0;
    ctx->sampler_pos += b->n_slots;
    return KBBQ_OK;
}

// ---- pass 2 --------------------------------------------------------------------------------------------------------
int kbbq_set_thresholds(kbbq_ctx* ctx, const int* thresholds) {
    if (int rc = use(ctx)) return rc;
    if (!thresholds) return fail(KBBQ_EINVAL, "null thresholds");
    int tmp[KBBQ_MAX_K + 1] = {0};
    std::memcpy(tmp, thresholds, sizeof(int) * (ctx->k + 1));
    CU(cudaMemcpyAsync(ctx->d_thresholds, tmp, sizeof tmp, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_thresholds = true;
    return KBBQ_OK;
}

int kbbq_compute_thresholds(kbbq_ctx* ctx, long double alpha, int* thresholds_out) {
    double fpr = 0;
    if (int rc = kbbq_bloom_fprate(ctx, KBBQ_FILTER_SAMPLED, &fpr)) return rc;
    int th[KBBQ_MAX_K + 1] = {0};
    host::calculate_thresholds(ctx->k, host::calculate_phit(fpr, alpha), th);
    if (thresholds_out) std::memcpy(thresholds_out, th, sizeof(int) * (ctx->k + 1));
    return kbbq_set_thresholds(ctx, th);
}

static int fetch_bits(kbbq_ctx* ctx, kbbq_dbatch* b, uint32_t* host_bits) {
    if (!host_bits || !b->n_words) return KBBQ_OK;
    CU(cudaMemcpyAsync(host_bits, b->d_err_bits, b->n_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return check_quals(ctx);
}

int kbbq_pass2_find_trusted(kbbq_ctx* ctx, const kbbq_dbatch* cb, uint32_t* err_bits) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, KBBQ_FILTER_SAMPLED)) return rc;
    if (int rc = filter_ok(ctx, KBBQ_FILTER_TRUSTED)) return rc;
    if (!ctx->have_thresholds) return fail(KBBQ_ESTATE, "thresholds not set");
    kbbq_dbatch* b = const_cast<kbbq_dbatch*>(cb);
    if (!b || b->k != ctx->k) return fail(KBBQ_EINVAL, "batch was uploaded for another k");
    if (err_bits && !b->d_err_bits) {
        CU(cudaMallocAsync(&b->d_err_bits, std::max<size_t>(b->n_words, 1) * 4, ctx->stream));
    }
    if (err_bits) CU(cudaMemsetAsync(b->d_err_bits, 0, std::max<size_t>(b->n_words, 1) * 4, ctx->stream));
    if (int rc = need_qmask(ctx, b)) return rc;
    pass_begin(ctx);
    if (b->n_reads) {
        mark_kernel(ctx, "pass2_trusted_kernel");
        const int grid = read_grid(ctx, b->n_reads);
        const int mode = insert_mode(ctx, KBBQ_FILTER_TRUSTED, b->n_slots);
        Staging st;
        if (mode == INS_STAGED)
            if (int rc = staging_begin(ctx, KBBQ_FILTER_TRUSTED, b->n_slots, grid, st)) return rc;
        const StageSink sink = st.sink;
        const BloomView sv = ctx->filt[KBBQ_FILTER_SAMPLED].view(), tv = ctx->filt[KBBQ_FILTER_TRUSTED].view();
        uint32_t* eo = err_bits ? b->d_err_bits : nullptr;
        // one hash computation per k-mer for both filters when their salts agree (same seed: always, in the reference)
        // knowledge carried between the passes of one batch (see kbbq_ctx::epoch): KBBQ_NO_KNOWN_BITS=1 switches it off
        const bool use_known = getenv("KBBQ_NO_KNOWN_BITS") == nullptr;
        const uint32_t* known = use_known && b->d_sample_bits && b->sample_epoch == ctx->epoch[KBBQ_FILTER_SAMPLED] ? b->d_sample_bits : nullptr;
        if (use_known && !b->d_ins_bits) CU(cudaMallocAsync(&b->d_ins_bits, std::max<size_t>(b->n_words, 1) * 4, ctx->stream));
        // a batch with a read beyond KBBQ_MAX_READ_LEN (the statically sized masks): masks in dynamic shared memory, no
        // parked hashes (8 bytes per k-mer and warp would not fit)
        const bool long_batch = b->max_len > KBBQ_MAX_READ_LEN;
        const int maxw = ((int)b->max_len + 31) / 32;
        const bool shared_hash = !long_batch && sv.salt0 == tv.salt0 && sv.salt1 == tv.salt1 && !getenv("KBBQ_P2_SEPARATE_HASH");
        const int max_kmers = std::max(1, (int)b->max_len - ctx->k + 1);
        const size_t smem = (shared_hash ? (size_t)KBBQ_WARPS_PER_BLOCK * max_kmers * sizeof(uint2) : 0) + (mode == INS_STAGED ? st.smem() : 0) +
                            (long_batch ? (size_t)KBBQ_WARPS_PER_BLOCK * 2 * (maxw + 1) * sizeof(uint32_t) : 0);
        if (smem > 200 * 1024) return fail(KBBQ_ERANGE, "read too long for pass 2's shared-memory masks");
        const int threads = KBBQ_WARPS_PER_BLOCK * 32;
#define P2_LAUNCH_L(M, S, L)                                                                                                         \
    do {                                                                                                                            \
        if (smem > 48 * 1024)                                                                                                       \
            CU(cudaFuncSetAttribute(pass2_trusted_kernel<M, S, PF, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        pass2_trusted_kernel<M, S, PF, L><<<grid, threads, smem, ctx->stream>>>(b->view(), sv, tv, ctx->k, ctx->d_thresholds, eo,   \
                                                                               sink, max_kmers, known, b->d_ins_bits, maxw);        \
    } while (0)
#define P2_LAUNCH(M, S) P2_LAUNCH_L(M, S, false)
        // KBBQ_P2_DEEP=1 (measurement): two queries in flight per warp (QMODE 2).  It wins 7 % in pass 3's scan, which is
        // latency bound at 40 % issue utilisation; this kernel is at 60 % and spills at 64 registers with 64 probes in flight:
        // 43.8 -> 50.0 ms on the config-3 shape, so it stays off here.
        // KBBQ_PREFETCH=1 (measurement): prefetch the blocks of all of a read's k-mers into the L2 before probing.  Measured
        // on the config-3 shape it LOSES -- 44 -> 68 ms here, 37 -> 53 ms in pass 3's scan -- so it stays off.
        const bool big = ctx->filt[KBBQ_FILTER_SAMPLED].geom.table_bits / 8 > (size_t)ctx->l2_bytes * 3 / 4;
        const char* deep_env = getenv("KBBQ_P2_DEEP");
        const int qmode = !shared_hash ? 0 : (big && getenv("KBBQ_PREFETCH")) ? 1 : (deep_env && atoi(deep_env) != 0) ? 2 : 0;
        if (qmode == 2) {
            constexpr int PF = 2;
            if (mode == INS_STAGED) P2_LAUNCH(INS_STAGED, true);
            else if (mode == INS_TEST_FIRST) P2_LAUNCH(INS_TEST_FIRST, true);
            else P2_LAUNCH(INS_DIRECT, true);
        } else if (qmode == 1) {
            constexpr int PF = 1;
            if (mode == INS_STAGED) P2_LAUNCH(INS_STAGED, true);
            else if (mode == INS_TEST_FIRST) P2_LAUNCH(INS_TEST_FIRST, true);
            else P2_LAUNCH(INS_DIRECT, true);
        } else if (shared_hash) {
            constexpr int PF = 0;
            if (mode == INS_STAGED) P2_LAUNCH(INS_STAGED, true);
            else if (mode == INS_TEST_FIRST) P2_LAUNCH(INS_TEST_FIRST, true);
            else P2_LAUNCH(INS_DIRECT, true);
        } else if (long_batch) {
            constexpr int PF = 0;
            if (mode == INS_STAGED) P2_LAUNCH_L(INS_STAGED, false, true);
            else if (mode == INS_TEST_FIRST) P2_LAUNCH_L(INS_TEST_FIRST, false, true);
            else P2_LAUNCH_L(INS_DIRECT, false, true);
        } else {
            constexpr int PF = 0;
            if (mode == INS_STAGED) P2_LAUNCH(INS_STAGED, false);
            else if (mode == INS_TEST_FIRST) P2_LAUNCH(INS_TEST_FIRST, false);
            else P2_LAUNCH(INS_DIRECT, false);
        }
#undef P2_LAUNCH
#undef P2_LAUNCH_L
        if (int rc = check_launch(ctx, "pass2_trusted_kernel")) return rc;
        if (mode == INS_STAGED) {
            if (int rc = staging_apply(ctx, KBBQ_FILTER_TRUSTED, st)) return rc;
        }
    }
    pass_end(ctx);
    b->ins_epoch = b->d_ins_bits && b->n_reads ? ctx->epoch[KBBQ_FILTER_TRUSTED] : 0;
    return fetch_bits(ctx, b, err_bits);
}

// ---- pass 3 --------------------------------------------------------------------------------------------------------
int kbbq_pass3_covariates(kbbq_ctx* ctx, const kbbq_dbatch* cb, uint32_t* err_bits) {
    if (int rc = use(ctx)) return rc;
    if (int rc = filter_ok(ctx, KBBQ_FILTER_TRUSTED)) return rc;
    kbbq_dbatch* b = const_cast<kbbq_dbatch*>(cb);
    if (!b || b->k != ctx->k) return fail(KBBQ_EINVAL, "batch was uploaded for another k");
    if (!b->d_err_bits) CU(cudaMallocAsync(&b->d_err_bits, std::max<size_t>(b->n_words, 1) * 4, ctx->stream));
    if (!b->d_mod) CU(cudaMallocAsync(&b->d_mod, std::max<size_t>(b->n_reads, 1) * 4, ctx->stream));
    // the scan reads bases and N masks only: after kbbq_batch_upload_async it runs while the last of the qualities are
    // still crossing the link, and the wait for them sits between the scan and the repairs (a streamed batch has
    // everything behind one event)
    if (b->all_on_copy)
        if (int rc = need_quals(ctx, b)) return rc;
    pass_begin(ctx);
    if (b->n_reads) {
        if (!b->d_tflags) CU(cudaMallocAsync(&b->d_tflags, std::max<size_t>(b->n_words, 1) * 4, ctx->stream));
        if (!b->d_repair) CU(cudaMallocAsync(&b->d_repair, (size_t)b->n_reads * 4, ctx->stream));
        CU(cudaMemsetAsync(ctx->d_work, 0, 2 * sizeof(unsigned int), ctx->stream));  // [0] repair-list length, [1] work counter
        const BloomView tv = ctx->filt[KBBQ_FILTER_TRUSTED].view();
        mark_kernel(ctx, "pass3_scan_kernel");
        const uint32_t* known = b->d_ins_bits && b->ins_epoch == ctx->epoch[KBBQ_FILTER_TRUSTED] && !getenv("KBBQ_NO_KNOWN_BITS") ? b->d_ins_bits : nullptr;
        const int max_kmers = std::max(1, (int)b->max_len - ctx->k + 1);
        // A filter beyond the L2: two rounds of probes in flight per warp (DEEP; KBBQ_SCAN_DEEP=0 switches it off).
        // KBBQ_PREFETCH=1 (measurement; off because it loses, see pass 2).
        const bool big = ctx->filt[KBBQ_FILTER_TRUSTED].geom.table_bits / 8 > (size_t)ctx->l2_bytes * 3 / 4;
        const char* deep_env = getenv("KBBQ_SCAN_DEEP");
        if (big && getenv("KBBQ_PREFETCH") && b->max_len <= KBBQ_MAX_READ_LEN)
            pass3_scan_kernel<true, false><<<read_grid(ctx, b->n_reads), KBBQ_WARPS_PER_BLOCK * 32, (size_t)KBBQ_WARPS_PER_BLOCK * max_kmers * sizeof(uint2), ctx->stream>>>(
                b->view(), tv, ctx->k, b->d_tflags, b->d_err_bits, b->d_mod, b->d_repair, ctx->d_work, known, max_kmers);
        else if (deep_env ? atoi(deep_env) != 0 : big)
            pass3_scan_kernel<false, true><<<read_grid(ctx, b->n_reads), KBBQ_WARPS_PER_BLOCK * 32, 0, ctx->stream>>>(
                b->view(), tv, ctx->k, b->d_tflags, b->d_err_bits, b->d_mod, b->d_repair, ctx->d_work, known, max_kmers);
        else
            pass3_scan_kernel<false, false><<<read_grid(ctx, b->n_reads), KBBQ_WARPS_PER_BLOCK * 32, 0, ctx->stream>>>(
                b->view(), tv, ctx->k, b->d_tflags, b->d_err_bits, b->d_mod, b->d_repair, ctx->d_work, known, max_kmers);
        if (int rc = check_launch(ctx, "pass3_scan_kernel")) return rc;
        if (int rc = need_quals(ctx, b)) return rc;
        mark_kernel(ctx, "pass3_repair_kernel");
        if (b->max_len <= KBBQ_MAX_READ_LEN) {
            pass3_repair_kernel<<<read_grid(ctx, b->n_reads), KBBQ_WARPS_PER_BLOCK * 32, 0, ctx->stream>>>(
                b->view(), tv, ctx->k, b->d_tflags, b->d_err_bits, b->d_mod, b->d_repair, ctx->d_work, ctx->d_work + 1);
        } else {
            // a batch with long reads (kbbq_long.cu): its repair list is split, the reads of up to KBBQ_MAX_READ_LEN bases
            // go through the common kernel at full speed and only the long ones through the slow one
            uint32_t* split = nullptr;
            unsigned int* counts = nullptr;
            CU(cudaMallocAsync(&split, (size_t)b->n_reads * 4, ctx->stream));
            CU(cudaMallocAsync(&counts, 4 * sizeof(unsigned int), ctx->stream));
            cudaError_t e = launch_split_repair_list(b->d_len, b->d_repair, ctx->d_work, b->n_reads, split, counts, ctx->sm_count, ctx->stream);
            ++ctx->launches;
            if (e == cudaSuccess) {
                pass3_repair_kernel<<<read_grid(ctx, b->n_reads), KBBQ_WARPS_PER_BLOCK * 32, 0, ctx->stream>>>(
                    b->view(), tv, ctx->k, b->d_tflags, b->d_err_bits, b->d_mod, split, counts, counts + 1);
                e = cudaGetLastError();
                ++ctx->launches;
            }
            if (e == cudaSuccess)
                e = launch_repair_long(b->view(), tv, ctx->k, b->d_tflags, b->d_err_bits, b->d_mod, split + (b->n_reads - 1), counts + 2,
                                       counts + 3, b->max_len, ctx->sm_count, ctx->stream);
            cudaFreeAsync(split, ctx->stream);
            cudaFreeAsync(counts, ctx->stream);
            if (e != cudaSuccess) return fail(KBBQ_ECUDA, std::string("pass 3 repairs of a long-read batch: ") + cudaGetErrorString(e));
        }
        if (int rc = check_launch(ctx, "pass3_repair_kernel")) return rc;
        mark_kernel(ctx, "covariate_count_kernel");
        if (int rc = launch_count(ctx, b, b->d_err_bits, b->d_mod, b->n_reads)) return rc;
    }
    pass_end(ctx);
    return fetch_bits(ctx, b, err_bits);
}

int kbbq_consume_errors(kbbq_ctx* ctx, const kbbq_dbatch* cb, const uint32_t* err_bits) {
    return kbbq_consume_errors_n(ctx, cb, err_bits, cb ? cb->n_reads : 0);
}

int kbbq_consume_errors_n(kbbq_ctx* ctx, const kbbq_dbatch* cb, const uint32_t* err_bits, uint32_t n_first) {
    if (int rc = use(ctx)) return rc;
    kbbq_dbatch* b = const_cast<kbbq_dbatch*>(cb);
    if (!b || !err_bits) return fail(KBBQ_EINVAL, "null argument");
    if (n_first > b->n_reads) return fail(KBBQ_EINVAL, "n_first exceeds the batch");
    if (!b->d_err_bits) CU(cudaMallocAsync(&b->d_err_bits, std::max<size_t>(b->n_words, 1) * 4, ctx->stream));
    if (b->n_words) CU(cudaMemcpyAsync(b->d_err_bits, err_bits, b->n_words * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = need_quals(ctx, b)) return rc;
    pass_begin(ctx);
    if (n_first) {
        mark_kernel(ctx, "covariate_count_kernel");
        if (int rc = launch_count(ctx, b, b->d_err_bits, nullptr, n_first)) return rc;
    }
    pass_end(ctx);
    CU(cudaStreamSynchronize(ctx->stream));  // err_bits is the caller's buffer
    return check_quals(ctx);
}

int kbbq_fixed_errors(kbbq_ctx* ctx, const kbbq_dbatch* cb, const kbbq_dbatch* fixed, uint32_t n_first, uint32_t* err_bits) {
    if (int rc = use(ctx)) return rc;
    kbbq_dbatch* b = const_cast<kbbq_dbatch*>(cb);
    if (!b || !fixed) return fail(KBBQ_EINVAL, "null argument");
    if (n_first > b->n_reads || n_first > fixed->n_reads) return fail(KBBQ_EINVAL, "n_first exceeds a batch");
    if (!b->d_err_bits) CU(cudaMallocAsync(&b->d_err_bits, std::max<size_t>(b->n_words, 1) * 4, ctx->stream));
    if (int rc = need_quals(ctx, b)) return rc;
    pass_begin(ctx);
    if (b->n_reads) {
        mark_kernel(ctx, "fixed_diff_kernel");
        fixed_diff_kernel<<<read_grid(ctx, b->n_reads), 256, 0, ctx->stream>>>(b->view(), fixed->view(), n_first, b->d_err_bits);
        if (int rc = check_launch(ctx, "fixed_diff_kernel")) return rc;
    }
    if (n_first) {
        mark_kernel(ctx, "covariate_count_kernel");
        if (int rc = launch_count(ctx, b, b->d_err_bits, nullptr, n_first)) return rc;
    }
    pass_end(ctx);
    return fetch_bits(ctx, b, err_bits);
}

int kbbq_batch_total_bases(kbbq_ctx* ctx, const kbbq_dbatch* b, uint64_t* total) {
    if (int rc = use(ctx)) return rc;
    if (!b || !total) return fail(KBBQ_EINVAL, "null argument");
    unsigned long long* d = nullptr;
    CU(cudaMallocAsync(&d, 8, ctx->stream));
    CU(cudaMemsetAsync(d, 0, 8, ctx->stream));
    if (b->n_reads) {
        const int grid = (int)std::min<uint64_t>(((uint64_t)b->n_reads + 255) / 256, (uint64_t)ctx->sm_count * 8);
        total_bases_kernel<<<grid, 256, 0, ctx->stream>>>(b->d_len, b->n_reads, d);
        if (int rc = check_launch(ctx, "total_bases_kernel")) return rc;
    }
    CU(cudaMemcpyAsync(total, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaFreeAsync(d, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KBBQ_OK;
}

int kbbq_covariates_reset(kbbq_ctx* ctx) {
    if (int rc = use(ctx)) return rc;
    CU(cudaMemsetAsync(ctx->d_counts, 0, ctx->n_counts() * 8, ctx->stream));
    return KBBQ_OK;
}

int kbbq_covariates_download(kbbq_ctx* ctx, uint64_t* rgc, uint64_t* qc, uint64_t* cyc, uint64_t* dic) {
    if (int rc = use(ctx)) return rc;
    const CountView c = ctx->counts();
    if (rgc) CU(cudaMemcpyAsync(rgc, c.rgc, ctx->n_rgc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (qc) CU(cudaMemcpyAsync(qc, c.qc, ctx->n_qc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (cyc) CU(cudaMemcpyAsync(cyc, c.cyc, ctx->n_cyc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (dic) CU(cudaMemcpyAsync(dic, c.dic, ctx->n_dic * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KBBQ_OK;
}

int kbbq_report_tables_host(const char* path, int n_rg, int max_cycle, const uint64_t* rgc, const uint64_t* qc, const uint64_t* cyc,
                            const uint64_t* dic, const char* const* rg_names) {
    if (!path || !rgc || !qc || !cyc || !dic || n_rg < 1 || max_cycle < 1) return fail(KBBQ_EINVAL, "bad argument");
    if (!host::write_gatk_report(path, n_rg, max_cycle, rgc, qc, cyc, dic, rg_names)) return fail(KBBQ_EINVAL, std::string("cannot write ") + path);
    return KBBQ_OK;
}

int kbbq_covariates_report(kbbq_ctx* ctx, const char* path, const char* const* rg_names) {
    if (int rc = use(ctx)) return rc;
    std::vector<uint64_t> counts(ctx->n_counts());
    CU(cudaMemcpyAsync(counts.data(), ctx->d_counts, counts.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const uint64_t* rgc = counts.data();
    return kbbq_report_tables_host(path, ctx->n_rg, ctx->max_cycle, rgc, rgc + ctx->n_rgc, rgc + ctx->n_rgc + ctx->n_qc,
                                   rgc + ctx->n_rgc + ctx->n_qc + ctx->n_cyc, rg_names);
}

void* kbbq_covariates_device(kbbq_ctx* ctx, uint64_t* n_words) {
    if (!ctx) return nullptr;
    if (n_words) *n_words = ctx->n_counts();
    return ctx->d_counts;
}

// ---- model ----------------------------------------------------------------------------------------------------------
int kbbq_set_model(kbbq_ctx* ctx, const int32_t* meanq, const int32_t* rgdq, const int32_t* qdq, const int32_t* cydq,
                   const int32_t* didq) {
    if (int rc = use(ctx)) return rc;
    if (!meanq || !rgdq || !qdq || !cydq || !didq) return fail(KBBQ_EINVAL, "null model table");
    const int R = ctx->n_rg, C = ctx->max_cycle;
    // fold dq_t into the three tables CReadData::recalibrate walks (readutils.cc:579-587)
    std::vector<int16_t> base((size_t)R * KBBQ_NQ);
    std::vector<int8_t> cy((size_t)R * KBBQ_NQ * 2 * C), di((size_t)R * KBBQ_NQ * 16);
    for (int r = 0; r < R; ++r)
        for (int q = 0; q < KBBQ_NQ; ++q) base[(size_t)r * KBBQ_NQ + q] = (int16_t)(meanq[r] + rgdq[r] + qdq[(size_t)r * KBBQ_NQ + q]);
    for (size_t i = 0; i < cy.size(); ++i) cy[i] = (int8_t)cydq[i];  // |dq| <= 93
    for (size_t i = 0; i < di.size(); ++i) di[i] = (int8_t)didq[i];
    CU(cudaMemcpyAsync(ctx->d_base, base.data(), base.size() * 2, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_cy, cy.data(), cy.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_di, di.data(), di.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_model = true;
    return KBBQ_OK;
}

int kbbq_fit_model(kbbq_ctx* ctx, int32_t* meanq, int32_t* rgdq, int32_t* qdq, int32_t* cydq, int32_t* didq) {
    if (int rc = use(ctx)) return rc;
    const int R = ctx->n_rg, C = ctx->max_cycle;
    std::vector<uint64_t> counts(ctx->n_counts());
    CU(cudaMemcpyAsync(counts.data(), ctx->d_counts, counts.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::vector<int32_t> m(R, 0), g(R, 0), q((size_t)R * KBBQ_NQ, 0), cy((size_t)R * KBBQ_NQ * 2 * C, 0), di((size_t)R * KBBQ_NQ * 16, 0);
    const uint64_t* rgc = counts.data();
    const uint64_t* qc = rgc + ctx->n_rgc;
    const uint64_t* cyc = qc + ctx->n_qc;
    const uint64_t* dic = cyc + ctx->n_cyc;
    const int nt = host_threads();
    host::fit_model(R, C, rgc, qc, cyc, dic, m.data(), g.data(), q.data(), cy.data(), di.data(), nt);
    if (meanq) std::memcpy(meanq, m.data(), m.size() * 4);
    if (rgdq) std::memcpy(rgdq, g.data(), g.size() * 4);
    if (qdq) std::memcpy(qdq, q.data(), q.size() * 4);
    if (cydq) std::memcpy(cydq, cy.data(), cy.size() * 4);
    if (didq) std::memcpy(didq, di.data(), di.size() * 4);
    return kbbq_set_model(ctx, m.data(), g.data(), q.data(), cy.data(), di.data());
}

// ---- pass 4 --------------------------------------------------------------------------------------------------------
int kbbq_pass4_recalibrate(kbbq_ctx* ctx, const kbbq_dbatch* cb, uint8_t* qual_out) {
    if (int rc = use(ctx)) return rc;
    if (!ctx->have_model) return fail(KBBQ_ESTATE, "model not set");
    kbbq_dbatch* b = const_cast<kbbq_dbatch*>(cb);
    if (!b) return fail(KBBQ_EINVAL, "null batch");
    if (!b->d_newqual) {
        CU(cudaMallocAsync(&b->d_newqual, std::max<size_t>((size_t)b->n_words * 32, 16), ctx->stream));
        CU(cudaMemsetAsync(b->d_newqual, 0, std::max<size_t>((size_t)b->n_words * 32, 16), ctx->stream));
    }
    if (int rc = need_quals(ctx, b)) return rc;
    pass_begin(ctx);
    const bool fetch = qual_out && b->n_words;
    if (b->n_reads) {
        mark_kernel(ctx, "pass4_apply_kernel");
        ModelView mv{ctx->d_base, ctx->d_cy, ctx->d_di, ctx->max_cycle};
        // lookup tables in shared memory when they leave room for three blocks per SM (two read groups of 150-base reads:
        // 63 KB); larger models keep the global-memory lookups
        size_t p4_smem = p4_smem_bytes(ctx->n_rg, ctx->max_cycle);
        if (p4_smem > 72 * 1024 || getenv("KBBQ_P4_GLOBAL")) p4_smem = 0;
        if (p4_smem && !ctx->p4_attr_set) {
            CU(cudaFuncSetAttribute(pass4_apply_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
            ctx->p4_attr_set = true;
        }
        FlatMap fm;
        if (int rc = flat_map(ctx, b, fm)) return rc;
        const int blocks_per_sm = p4_smem ? (int)std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (p4_smem + 1024))) : 8;
        auto p4_grid = [&](unsigned long long n_half) {
            const long long need = (long long)((n_half + 255) / 256);
            return (int)std::max(1ll, std::min(need, (long long)ctx->sm_count * blocks_per_sm));
        };
        // with a host destination the batch goes in four pieces: piece i travels on the copy stream while piece i+1 is
        // looked up (the D2H copy is several times the kernel, so this hides the kernel, not the copy)
        const int parts = fetch && b->n_reads >= (1u << 16) ? 4 : 1;
        for (int i = 0; i < parts; ++i) {
            const uint64_t w0 = parts == 1 ? 0 : b->part_word[i], w1 = parts == 1 ? b->n_words : b->part_word[i + 1];
            if (w1 > w0) {
                if (p4_smem)
                    pass4_apply_kernel<true><<<p4_grid(2 * (w1 - w0)), 256, p4_smem, ctx->stream>>>(b->view(), fm, mv, ctx->n_rg, b->d_newqual, 2 * w0, 2 * w1);
                else
                    pass4_apply_kernel<false><<<p4_grid(2 * (w1 - w0)), 256, 0, ctx->stream>>>(b->view(), fm, mv, ctx->n_rg, b->d_newqual, 2 * w0, 2 * w1);
                if (int rc = check_launch(ctx, "pass4_apply_kernel")) return rc;
            }
            if (parts > 1) {
                CU(cudaEventRecord(ctx->evp[i], ctx->stream));
                CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->evp[i], 0));
                if (w1 > w0)
                    CU(cudaMemcpyAsync(qual_out + w0 * 32, b->d_newqual + w0 * 32, (size_t)(w1 - w0) * 32, cudaMemcpyDeviceToHost, ctx->copy_stream));
            }
        }
        pass_end(ctx);
        if (parts > 1) {
            CU(cudaStreamSynchronize(ctx->copy_stream));
            return check_quals(ctx);
        }
    } else {
        pass_end(ctx);
    }
    if (fetch) {
        CU(cudaMemcpyAsync(qual_out, b->d_newqual, (size_t)b->n_words * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return check_quals(ctx);
    }
    return KBBQ_OK;
}

// ---- synthetic batches (measurement tooling) -------------------------------------------------------------------------
int kbbq_synth_host(const kbbq_synth_params* p, uint64_t first_read, uint64_t n_reads, char* seq, uint8_t* qual, int32_t* rg,
                    uint8_t* second, int n_threads) {
    SynthSpec s;
    if (int rc = synth_spec(p, s)) return rc;
    auto work = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t r = lo; r < hi; ++r) {
            const SynthRead rd = synth_read(s, first_read + r);
            for (uint32_t i = 0; i < s.read_len; ++i) {
                int code, q;
                synth_base(s, rd, i, code, q);
                seq[r * s.read_len + i] = "ACGT"[code];
                qual[r * s.read_len + i] = (uint8_t)q;
            }
            if (rg) rg[r] = (int32_t)(((first_read + r) >> 1) % s.n_rg);
            if (second) second[r] = (uint8_t)((first_read + r) & 1);
        }
    };
    if (n_threads <= 1 || n_reads < 4096) {
        work(0, n_reads);
    } else {
        std::vector<std::thread> th;
        const uint64_t per = (n_reads + n_threads - 1) / n_threads;
        for (int t = 0; t < n_threads; ++t) {
            const uint64_t lo = std::min<uint64_t>(n_reads, (uint64_t)t * per), hi = std::min<uint64_t>(n_reads, lo + per);
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto& t : th) t.join();
    }
    return KBBQ_OK;
}

int kbbq_synth_batch(kbbq_ctx* ctx, const kbbq_synth_params* p, uint64_t first_read, uint32_t n_reads, kbbq_dbatch** out) {
    if (int rc = use(ctx)) return rc;
    if (!out) return fail(KBBQ_EINVAL, "null output");
    SynthSpec s;
    if (int rc = synth_spec(p, s)) return rc;
    if ((int)s.read_len > ctx->max_cycle) return fail(KBBQ_ERANGE, "read longer than the context's max_cycle");
    if ((int)s.n_rg > ctx->n_rg) return fail(KBBQ_ERANGE, "read group id >= n_rg");
    const uint32_t wpr = (s.read_len + 31) / 32;
    const uint64_t n_words = (uint64_t)n_reads * wpr;
    if (n_words > 0xFFFFFFFFull) return fail(KBBQ_ERANGE, "batch too large: split it");
    kbbq_dbatch* b = new (std::nothrow) kbbq_dbatch();
    if (!b) return fail(KBBQ_ENOMEM, "out of host memory");
    b->n_reads = n_reads; b->max_len = s.read_len; b->n_words = n_words; b->k = ctx->k;
    for (int i = 0; i <= 4; ++i) {
        b->part_read[i] = (uint32_t)((uint64_t)n_reads * i / 4);
        b->part_word[i] = (uint64_t)b->part_read[i] * wpr;
    }
    const size_t nw = std::max<uint64_t>(n_words, 1);
    void** ptrs[] = {(void**)&b->d_bases, (void**)&b->d_nmask, (void**)&b->d_qual, (void**)&b->d_qmask, (void**)&b->d_word_off,
                     (void**)&b->d_len, (void**)&b->d_rg, (void**)&b->d_second, (void**)&b->d_slot_off};
    const size_t bytes[] = {(nw + 1) * 8, (nw + 1) * 4, nw * 32, nw * 8, ((size_t)n_reads + 1) * 4, std::max<size_t>(n_reads, 4) * 4,
                            std::max<size_t>(n_reads, 8) * 2, std::max<size_t>(n_reads, 16), ((size_t)n_reads + 1) * 8};
    for (int i = 0; i < 9; ++i) {
        cudaError_t e = cudaMallocAsync(ptrs[i], bytes[i], ctx->stream);
        if (e == cudaSuccess && i < 2) e = cudaMemsetAsync((char*)*ptrs[i] + (bytes[i] - (i ? 4 : 8)), 0, i ? 4 : 8, ctx->stream);  // the padding word
        if (e != cudaSuccess) {
            kbbq_batch_free(ctx, b);
            return fail(e == cudaErrorMemoryAllocation ? KBBQ_ENOMEM : KBBQ_ECUDA, std::string("synth batch: ") + cudaGetErrorString(e));
        }
    }
    if (n_reads) {
        synth_reads_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, ctx->stream>>>(s, first_read, n_reads, wpr, b->d_bases, b->d_nmask, b->d_qual,
                                                                                      b->d_qmask, b->d_word_off, b->d_len, b->d_rg, b->d_second);
        if (int rc = check_launch(ctx, "synth_reads_kernel")) { kbbq_batch_free(ctx, b); return rc; }
        if (int rc = slot_prefix(ctx, b->d_len, n_reads, b->d_slot_off)) { kbbq_batch_free(ctx, b); return rc; }
    } else {
        CU(cudaMemsetAsync(b->d_word_off, 0, 4, ctx->stream));
        CU(cudaMemsetAsync(b->d_slot_off, 0, 8, ctx->stream));
    }
    b->n_slots = (int)s.read_len >= ctx->k ? (uint64_t)n_reads * (s.read_len - ctx->k + 1) : 0;
    b->uniform_wpr = n_reads ? wpr : 0;
    CU(cudaStreamSynchronize(ctx->stream));
    *out = b;
    return KBBQ_OK;
}

int kbbq_batch_extent(const kbbq_dbatch* b, uint32_t* n_reads, uint64_t* n_words, uint32_t* max_len) {
    if (!b) return fail(KBBQ_EINVAL, "null batch");
    if (n_reads) *n_reads = b->n_reads;
    if (n_words) *n_words = b->n_words;
    if (max_len) *max_len = b->max_len;
    return KBBQ_OK;
}

int kbbq_batch_download(kbbq_ctx* ctx, const kbbq_dbatch* b, uint64_t* bases, uint32_t* nmask, uint8_t* qual, uint32_t* word_off,
                        uint32_t* len, uint16_t* rg, uint8_t* second, uint32_t* qmask) {
    if (int rc = use(ctx)) return rc;
    if (!b) return fail(KBBQ_EINVAL, "null batch");
    if (int rc = need_quals(ctx, const_cast<kbbq_dbatch*>(b))) return rc;
    const size_t nw = b->n_words, n = b->n_reads;
    if (bases && nw) CU(cudaMemcpyAsync(bases, b->d_bases, nw * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (nmask && nw) CU(cudaMemcpyAsync(nmask, b->d_nmask, nw * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (qual && nw) CU(cudaMemcpyAsync(qual, b->d_qual, nw * 32, cudaMemcpyDeviceToHost, ctx->stream));
    if (qmask && nw) CU(cudaMemcpyAsync(qmask, b->d_qmask, nw * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (word_off) CU(cudaMemcpyAsync(word_off, b->d_word_off, (n + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (len && n) CU(cudaMemcpyAsync(len, b->d_len, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (rg && n) CU(cudaMemcpyAsync(rg, b->d_rg, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    if (second && n) CU(cudaMemcpyAsync(second, b->d_second, n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KBBQ_OK;
}

}  // extern "C"
