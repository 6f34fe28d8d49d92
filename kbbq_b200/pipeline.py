"""The pass sequence of kbbq's main() (src/kbbq/kbbq.cc:229-457) driven through the C ABI.

`Recalibrator` is what the reference's main() is to its pass drivers: it sizes the two filters, seeds the sampler,
runs pass 1 (subsample), derives the thresholds, runs pass 2 (trusted k-mers), pass 3 (errors + covariates), fits
the model on the host and runs pass 4 (apply).  Read batches stay resident in HBM across the passes (the reference
re-opens and re-parses its input five times, kbbq.cc:232,258,336,365,456).

Multi-GPU (one process per GPU): reads are sharded over ranks; every rank builds partial filters and partial count
tables over its shard, and the partial results are merged after each pass -- bitwise OR of the filter bit arrays,
sum of the insert counters and of the count tables -- so every rank holds the same merged state the single-GPU run
would hold.  The sampling stream is ONE serial stream over the whole input (htsiter.cc:113-129): each rank seeks to
the ordinal of its first k-mer slot.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from . import api


@dataclass
class RunStats:
    pass_ms: dict = field(default_factory=dict)   # device time per pass (CUDA events), summed over batches
    kernel_ms: dict = field(default_factory=dict) # device time per kernel (events between launches), summed over batches
    host_ms: dict = field(default_factory=dict)   # host maths between passes
    merge_ms: dict = field(default_factory=dict)  # collectives
    sampled_inserted: int = 0
    trusted_inserted: int = 0
    thresholds: np.ndarray | None = None
    launches: int = 0


class Recalibrator:
    SAMPLER_FPR = 0.01     # kbbq.cc:155
    TRUSTED_FPR = 0.0005   # kbbq.cc:156

    def __init__(self, k: int, n_rg: int, max_cycle: int, genomelen: int, coverage: int, seed: int,
                 alpha=None, device: int = 0, comm=None):
        self.k, self.n_rg, self.max_cycle = k, n_rg, max_cycle
        self.genomelen, self.coverage, self.seed = genomelen, coverage, seed
        # alpha = 7.0l / coverage (kbbq.cc:251); approx_kmers = genomelen*coverage*alpha in long double (kbbq.cc:264)
        self.alpha_ld = np.longdouble(7.0) / np.longdouble(coverage) if alpha is None else np.longdouble(alpha)
        self.approx_kmers = int(np.longdouble(genomelen) * np.longdouble(coverage) * self.alpha_ld)
        self.ctx = api.Context(k, n_rg, max_cycle, device)
        self.comm = comm
        self.batches: list[api.DeviceBatch] = []
        self.slot_base = 0
        self._create_filters()

    def _create_filters(self):
        """The two filters.  When the merges are to go through the NVSwitch (Comm.wants_nvls) they are created as shareable
        virtual-memory-API memory and their multicast objects are set up right away (collectively); if any rank cannot
        have that, every rank re-creates them as ordinary device memory, which the peer-memory merge can export."""
        import os

        want = self.comm is not None and self.comm.wants_nvls()
        had = os.environ.get("KBBQ_FILTER_VMM")
        if want:
            os.environ["KBBQ_FILTER_VMM"] = "1"
        try:
            self.ctx.bloom_create(api.FILTER_SAMPLED, self.approx_kmers, self.SAMPLER_FPR)
            self.ctx.bloom_create(api.FILTER_TRUSTED, self.approx_kmers, self.TRUSTED_FPR)
        finally:
            if want and had is None:
                del os.environ["KBBQ_FILTER_VMM"]
        if want:
            ok = self.comm.mc_ready(self.ctx, api.FILTER_SAMPLED)
            ok = self.comm.mc_ready(self.ctx, api.FILTER_TRUSTED) and ok
            if not ok:
                self.ctx.bloom_create(api.FILTER_SAMPLED, self.approx_kmers, self.SAMPLER_FPR)
                self.ctx.bloom_create(api.FILTER_TRUSTED, self.approx_kmers, self.TRUSTED_FPR)

    # ---- data ------------------------------------------------------------------------------------------------------
    def add_batch(self, pb: api.PackedBatch) -> api.DeviceBatch:
        db = self.ctx.upload(pb)
        self.batches.append(db)
        return db

    def add_resident(self, db: api.DeviceBatch) -> api.DeviceBatch:
        """adopt a batch that is already in device memory (Context.synth_batch)"""
        self.batches.append(db)
        return db

    def clear_batches(self):
        for b in self.batches:
            b.free()
        self.batches = []

    @property
    def local_slots(self) -> int:
        return sum(b.slots for b in self.batches)

    def reset(self):
        """Forget everything the passes accumulated (filters, counts); keep the resident batches."""
        self.ctx.bloom_clear(api.FILTER_SAMPLED)
        self.ctx.bloom_clear(api.FILTER_TRUSTED)
        self.ctx.covariates_reset()

    # ---- the passes ----------------------------------------------------------------------------------------------
    def run(self, fetch_quals: bool = False, want_errors: bool = False, qual_bufs=None, host_batches=None,
            timing: bool = True) -> RunStats:
        """`host_batches`: pinned PackedBatches to stream in instead of batches that are already resident -- batch i+1
        travels (kbbq_batch_upload_stream) while pass 1 runs on batch i; they are resident afterwards.  `timing=False`
        skips the per-pass event queries, which would make the host wait for every pass before queuing the next."""
        st = RunStats()
        ctx, comm = self.ctx, self.comm
        launches0 = ctx.launches

        def timed_pass(name, fn):
            total = 0.0
            outs = []
            for b in self.batches:
                outs.append(fn(b))
                if timing:
                    total += ctx.last_pass_ms()
                    for kn, kms in ctx.last_pass_kernels().items():
                        st.kernel_ms[kn] = st.kernel_ms.get(kn, 0.0) + kms
            st.pass_ms[name] = total
            return outs

        # pass 1: every rank seeks to the ordinal of its first slot in the one global sampling stream
        if comm is not None:
            local = self.local_slots if host_batches is None else sum(
                int(np.maximum(pb.len[: pb.n_reads].astype(np.int64) - self.k + 1, 0).sum()) for pb in host_batches)
            self.slot_base = comm.exclusive_prefix(local)
        ctx.sampler_seed(self.seed, float(self.alpha_ld))
        ctx.sampler_seek(self.slot_base)
        if host_batches is not None:
            assert not self.batches, "host_batches replace the resident batches"
            for pb in host_batches:  # the sampling stream advances batch by batch inside the library
                db = ctx.upload_stream(pb)
                self.batches.append(db)
                ctx.pass1(db)
            st.pass_ms["pass1"] = 0.0
        else:
            timed_pass("pass1", ctx.pass1)
        if comm is not None:
            t0 = time.perf_counter()
            comm.merge_filter(ctx, api.FILTER_SAMPLED)
            st.merge_ms["sampled"] = (time.perf_counter() - t0) * 1e3
            self._note_merge_kernel(st)
        # fprate -> phit -> thresholds (kbbq.cc:304-313)
        t0 = time.perf_counter()
        st.thresholds = ctx.compute_thresholds(self.alpha_ld)
        st.host_ms["thresholds"] = (time.perf_counter() - t0) * 1e3
        # pass 2
        e2 = timed_pass("pass2", lambda b: ctx.pass2(b, want_errors))
        if comm is not None:
            t0 = time.perf_counter()
            comm.merge_filter(ctx, api.FILTER_TRUSTED)
            st.merge_ms["trusted"] = (time.perf_counter() - t0) * 1e3
            self._note_merge_kernel(st)
        # pass 3
        e3 = timed_pass("pass3", lambda b: ctx.pass3(b, want_errors))
        if comm is not None:
            t0 = time.perf_counter()
            comm.merge_counts(ctx)
            st.merge_ms["counts"] = (time.perf_counter() - t0) * 1e3
        # model fit on the host in long double (covariateutils.cc:204-230)
        t0 = time.perf_counter()
        ctx.fit_model_quiet()
        st.host_ms["fit"] = (time.perf_counter() - t0) * 1e3
        # pass 4
        bufs = iter(qual_bufs) if qual_bufs is not None else None
        quals = timed_pass("pass4", lambda b: ctx.pass4(b, out=next(bufs) if bufs is not None else None, fetch=fetch_quals))
        ctx.sync()
        st.launches = ctx.launches - launches0
        self.last_errors2, self.last_errors3, self.last_quals = e2, e3, quals
        return st

    def _note_merge_kernel(self, st: RunStats) -> None:
        """device time of the peer-memory merge kernel, when that is what merged the filter (it is timed like a pass)"""
        for kn, kms in self.ctx.last_pass_kernels().items():
            if kn in ("or_allreduce_p2p_kernel", "or_allreduce_mc_kernel"):
                st.kernel_ms[kn] = st.kernel_ms.get(kn, 0.0) + kms

    def close(self):
        self.clear_batches()
        self.ctx.close()
