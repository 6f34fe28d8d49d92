"""Parity tests proper: the sm_100a kernels, called through the C ABI, against the oracle and the golden fixtures.
Bit-exact everywhere: Bloom bit arrays, insert counters, error flags, covariate tables, model tables, new qualities."""
import ctypes
import os

import numpy as np
import pytest

import common
from kbbq_b200 import api, synth
from oracle import port as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", common.PIPES)
def test_gpu_pipeline_matches_golden(name):
    rs, g = common.load_pipe(name)
    got = common.run_gpu(rs, int(g["k"]), int(g["genomelen"]), int(g["coverage"]), int(g["seed"]))
    common.compare(got, g, what=name + ": ")
    # rng, insert, trusted, scan, repair, covariate counts, apply (+ the word->read map of a batch with ragged reads)
    assert got["stats"].launches == (8 if name == "pipe_k21_ragged" else 7)


@pytest.mark.parametrize("k,cov,L,kw", [
    (32, 20, 100, dict()),                                   # config 1 shape, scaled
    (32, 50, 150, dict(n_rg=3, n_frac=0.002)),               # config 2 shape: several read groups, some N
    (21, 30, 150, dict(err_scale=3.0)),                      # config 5 k sweep
    (31, 30, 150, dict(ragged=True, min_len=62, n_frac=0.004)),
    (32, 20, 64, dict(err_scale=3.0)),                       # len == 2k
    (16, 30, 250, dict(err_scale=6.0)),                      # recursion-heavy
])
def test_gpu_pipeline_matches_port(k, cov, L, kw):
    G = 120000
    rs = synth.make_reads(G, cov, L, seed=200 + k, **kw)
    want = common.run_port(rs, k, G, cov, 4242)
    got = common.run_gpu(rs, k, G, cov, 4242)
    common.compare(got, want)


def test_batch_split_is_invisible():
    """One serial sampling stream: splitting the input into batches must not change a single bit."""
    G, cov, L, k = 60000, 20, 100, 32
    rs = synth.make_reads(G, cov, L, seed=31, ragged=True, min_len=70)
    one = common.run_gpu(rs, k, G, cov, 99, n_batches=1)
    five = common.run_gpu(rs, k, G, cov, 99, n_batches=5)
    common.compare(five, one)


def test_raw_key_insert_query_matches_port():
    rng = np.random.default_rng(3)
    keys = rng.integers(0, 2**64, 200000, dtype=np.uint64)
    probe = np.concatenate([keys[:50000], rng.integers(0, 2**64, 150000, dtype=np.uint64)])
    ctx = api.Context(32, 1, 100)
    for which, fpr in ((api.FILTER_SAMPLED, 0.01), (api.FILTER_TRUSTED, 0.0005)):
        ctx.bloom_create(which, 150000, fpr)
        ctx.bloom_insert_keys(which, keys)
        pb = P.PortBloom(150000, fpr)
        for x in keys:
            pb.insert(int(x))
        assert ctx.bloom_inserted(which) == pb.inserted
        assert np.array_equal(ctx.bloom_words(which), pb.words())
        want = np.array([pb.contains(int(x)) for x in probe], dtype=np.uint8)
        assert np.array_equal(ctx.bloom_query_keys(which, probe), want)
        assert ctx.bloom_fprate(which) == pb.fprate()
    ctx.close()


def test_reference_unit_test_vector_on_gpu():
    """src/kbbq/test.cc:26-59 through the C ABI."""
    kat = common.load_kat()["test_cc"]
    ctx = api.Context(32, 1, 100)
    ctx.bloom_create(0, 100, 0.05)
    ctx.bloom_insert_keys(0, np.array([21345423534512], dtype=np.uint64))
    hits = ctx.bloom_query_keys(0, np.array([21345423534512, 123543451243], dtype=np.uint64))
    assert bool(hits[0]) is True and bool(hits[1]) == kat["contains_gibberish"]
    assert [str(int(x)) for x in ctx.bloom_words(0)] == kat["words"]
    ctx.close()


def test_fixed_mode_consume_matches_port():
    rs = synth.make_reads(50000, 20, 100, seed=5, n_rg=2, n_frac=0.002)
    errors = (rs.seq != rs.clean_seq).astype(np.uint8)
    want = P.PortPipeline(rs, 32).consume(errors)
    ctx = api.Context(32, rs.n_rg, rs.max_len)
    pb = api.PackedBatch(rs)
    db = ctx.upload(pb)
    ctx.consume_errors(db, pb.pack_bits(errors))
    got = ctx.covariates()
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    dq = ctx.fit_model()
    dq_want = P.PortPipeline(rs, 32).get_dqs(want)
    for a, b in zip(dq, dq_want):
        assert np.array_equal(a, b)
    newq = pb.unpack_qual(ctx.pass4(db), rs.total_bases)
    assert np.array_equal(newq, P.PortPipeline(rs, 32).recalibrate(dq_want))
    db.free()
    ctx.close()


def test_filter_or_merge_equals_single_build():
    """Partial filters built from disjoint shards OR-merge into the filter of the whole input (multi-GPU merge step)."""
    G, cov, L, k = 60000, 20, 100, 32
    rs = synth.make_reads(G, cov, L, seed=77)
    whole = common.run_gpu(rs, k, G, cov, 5)
    alpha = np.longdouble(7) / np.longdouble(cov)
    approx = int(np.longdouble(G) * cov * alpha)
    half = rs.n // 2
    ctxs, dbs, base = [], [], 0
    for lo, hi in ((0, half), (half, rs.n)):
        c = api.Context(k, 1, L)
        c.bloom_create(0, approx, 0.01)
        d = c.upload(api.PackedBatch(rs.subset(lo, hi)))
        c.sampler_seed(5, float(alpha))
        c.sampler_seek(base)
        c.pass1(d)
        base += d.slots
        ctxs.append(c)
        dbs.append(d)
    g = ctxs[0].bloom_info(0)
    ctxs[1].sync()
    ctxs[0].bloom_or_from(0, ctxs[1].bloom_device_words(0), 0, g.table_bits // 64)
    assert np.array_equal(ctxs[0].bloom_words(0), whole["sampled_words"])
    assert ctxs[0].bloom_inserted(0) + ctxs[1].bloom_inserted(0) == whole["sampled_inserted"]
    for d in dbs:
        d.free()
    for c in ctxs:
        c.close()


def test_full_size_properties():
    """Config-2-sized input (1.5 M reads x 150): properties that need no oracle run.
    Determinism across runs, insert counters consistent with set bits, every quality in range, q<6 untouched."""
    G, cov, L, k = 4_600_000, 50, 150, 32
    rs = synth.make_reads(G, cov, L, seed=1, n_rg=2)
    a = common.run_gpu(rs, k, G, cov, 2854027472, n_batches=3)
    b = common.run_gpu(rs, k, G, cov, 2854027472, n_batches=1)
    common.compare(a, b)
    assert a["slots"] == rs.n * (L - k + 1)
    assert int(a["rgc"][:, 1].sum()) == rs.total_bases
    assert int(a["qc"][..., 1].sum()) == rs.total_bases and int(a["cyc"][..., 1].sum()) == rs.total_bases
    assert int(a["qc"][..., 0].sum()) == int(a["rgc"][:, 0].sum()) == int(a["pass3_errors"].sum())
    assert a["newqual"].max() <= 93
    low = rs.qual < 6
    assert np.array_equal(a["newqual"][low], rs.qual[low])
    nbits = int(np.unpackbits(a["sampled_words"].view(np.uint8)).sum())
    assert 0 < nbits <= 7 * a["sampled_inserted"]


def _respray(rs, n_rg, qmax, seed):
    """Replace the binned qualities / few read groups of a synthetic set by many distinct ones."""
    rng = np.random.default_rng(seed)
    rs.qual[:] = rng.integers(0, qmax + 1, rs.qual.shape[0], dtype=np.uint8)
    rs.rg[:] = rng.integers(0, n_rg, rs.n).astype(rs.rg.dtype)
    rs.n_rg = n_rg
    return rs


def test_covariate_rows_beyond_the_private_table():
    """covariate_count_kernel keeps ~40 (read group, quality) rows per block in shared memory and privatises 64 read
    groups; 70 read groups x 94 qualities = 6580 rows force the global-atomic fall-through of every table."""
    rs = _respray(synth.make_reads(40000, 20, 100, seed=9, n_frac=0.003), 70, 93, 1)
    errors = (rs.seq != rs.clean_seq).astype(np.uint8)
    want = P.PortPipeline(rs, 32).consume(errors)
    ctx = api.Context(32, rs.n_rg, rs.max_len)
    pb = api.PackedBatch(rs)
    db = ctx.upload(pb)
    ctx.consume_errors(db, pb.pack_bits(errors))
    for a, b, name in zip(ctx.covariates(), want, ("rgc", "qc", "cyc", "dic")):
        assert np.array_equal(a, b), name
    dq = ctx.fit_model()
    dq_want = P.PortPipeline(rs, 32).get_dqs(want)
    for a, b in zip(dq, dq_want):
        assert np.array_equal(a, b)
    assert np.array_equal(pb.unpack_qual(ctx.pass4(db), rs.total_bases), P.PortPipeline(rs, 32).recalibrate(dq_want))
    db.free()
    ctx.close()


def test_pipeline_with_many_distinct_qualities():
    """Un-binned qualities 0..45 in 3 read groups (138 rows): whole pipeline vs the port, tie-breaks on quality included."""
    G, cov, L, k = 60000, 25, 150, 32
    rs = _respray(synth.make_reads(G, cov, L, seed=17, err_scale=3.0, n_frac=0.002), 3, 45, 2)
    common.compare(common.run_gpu(rs, k, G, cov, 31337), common.run_port(rs, k, G, cov, 31337))


def test_longest_reads_of_the_static_scratch():
    """512-base reads (KBBQ_MAX_READ_LEN: 16 words, the largest the statically sized shared-memory scratch holds) go
    through every pass."""
    G, cov, L, k = 50000, 30, 512, 32
    rs = synth.make_reads(G, cov, L, seed=23, err_scale=2.0, ragged=True, min_len=400, n_rg=2)
    common.compare(common.run_gpu(rs, k, G, cov, 7), common.run_port(rs, k, G, cov, 7))


@pytest.mark.parametrize("k,cov,L,kw", [
    (32, 14, 513, dict(err_scale=2.0, ragged=True, min_len=505)),                      # one base more: 17 words
    (32, 12, 700, dict(err_scale=3.0, ragged=True, min_len=300, n_frac=0.001, n_rg=2)),  # mixed with short reads
    (25, 16, 3000, dict(err_scale=0.7, ragged=True, min_len=100)),                     # ~94 words, several chunks of pass 1's list
    (21, 14, 20000, dict(err_scale=0.5, ragged=True, min_len=9000)),                    # the repair kernel down to one warp per block
])
def test_reads_beyond_512_bases_take_the_long_path(k, cov, L, kw):
    """The reference has no read-length limit (readutils.cc:64-104).  A batch with a read of more than 512 bases is
    routed to the kernels' LONG instantiations (masks and error-calling scratch in dynamic shared memory sized for the
    batch, the position list of pass 1 in pieces); every intermediate must still be the port's."""
    G = 60000
    rs = synth.make_reads(G, cov, L, seed=200 + k, **kw)
    assert int(np.diff(rs.off).max()) > 512
    common.compare(common.run_gpu(rs, k, G, cov, 77), common.run_port(rs, k, G, cov, 77))


def test_long_reads_with_every_insert_mode(monkeypatch):
    G, cov, L, k = 40000, 16, 900, 32
    rs = synth.make_reads(G, cov, L, seed=9, err_scale=0.7, ragged=True, min_len=200)
    want = common.run_port(rs, k, G, cov, 5)
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("KBBQ_INSERT_MODE", mode)
        monkeypatch.setenv("KBBQ_STAGE_SHIFT", "4")
        common.compare(common.run_gpu(rs, k, G, cov, 5), want, what=f"insert mode {mode}: ")


def test_the_longest_read_the_abi_takes():
    """65535 bases: one warp per block in the repair kernel, 2049 words of every mask."""
    G, k = 200000, 32
    rs = synth.make_reads(G, 1, 65535, seed=4, n_reads=6, err_scale=0.3)
    common.compare(common.run_gpu(rs, k, G, 2, 3), common.run_port(rs, k, G, 2, 3))


def test_empty_batch_and_reads_shorter_than_k():
    """An empty batch is a no-op for every pass; reads shorter than k consume no draw (htsiter.cc:100-105), are
    skipped by passes 2-3 (the reference underflows there, bloom.cc:30) and still counted and recalibrated."""
    G, cov, L, k = 30000, 20, 100, 32
    rs = synth.make_reads(G, cov, L, seed=3, ragged=True, min_len=10)
    assert (np.diff(rs.off) < k).any()
    whole = common.run_gpu(rs, k, G, cov, 11, n_batches=1)
    split = common.run_gpu(rs, k, G, cov, 11, n_batches=4)
    common.compare(split, whole)
    lens = np.diff(rs.off).astype(np.int64)
    assert whole["slots"] == int(np.maximum(lens - k + 1, 0).sum())
    short = np.repeat(lens < k, lens)
    assert not whole["pass2_errors"][short].any() and not whole["pass3_errors"][short].any()
    assert int(whole["rgc"][:, 1].sum()) == rs.total_bases
    # the port on the reads the reference can process (len >= k) gives the same filters
    keep = np.flatnonzero(lens >= k)
    sub = synth.ReadSet(seq=np.concatenate([rs.seq[rs.off[i]:rs.off[i + 1]] for i in keep]),
                        off=np.concatenate([[0], np.cumsum(lens[keep])]).astype(np.uint64),
                        qual=np.concatenate([rs.qual[rs.off[i]:rs.off[i + 1]] for i in keep]),
                        rg=rs.rg[keep], second=rs.second[keep], n_rg=rs.n_rg)
    # (dropping short reads does not move the others' draws: they own no slot)
    want = common.run_port(sub, k, G, cov, 11)
    assert np.array_equal(whole["sampled_words"], want["sampled_words"]) and np.array_equal(whole["trusted_words"], want["trusted_words"])
    # empty batch
    ctx = api.Context(k, 1, L)
    ctx.bloom_create(0, 1000, 0.01)
    ctx.bloom_create(1, 1000, 0.0005)
    db = ctx.upload(api.PackedBatch(rs.subset(0, 0)))
    ctx.sampler_seed(1, 0.5)
    ctx.pass1(db)
    ctx.set_thresholds(np.zeros(k + 1, dtype=np.int32))
    ctx.pass2(db)
    ctx.pass3(db)
    ctx.fit_model()
    ctx.pass4(db)
    assert ctx.bloom_inserted(0) == 0 and ctx.bloom_inserted(1) == 0 and int(ctx.covariates()[0].sum()) == 0
    db.free()
    ctx.close()


def test_human_scale_filters_address_every_block():
    """BASELINE.json configs[3]: a 3.1 Gb genome sizes the two filters at 26 GB and 43 GB (kbbq.cc:264-266), i.e. tables
    far beyond 4 GiB and 2^29 blocks.  No oracle can hold them on this host, so the check is structural: geometry as the
    reference computes it, every inserted key found again, each key's 64-byte block (hash_ap(key, salt_[0]) % n_blocks,
    bloom.hh:99-106) holding exactly its pattern's bits when read back from device memory, and nothing else set."""
    import torch
    from kbbq_b200.dist import device_tensor
    from oracle import port

    free, _ = torch.cuda.mem_get_info(0)
    if free < 80 << 30:
        pytest.skip("needs ~70 GB of free HBM")
    genomelen, cov = 3_100_000_000, 30
    alpha = np.longdouble(7) / np.longdouble(cov)
    approx = int(np.longdouble(genomelen) * cov * alpha)
    rng = np.random.default_rng(11)
    keys = np.unique(rng.integers(0, 2**64, 4096, dtype=np.uint64))
    others = rng.integers(0, 2**64, 100000, dtype=np.uint64)
    ctx = api.Context(32, 1, 150)
    for which, fpr, n_salts in ((api.FILTER_SAMPLED, 0.01, 7), (api.FILTER_TRUSTED, 0.0005, 11)):
        g = ctx.bloom_create(which, approx, fpr)
        assert g.n_salts == n_salts and g.table_bits == g.n_blocks * 512 and g.table_bits // 8 > (20 << 30)
        want = api.bloom_geometry(approx, fpr)
        assert (g.table_bits, g.n_blocks, list(g.salts[:n_salts])) == (want.table_bits, want.n_blocks, list(want.salts[:n_salts]))
        ctx.bloom_insert_keys(which, keys)
        assert ctx.bloom_inserted(which) == len(keys)
        assert ctx.bloom_query_keys(which, keys).all()
        assert ctx.bloom_query_keys(which, others).mean() < 1e-3     # a near-empty table answers "no"
        pat = api.bloom_patterns(g).reshape(65536, 8)
        words = device_tensor(ctx.bloom_device_words(which), g.table_bits // 8, 0).view(torch.int64)
        salt0, salt1 = int(g.salts[0]), int(g.salts[1])
        blocks = {}
        for key in keys.tolist():
            blk = port.lib().ko_hash_ap8(key, salt0) % g.n_blocks
            p = port.lib().ko_hash_ap8(key, salt1) & 0xFFFF
            blocks[blk] = blocks.get(blk, np.zeros(8, dtype=np.uint64)) | pat[p]
        assert max(blocks) * 64 > (16 << 30)                          # byte offsets beyond 16 GiB are exercised ...
        for blk, bits in blocks.items():
            got = words[blk * 8: blk * 8 + 8].cpu().numpy().view(np.uint64)
            assert np.array_equal(got, bits), blk
        # ... and the table holds nothing but those bits
        total = int(sum(int(np.unpackbits(b.view(np.uint8)).sum()) for b in blocks.values()))
        nz = 0
        step = 1 << 27                                                 # 1 GiB of int64 words at a time
        for lo in range(0, words.numel(), step):
            w = words[lo: lo + step]
            nzw = w[w != 0]
            if nzw.numel():
                nz += int(np.unpackbits(nzw.cpu().numpy().view(np.uint8)).sum())
        assert nz == total
    ctx.close()


@pytest.mark.parametrize("mode,shift,cap", [("1", None, None), ("2", None, None), ("2", "6", None), ("2", "6", "256"), ("2", "4", "128")])
def test_insert_modes_for_large_filters_give_the_same_bits(mode, shift, cap, monkeypatch):
    """Filters larger than the L2 are filled through other code paths (kbbq_kernels.cuh INS_TEST_FIRST: read before OR;
    INS_STAGED: (block, pattern) items staged into one region per slice of the table, then applied region after region):
    forced here on a small input -- with the slices shrunk so that there are hundreds of them (KBBQ_STAGE_SHIFT), and with
    regions so small that most items take the overflow path (KBBQ_STAGE_CAP) -- every intermediate and the filters'
    words must still equal the oracle's."""
    monkeypatch.setenv("KBBQ_INSERT_MODE", mode)
    if shift:
        monkeypatch.setenv("KBBQ_STAGE_SHIFT", shift)
    if cap:
        monkeypatch.setenv("KBBQ_STAGE_CAP", cap)
    G, cov, L, k = 80000, 30, 150, 32
    rs = synth.make_reads(G, cov, L, seed=61, n_rg=2, n_frac=0.002, ragged=True, min_len=40)
    common.compare(common.run_gpu(rs, k, G, cov, 2024, n_batches=2), common.run_port(rs, k, G, cov, 2024))


def test_quality_masks_derived_on_the_device(monkeypatch):
    """kbbq_batch::qmask is optional: without it the library derives the qual <= 2 / qual <= 6 masks from the quality
    array on the device (qual_masks_kernel) and pass 2 waits for the qualities."""
    monkeypatch.setenv("KBBQ_NO_HOST_QMASK", "1")
    G, cov, L, k = 60000, 25, 100, 32
    rs = synth.make_reads(G, cov, L, seed=71, n_rg=2, ragged=True, min_len=50)
    assert api.PackedBatch(rs).qmask is None
    common.compare(common.run_gpu(rs, k, G, cov, 5150, n_batches=2), common.run_port(rs, k, G, cov, 5150))


def test_upload_rejects_out_of_range_metadata():
    """KBBQ_ERANGE for a read group id the context has no tables for and for a read longer than the batch claims; the
    failed upload leaves the context usable."""
    rs = synth.make_reads(20000, 10, 100, seed=4, n_rg=2)
    ctx = api.Context(32, 1, 100)             # tables for ONE read group
    with pytest.raises(api.KbbqError) as e:
        ctx.upload(api.PackedBatch(rs))
    assert e.value.code == -5 and "read group" in str(e.value)
    pb = api.PackedBatch(rs)
    pb.rg[:] = 0
    pb.max_len = 50                            # the reads are 100 long
    with pytest.raises(api.KbbqError) as e:
        ctx.upload(pb)
    assert e.value.code == -5
    pb.max_len = 100
    db = ctx.upload(pb)
    assert db.slots == rs.n * (100 - 32 + 1)
    db.free()
    ctx.close()


@pytest.mark.parametrize("case", range(12))
def test_randomised_shapes_match_port(case):
    """A deterministic sweep over k, read length, coverage, error rate, N rate, ragged lengths and read-group count
    (tests/common.py:random_case; tests/test_oracle.py pins the port to the unmodified reference on the same cases):
    every intermediate of the GPU run must equal the oracle port's, whatever path through get_errors the reads take."""
    rs, k, G, cov, seed, n_batches, what = common.random_case(case)
    common.compare(common.run_gpu(rs, k, G, cov, seed, n_batches=n_batches), common.run_port(rs, k, G, cov, seed), what=what)


def test_filter_save_and_load_round_trip(tmp_path):
    """kbbq_bloom_save / kbbq_bloom_load: a filter built by pass 1 survives the file and serves pass 2 in a new context
    exactly as the original does (same words, counter, false-positive rate, thresholds, trusted filter)."""
    G, cov, L, k, seed = 50000, 20, 100, 32, 77
    rs = synth.make_reads(G, cov, L, seed=13, n_frac=0.001)
    want = common.run_port(rs, k, G, cov, seed)
    alpha = np.longdouble(7) / np.longdouble(cov)
    approx = int(np.longdouble(G) * cov * alpha)
    a = api.Context(k, 1, L)
    a.bloom_create(api.FILTER_SAMPLED, approx, 0.01)
    da = a.upload(api.PackedBatch(rs))
    a.sampler_seed(seed, float(alpha))
    a.pass1(da)
    path = str(tmp_path / "sampled.kbf")
    a.bloom_save(api.FILTER_SAMPLED, path)
    assert os.path.getsize(path) == 8 + ctypes.sizeof(api.BloomGeom) + 8 + a.bloom_info(0).table_bits // 8
    da.free()
    a.close()
    b = api.Context(k, 1, L)
    b.bloom_load(api.FILTER_SAMPLED, path)
    assert np.array_equal(b.bloom_words(0), want["sampled_words"]) and b.bloom_inserted(0) == want["sampled_inserted"]
    assert b.bloom_fprate(0) == want["sampled_fpr"]
    b.bloom_create(api.FILTER_TRUSTED, approx, 0.0005)
    db = b.upload(api.PackedBatch(rs))
    assert np.array_equal(b.compute_thresholds(alpha), want["thresholds"])
    b.pass2(db)
    assert np.array_equal(b.bloom_words(1), want["trusted_words"]) and b.bloom_inserted(1) == want["trusted_inserted"]
    with open(path, "r+b") as fh:
        fh.write(b"NOTAKBBQ")
    with pytest.raises(api.KbbqError):
        b.bloom_load(api.FILTER_SAMPLED, path)
    db.free()
    b.close()


@pytest.mark.parametrize("ragged", [False, True])
def test_fixed_mode_on_the_device_matches_port(ragged):
    """kbbq_fixed_errors: the --fixed comparison (kbbq.cc:367-378) as a kernel over two resident batches -- flags and
    tables against the oracle fed with the character comparison; the fixed file may end early (n_first)."""
    kw = dict(ragged=True, min_len=40) if ragged else {}
    rs = synth.make_reads(50000, 20, 100, seed=6, n_rg=2, n_frac=0.003, **kw)
    fx = synth.ReadSet(seq=rs.clean_seq, off=rs.off, qual=rs.qual, rg=rs.rg, second=rs.second, n_rg=rs.n_rg)
    n_first = rs.n - 37
    errors = (rs.seq != rs.clean_seq).astype(np.uint8)
    first = rs.subset(0, n_first)
    want = P.PortPipeline(first, 32).consume(errors[: first.total_bases], maxlen=rs.max_len)
    ctx = api.Context(32, rs.n_rg, rs.max_len)
    pb, pf = api.PackedBatch(rs), api.PackedBatch(fx)
    db, df = ctx.upload(pb), ctx.upload(pf)
    bits = ctx.fixed_errors(db, df, n_first, want_errors=True)
    flags = pb.unpack_bits(bits, rs.total_bases)
    assert np.array_equal(flags[: first.total_bases], errors[: first.total_bases])
    assert not flags[first.total_bases :].any()  # reads without a partner carry no flags
    for a, b in zip(ctx.covariates(), want):
        assert np.array_equal(a, b)
    assert ctx.total_bases(db) == rs.total_bases and ctx.total_bases(df) == rs.total_bases
    db.free()
    df.free()
    ctx.close()


def test_known_present_kmers_change_nothing(monkeypatch):
    """Pass 2 skips the probes of k-mers pass 1 sampled from the same batch, pass 3's scan those pass 2 inserted (a Bloom
    filter has no false negatives); with the shortcut switched off the run must give the same bits -- and clearing a
    filter between the passes must switch it off by itself (kbbq_ctx::epoch)."""
    G, cov, L, k = 80000, 30, 150, 32
    rs = synth.make_reads(G, cov, L, seed=62, n_rg=2, n_frac=0.002)
    with_bits = common.run_gpu(rs, k, G, cov, 77, n_batches=2)
    monkeypatch.setenv("KBBQ_NO_KNOWN_BITS", "1")
    without = common.run_gpu(rs, k, G, cov, 77, n_batches=2)
    common.compare(with_bits, without)
    monkeypatch.delenv("KBBQ_NO_KNOWN_BITS")
    # pass 1 on a batch, then the sampled filter is cleared and rebuilt from OTHER reads: the batch's sample bits must
    # not be trusted any more
    from kbbq_b200.pipeline import Recalibrator
    other = synth.make_reads(G, cov, L, seed=63, n_rg=2)
    rc = Recalibrator(k, 2, L, G, cov, 77)
    db, do = rc.ctx.upload(api.PackedBatch(rs)), rc.ctx.upload(api.PackedBatch(other))
    rc.ctx.sampler_seed(77, float(rc.alpha_ld))
    rc.ctx.pass1(db)
    rc.ctx.bloom_clear(api.FILTER_SAMPLED)
    rc.ctx.sampler_seek(0)
    rc.ctx.pass1(do)
    th = rc.ctx.compute_thresholds(rc.alpha_ld)
    bits = rc.ctx.pass2(db, want_errors=True)
    want = P.PortPipeline(rs, k)
    sampled = P.PortBloom(rc.approx_kmers, 0.01)
    P.PortPipeline(other, k).subsample(sampled, float(rc.alpha_ld), 77)
    pb = api.PackedBatch(rs)
    assert np.array_equal(pb.unpack_bits(bits, rs.total_bases), want.infer_errors(sampled, th))
    rc.close()


def test_streamed_upload_gives_the_resident_answer():
    """kbbq_batch_upload_stream: every array on the side stream, nothing waited for; the passes wait on the device.  Two
    pinned batches fed that way (pipeline.run(host_batches=...)) must give what the ordinary upload gives."""
    G, cov, L, k = 80000, 30, 150, 32
    rs = synth.make_reads(G, cov, L, seed=64, n_rg=2, n_frac=0.002)
    want = common.run_gpu(rs, k, G, cov, 99, n_batches=2)
    from kbbq_b200.pipeline import Recalibrator
    rc = Recalibrator(k, rs.n_rg, L, G, cov, 99)
    half = rs.n // 2
    parts = [rs.subset(0, half), rs.subset(half, rs.n)]
    pbs = [api.PackedBatch(p, pinned=True) for p in parts]
    rc.run(fetch_quals=True, want_errors=True, host_batches=pbs, timing=False)
    assert np.array_equal(rc.ctx.bloom_words(0), want["sampled_words"]) and np.array_equal(rc.ctx.bloom_words(1), want["trusted_words"])
    newq = np.concatenate([pb.unpack_qual(q, p.total_bases) for pb, q, p in zip(pbs, rc.last_quals, parts)])
    assert np.array_equal(newq, want["newqual"])
    e3 = np.concatenate([pb.unpack_bits(e, p.total_bases) for pb, e, p in zip(pbs, rc.last_errors3, parts)])
    assert np.array_equal(e3, want["pass3_errors"])
    for a, b in zip(rc.ctx.covariates(), (want["rgc"], want["qc"], want["cyc"], want["dic"])):
        assert np.array_equal(a, b)
    rc.close()


def test_quality_above_93_in_a_hand_filled_batch_is_reported_not_indexed():
    """kbbq_batch is a public struct: a caller may fill it without kbbq_pack_reads.  A quality above 93 would index the
    count and model tables out of bounds: the upload flags it on the device, the count and apply kernels skip
    the base, and the next synchronising call returns KBBQ_ERANGE."""
    rs = synth.make_reads(20000, 10, 100, seed=7, n_rg=1)
    pb = api.PackedBatch(rs)
    pb.qual[5] = 200
    ctx = api.Context(32, 1, 100)
    db = ctx.upload(pb)
    flags = np.zeros(rs.total_bases, dtype=np.uint8)
    with pytest.raises(api.KbbqError) as e:
        ctx.consume_errors(db, pb.pack_bits(flags))
        ctx.sync()
    assert e.value.code == -5 and "quality above 93" in str(e.value)
    rgc, qc, cyc, dic = ctx.covariates()
    assert int(rgc[0, 1]) == rs.total_bases - 1  # every base but the bad one was counted, nothing out of bounds
    db.free()
    ctx.close()
