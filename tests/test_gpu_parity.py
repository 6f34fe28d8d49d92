"""Parity tests proper: the sm_100a kernels, called through the C ABI, against the oracle and the golden fixtures.
Bit-exact everywhere: Bloom bit arrays, insert counters, error flags, covariate tables, model tables, new qualities."""
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
    assert got["stats"].launches == 6  # rng, insert, trusted, errors, covariate counts, apply


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
