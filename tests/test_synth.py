"""The synthetic-read generator (kbbq_synth.cuh: measurement tooling) and the device digests: the device batch is the
host generator's reads, the digest kernel is its numpy restatement, and a pipeline over a device-made batch matches the
oracle on the host-made twin."""
import numpy as np
import pytest

import common
from kbbq_b200 import api
from kbbq_b200.pipeline import Recalibrator


def _params(G, L=150, n_rg=2, seed=5):
    return api.SynthParams(seed=seed, genome_seed=77, genome_len=G, read_len=L, n_rg=n_rg, err_scale=1.5)


def test_synth_host_is_a_pure_function_of_the_read_index():
    sp = _params(50_000)
    whole = api.synth_host(sp, 0, 3000, n_threads=4)
    part = api.synth_host(sp, 1000, 500, n_threads=1)
    assert np.array_equal(whole.seq[1000 * 150 : 1500 * 150], part.seq)
    assert np.array_equal(whole.qual[1000 * 150 : 1500 * 150], part.qual)
    assert np.array_equal(whole.rg[1000:1500], part.rg) and np.array_equal(whole.second[1000:1500], part.second)
    assert set(np.unique(whole.qual)) <= {2, 11, 25, 37, 40} and set(np.unique(whole.seq)) <= set(b"ACGT")
    assert whole.rg.max() == 1 and list(whole.second[:4]) == [0, 1, 0, 1]


def test_digest_words_ignores_zero_words_and_sees_positions():
    a = np.zeros(1000, dtype=np.uint64)
    assert api.digest_words(a) == (0, 0)
    a[17] = 5
    b = np.zeros(1000, dtype=np.uint64)
    b[18] = 5
    assert api.digest_words(a) != api.digest_words(b)
    assert api.digest_words(a) == api.digest_words(np.concatenate([a, np.zeros(77, dtype=np.uint64)]))


@pytest.mark.gpu
@pytest.mark.parametrize("L", [150, 100, 64])
def test_device_batch_is_the_host_generators_reads(L):
    sp = _params(40_000, L=L)
    ctx = api.Context(32, 2, L)
    db = ctx.synth_batch(sp, 123, 4000)
    got = ctx.download(db)
    want = api.PackedBatch(api.synth_host(sp, 123, 4000))
    for name in ("bases", "nmask", "qual", "word_off", "len", "rg", "second", "qmask"):
        assert np.array_equal(getattr(got, name), getattr(want, name)), name
    assert db.slots == 4000 * (L - 32 + 1)
    db.free()
    ctx.close()


@pytest.mark.gpu
def test_pipeline_on_a_device_made_batch_matches_the_oracle():
    G, cov, L, k, seed = 120_000, 30, 150, 32, 4242
    sp = _params(G)
    n = G * cov // L
    rs = api.synth_host(sp, 0, n)
    want = common.run_port(rs, k, G, cov, seed)
    rc = Recalibrator(k, 2, L, G, cov, seed)
    db = rc.ctx.synth_batch(sp, 0, n)
    rc.add_resident(db)
    rc.run(fetch_quals=True, want_errors=True)
    ctx = rc.ctx
    assert ctx.bloom_inserted(0) == want["sampled_inserted"] and ctx.bloom_inserted(1) == want["trusted_inserted"]
    assert np.array_equal(ctx.bloom_words(0), want["sampled_words"]) and np.array_equal(ctx.bloom_words(1), want["trusted_words"])
    assert ctx.bloom_digest(1) == api.digest_words(want["trusted_words"])
    rgc, qc, cyc, dic = ctx.covariates()
    assert np.array_equal(qc, want["qc"]) and np.array_equal(dic, want["dic"]) and np.array_equal(rgc, want["rgc"])
    assert ctx.covariates_digest() == api.digest_words(np.concatenate([a.reshape(-1) for a in (rgc, qc, cyc, dic)]))
    pb = api.PackedBatch(rs)
    newq = pb.unpack_qual(rc.last_quals[0], rs.total_bases)
    assert np.array_equal(newq, want["newqual"])
    assert ctx.newqual_digest(db) == api.digest_words(rc.last_quals[0])
    assert np.array_equal(pb.unpack_bits(rc.last_errors3[0], rs.total_bases), want["pass3_errors"])
    rc.close()
