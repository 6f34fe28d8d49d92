"""Parity at the BASELINE configs' REAL sizes: the GPU path against the unmodified reference
(oracle/_ref) on the whole input, every intermediate, bit for bit -- config 1 (200 k reads) and config 2 (1.53 M reads)
against a live reference run on this box's host, config 3 (12.8 M reads) against the digests a committed script made
from one reference run (tests/golden/make_fullsize_golden.py -> tests/golden/c3_digests.json)."""
import json
import os

import numpy as np
import pytest

import common
from kbbq_b200 import api, synth
from kbbq_b200.pipeline import Recalibrator
from oracle import refharness as R

pytestmark = pytest.mark.gpu


def _want(rs, k, G, cov, seed):
    """the reference when it was built here and travelled to this box, else its restatement (which tests/test_oracle.py
    pins to the reference)"""
    return common.run_ref(rs, k, G, cov, seed) if R.available() else common.run_port(rs, k, G, cov, seed)


def test_config1_full_size_matches_reference():
    """BASELINE.json configs[0]: 1 Mb genome, 20X, 100 bp reads, k=32 -- the whole 200 000-read input."""
    G, cov, L, k = 1_000_000, 20, 100, 32
    rs = synth.make_reads(G, cov, L, seed=1)
    assert rs.n == 200_000
    want = _want(rs, k, G, cov, 2854027472)
    got = common.run_gpu(rs, k, G, cov, 2854027472)
    common.compare(got, want, what="config 1 full size: ")


def test_config2_full_size_matches_reference():
    """BASELINE.json configs[1]: 4.6 Mb genome, 50X, 150 bp reads, 2 read groups -- the whole 1 533 333-read input,
    uploaded as three batches (the >65 536-read pass-4 hand-off and multi-batch paths at their real sizes)."""
    G, cov, L, k = 4_600_000, 50, 150, 32
    rs = synth.make_reads(G, cov, L, seed=1, n_rg=2, keep_clean=False)
    assert rs.n == 1_533_333
    want = _want(rs, k, G, cov, 2854027472)
    got = common.run_gpu(rs, k, G, cov, 2854027472, n_batches=3)
    common.compare(got, want, what="config 2 full size: ")


def _hex(d):
    return "%016x%016x" % d


def test_config3_full_size_matches_reference_digests():
    """BASELINE.json configs[2] -- bench.py's workload: 64 Mb genome, 30X, 12.8 M reads made on the device.  Filters
    (537 + 886 MB, HBM resident: the staged-insert and large-table probe paths at their real sizes), insert counters,
    thresholds, count tables and every recalibrated quality against one run of the unmodified reference."""
    if not os.path.exists(os.path.join(common.GOLDEN, "c3_digests.json")):
        pytest.skip("tests/golden/c3_digests.json has not been generated (tests/golden/make_fullsize_golden.py)")
    with open(os.path.join(common.GOLDEN, "c3_digests.json")) as fh:
        g = json.load(fh)
    G, cov, L, k = g["genome"], g["coverage"], g["read_len"], g["k"]
    rc = Recalibrator(k, g["n_rg"], L, G, cov, g["seed"])
    sp = api.SynthParams(seed=g["synth_seed"], genome_seed=g["genome_seed"], genome_len=G, read_len=L, n_rg=g["n_rg"], err_scale=g["err_scale"])
    db = rc.ctx.synth_batch(sp, 0, g["n_reads"])
    rc.add_resident(db)
    st = rc.run(fetch_quals=False)
    ctx = rc.ctx
    assert rc.local_slots == g["slots"]
    assert ctx.bloom_inserted(api.FILTER_SAMPLED) == g["sampled_inserted"]
    assert _hex(ctx.bloom_digest(api.FILTER_SAMPLED)) == g["sampled_digest"], "sampled filter bits differ"
    assert [int(x) for x in st.thresholds] == g["thresholds"]
    assert ctx.bloom_inserted(api.FILTER_TRUSTED) == g["trusted_inserted"]
    assert _hex(ctx.bloom_digest(api.FILTER_TRUSTED)) == g["trusted_digest"], "trusted filter bits differ"
    rgc = ctx.covariates()[0]
    assert int(rgc[:, 0].sum()) == g["errors_total"]
    assert _hex(ctx.covariates_digest()) == g["counts_digest"], "covariate count tables differ"
    assert _hex(ctx.newqual_digest(db)) == g["newqual_digest"], "recalibrated qualities differ"
    rc.close()
