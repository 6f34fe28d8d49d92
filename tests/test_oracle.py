"""The oracle port (oracle/kbbq_oracle.c) against the golden vectors made from the unmodified reference, and --
where oracle/_ref exists (authoring container, or a box the prebuilt .so travelled to) -- against the reference itself."""
import numpy as np
import pytest

import common
from kbbq_b200 import synth
from oracle import port as P
from oracle import refharness as R

KAT = common.load_kat()


def test_kmer_known_answers():
    for e in KAT["kmer"]:
        got = P.kmer_of(e["k"], e["s"].encode())
        assert (got["fwd"], got["rc"], got["get"], got["size"], got["valid"]) == (
            int(e["fwd"]), int(e["rc"]), int(e["get"]), int(e["size"]), e["valid"]), e


def test_bloom_geometry_salts_patterns_hashes():
    for e in KAT["bloom"]:
        if e["table_bits"] > 2_000_000_000:
            continue  # geometry of the human-scale filter is checked through the product's host helper
        b = P.PortBloom(e["projected"], e["fpr"])
        assert (b.table_bits, b.n_blocks, b.n_salts, b.random_seed, b.optimal_k, b.optimal_table_size) == (
            e["table_bits"], e["n_blocks"], e["n_salts"], int(e["random_seed"]), e["optimal_k"], e["optimal_table_size"])
        assert [int(x) for x in b.salts] == e["salts"]
        pat = b.patterns()
        assert [str(int(x)) for x in pat[:8]] == e["pattern0"]
        assert [str(int(x)) for x in pat[36456 * 8:36456 * 8 + 8]] == e["pattern36456"]
        assert str(int(np.bitwise_xor.reduce(pat))) == e["pattern_xor"]
        assert int(np.unpackbits(pat.view(np.uint8)).sum()) == e["pattern_popcount"]
        for key, (h0, h1) in e["hashes"].items():
            assert P.lib().ko_hash_ap8(int(key), int(b.salts[0])) == h0
            assert P.lib().ko_hash_ap8(int(key), int(b.salts[1])) == h1


def test_reference_unit_test_vector():
    """src/kbbq/test.cc:26-59, the reference's only unit-level check."""
    b = P.PortBloom(100, 0.05)
    b.insert(21345423534512)
    assert b.contains(21345423534512) == KAT["test_cc"]["contains_kmer"] is True
    assert b.contains(123543451243) == KAT["test_cc"]["contains_gibberish"]
    assert [str(int(x)) for x in b.words()] == KAT["test_cc"]["words"]


def test_rng_stream_and_bernoulli():
    st, out = P.rng_outputs(KAT["rng"]["seed"], 0, 8)
    assert [str(int(x)) for x in st] == KAT["rng"]["state"]
    assert [str(int(x)) for x in out] == KAT["rng"]["first"]
    assert str(int(P.rng_outputs(12345, 1000000, 1)[1][0])) == KAT["rng"]["at_1e6"]
    assert "".join(map(str, P.bernoulli(12345, 0.35, 64))) == KAT["rng"]["bernoulli_0.35"]
    assert [str(int(x)) for x in P.rng_outputs(2854027472, 0, 0)[0]] == KAT["rng"]["seed_2854027472_state"]


def test_thresholds_known_answers():
    for key, want in KAT["thresholds"].items():
        k, p = key.split(":")
        out = np.zeros(int(k) + 1, dtype=np.int32)
        P.lib().ko_thresholds(int(k), np.longdouble(float(p)), out)
        assert [int(x) for x in out] == want


@pytest.mark.parametrize("name", common.PIPES)
def test_port_pipeline_matches_golden(name):
    rs, g = common.load_pipe(name)
    got = common.run_port(rs, int(g["k"]), int(g["genomelen"]), int(g["coverage"]), int(g["seed"]))
    common.compare(got, g, what=name + ": ")
    assert np.array_equal(got["seq_after_diff"], g["seq_after_diff"])


needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("k,cov,L,kw", [
    (32, 20, 100, dict()),
    (21, 20, 100, dict(n_frac=0.003, n_rg=3)),
    (31, 30, 150, dict(err_scale=4.0)),
    (32, 25, 130, dict(ragged=True, min_len=64, n_frac=0.001)),
    (11, 20, 60, dict(err_scale=2.0)),
    (32, 12, 700, dict(err_scale=3.0, ragged=True, min_len=300, n_frac=0.001)),  # reads beyond 512 bases (the GPU's long-read path)
    (25, 16, 3000, dict(err_scale=0.7, ragged=True, min_len=100)),
])
def test_port_matches_reference(k, cov, L, kw):
    G = 60000
    rs = synth.make_reads(G, cov, L, seed=100 + k, **kw)
    alpha_ld = np.longdouble(7.0) / np.longdouble(cov)
    approx = int(np.longdouble(G) * cov * alpha_ld)
    rp = R.RefPipeline(rs, k)
    rsam, rtru = R.RefBloom(approx, 0.01), R.RefBloom(approx, 0.0005)
    want = {"slots": rp.subsample(rsam, float(alpha_ld), 99), "sampled_inserted": rsam.inserted,
            "sampled_words": rsam.words(), "sampled_fpr": rsam.fprate()}
    want["thresholds"] = rp.thresholds(rsam, 7.0, float(cov))
    want["pass2_errors"] = rp.infer_errors(rsam, want["thresholds"])
    rp.find_trusted(rtru, rsam, want["thresholds"])
    want["trusted_inserted"], want["trusted_words"] = rtru.inserted, rtru.words()
    want["pass3_errors"], seq_after = rp.get_errors(rtru, want_seq=True)
    want["rgc"], want["qc"], want["cyc"], want["dic"] = rp.covariates(rtru)
    want["meanq"], want["rgdq"], want["qdq"], want["cydq"], want["didq"] = rp.get_dqs()
    want["newqual"] = rp.recalibrate()
    got = common.run_port(rs, k, G, cov, 99)
    common.compare(got, want)
    assert np.array_equal(got["seq_after_diff"], np.flatnonzero(seq_after != rs.seq))


@needs_ref
@pytest.mark.parametrize("case", range(12))
def test_port_matches_reference_on_the_randomised_sweep(case):
    """The shapes the GPU is checked on against the port (tests/test_gpu_parity.py) are the shapes the port is checked on
    against the unmodified reference here: the chain reference -> port -> kernels has no untested link."""
    rs, k, G, cov, seed, _, what = common.random_case(case)
    want = common.run_ref(rs, k, G, cov, seed)
    got = common.run_port(rs, k, G, cov, seed)
    common.compare(got, want, what=what)
    assert np.array_equal(got["seq_after_diff"], want["seq_after_diff"])


@needs_ref
def test_port_fixed_mode_matches_reference():
    """--fixed (kbbq.cc:367-378): errors = read differs from the pre-corrected read, then consume_read."""
    rs = synth.make_reads(30000, 20, 100, seed=5, n_rg=2, n_frac=0.002)
    errors = (rs.seq != rs.clean_seq).astype(np.uint8)
    rp = R.RefPipeline(rs, 32)
    want = rp.consume_with_errors(errors)
    got = P.PortPipeline(rs, 32).consume(errors)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    d_ref = rp.get_dqs()
    d_port = P.PortPipeline(rs, 32).get_dqs(got)
    for a, b in zip(d_port, d_ref):
        assert np.array_equal(a, b)
