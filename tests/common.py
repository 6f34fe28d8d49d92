"""Shared helpers for the parity tests: golden fixtures, dense/sparse conversion, the three pipeline runners."""
import json
import os

import numpy as np

from kbbq_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
PIPES = ["pipe_k32", "pipe_k21_ragged", "pipe_k31_2rg"]


def load_kat():
    with open(os.path.join(GOLDEN, "kat.json")) as fh:
        return json.load(fh)


def load_pipe(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    rs = synth.ReadSet(seq=g["seq"], off=g["off"], qual=g["qual"], rg=g["rg"], second=g["second"], n_rg=int(g["n_rg"]))
    for nm in ("qc", "cyc", "dic", "qdq", "cydq", "didq"):
        a = np.zeros(int(np.prod(g[nm + "_shape"])), dtype=g[nm + "_val"].dtype)
        a[g[nm + "_idx"]] = g[nm + "_val"]
        g[nm] = a.reshape(tuple(g[nm + "_shape"]))
    total = rs.total_bases
    g["pass2_errors"] = np.unpackbits(g["pass2_errors"])[:total]
    g["pass3_errors"] = np.unpackbits(g["pass3_errors"])[:total]
    return rs, g


def set_bits(words: np.ndarray) -> np.ndarray:
    return np.flatnonzero(np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")).astype(np.int64)


def pad_cycle(a: np.ndarray, ml: int, axis: int = 3) -> np.ndarray:
    """grow/shrink the cycle axis of a dense table to ml entries (tables from different runs use different extents)."""
    if a.shape[axis] == ml:
        return a
    if a.shape[axis] > ml:
        idx = [slice(None)] * a.ndim
        idx[axis] = slice(0, ml)
        assert not a[tuple([slice(None)] * axis + [slice(ml, None)])].any()
        return a[tuple(idx)]
    pad = [(0, 0)] * a.ndim
    pad[axis] = (0, ml - a.shape[axis])
    return np.pad(a, pad)


def run_port(rs, k, genomelen, coverage, seed):
    """The whole pipeline through the oracle port; returns a dict shaped like the golden fixtures."""
    from oracle import port as P

    alpha_ld = np.longdouble(7.0) / np.longdouble(coverage)
    approx = int(np.longdouble(genomelen) * np.longdouble(coverage) * alpha_ld)
    pp = P.PortPipeline(rs, k)
    sampled, trusted = P.PortBloom(approx, 0.01), P.PortBloom(approx, 0.0005)
    out = {}
    out["slots"] = pp.subsample(sampled, float(alpha_ld), seed)
    out["sampled_inserted"] = sampled.inserted
    out["sampled_words"] = sampled.words()
    out["sampled_fpr"] = sampled.fprate()
    out["thresholds"] = pp.thresholds(sampled, 7.0, float(coverage))
    out["pass2_errors"] = pp.infer_errors(sampled, out["thresholds"])
    pp.find_trusted(trusted, sampled, out["thresholds"])
    out["trusted_inserted"] = trusted.inserted
    out["trusted_words"] = trusted.words()
    e3, seq_after = pp.get_errors(trusted, want_seq=True)
    out["pass3_errors"] = e3
    out["seq_after_diff"] = np.flatnonzero(seq_after != rs.seq).astype(np.int64)
    tables = pp.consume(e3, seq_after)
    out["rgc"], out["qc"], out["cyc"], out["dic"] = tables
    dqs = pp.get_dqs(tables)
    out["meanq"], out["rgdq"], out["qdq"], out["cydq"], out["didq"] = dqs
    out["newqual"] = pp.recalibrate(dqs)
    return out


def run_gpu(rs, k, genomelen, coverage, seed, n_batches=1, max_cycle=None):
    """The whole pipeline through the C ABI on cuda:0; returns a dict shaped like run_port's."""
    from kbbq_b200 import api
    from kbbq_b200.pipeline import Recalibrator

    ml = rs.max_len if max_cycle is None else max_cycle
    rc = Recalibrator(k, rs.n_rg, ml, genomelen, coverage, seed)
    bounds = np.linspace(0, rs.n, n_batches + 1).astype(int)
    subsets = [rs.subset(int(bounds[i]), int(bounds[i + 1])) for i in range(n_batches)]
    packed = [api.PackedBatch(s) for s in subsets]
    for pb in packed:
        rc.add_batch(pb)
    st = rc.run(fetch_quals=True, want_errors=True)
    ctx = rc.ctx
    out = {"stats": st}
    out["slots"] = rc.local_slots
    out["sampled_inserted"] = ctx.bloom_inserted(api.FILTER_SAMPLED)
    out["sampled_words"] = ctx.bloom_words(api.FILTER_SAMPLED)
    out["sampled_fpr"] = ctx.bloom_fprate(api.FILTER_SAMPLED)
    out["thresholds"] = st.thresholds
    out["trusted_inserted"] = ctx.bloom_inserted(api.FILTER_TRUSTED)
    out["trusted_words"] = ctx.bloom_words(api.FILTER_TRUSTED)
    out["pass2_errors"] = np.concatenate([pb.unpack_bits(e, s.total_bases) for pb, e, s in zip(packed, rc.last_errors2, subsets)])
    out["pass3_errors"] = np.concatenate([pb.unpack_bits(e, s.total_bases) for pb, e, s in zip(packed, rc.last_errors3, subsets)])
    out["rgc"], out["qc"], out["cyc"], out["dic"] = ctx.covariates()
    out["meanq"], out["rgdq"], out["qdq"], out["cydq"], out["didq"] = ctx.fit_model()
    out["newqual"] = np.concatenate([pb.unpack_qual(q, s.total_bases) for pb, q, s in zip(packed, rc.last_quals, subsets)])
    rc.close()
    return out


def compare(got, want, ml=None, what=""):
    """bit-exact comparison of two pipeline result dicts (golden fixtures carry set-bit lists instead of words)."""
    for key in ("slots", "sampled_inserted", "trusted_inserted"):
        assert int(got[key]) == int(want[key]), f"{what}{key}: {got[key]} != {want[key]}"
    for f in ("sampled", "trusted"):
        if f + "_words" in want:
            assert np.array_equal(got[f + "_words"], want[f + "_words"]), f"{what}{f} filter bits differ"
        else:
            assert np.array_equal(set_bits(got[f + "_words"]), want[f + "_bits"]), f"{what}{f} filter bits differ"
    assert float(got["sampled_fpr"]) == float(want["sampled_fpr"]), f"{what}fprate"
    assert np.array_equal(got["thresholds"], want["thresholds"]), f"{what}thresholds"
    for key in ("pass2_errors", "pass3_errors", "newqual"):
        nd = int(np.count_nonzero(np.asarray(got[key]) != np.asarray(want[key])))
        assert nd == 0, f"{what}{key}: {nd} positions differ"
    ml = want["cyc"].shape[3] if ml is None else ml
    for key in ("rgc", "qc", "dic", "meanq", "rgdq", "qdq", "didq"):
        assert np.array_equal(got[key], want[key]), f"{what}{key} differs"
    assert np.array_equal(pad_cycle(got["cyc"], ml), pad_cycle(want["cyc"], ml)), f"{what}cycle counts differ"
    assert np.array_equal(pad_cycle(got["cydq"], ml), pad_cycle(want["cydq"], ml)), f"{what}cycle dq differ"


def random_case(case):
    """The deterministic shape sweep shared by the GPU-vs-port test (tests/test_gpu_parity.py) and the
    port-vs-reference test (tests/test_oracle.py): (rs, k, genomelen, coverage, seed, n_batches, description)."""
    from kbbq_b200 import synth

    rng = np.random.default_rng(9000 + case)
    k = int(rng.choice([12, 16, 21, 25, 31, 32]))
    L = int(rng.choice([2 * k, 75, 100, 151, 200, 300]))
    L = max(L, 2 * k)
    cov = int(rng.choice([8, 15, 30, 60]))
    G = int(rng.choice([20000, 50000]))
    kw = dict(err_scale=float(rng.choice([0.5, 1.5, 4.0, 8.0])), n_frac=float(rng.choice([0.0, 0.001, 0.01])),
              n_rg=int(rng.choice([1, 2, 5])), ragged=bool(rng.integers(0, 2)))
    if kw["ragged"]:
        kw["min_len"] = int(rng.integers(k, L + 1))
    rs = synth.make_reads(G, cov, L, seed=300 + case, **kw)
    if case % 3 == 0:  # un-binned qualities: more rows, more quality tie-breaks
        rs.qual[:] = np.random.default_rng(case).integers(0, 42, rs.qual.shape[0], dtype=np.uint8)
    seed = int(rng.integers(1, 2**32))
    return rs, k, G, cov, seed, int(rng.integers(1, 4)), f"case {case} k={k} L={L} cov={cov} {kw}: "


def run_ref(rs, k, genomelen, coverage, seed):
    """The whole pipeline through the UNMODIFIED reference (oracle/_ref); a dict shaped like run_port's."""
    from oracle import refharness as R

    alpha_ld = np.longdouble(7.0) / np.longdouble(coverage)
    approx = int(np.longdouble(genomelen) * coverage * alpha_ld)
    rp = R.RefPipeline(rs, k)
    rsam, rtru = R.RefBloom(approx, 0.01), R.RefBloom(approx, 0.0005)
    want = {"slots": rp.subsample(rsam, float(alpha_ld), seed), "sampled_inserted": rsam.inserted,
            "sampled_words": rsam.words(), "sampled_fpr": rsam.fprate()}
    want["thresholds"] = rp.thresholds(rsam, 7.0, float(coverage))
    want["pass2_errors"] = rp.infer_errors(rsam, want["thresholds"])
    rp.find_trusted(rtru, rsam, want["thresholds"])
    want["trusted_inserted"], want["trusted_words"] = rtru.inserted, rtru.words()
    want["pass3_errors"], seq_after = rp.get_errors(rtru, want_seq=True)
    want["seq_after_diff"] = np.flatnonzero(seq_after != rs.seq).astype(np.int64)
    want["rgc"], want["qc"], want["cyc"], want["dic"] = rp.covariates(rtru)
    want["meanq"], want["rgdq"], want["qdq"], want["cydq"], want["didq"] = rp.get_dqs()
    want["newqual"] = rp.recalibrate()
    return want
