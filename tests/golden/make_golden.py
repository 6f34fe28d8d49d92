"""Generate the golden fixtures in this directory from the UNMODIFIED reference (oracle/_ref).

Run in the authoring container only (needs /root/reference to build oracle/_ref):
    make -C oracle ref && python tests/golden/make_golden.py
The fixtures are small .npz/.json files that travel with the repo, so that the oracle port and the CUDA path
can be checked against reference outputs on boxes where /root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kbbq_b200 import synth  # noqa: E402
from oracle import refharness as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sparse(a):
    a = np.ascontiguousarray(a)
    idx = np.flatnonzero(a.reshape(-1))
    return idx.astype(np.int64), a.reshape(-1)[idx]


def pipeline_fixture(name, rs, k, cov, seed):
    genomelen = int(name_to_genomelen[name])
    alpha = float(np.longdouble(7.0) / np.longdouble(cov))
    approx = int(np.longdouble(genomelen) * cov * (np.longdouble(7.0) / np.longdouble(cov)))  # kbbq.cc:264
    p = R.RefPipeline(rs, k)
    sampled, trusted = R.RefBloom(approx, 0.01), R.RefBloom(approx, 0.0005)
    slots = p.subsample(sampled, alpha, seed)
    th = p.thresholds(sampled, 7.0, float(cov))
    e2 = p.infer_errors(sampled, th)
    p.find_trusted(trusted, sampled, th)
    e3, seq_after = p.get_errors(trusted, want_seq=True)
    rgc, qc, cyc, dic = p.covariates(trusted)
    meanq, rgdq, qdq, cydq, didq = p.get_dqs()
    newq = p.recalibrate()
    out = dict(
        k=k, coverage=cov, genomelen=genomelen, seed=seed, alpha=alpha, approx_kmers=approx, n_rg=rs.n_rg,
        seq=rs.seq, off=rs.off, qual=rs.qual, rg=rs.rg, second=rs.second,
        slots=slots, sampled_inserted=sampled.inserted, sampled_fpr=sampled.fprate(),
        sampled_bits=np.flatnonzero(np.unpackbits(sampled.words().view(np.uint8), bitorder="little")).astype(np.int64),
        sampled_table_bits=sampled.table_bits, thresholds=th,
        pass2_errors=np.packbits(e2), trusted_inserted=trusted.inserted,
        trusted_bits=np.flatnonzero(np.unpackbits(trusted.words().view(np.uint8), bitorder="little")).astype(np.int64),
        trusted_table_bits=trusted.table_bits,
        pass3_errors=np.packbits(e3), seq_after_diff=np.flatnonzero(seq_after != rs.seq).astype(np.int64),
        rgc=rgc, meanq=meanq, rgdq=rgdq, newqual=newq,
    )
    for nm, a in (("qc", qc), ("cyc", cyc), ("dic", dic), ("qdq", qdq), ("cydq", cydq), ("didq", didq)):
        i, v = sparse(a)
        out[nm + "_idx"], out[nm + "_val"], out[nm + "_shape"] = i, v, np.array(a.shape)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "reads", rs.n, "slots", slots, "sampled", sampled.inserted, "trusted", trusted.inserted, "errors", int(e3.sum()))


name_to_genomelen = {}


def main():
    # ---- known-answer vectors, regenerated from the reference build ----
    kat = {}
    kat["kmer"] = []
    for k, s in [(32, "ACGT" * 8), (21, "GATTACAGATTACAGATTACAGATTACAGATTACAGG"), (31, "GATTACAGATTACAGATTACAGATTACAGATTACAGG"),
                 (32, "GATTACAGATTACAGATTACAGATTACAGATTACAGG"), (32, "GATTACAGATTACAGATTACAGATTACAGATTACAGGN"),
                 (5, "acgtRYacgt0123"), (32, "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT"), (1, "G"), (2, "GA")]:
        r = R.ref_kmer(k, s.encode())
        kat["kmer"].append({"k": k, "s": s, **{a: (str(b) if a != "valid" else bool(b)) for a, b in r.items()}})
    kat["bloom"] = []
    for proj, fpr in [(7000000, 0.01), (7000000, 0.0005), (100, 0.05), (32200000, 0.01), (448000000, 0.0005), (12345, 1e-6)]:
        b = R.RefBloom(proj, fpr)
        keys = [0, 1, 21345423534512, 0x1B1B1B1B1B1B1B1B, 0xFFFFFFFFFFFFFFFF, 123543451243]
        pat = b.patterns()
        kat["bloom"].append({
            "projected": proj, "fpr": fpr, "table_bits": b.table_bits, "n_blocks": b.n_blocks, "n_salts": b.n_salts,
            "random_seed": str(b.random_seed), "salts": [int(x) for x in b.salts], "optimal_k": b.optimal_k,
            "optimal_table_size": b.optimal_table_size,
            "hashes": {str(key): list(b.hashes(key)) for key in keys},
            "pattern0": [str(int(x)) for x in pat[:8]], "pattern36456": [str(int(x)) for x in pat[36456 * 8:36456 * 8 + 8]],
            "pattern_xor": str(int(np.bitwise_xor.reduce(pat))), "pattern_popcount": int(np.unpackbits(pat.view(np.uint8)).sum()),
        })
    # test.cc:26-59: insert 21345423534512 into (100, .05) and it must be contained
    b = R.RefBloom(100, 0.05)
    b.insert(21345423534512)
    kat["test_cc"] = {"contains_kmer": b.contains(21345423534512), "contains_gibberish": b.contains(123543451243),
                      "words": [str(int(x)) for x in b.words()]}
    st, out = R.ref_rng(12345, 0, 8)
    kat["rng"] = {"seed": 12345, "state": [str(int(x)) for x in st], "first": [str(int(x)) for x in out],
                  "at_1e6": str(int(R.ref_rng(12345, 1000000, 1)[1][0])),
                  "bernoulli_0.35": "".join(map(str, R.ref_bernoulli(12345, 0.35, 64))),
                  "seed_2854027472_state": [str(int(x)) for x in R.ref_rng(2854027472, 0, 0)[0]]}
    kat["thresholds"] = {}
    for k, p in [(32, 0.578922), (21, 0.3), (31, 0.0999), (32, 0.95), (5, 0.5)]:
        o = np.zeros(k + 1, dtype=np.int32)
        R.lib().ref_thresholds(k, p, o)
        kat["thresholds"][f"{k}:{p}"] = [int(x) for x in o]
    with open(os.path.join(HERE, "kat.json"), "w") as fh:
        json.dump(kat, fh, indent=1)
    print("kat.json written")

    # ---- small end-to-end pipelines ----
    name_to_genomelen.update({"pipe_k32": 3000, "pipe_k21_ragged": 3000, "pipe_k31_2rg": 4000})
    pipeline_fixture("pipe_k32", synth.make_reads(3000, 20, 100, 11), 32, 20, 12345)
    pipeline_fixture("pipe_k21_ragged", synth.make_reads(3000, 30, 120, 12, ragged=True, min_len=60, n_frac=0.004), 21, 30, 777)
    pipeline_fixture("pipe_k31_2rg", synth.make_reads(4000, 25, 150, 13, n_rg=2, n_frac=0.002, err_scale=3.0), 31, 25, 2854027472)


if __name__ == "__main__":
    main()
