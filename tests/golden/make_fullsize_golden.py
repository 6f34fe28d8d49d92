"""Generates tests/golden/c3_digests.json: what the UNMODIFIED reference (oracle/_ref) computes for the whole config-3
input of bench.py -- 12.8 M reads x 150 bp of the synthetic stream (kbbq_synth_host, the host twin of the device
generator), 64 Mb genome, 30X, k = 32 -- reduced to digests so that the fixture is a few hundred bytes:

  filters        the library's order-independent 128-bit digest (api.digest_words) of each bit array + insert counters
  thresholds     as a list
  count tables   digest of rg|q|cycle|dinuc laid out back to back as the library holds them
  new qualities  digest of the packed-layout quality array (32 bytes per word, zero padded)

Run here (needs /root/reference compiled into oracle/_ref; ~10 minutes, ~20 GB of RAM):
    python tests/golden/make_fullsize_golden.py [--genome 64000000 --coverage 30]
tests/test_full_size.py::test_config3_full_size_matches_reference_digests compares the GPU run with it.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from kbbq_b200 import api  # noqa: E402
from oracle import refharness as R  # noqa: E402

K, READ_LEN, N_RG, SEED, ERR_SCALE = 32, 150, 2, 2854027472, 1.5


def hexd(d):
    return "%016x%016x" % d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=64_000_000)
    ap.add_argument("--coverage", type=int, default=30)
    ap.add_argument("--out", default=os.path.join(HERE, "c3_digests.json"))
    args = ap.parse_args()
    G, cov = args.genome, args.coverage
    n = G * cov // READ_LEN
    sp = api.SynthParams(seed=1000, genome_seed=999, genome_len=G, read_len=READ_LEN, n_rg=N_RG, err_scale=ERR_SCALE)
    t0 = time.time()
    rs = api.synth_host(sp, 0, n, n_threads=os.cpu_count() or 1)
    print("reads made", time.time() - t0, flush=True)
    alpha_ld = np.longdouble(7.0) / np.longdouble(cov)
    approx = int(np.longdouble(G) * cov * alpha_ld)
    rp = R.RefPipeline(rs, K)
    sampled, trusted = R.RefBloom(approx, 0.01), R.RefBloom(approx, 0.0005)
    out = {"genome": G, "coverage": cov, "read_len": READ_LEN, "k": K, "n_rg": N_RG, "seed": SEED, "err_scale": ERR_SCALE,
           "synth_seed": 1000, "genome_seed": 999, "n_reads": n}
    out["slots"] = rp.subsample(sampled, float(alpha_ld), SEED)
    print("pass 1", time.time() - t0, flush=True)
    out["sampled_inserted"] = sampled.inserted
    out["sampled_digest"] = hexd(api.digest_words(sampled.words()))
    th = rp.thresholds(sampled, 7.0, float(cov))
    out["thresholds"] = [int(x) for x in th]
    rp.find_trusted(trusted, sampled, th)
    print("pass 2", time.time() - t0, flush=True)
    out["trusted_inserted"] = trusted.inserted
    out["trusted_digest"] = hexd(api.digest_words(trusted.words()))
    rgc, qc, cyc, dic = rp.covariates(trusted)
    print("pass 3", time.time() - t0, flush=True)
    out["counts_digest"] = hexd(api.digest_words(np.concatenate([a.reshape(-1) for a in (rgc, qc, cyc, dic)])))
    out["errors_total"] = int(rgc[:, 0].sum())
    rp.get_dqs()
    newq = rp.recalibrate()
    print("pass 4", time.time() - t0, flush=True)
    wpr = (READ_LEN + 31) // 32
    padded = np.zeros((n, wpr * 32), dtype=np.uint8)
    padded[:, :READ_LEN] = newq.reshape(n, READ_LEN)
    out["newqual_digest"] = hexd(api.digest_words(padded))
    out["stages_s"] = {k: float(v) for k, v in rp.times.items()}
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
