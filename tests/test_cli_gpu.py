"""End to end through the C++ host mirror and the kbbq CLI (kbbq_b200/host): the recalibrated FASTQ must be byte for
byte the reference CLI's (oracle/_ref/kbbq_ref run with its seed pinned; outputs committed under tests/golden/)."""
import gzip
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "kbbq_b200", "host", "kbbq")
GOLD = os.path.join(ROOT, "tests", "golden")
pytestmark = pytest.mark.gpu


def _gunzip_to(src, dst):
    with gzip.open(src) as f, open(dst, "wb") as o:
        o.write(f.read())


def _run(args, tmp_path):
    out = tmp_path / "out.gz"
    with open(out, "wb") as fo:
        r = subprocess.run([CLI] + args, stdout=fo, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    return gzip.open(out).read(), r.stderr.decode()


def test_cli_matches_reference_cli(tmp_path):
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.dirname(CLI)])
    fq = str(tmp_path / "in.fq")
    _gunzip_to(os.path.join(GOLD, "cli_in.fq.gz"), fq)
    got, log = _run(["--seed", "2854027472", "-k", "31", "-g", "4000", "-c", "25", fq], tmp_path)
    want = gzip.open(os.path.join(GOLD, "cli_out.fq.gz")).read()
    assert "Seed: 2854027472" in log and "Sampled 20870 valid kmers." in log
    assert got == want
    # gzip input gives the same answer (the reference reads FASTQ through BGZF either way)
    got2, _ = _run(["--seed", "2854027472", "-k", "31", "-g", "4000", "-c", "25", os.path.join(GOLD, "cli_in.fq.gz")], tmp_path)
    assert got2 == want
    # coverage estimated by the extra pass (kbbq.cc:229-250): 666*150/4000 = 24 != 25, so the output differs from -c 25
    got3, log3 = _run(["--seed", "2854027472", "-k", "31", "-g", "4000", fq], tmp_path)
    assert "Estimated coverage: 24" in log3


def test_cli_fixed_mode_matches_reference_cli(tmp_path):
    fq, fx = str(tmp_path / "in.fq"), str(tmp_path / "fixed.fq")
    _gunzip_to(os.path.join(GOLD, "cli_in.fq.gz"), fq)
    _gunzip_to(os.path.join(GOLD, "cli_fixed.fq.gz"), fx)
    got, _ = _run(["-k", "31", "-f", fx, fq], tmp_path)
    assert got == gzip.open(os.path.join(GOLD, "cli_fixed_out.fq.gz")).read()


def test_cli_against_live_reference_cli(tmp_path):
    """Where the prebuilt reference CLI travelled with the repo: a fresh, larger input through both binaries."""
    ref = os.path.join(ROOT, "oracle", "_ref", "kbbq_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/kbbq_ref not present")
    import sys
    sys.path.insert(0, ROOT)
    from kbbq_b200 import synth

    rs = synth.make_reads(100000, 20, 100, seed=77, n_rg=3, n_frac=0.001)
    fq = str(tmp_path / "big.fq")
    synth.to_fastq(rs, fq)
    args = ["-k", "32", "-g", "100000", "-c", "20", fq]
    with open(tmp_path / "ref.gz", "wb") as fo:
        r = subprocess.run([ref] + args, stdout=fo, stderr=subprocess.PIPE)
    assert r.returncode == 0
    got, _ = _run(["--seed", "2854027472"] + args, tmp_path)
    assert got == gzip.open(tmp_path / "ref.gz").read()


def _bam_records(data):
    """Split an uncompressed BAM stream into (header bytes, [record bytes])."""
    import struct
    assert data[:4] == b"BAM\1"
    p = 4
    (l_text,) = struct.unpack_from("<i", data, p); p += 4 + l_text
    (n_ref,) = struct.unpack_from("<i", data, p); p += 4
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", data, p); p += 4 + l_name + 4
    head, recs = data[:p], []
    while p < len(data):
        (bs,) = struct.unpack_from("<i", data, p)
        recs.append(data[p + 4:p + 4 + bs])
        p += 4 + bs
    return head, recs


@pytest.mark.parametrize("inp,args,gold", [
    ("cli_in.bam", ["-k", "31", "-c", "25", "--set-oq"], "cli_bam_out.bam.gz"),           # genome length from the header
    ("cli_in.bam", ["-k", "31", "-c", "25", "-g", "4000", "--use-oq"], "cli_bam_useoq_out.bam.gz"),
    ("cli_in_nooq.bam", ["-k", "31", "-c", "25", "--set-oq"], "cli_bam_nooq_out.bam.gz"),  # OQ tag appended
])
def test_cli_bam_matches_reference_cli(tmp_path, inp, args, gold):
    """BASELINE.json configs[1]'s input format: BAM with read groups, reverse-strand records and --set-oq / --use-oq
    (htsiter.cc:11-35, readutils.cc:13-61); the uncompressed output stream must be the reference CLI's byte for byte."""
    got, log = _run(["--seed", "2854027472"] + args + [os.path.join(GOLD, inp)], tmp_path)
    want = gzip.open(os.path.join(GOLD, gold)).read()
    if got != want:  # say where
        gh, gr = _bam_records(got)
        wh, wr = _bam_records(want)
        assert gh == wh, "BAM header differs"
        assert len(gr) == len(wr)
        bad = [i for i, (a, b) in enumerate(zip(gr, wr)) if a != b]
        assert not bad, "first differing record %d of %d" % (bad[0], len(bad))
    assert got == want
    if "-g" not in args:
        assert "Genome length is 4000 bp." in log


def test_cli_bam_against_live_reference_cli(tmp_path):
    ref = os.path.join(ROOT, "oracle", "_ref", "kbbq_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/kbbq_ref not present")
    import sys
    sys.path.insert(0, ROOT)
    from kbbq_b200 import synth

    rs = synth.make_reads(100000, 20, 100, seed=78, n_rg=3, n_frac=0.001)
    bam = str(tmp_path / "big.bam")
    synth.to_bam(rs, bam, genome_len=100000, rev_every=5, with_oq=True)
    for extra in (["--set-oq"], ["--use-oq", "--set-oq"]):
        args = ["-k", "32", "-c", "20"] + extra + [bam]
        with open(tmp_path / "ref.gz", "wb") as fo:
            r = subprocess.run([ref] + args, stdout=fo, stderr=subprocess.PIPE)
        assert r.returncode == 0
        got, _ = _run(["--seed", "2854027472"] + args, tmp_path)
        assert got == gzip.open(tmp_path / "ref.gz").read()


def test_cli_fixed_mode_with_a_shorter_fixed_file(tmp_path):
    """main()'s --fixed loop stops consuming where the shorter file ends (kbbq.cc:370); every read is still
    recalibrated.  Checked against the live reference CLI."""
    ref = os.path.join(ROOT, "oracle", "_ref", "kbbq_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/kbbq_ref not present")
    import sys
    sys.path.insert(0, ROOT)
    import numpy as np
    from kbbq_b200 import synth

    rs = synth.make_reads(30000, 20, 100, seed=91, n_rg=2, n_frac=0.002)
    fq, fx = str(tmp_path / "in.fq"), str(tmp_path / "fixed.fq")
    synth.to_fastq(rs, fq)
    clean = rs.subset(0, rs.n * 2 // 3)
    clean.seq = np.ascontiguousarray(clean.clean_seq)
    synth.to_fastq(clean, fx)
    args = ["-k", "32", "-f", fx, fq]
    with open(tmp_path / "ref.gz", "wb") as fo:
        r = subprocess.run([ref] + args, stdout=fo, stderr=subprocess.PIPE)
    assert r.returncode == 0
    got, _ = _run(args, tmp_path)
    assert got == gzip.open(tmp_path / "ref.gz").read()
