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


def test_cli_long_reads_against_live_reference_cli(tmp_path):
    """Reads of up to 2500 bases (the reference has no length limit) through both binaries, sharded over two GPUs where
    the box has them."""
    ref = os.path.join(ROOT, "oracle", "_ref", "kbbq_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/kbbq_ref not present")
    import sys
    sys.path.insert(0, ROOT)
    from kbbq_b200 import synth

    rs = synth.make_reads(60000, 16, 2500, seed=12, err_scale=0.7, ragged=True, min_len=90, n_rg=2)
    fq = str(tmp_path / "long.fq")
    synth.to_fastq(rs, fq)
    args = ["-k", "32", "-g", "60000", "-c", "16", fq]
    with open(tmp_path / "ref.gz", "wb") as fo:
        r = subprocess.run([ref] + args, stdout=fo, stderr=subprocess.PIPE)
    assert r.returncode == 0
    want = gzip.open(tmp_path / "ref.gz").read()
    got, _ = _run(["--seed", "2854027472"] + args, tmp_path)
    assert got == want
    import torch
    if torch.cuda.device_count() >= 2:
        got2, _ = _run(["--seed", "2854027472", "--gpus", "2"] + args, tmp_path)
        assert got2 == want


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


def _run_env(args, tmp_path, env):
    out = tmp_path / "out_env.gz"
    with open(out, "wb") as fo:
        r = subprocess.run([CLI] + args, stdout=fo, stderr=subprocess.PIPE, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stderr.decode()
    return gzip.open(out).read(), r.stderr.decode()


def _big_fastq(tmp_path, n_frac=0.001, seed=79):
    import sys
    sys.path.insert(0, ROOT)
    from kbbq_b200 import synth

    rs = synth.make_reads(200000, 20, 100, seed=seed, n_rg=3, n_frac=n_frac)
    fq = str(tmp_path / "big.fq")
    synth.to_fastq(rs, fq)
    return rs, fq


def test_cli_streamed_through_the_device_gives_the_resident_answer(tmp_path):
    """Inputs larger than device memory are streamed through it once per pass, batch i+1 travelling while batch i is
    computed (VERDICT r1 'missing' 1).  KBBQ_RESIDENT_BYTES=0 forces that path on a small input; the output must not
    change by a byte, nor may the host-side or device-side --fixed comparison differ."""
    rs, fq = _big_fastq(tmp_path)
    args = ["--seed", "2854027472", "-k", "32", "-g", "200000", "-c", "20", fq]
    want, _ = _run(args, tmp_path)
    got, _ = _run_env(args, tmp_path, {"KBBQ_RESIDENT_BYTES": "0"})
    assert got == want
    ref = os.path.join(ROOT, "oracle", "_ref", "kbbq_ref")
    if os.path.exists(ref):
        with open(tmp_path / "ref.gz", "wb") as fo:
            r = subprocess.run([ref] + args[2:], stdout=fo, stderr=subprocess.PIPE)
        assert r.returncode == 0
        assert got == gzip.open(tmp_path / "ref.gz").read()


def test_cli_fixed_mode_device_and_host_comparisons_agree(tmp_path):
    import numpy as np
    import sys
    sys.path.insert(0, ROOT)
    from kbbq_b200 import synth

    rs, fq = _big_fastq(tmp_path, n_frac=0.002, seed=80)
    fx = str(tmp_path / "fixed.fq")
    clean = rs.subset(0, rs.n)
    clean.seq = np.ascontiguousarray(clean.clean_seq)
    synth.to_fastq(clean, fx)
    args = ["-k", "32", "-f", fx, fq]
    dev, _ = _run(args, tmp_path)
    host, _ = _run_env(args, tmp_path, {"KBBQ_FIXED_ON_HOST": "1"})
    streamed, _ = _run_env(args, tmp_path, {"KBBQ_RESIDENT_BYTES": "0"})
    assert dev == host == streamed
    ref = os.path.join(ROOT, "oracle", "_ref", "kbbq_ref")
    if os.path.exists(ref):
        with open(tmp_path / "ref.gz", "wb") as fo:
            r = subprocess.run([ref] + args, stdout=fo, stderr=subprocess.PIPE)
        assert r.returncode == 0
        assert dev == gzip.open(tmp_path / "ref.gz").read()


def test_cli_sharded_over_gpus_gives_the_single_gpu_answer(tmp_path):
    """--gpus N: reads dealt to N GPUs inside the C++ host, partial filters merged over peer memory, count tables summed
    (SURVEY.md section 8(e)); byte-identical output, resident and streamed."""
    import sys
    sys.path.insert(0, ROOT)
    from kbbq_b200 import api

    n = api.lib().kbbq_device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs on the box")
    rs, fq = _big_fastq(tmp_path)
    args = ["--seed", "2854027472", "-k", "32", "-g", "200000", "-c", "20", fq]
    want, _ = _run(args, tmp_path)
    for g in sorted({2, min(n, 4), n}):
        got, _ = _run(["--gpus", str(g)] + args, tmp_path)
        assert got == want, f"--gpus {g}"
    got, _ = _run_env(["--gpus", "2"] + args, tmp_path, {"KBBQ_RESIDENT_BYTES": "0"})
    assert got == want


def test_cli_writes_a_gatk_style_report(tmp_path):
    rs, fq = _big_fastq(tmp_path)
    rep = str(tmp_path / "recal.table")
    _run(["--seed", "2854027472", "-k", "32", "-g", "200000", "-c", "20", "--report", rep, fq], tmp_path)
    text = open(rep).read()
    assert text.startswith("#:GATKReport.v1.1:3") and "RecalTable2" in text
    obs = [int(line.split()[4]) for line in text.split("#:GATKTable:RecalTable0:")[1].split("\n\n")[0].splitlines()[2:]]
    assert sum(obs) == rs.total_bases


@pytest.mark.parametrize("kind", ["fastq", "bam"])
def test_reference_main_bound_to_the_c_abi_matches_the_reference(tmp_path, kind):
    """oracle/_ref/kbbq_ref_b200 = the reference's UNMODIFIED main(), file classes, threshold maths and model fit linked
    with integration/recalibrateutils_b200.cc (the four pass drivers over include/kbbq_b200.h) instead of its
    recalibrateutils.cc: the drop-in boundary, compiled.  Its output must be the reference CLI's, byte for byte."""
    ref = os.path.join(ROOT, "oracle", "_ref", "kbbq_ref")
    bound = os.path.join(ROOT, "oracle", "_ref", "kbbq_ref_b200")
    if not (os.path.exists(ref) and os.path.exists(bound)):
        pytest.skip("oracle/_ref/kbbq_ref{,_b200} not present (built where /root/reference exists)")
    import sys
    sys.path.insert(0, ROOT)
    from kbbq_b200 import synth

    rs = synth.make_reads(100000, 20, 100, seed=81, n_rg=3, n_frac=0.001)
    if kind == "fastq":
        inp = str(tmp_path / "in.fq")
        synth.to_fastq(rs, inp)
        args = ["-k", "32", "-g", "100000", "-c", "20", inp]
    else:
        inp = str(tmp_path / "in.bam")
        synth.to_bam(rs, inp, genome_len=100000, rev_every=5, with_oq=True)
        args = ["-k", "32", "-c", "20", "--set-oq", inp]
    outs = []
    for exe in (ref, bound):
        with open(tmp_path / "o.gz", "wb") as fo:
            r = subprocess.run([exe] + args, stdout=fo, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr.decode()
        outs.append((gzip.open(tmp_path / "o.gz").read(), r.stderr.decode()))
    assert outs[0][0] == outs[1][0]
    # same seed (time()/getpid() pinned in both), same counters reported by main()
    pick = lambda log: [l.split("]", 1)[1] for l in log.splitlines() if "Sampled" in l or "Seed" in l or "false positive" in l]
    assert pick(outs[0][1]) == pick(outs[1][1])
