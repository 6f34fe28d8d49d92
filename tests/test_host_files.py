"""CPU-side checks of the host mirror's file layer (kbbq_b200/host: FastqFile, BamFile, the BGZF writer) through
tests/hosttool -- no GPU: records are read into the ReadStore, and rewritten with the qualities the reference CLI
produced (taken from the golden outputs); the result must be the golden stream byte for byte."""
import gzip
import os
import struct
import subprocess

import numpy as np
import pytest

from kbbq_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
TOOL = os.path.join(ROOT, "tests", "hosttool", "hosttool")


@pytest.fixture(scope="module")
def tool():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "kbbq_b200", "host"), "hosttool"])
    return TOOL


def _records(data):
    p = 4
    (lt,) = struct.unpack_from("<i", data, p); p += 4 + lt
    (nr,) = struct.unpack_from("<i", data, p); p += 4
    for _ in range(nr):
        (ln,) = struct.unpack_from("<i", data, p); p += 4 + ln + 4
    head, out = data[:p], []
    while p < len(data):
        (bs,) = struct.unpack_from("<i", data, p)
        out.append(data[p + 4:p + 4 + bs])
        p += 4 + bs
    return head, out


def _read_orientation_quals(rec):
    lname = rec[8]
    ncig, flag, lseq = struct.unpack_from("<HHi", rec, 12)
    qoff = 32 + lname + 4 * ncig + (lseq + 1) // 2
    q = rec[qoff:qoff + lseq]
    return bytes(c + 33 for c in (q[::-1] if flag & 16 else q))


def test_bam_reader_returns_the_read_set(tool):
    """cli_in.bam was written from this read set with every 3rd record reverse strand (tests/golden/make_cli_golden.py):
    what BamFile hands out (readutils.hh:30-42, readutils.cc:13-61) must be the read set itself."""
    rs = synth.make_reads(4000, 25, 150, seed=11, n_rg=2, n_frac=0.002)
    out = subprocess.run([tool, "dump", os.path.join(GOLD, "cli_in.bam")], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[-1] == "#ref\t4000" and len(out) == rs.n + 1
    for i, line in enumerate(out[:-1]):
        name, seq, q, rg, sec = line.split("\t")
        b0, b1 = int(rs.off[i]), int(rs.off[i + 1])
        assert name == "r%d" % i
        assert seq.encode() == rs.seq[b0:b1].tobytes()
        assert q.encode() == (rs.qual[b0:b1] + 33).tobytes()
        assert int(rg) == rs.rg[i] and int(sec) == rs.second[i]
    # --use-oq reads the qualities from the OQ tag (qual+1 in this file)
    out = subprocess.run([tool, "dump", os.path.join(GOLD, "cli_in.bam"), "use_oq"], capture_output=True, text=True, check=True).stdout.splitlines()
    q0 = out[0].split("\t")[2].encode()
    assert q0 == (np.minimum(rs.qual[:150].astype(np.int32) + 1, 93).astype(np.uint8) + 33).tobytes()


@pytest.mark.parametrize("inp,gold,flags", [
    ("cli_in.bam", "cli_bam_out.bam.gz", ["set_oq"]),
    ("cli_in.bam", "cli_bam_useoq_out.bam.gz", ["use_oq"]),
    ("cli_in_nooq.bam", "cli_bam_nooq_out.bam.gz", ["set_oq"]),
])
def test_bam_writer_reproduces_reference_stream(tool, tmp_path, inp, gold, flags):
    want = gzip.open(os.path.join(GOLD, gold)).read()
    _, wr = _records(want)
    qf = tmp_path / "q.txt"
    qf.write_bytes(b"\n".join(_read_orientation_quals(r) for r in wr) + b"\n")
    out = tmp_path / "out.bam"
    subprocess.run([tool, "rewrite", os.path.join(GOLD, inp), str(qf), str(out)] + flags, check=True)
    raw = out.read_bytes()
    assert gzip.decompress(raw) == want
    # BGZF framing: every member carries the BC extra field and the file ends with the 28-byte EOF block
    assert raw[:4] == b"\x1f\x8b\x08\x04" and raw[12:14] == b"BC"
    assert raw[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def test_fastq_writer_moves_comment_to_plus_line(tool, tmp_path):
    fq = tmp_path / "in.fq"
    fq.write_bytes(b"@r0/1 some comment\nACGTN\n+\nIIII#\n@r1/2_RG:Z:x\nGGCC\n+r1\n5555\n")
    (tmp_path / "q.txt").write_bytes(b"ABCDE\nFGHI\n")
    out = tmp_path / "out.gz"
    subprocess.run([tool, "rewrite", str(fq), str(tmp_path / "q.txt"), str(out)], check=True)
    assert gzip.decompress(out.read_bytes()) == b"@r0/1\nACGTN\n+some comment\nABCDE\n@r1/2_RG:Z:x\nGGCC\n+\nFGHI\n"
    dump = subprocess.run([tool, "dump", str(fq)], capture_output=True, text=True, check=True).stdout.splitlines()
    assert dump[0].split("\t") == ["r0", "ACGTN", "IIII#", "0", "0"]
    assert dump[1].split("\t") == ["r1", "GGCC", "5555", "1", "1"]


def _dump(tool, path, threads):
    env = dict(os.environ, HOSTTOOL_THREADS=str(threads))
    return subprocess.run([tool, "dump", str(path)], capture_output=True, text=True, check=True, env=env).stdout


def test_parallel_fastq_parse_equals_serial(tool, tmp_path):
    """A plain 4-line FASTQ large enough to be cut into ranges (>= 1 MiB per thread) parses to the same records with
    1 and 8 threads, read groups numbered in file order although they first appear in different ranges; qualities that
    begin with '@' or '+' do not confuse the search for record starts."""
    rng = np.random.default_rng(5)
    n, L = 30000, 100
    with open(tmp_path / "big.fq", "wb") as fh:
        for i in range(n):
            seq = bytes(rng.choice(list(b"ACGTN"), L, p=[.245, .245, .245, .245, .02]).astype(np.uint8))
            qual = bytearray((rng.integers(0, 41, L) + 33).astype(np.uint8).tobytes())
            if i % 7 == 0:
                qual[0] = ord("@")
            if i % 11 == 0:
                qual[0] = ord("+")
            rg = "late" if i > 2 * n // 3 else ("mid" if i > n // 3 else "early")
            fh.write(b"@r%d/%d_RG:Z:%s some comment %d\n" % (i, 1 + (i & 1), rg.encode(), i) + seq + b"\n+\n" + bytes(qual) + b"\n")
    assert os.path.getsize(tmp_path / "big.fq") > 4 << 20
    one, eight = _dump(tool, tmp_path / "big.fq", 1), _dump(tool, tmp_path / "big.fq", 8)
    assert one == eight and one.count("\n") == n
    lines = one.splitlines()
    assert lines[0].split("\t")[3] == "0" and lines[n // 2].split("\t")[3] == "1" and lines[-1].split("\t")[3] == "2"
    # bgzip-style input (independent 64 KiB members) goes through the parallel inflater and gives the same records
    bam_like = tmp_path / "big.fq.gz"
    raw = (tmp_path / "big.fq").read_bytes()
    import struct, zlib
    with open(bam_like, "wb") as fh:
        for p in range(0, len(raw), 0xFF00):
            chunk = raw[p:p + 0xFF00]
            co = zlib.compressobj(1, zlib.DEFLATED, -15)
            comp = co.compress(chunk) + co.flush()
            fh.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", 18 + len(comp) + 8 - 1) + comp + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
        fh.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    assert _dump(tool, bam_like, 8) == one
    # and plain gzip (one serial stream) too
    with gzip.open(tmp_path / "plain.fq.gz", "wb", compresslevel=1) as g:
        g.write(raw)
    assert _dump(tool, tmp_path / "plain.fq.gz", 8) == one


def test_multi_line_and_crlf_fastq_fall_back_to_kseq_semantics(tool, tmp_path):
    """Anything but plain 4-line records takes the serial parser: wrapped sequence/quality lines, CRLF, FASTA-style junk."""
    body = b"".join(b"@x%d c%d\r\nACGTACGTAC\r\nGTACGTNN\r\n+\r\nIIIIIIIIII\r\nIIIIII##\r\n" % (i, i) for i in range(40000))
    (tmp_path / "wrapped.fq").write_bytes(body)
    assert os.path.getsize(tmp_path / "wrapped.fq") > 2 << 20
    out = _dump(tool, tmp_path / "wrapped.fq", 8).splitlines()
    assert len(out) == 40000
    assert out[0].split("\t")[:3] == ["x0", "ACGTACGTACGTACGTNN", "IIIIIIIIIIIIIIII##"]
    assert out[-1].split("\t")[:3] == ["x39999", "ACGTACGTACGTACGTNN", "IIIIIIIIIIIIIIII##"]
    (tmp_path / "short_qual.fq").write_bytes(b"@a\nACGT\n+\nIII\n")
    r = subprocess.run([tool, "dump", str(tmp_path / "short_qual.fq")], capture_output=True, text=True)
    assert r.returncode == 1 and "truncated quality" in r.stderr


def test_cli_error_paths_without_a_gpu(tmp_path):
    """The CLI's argument and input errors (kbbq.cc:98-190) surface before any GPU work; with a valid input and no GPU
    it fails loudly instead of falling back to a CPU path."""
    cli = os.path.join(ROOT, "kbbq_b200", "host", "kbbq")
    subprocess.check_call(["make", "-s", "-C", os.path.dirname(cli)])
    (tmp_path / "tiny.fq").write_bytes(b"@ab\n" + b"ACGT" * 9 + b"\n+\n" + b"I" * 36 + b"\n")
    (tmp_path / "empty.fq").write_bytes(b"")

    def run(*args):
        r = subprocess.run([cli] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=tmp_path)
        return r.returncode, r.stderr.decode()

    rc, err = run("-k", "40", "tiny.fq")
    assert rc == 1 and "k must be <= 32" in err
    rc, err = run("nonexistent.fq")
    assert rc == 1 and "Error opening file nonexistent.fq" in err
    rc, err = run("-g", "100", "empty.fq")
    assert rc == 1 and "File format must be bam, cram, or fastq" in err
    rc, err = run("--bogus", "tiny.fq")
    assert rc == 1 and "Unknown argument" in err
    # a SAM text file starts with '@' like a FASTQ record: it is recognised and refused like the reference refuses it
    (tmp_path / "x.sam").write_bytes(b"@HD\tVN:1.6\n@SQ\tSN:c\tLN:10\nr\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII\n")
    rc, err = run("-g", "100", "x.sam")
    assert rc == 1 and "File format must be bam, cram, or fastq" in err
    (tmp_path / "x.cram").write_bytes(b"CRAM\x03\x00" + b"\0" * 20)
    rc, err = run("-g", "100", "x.cram")
    assert rc == 1 and "CRAM input is not supported" in err
    rc, err = run("-g", "100")  # the reference's default "-": standard input is refused with a message, not a crash
    assert rc == 1 and "standard input" in err
    rc, err = run("--gpus", "-1", "tiny.fq")
    assert rc == 1 and "gpus must be >= 0" in err
    # a FASTQ whose name line is a single character cannot name a read group or a mate: an error message, not std::terminate
    (tmp_path / "many.fq").write_bytes(b"".join(b"@\n" + b"ACGT" * 9 + b"\n+\n" + b"I" * 36 + b"\n" for _ in range(6000)))
    r = subprocess.run([cli, "-g", "100", "many.fq"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=tmp_path)
    assert r.returncode in (0, 1)  # never a signal
    # a read beyond the ABI's 65 535 bases: refused while packing, with the limit named, whatever the box has
    (tmp_path / "huge.fq").write_bytes(b"@r0\n" + b"ACGT" * 17500 + b"\n+\n" + b"I" * 70000 + b"\n")
    rc, err = run("-g", "100000", "-c", "5", "huge.fq")
    assert rc == 1 and "KBBQ_ABI_MAX_READ_LEN" in err
    from kbbq_b200 import api
    rc, err = run("tiny.fq")
    if api.lib().kbbq_device_count() == 0:  # the GPU context comes first: no device, no run
        assert rc == 1 and "no CPU path" in err
        rc, err = run("-g", "100", "-c", "5", "tiny.fq")
        assert rc == 1 and "no CPU path" in err
    else:
        assert rc == 1 and "--genomelen must be specified" in err


def _inflate_pieces(blob):
    import zlib
    out, p = [], 0
    while p < len(blob):
        n = int.from_bytes(blob[p:p + 4], "little")
        d = zlib.decompressobj(-15)
        piece = d.decompress(blob[p + 4:p + 4 + n]) + d.flush()
        assert d.eof and d.unused_data == b""  # one complete final block, nothing after it
        out.append(piece)
        p += 4 + n
    return b"".join(out)


def _deflate_cases():
    rng = np.random.default_rng(5)
    fib = [1, 1]
    while len(fib) < 40:
        fib.append(fib[-1] + fib[-2])
    deep = b"".join(bytes([i]) * min(f, 40000) for i, f in enumerate(fib[:24]))  # Fibonacci counts: plain Huffman depth > 15
    rec = b"".join(b"@SRR1234567.%d %d length=100\n" % (i, i) + bytes(rng.choice(list(b"ACGT"), 100).tolist()) + b"\n+\n"
                   + bytes(rng.choice(list(b"#,:FI"), 100, p=[.01, .04, .1, .35, .5]).tolist()) + b"\n" for i in range(3000))
    return {
        "empty": b"", "one": b"A", "two": b"AB", "same": b"x" * 200000, "random": rng.bytes(150000),
        "short_edges": b"".join(bytes(rng.integers(65, 70, n).tolist()) for n in (15, 16, 31, 32, 33)),
        "deep": bytes(rng.permutation(np.frombuffer(deep, dtype=np.uint8)).tolist()), "fastq": rec,
        "periodic": (b"ACGTTGCA" * 9 + b"N") * 3000, "far": rng.bytes(40000) * 3,
        "bam_like": rng.integers(0, 16, 120000).astype(np.uint8).tobytes() + bytes(200) + rng.integers(0, 4, 9000).astype(np.uint8).tobytes(),
    }


@pytest.mark.parametrize("min_match", [8, 20, 1 << 30])
@pytest.mark.parametrize("block", [65280, 65535, 4096, 37])
def test_sink_deflate_encoder_round_trips_through_zlib(tool, tmp_path, block, min_match):
    """fast_deflate.cc against zlib's inflater: every shape of block (empty, one symbol, incompressible -> stored,
    code depths beyond the 15-bit limit, matches at every distance up to the window, the last bytes of a block)."""
    for name, data in _deflate_cases().items():
        if block == 37 and len(data) > 40000:
            data = data[:40000]
        (tmp_path / "in").write_bytes(data)
        subprocess.run([tool, "deflate", str(tmp_path / "in"), str(tmp_path / "out"), str(block), str(min_match)], check=True)
        blob = (tmp_path / "out").read_bytes()
        assert _inflate_pieces(blob) == data, name
        if name in ("same", "periodic") and block >= 4096 and min_match <= 20:
            assert len(blob) < len(data) // 20, name  # matches are found and used
        if name == "random":
            assert len(blob) <= len(data) + 9 * (len(data) // block + 1), name  # never worse than stored blocks
        if name == "fastq" and block >= 65280:
            assert len(blob) < 0.5 * len(data)


@pytest.mark.parametrize("level", [None, "huff", "1", "6"])
def test_bgzf_writer_modes_all_inflate_to_the_same_stream(tool, tmp_path, level):
    want = gzip.open(os.path.join(GOLD, "cli_bam_out.bam.gz")).read()
    _, wr = _records(want)
    (tmp_path / "q.txt").write_bytes(b"\n".join(_read_orientation_quals(r) for r in wr) + b"\n")
    env = dict(os.environ)
    env.pop("KBBQ_BGZF_LEVEL", None)
    if level:
        env["KBBQ_BGZF_LEVEL"] = level
    subprocess.run([tool, "rewrite", os.path.join(GOLD, "cli_in.bam"), str(tmp_path / "q.txt"), str(tmp_path / "o.bam"), "set_oq"], check=True, env=env)
    raw = (tmp_path / "o.bam").read_bytes()
    assert gzip.decompress(raw) == want
    p = 0
    while p < len(raw):  # every member is a BGZF block whose BSIZE field is its length
        assert raw[p:p + 4] == b"\x1f\x8b\x08\x04" and raw[p + 12:p + 14] == b"BC"
        p += int.from_bytes(raw[p + 16:p + 18], "little") + 1
    assert p == len(raw)


def test_sink_deflate_encoder_on_mixed_content(tool, tmp_path):
    """2 MB of segments with very different statistics (random bytes, two-symbol runs, long repeats, zero runs, text with
    a shifting alphabet) cut into small blocks, so that hundreds of differently shaped code tables are built and read back."""
    rng = np.random.default_rng(11)
    parts = []
    while sum(len(p) for p in parts) < (2 << 20):
        kind = int(rng.integers(0, 6))
        n = int(rng.integers(1, 30000))
        if kind == 0:
            parts.append(rng.bytes(n))
        elif kind == 1:
            parts.append(bytes(rng.choice([65, 66], n, p=[0.97, 0.03]).astype(np.uint8).tolist()))
        elif kind == 2:
            motif = rng.bytes(int(rng.integers(1, 300)))
            parts.append((motif * (n // len(motif) + 1))[:n])
        elif kind == 3:
            parts.append(bytes(n))
        elif kind == 4:
            lo = int(rng.integers(0, 200))
            parts.append(bytes((rng.integers(0, int(rng.integers(2, 56)), n) + lo).astype(np.uint8).tolist()))
        else:
            parts.append(bytes(np.minimum(rng.geometric(0.05, n), 255).astype(np.uint8).tolist()))
    data = b"".join(parts)
    (tmp_path / "in").write_bytes(data)
    for block, min_match in ((1000, 8), (4096, 12), (65280, 8), (65535, 258)):
        subprocess.run([tool, "deflate", str(tmp_path / "in"), str(tmp_path / "out"), str(block), str(min_match)], check=True)
        assert _inflate_pieces((tmp_path / "out").read_bytes()) == data, (block, min_match)


def test_session_packs_ahead_of_its_contexts(tool, tmp_path):
    """kbbq::Session(background): the contexts are created on a side thread and Session::prepack cuts the input into
    batches and packs every GPU's first batch without one -- on a box without a GPU the packing still happens and the
    missing device is reported, loudly, only when a context is asked for."""
    from kbbq_b200 import api, synth
    rs = synth.make_reads(50000, 12, 120, seed=8, ragged=True, min_len=20, n_rg=2)
    fq = tmp_path / "in.fq"
    synth.to_fastq(rs, str(fq))
    k, ngpus = 25, 3
    out = subprocess.run([tool, "prepack", str(fq), str(k), str(ngpus)], capture_output=True, text=True, check=True).stdout.splitlines()
    rows = [tuple(int(x) for x in ln.split("\t")) for ln in out if not ln.startswith("contexts")]
    assert [r[2] for r in rows] == [i % ngpus for i in range(len(rows))]           # dealt round-robin
    assert rows[0][0] == 0 and rows[-1][1] == rs.n and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    lens = np.diff(rs.off).astype(np.int64)
    for lo, hi, _gpu, n_words, slots, prepacked in rows:
        assert prepacked == 1                                                        # one batch per GPU here: all packed ahead
        assert n_words == api.PackedBatch(rs.subset(lo, hi)).n_words
        assert slots == int(np.maximum(lens[lo:hi] - k + 1, 0).sum())
    ctx = [ln for ln in out if ln.startswith("contexts")][0].split("\t")
    if api.lib().kbbq_device_count() >= ngpus:
        assert ctx[1] == str(ngpus)
    else:
        assert ctx[1] == "failed"
