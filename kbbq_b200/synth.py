"""Synthetic read sets for parity tests and bench.py.

Random i.i.d. genome, uniformly placed fixed-length reads on either strand,
binned qualities drawn from {2, 11, 25, 37, 40}, substitution errors injected at
``c * 10^(-q/10)`` (deliberately mis-calibrated), optional N bases, optional
ragged lengths, several read groups and first/second-of-pair flags.

Everything is returned as flat numpy arrays in the layout the host packer
(`kbbq_b200.api.pack_reads`) and the oracle harness both take:
  seq   uint8[total]  ASCII bases, reads concatenated without separators
  off   uint64[n+1]   base offsets
  qual  uint8[total]  phred values (no +33)
  rg    int32[n]      dense read-group id
  second uint8[n]     1 = second of pair
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_QVALS = np.array([2, 11, 25, 37, 40], dtype=np.uint8)
_QPROB = [0.01, 0.04, 0.10, 0.35, 0.50]


@dataclass
class ReadSet:
    seq: np.ndarray
    off: np.ndarray
    qual: np.ndarray
    rg: np.ndarray
    second: np.ndarray
    n_rg: int
    clean_seq: np.ndarray | None = None  # the same reads before error injection (--fixed input)

    @property
    def n(self) -> int:
        return int(self.off.shape[0] - 1)

    @property
    def total_bases(self) -> int:
        return int(self.off[-1])

    @property
    def max_len(self) -> int:
        return int(np.diff(self.off).max()) if self.n else 0

    def subset(self, lo: int, hi: int) -> "ReadSet":
        b0, b1 = int(self.off[lo]), int(self.off[hi])
        return ReadSet(
            seq=np.ascontiguousarray(self.seq[b0:b1]),
            off=np.ascontiguousarray(self.off[lo : hi + 1] - self.off[lo]),
            qual=np.ascontiguousarray(self.qual[b0:b1]),
            rg=np.ascontiguousarray(self.rg[lo:hi]),
            second=np.ascontiguousarray(self.second[lo:hi]),
            n_rg=self.n_rg,
            clean_seq=None if self.clean_seq is None else np.ascontiguousarray(self.clean_seq[b0:b1]),
        )


def make_reads(
    genome_len: int,
    coverage: int,
    read_len: int,
    seed: int = 1,
    *,
    n_rg: int = 1,
    err_scale: float = 1.5,
    n_frac: float = 0.0,
    ragged: bool = False,
    min_len: int | None = None,
    n_reads: int | None = None,
    chunk: int = 1 << 18,
    keep_clean: bool = True,
    genome_seed: int | None = None,
) -> ReadSet:
    """Appendix-C generator.  ``ragged`` draws lengths uniformly in [min_len, read_len].
    ``genome_seed``: draw the genome from its own stream, so that several calls with different ``seed`` (the shards of
    a multi-GPU run) sample reads from the SAME genome."""
    rng = np.random.default_rng(seed)
    grng = rng if genome_seed is None else np.random.default_rng(genome_seed)
    genome = grng.integers(0, 4, genome_len, dtype=np.uint8)
    n = genome_len * coverage // read_len if n_reads is None else n_reads
    if ragged:
        lo = read_len // 2 if min_len is None else min_len
        lens = rng.integers(lo, read_len + 1, n).astype(np.uint64)
    else:
        lens = np.full(n, read_len, dtype=np.uint64)
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    total = int(off[-1])
    seq = np.empty(total, dtype=np.uint8)
    clean = np.empty(total if keep_clean else 0, dtype=np.uint8)
    qual = np.empty(total, dtype=np.uint8)
    ar = np.arange(read_len, dtype=np.int64)
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        m = c1 - c0
        starts = rng.integers(0, genome_len - read_len + 1, m)
        strand = rng.integers(0, 2, m)
        codes = genome[starts[:, None] + ar[None, :]]
        codes = np.where(strand[:, None] == 1, 3 - codes[:, ::-1], codes)
        q = _QVALS[rng.choice(5, size=(m, read_len), p=_QPROB)]
        err = rng.random((m, read_len)) < (10.0 ** (-q.astype(np.float64) / 10.0)) * err_scale
        sub = rng.integers(1, 4, (m, read_len), dtype=np.uint8)
        noisy = np.where(err, (codes + sub) % 4, codes).astype(np.uint8)
        chars = _ACGT[noisy]
        cchars = _ACGT[codes]
        if n_frac > 0:
            isn = rng.random((m, read_len)) < n_frac
            chars = np.where(isn, np.uint8(ord("N")), chars)
            cchars = np.where(isn, np.uint8(ord("N")), cchars)
        if ragged:
            keep = ar[None, :] < lens[c0:c1, None].astype(np.int64)
            seq[int(off[c0]) : int(off[c1])] = chars[keep]
            if keep_clean:
                clean[int(off[c0]) : int(off[c1])] = cchars[keep]
            qual[int(off[c0]) : int(off[c1])] = q[keep]
        else:
            seq[int(off[c0]) : int(off[c1])] = chars.reshape(-1)
            if keep_clean:
                clean[int(off[c0]) : int(off[c1])] = cchars.reshape(-1)
            qual[int(off[c0]) : int(off[c1])] = q.reshape(-1)
    idx = np.arange(n)
    second = (idx & 1).astype(np.uint8)
    rg = ((idx // 2) % n_rg).astype(np.int32) if n_rg > 1 else np.zeros(n, dtype=np.int32)
    return ReadSet(seq=seq, off=off, qual=qual, rg=rg, second=second, n_rg=n_rg, clean_seq=clean if keep_clean else None)


def to_fastq(rs: ReadSet, path: str) -> None:
    """Write the read set as FASTQ; names r<i>/1, r<i>/2 (exercises `second`, readutils.cc:89-95)."""
    with open(path, "wb") as fh:
        for i in range(rs.n):
            b0, b1 = int(rs.off[i]), int(rs.off[i + 1])
            name = b"@r%d/%d" % (i, 1 + int(rs.second[i]))
            if rs.n_rg > 1:
                name += b"_RG:Z:rg%d" % int(rs.rg[i])
            fh.write(name + b"\n" + rs.seq[b0:b1].tobytes() + b"\n+\n" + (rs.qual[b0:b1] + 33).tobytes() + b"\n")


_NT16 = np.full(256, 15, dtype=np.uint8)
for _c, _v in zip(b"=ACMGRSVTWYHKDBN", range(16)):
    _NT16[_c] = _v
_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b


def to_bam(rs: ReadSet, path: str, *, genome_len: int = 0, rev_every: int = 0, with_oq: bool = False,
           seq_source: np.ndarray | None = None, block: int = 0xFF00, first_name: bytes | None = None) -> None:
    """Write the read set as an unaligned BAM (BGZF blocks + EOF block): `RG:Z:rg<id>` on every record, flags
    0x1|0x40 / 0x1|0x80 by `second`, header `@SQ LN:` = genome_len (kbbq.cc:196-216 sums them).

    Every `rev_every`-th record is stored as a reverse-strand record (flag 0x10): its bases are written reverse
    complemented and its qualities reversed, so that what the reference reads back in forward orientation
    (readutils.hh:30-42, readutils.cc:36-39) is the read set's own read.  `with_oq` adds an `OQ:Z:` tag holding
    qual+1 (capped at 93) so --use-oq visibly differs from the QUAL field.  `first_name` names record 0 (a long
    name makes the reader's record buffer big enough that appending an OQ tag later never reallocates it: the
    reference keeps a pointer into that buffer across bam_aux_update_str, htsiter.cc:12-29).
    """
    import struct
    import zlib

    seq = rs.seq if seq_source is None else seq_source
    text = b"@HD\tVN:1.6\tSO:unsorted\n"
    if genome_len:
        text += b"@SQ\tSN:chrSynth\tLN:%d\n" % genome_len
    for g in range(rs.n_rg):
        text += b"@RG\tID:rg%d\tPU:rg%d\n" % (g, g)
    raw = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text)
    if genome_len:
        raw += struct.pack("<i", 1) + struct.pack("<i", 9) + b"chrSynth\0" + struct.pack("<i", genome_len)
    else:
        raw += struct.pack("<i", 0)
    for i in range(rs.n):
        b0, b1 = int(rs.off[i]), int(rs.off[i + 1])
        L = b1 - b0
        s = seq[b0:b1]
        q = rs.qual[b0:b1]
        flag = 0x1 | 0x4 | (0x80 if rs.second[i] else 0x40)
        if rev_every and i % rev_every == rev_every - 1:
            flag |= 0x10
            s = _COMP[s[::-1]]
            q = q[::-1]
        nib = _NT16[s]
        if L & 1:
            nib = np.concatenate([nib, np.zeros(1, np.uint8)])
        packed = ((nib[0::2] << 4) | nib[1::2]).astype(np.uint8).tobytes()
        name = (first_name if (i == 0 and first_name) else b"r%d" % i) + b"\0"
        aux = b""
        if with_oq:
            aux += b"OQZ" + (np.minimum(q.astype(np.int32) + 1, 93).astype(np.uint8) + 33).tobytes() + b"\0"
        aux += b"RGZrg%d\0" % int(rs.rg[i])
        core = struct.pack("<iiBBHHHiiii", -1, -1, len(name), 0, 4680, 0, flag, L, -1, -1, 0)
        body = core + name + packed + q.tobytes() + aux
        raw += struct.pack("<i", len(body)) + body
    with open(path, "wb") as fh:
        for p in range(0, len(raw), block):
            chunk = bytes(raw[p : p + block])
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            comp = co.compress(chunk) + co.flush()
            bsize = 18 + len(comp) + 8
            fh.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize - 1) + comp
                     + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
        fh.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
