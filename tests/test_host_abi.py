"""CPU-side checks of the product: the C-ABI library loads and exports what include/kbbq_b200.h declares, its
host-only helpers agree with the oracle / golden vectors, and the kernels' shared __host__ __device__ logic
(compiled for the host by tests/emu) reproduces the oracle's error calls.  No GPU compute is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import common
from kbbq_b200 import api, synth
from oracle import port as P

KAT = common.load_kat()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "kbbq_b200.h")).read()
    names = sorted(set(re.findall(r"\b(kbbq_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 40
    L = api.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_gpu_means_loud_failure():
    if api.lib().kbbq_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.KbbqError) as e:
        api.Context(32, 1, 100)
    assert e.value.code == -3 and "no CPU path" in str(e.value)


def test_geometry_salts_patterns_known_answers():
    for e in KAT["bloom"]:
        g = api.bloom_geometry(e["projected"], e["fpr"])
        assert (g.table_bits, g.n_blocks, g.n_salts, g.random_seed, g.optimal_k, g.optimal_table_size) == (
            e["table_bits"], e["n_blocks"], e["n_salts"], int(e["random_seed"]), e["optimal_k"], e["optimal_table_size"])
        assert list(g.salts[: g.n_salts]) == e["salts"]
        pat = api.bloom_patterns(g)
        assert [str(int(x)) for x in pat[:8]] == e["pattern0"]
        assert [str(int(x)) for x in pat[36456 * 8:36456 * 8 + 8]] == e["pattern36456"]
        assert str(int(np.bitwise_xor.reduce(pat))) == e["pattern_xor"]


def test_invalid_bloom_parameters_are_rejected():
    for proj, fpr, seed in [(0, 0.01, api.DEFAULT_BLOOM_SEED), (10, -0.1, api.DEFAULT_BLOOM_SEED), (10, 0.01, 0),
                            (10, float("inf"), api.DEFAULT_BLOOM_SEED), (10, 0.01, 0xFFFFFFFFFFFFFFFF)]:
        with pytest.raises(api.KbbqError) as e:
            api.bloom_geometry(proj, fpr, seed)
        assert e.value.code == -1


def test_rng_seeding_and_jump_ahead():
    assert [str(int(x)) for x in api.rng_state(12345)] == KAT["rng"]["state"]
    assert [str(int(x)) for x in api.rng_state(2854027472)] == KAT["rng"]["seed_2854027472_state"]
    for skip in (1, 255, 1024, 99999, 1000000):
        want = np.zeros(4, dtype=np.uint64)
        st, _ = P.rng_outputs(777, 0, 0)
        s = st.copy()
        for _ in range(skip):
            P.lib().ko_rng_next.argtypes = [C.c_void_p]
            P.lib().ko_rng_next(s.ctypes.data_as(C.c_void_p))
        assert np.array_equal(api.rng_state(777, skip), s), skip
    # additivity far beyond what can be stepped
    a = api.rng_state(5, 2**40 + 12345)
    assert not np.array_equal(a, api.rng_state(5, 2**40 + 12344))


def test_thresholds_and_fpp_match_port():
    for e in KAT["thresholds"].items():
        pass
    for proj, ins, alpha in [(1400000, 483000, np.longdouble(7) / 20), (1400000, 100000, np.longdouble(7) / 100),
                             (32200000, 9000000, np.longdouble(7) / 50), (21000, 14556, np.longdouble(7) / 20)]:
        pb = P.PortBloom(proj, 0.01)
        pb.set_words(pb.words(), ins)
        fpr = api.effective_fpp(pb.table_bits, ins, pb.n_salts)
        assert fpr == pb.fprate()
        for k in (21, 31, 32):
            want = np.zeros(k + 1, dtype=np.int32)
            P.lib().ko_thresholds_from_filter(pb.p, alpha, np.longdouble(1), k, want)
            assert np.array_equal(api.thresholds_from_fpr(fpr, alpha, k), want)


@pytest.mark.parametrize("name", common.PIPES)
def test_model_fit_matches_golden(name):
    _, g = common.load_pipe(name)
    got = api.fit_model_host(g["rgc"], g["qc"], g["cyc"], g["dic"])
    for a, key in zip(got, ("meanq", "rgdq", "qdq", "cydq", "didq")):
        assert np.array_equal(a, g[key]), key


def test_pack_roundtrip_and_limits():
    rs = synth.make_reads(5000, 10, 97, seed=3, ragged=True, min_len=1, n_frac=0.01)
    pb = api.PackedBatch(rs)
    codes = np.full(256, 4, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        codes[ch] = i
    pos = 0
    for r in range(0, rs.n, 7):
        b0, L, wo = int(rs.off[r]), int(rs.off[r + 1] - rs.off[r]), int(pb.word_off[r])
        want = codes[rs.seq[b0:b0 + L]]
        words = pb.bases[wo:wo + (L + 31) // 32]
        got = ((words[:, None] >> (2 * np.arange(32, dtype=np.uint64))[None, :]) & np.uint64(3)).reshape(-1)[:L]
        nm = np.unpackbits(pb.nmask[wo:wo + (L + 31) // 32].view(np.uint8), bitorder="little")[:L]
        assert np.array_equal(nm, (want == 4).astype(np.uint8))
        assert np.array_equal(got[want < 4].astype(np.uint8), want[want < 4])
        assert np.array_equal(pb.qual[wo * 32: wo * 32 + L], rs.qual[b0:b0 + L])
    flags = (np.arange(rs.total_bases) % 3 == 0).astype(np.uint8)
    assert np.array_equal(pb.unpack_bits(pb.pack_bits(flags), rs.total_bases), flags)
    long = synth.make_reads(5000, 1, 600, seed=1, n_reads=2)  # beyond 512 bases: packed like any other read
    assert api.PackedBatch(long).max_len == 600
    too_long = synth.make_reads(70000, 1, 65536, seed=1, n_reads=1)
    with pytest.raises(api.KbbqError) as e:
        api.PackedBatch(too_long)
    assert e.value.code == -5  # KBBQ_ERANGE: longer than KBBQ_ABI_MAX_READ_LEN
    empty = rs.subset(0, 0)
    assert api.PackedBatch(empty).n_words == 0


# ---- the kernels' shared logic, compiled for the host (tests/emu) ----------------------------------------------------
def _emu():
    E = C.CDLL(os.path.join(ROOT, "tests", "emu", "libkbbq_emu.so"))
    vp = C.c_void_p
    E.emu_canonical_at.restype = C.c_uint64
    E.emu_canonical_at.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    E.emu_fastmod.restype = C.c_uint32
    E.emu_fastmod.argtypes = [C.c_uint32, C.c_uint64]
    E.emu_get_errors.restype = C.c_uint64
    E.emu_get_errors.argtypes = [C.c_uint32] + [vp] * 5 + [C.c_int, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, vp]
    E.emu_pass2.argtypes = [C.c_uint32] + [vp] * 5 + [C.c_int, vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp,
                                                      C.c_uint64, C.c_uint32, C.c_uint32, vp, vp]
    return E


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def test_fastmod_is_exact():
    E = _emu()
    rng = np.random.default_rng(5)
    for n in [1, 2, 3, 263, 131154, 216294, 670000001, 4294967295]:
        for h in list(rng.integers(0, 2**32, 300)) + [0, 2**32 - 1, n - 1, n % 2**32]:
            assert E.emu_fastmod(int(h), n) == int(h) % n


def test_canonical_kmers_with_substitution():
    E = _emu()
    rng = np.random.default_rng(7)
    rs = synth.make_reads(20000, 5, 150, seed=9, n_frac=0.01)
    pb = api.PackedBatch(rs)
    for k in (1, 2, 5, 16, 21, 31, 32):
        for r in rng.integers(0, rs.n, 60):
            b0, L, wo = int(rs.off[r]), 150, int(pb.word_off[r])
            j = int(rng.integers(0, L - k + 1))
            p, c = int(rng.integers(0, L)), int(rng.integers(0, 4))
            s = bytearray(rs.seq[b0 + j:b0 + j + k].tobytes())
            if j <= p < j + k:
                s[p - j] = b"ACGT"[c]
            want = P.kmer_of(k, bytes(s))
            valid = C.c_int(0)
            key = E.emu_canonical_at(_ptr(pb.bases[wo:]), _ptr(pb.nmask[wo:]), L, j, k, p, c, C.byref(valid))
            assert bool(valid.value) == want["valid"]
            if want["valid"]:
                assert key == want["get"]


@pytest.mark.parametrize("k,cov,L,kw", [
    (32, 20, 100, dict(n_frac=0.002, n_rg=2)),
    (21, 20, 100, dict(err_scale=4.0)),
    (31, 14, 150, dict(ragged=True, min_len=62, n_frac=0.003)),
    (32, 20, 64, dict(err_scale=3.0)),       # len == 2k: shortest reads every k-mer helper still accepts
    (16, 30, 250, dict(err_scale=6.0)),      # long, error-rich reads: recursion on bad prefixes/suffixes
    (32, 12, 700, dict(err_scale=3.0, ragged=True, min_len=300, n_frac=0.001)),  # beyond 512 bases: the dynamically sized scratch
    (25, 16, 3000, dict(err_scale=0.7, ragged=True, min_len=100)),               # ... and ~94 words per read
])
def test_emulated_kernel_logic_matches_port(k, cov, L, kw):
    E = _emu()
    G = 40000
    rs = synth.make_reads(G, cov, L, seed=50 + k, **kw)
    pb = api.PackedBatch(rs)
    pp = P.PortPipeline(rs, k)
    approx = G * 7
    ps, pt = P.PortBloom(approx, 0.01), P.PortBloom(approx, 0.0005)
    alpha = np.longdouble(7) / np.longdouble(cov)
    pp.subsample(ps, float(alpha), 4242)
    th = pp.thresholds(ps, 7.0, float(cov))
    e2 = pp.infer_errors(ps, th)
    pp.find_trusted(pt, ps, th)
    sw, spat, tpat = ps.words(), ps.patterns(), pt.patterns()
    tw = np.zeros(pt.table_bits // 64, dtype=np.uint64)
    tins = np.zeros(1, dtype=np.uint64)
    eb = np.zeros(pb.n_words, dtype=np.uint32)
    E.emu_pass2(rs.n, _ptr(pb.bases), _ptr(pb.nmask), _ptr(pb.qual), _ptr(pb.word_off), _ptr(pb.len), k, _ptr(th),
                _ptr(sw), _ptr(spat), ps.n_blocks, int(ps.salts[0]), int(ps.salts[1]), _ptr(tw), _ptr(tpat), pt.n_blocks,
                int(pt.salts[0]), int(pt.salts[1]), _ptr(tins), _ptr(eb))
    assert np.array_equal(pb.unpack_bits(eb, rs.total_bases), e2)
    assert int(tins[0]) == pt.inserted and np.array_equal(tw, pt.words())
    e3, seq_after = pp.get_errors(pt, want_seq=True)
    eb3 = np.zeros(pb.n_words, dtype=np.uint32)
    mi = np.zeros(rs.n, dtype=np.int32)
    mb = np.zeros(rs.n, dtype=np.int32)
    E.emu_get_errors(rs.n, _ptr(pb.bases), _ptr(pb.nmask), _ptr(pb.qual), _ptr(pb.word_off), _ptr(pb.len), k, _ptr(tw),
                     _ptr(tpat), pt.n_blocks, int(pt.salts[0]), int(pt.salts[1]), _ptr(eb3), _ptr(mi), _ptr(mb))
    got = pb.unpack_bits(eb3, rs.total_bases)
    assert int(np.count_nonzero(got != e3)) == 0
    # the early-return quirk: reads whose seq differs after get_errors are exactly those the kernel logic reports
    changed = np.unique(np.searchsorted(rs.off, np.flatnonzero(seq_after != rs.seq), side="right") - 1)
    assert np.array_equal(changed, np.flatnonzero(mi >= 0))


def test_quality_masks_helper():
    """kbbq_quality_masks: per packed word, bit i of qmask[2w] is qual <= 2 and of qmask[2w+1] is qual <= 6."""
    rs = synth.make_reads(3000, 10, 77, seed=2, ragged=True, min_len=40)
    rs.qual[:] = np.random.default_rng(0).integers(0, 12, rs.qual.shape[0], dtype=np.uint8)
    for threads in (1, 4):
        pb = api.PackedBatch(rs, n_threads=threads)
        q = pb.qual.reshape(-1, 32)
        want2 = (q <= 2).astype(np.uint64) @ (1 << np.arange(32, dtype=np.uint64))
        want6 = (q <= 6).astype(np.uint64) @ (1 << np.arange(32, dtype=np.uint64))
        got = pb.qmask.reshape(-1, 2)
        assert np.array_equal(got[:, 0].astype(np.uint64), want2) and np.array_equal(got[:, 1].astype(np.uint64), want6)
    assert pb.as_struct().qmask is not None


def test_device_warmup_fails_loudly_without_a_gpu():
    L = api.lib()
    if L.kbbq_device_count() > 0:
        assert L.kbbq_device_warmup(0) == 0
    else:
        assert L.kbbq_device_warmup(0) == -3 and b"no CPU path" in L.kbbq_last_error()


def test_context_arguments_are_checked_before_any_device_is_touched():
    """kbbq_ctx_create: k, read groups and the longest read are validated first -- also that the tables the kernels index
    with 32-bit offsets stay below 2^31 entries (n_rg x 94 x 2 x max_cycle)."""
    import ctypes as C
    L = api.lib()
    ctx = C.c_void_p()
    assert L.kbbq_ctx_create(0, 33, 1, 100, C.byref(ctx)) == -1 and b"k must be" in L.kbbq_last_error()
    assert L.kbbq_ctx_create(0, 32, 0, 100, C.byref(ctx)) == -1
    assert L.kbbq_ctx_create(0, 32, 1, 65536, C.byref(ctx)) == -1  # longer than KBBQ_ABI_MAX_READ_LEN
    assert L.kbbq_ctx_create(0, 32, 400, 65535, C.byref(ctx)) == -5 and b"too large" in L.kbbq_last_error()


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/kbbq_b200.h must compile as C99 on its own, and a C program must link against
    the library and get the loud no-GPU failure (or a context) through it."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "kbbq_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "kbbq_b200.h"\n'
                   'int main(void) { kbbq_ctx* c = 0; kbbq_bloom_geom g; int rc;\n'
                   '  if (kbbq_bloom_geometry(7000000, 0.01, KBBQ_DEFAULT_BLOOM_SEED, &g)) return 2;\n'
                   '  printf("%llu %u\\n", (unsigned long long)g.table_bits, g.n_salts);\n'
                   '  rc = kbbq_ctx_create(0, 32, 1, 100, &c);\n'
                   '  if (rc == KBBQ_OK) kbbq_ctx_destroy(c); else printf("%d %s\\n", rc, kbbq_last_error());\n'
                   '  return 0; }\n')
    exe = tmp_path / "t"
    lib_dir = os.path.join(ROOT, "kbbq_b200")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src), "-L", lib_dir,
                           "-lkbbq_b200", "-Wl,-rpath," + lib_dir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0] == "67150848 7"   # Bloom(7000000, 0.01)
    if api.lib().kbbq_device_count() == 0:
        assert out[1].startswith("-3 ") and "no CPU path" in out[1]
