"""ctypes binding of include/kbbq_b200.h (kbbq_b200/libkbbq_b200.so).

Python is plumbing here: tests and bench.py drive the C ABI through this module exactly as a C++ host would.
There is no CPU fallback: if the library is missing, or there is no GPU, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KBBQ_B200_LIB", os.path.join(_HERE, "libkbbq_b200.so"))  # override: kernel experiments

FILTER_SAMPLED = 0
FILTER_TRUSTED = 1
DEFAULT_BLOOM_SEED = 0xA5A5A5A55A5A5A5A
MAX_READ_LEN = 65535
NQ = 94


class KbbqError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"kbbq_b200 error {code}: {msg}")
        self.code = code


class BloomGeom(C.Structure):
    _fields_ = [
        ("projected", C.c_uint64), ("fpr", C.c_double), ("seed", C.c_uint64), ("optimal_k", C.c_uint32),
        ("optimal_table_size", C.c_uint64), ("random_seed", C.c_uint64), ("table_bits", C.c_uint64),
        ("n_blocks", C.c_uint64), ("n_salts", C.c_uint32), ("salts", C.c_uint32 * 128),
    ]


class _Batch(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint32), ("max_len", C.c_uint32), ("n_words", C.c_uint64), ("bases", C.c_void_p),
        ("nmask", C.c_void_p), ("qual", C.c_void_p), ("word_off", C.c_void_p), ("len", C.c_void_p),
        ("rg", C.c_void_p), ("second", C.c_void_p), ("qmask", C.c_void_p),
    ]


class SynthParams(C.Structure):
    """kbbq_synth_params: synthetic reads as a pure function of (seed, read index) -- measurement tooling"""
    _fields_ = [("seed", C.c_uint64), ("genome_seed", C.c_uint64), ("genome_len", C.c_uint64), ("read_len", C.c_uint32),
                ("n_rg", C.c_uint32), ("err_scale", C.c_double)]


_lib = None


def lib():
    """Load the C-ABI library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C kbbq_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, u64, u32, i32, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_double
        L.kbbq_last_error.restype = C.c_char_p
        L.kbbq_device_count.restype = i32
        L.kbbq_device_warmup.argtypes = [i32]
        L.kbbq_bloom_geometry.argtypes = [u64, dbl, u64, C.POINTER(BloomGeom)]
        L.kbbq_bloom_patterns.argtypes = [C.POINTER(BloomGeom), vp]
        L.kbbq_effective_fpp.restype = dbl
        L.kbbq_effective_fpp.argtypes = [u64, u64, u32]
        L.kbbq_thresholds_from_fpr.argtypes = [dbl, C.c_longdouble, i32, vp]
        L.kbbq_rng_state.argtypes = [u32, u64, vp]
        L.kbbq_phit.argtypes = [dbl, C.c_longdouble, C.POINTER(C.c_longdouble)]
        L.kbbq_thresholds_from_p.argtypes = [i32, C.c_longdouble, vp]
        L.kbbq_default_seed.restype = u32
        L.kbbq_default_seed.argtypes = [u64, u32]
        L.kbbq_fit_model_host.argtypes = [i32, i32] + [vp] * 9
        L.kbbq_pack_size.argtypes = [u64, vp, C.POINTER(u64), C.POINTER(u32)]
        L.kbbq_pack_reads.argtypes = [u64] + [vp] * 8 + [i32]
        L.kbbq_quality_masks.argtypes = [u64, vp, vp, i32]
        L.kbbq_ctx_create.argtypes = [i32, i32, i32, i32, C.POINTER(vp)]
        L.kbbq_ctx_destroy.argtypes = [vp]
        L.kbbq_sync.argtypes = [vp]
        L.kbbq_stream.restype = vp
        L.kbbq_stream.argtypes = [vp]
        L.kbbq_last_pass_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.kbbq_last_pass_kernels.argtypes = [vp, i32, vp, vp]
        L.kbbq_launch_count.restype = u64
        L.kbbq_launch_count.argtypes = [vp]
        L.kbbq_bloom_create.argtypes = [vp, i32, u64, dbl, u64]
        L.kbbq_bloom_info.argtypes = [vp, i32, C.POINTER(BloomGeom)]
        L.kbbq_bloom_clear.argtypes = [vp, i32]
        L.kbbq_bloom_inserted.argtypes = [vp, i32, C.POINTER(u64)]
        L.kbbq_bloom_download.argtypes = [vp, i32, vp, C.POINTER(u64)]
        L.kbbq_bloom_upload.argtypes = [vp, i32, vp, u64]
        L.kbbq_bloom_insert_keys.argtypes = [vp, i32, vp, u64]
        L.kbbq_bloom_query_keys.argtypes = [vp, i32, vp, u64, vp]
        L.kbbq_bloom_fprate.argtypes = [vp, i32, C.POINTER(dbl)]
        L.kbbq_bloom_device_words.restype = vp
        L.kbbq_bloom_device_words.argtypes = [vp, i32]
        L.kbbq_bloom_device_counter.restype = vp
        L.kbbq_bloom_device_counter.argtypes = [vp, i32]
        L.kbbq_bloom_or_from.argtypes = [vp, i32, vp, u64, u64]
        L.kbbq_bloom_save.argtypes = [vp, i32, C.c_char_p]
        L.kbbq_bloom_load.argtypes = [vp, i32, C.c_char_p]
        L.kbbq_bloom_ipc_handle.argtypes = [vp, i32, vp]
        L.kbbq_bloom_or_allreduce_p2p.argtypes = [vp, i32, i32, i32, vp]
        L.kbbq_bloom_mc_supported.argtypes = [vp, i32]
        L.kbbq_bloom_mc_create.argtypes = [vp, i32, i32, C.POINTER(i32)]
        L.kbbq_bloom_mc_import.argtypes = [vp, i32, i32, i32]
        L.kbbq_bloom_mc_add_device.argtypes = [vp, i32]
        L.kbbq_bloom_mc_bind.argtypes = [vp, i32]
        L.kbbq_bloom_or_allreduce_mc.argtypes = [vp, i32, i32, i32]
        L.kbbq_batch_upload.argtypes = [vp, C.POINTER(_Batch), C.POINTER(vp)]
        L.kbbq_batch_upload_async.argtypes = [vp, C.POINTER(_Batch), C.POINTER(vp)]
        L.kbbq_batch_upload_stream.argtypes = [vp, C.POINTER(_Batch), C.POINTER(vp)]
        L.kbbq_batch_wait.argtypes = [vp, vp]
        L.kbbq_batch_free.argtypes = [vp, vp]
        L.kbbq_batch_slots.restype = u64
        L.kbbq_batch_slots.argtypes = [vp]
        L.kbbq_sampler_seed.argtypes = [vp, u32, dbl]
        L.kbbq_sampler_seek.argtypes = [vp, u64]
        L.kbbq_pass1_subsample.argtypes = [vp, vp]
        L.kbbq_set_thresholds.argtypes = [vp, vp]
        L.kbbq_compute_thresholds.argtypes = [vp, C.c_longdouble, vp]
        L.kbbq_pass2_find_trusted.argtypes = [vp, vp, vp]
        L.kbbq_pass3_covariates.argtypes = [vp, vp, vp]
        L.kbbq_consume_errors.argtypes = [vp, vp, vp]
        L.kbbq_consume_errors_n.argtypes = [vp, vp, vp, u32]
        L.kbbq_covariates_reset.argtypes = [vp]
        L.kbbq_covariates_download.argtypes = [vp, vp, vp, vp, vp]
        L.kbbq_covariates_device.restype = vp
        L.kbbq_covariates_device.argtypes = [vp, C.POINTER(u64)]
        L.kbbq_fit_model.argtypes = [vp] + [vp] * 5
        L.kbbq_set_model.argtypes = [vp] + [vp] * 5
        L.kbbq_pass4_recalibrate.argtypes = [vp, vp, vp]
        L.kbbq_batch_device_newqual.restype = vp
        L.kbbq_batch_device_newqual.argtypes = [vp]
        L.kbbq_covariates_report.argtypes = [vp, C.c_char_p, vp]
        L.kbbq_report_tables_host.argtypes = [C.c_char_p, i32, i32, vp, vp, vp, vp, vp]
        L.kbbq_fixed_errors.argtypes = [vp, vp, vp, u32, vp]
        L.kbbq_batch_total_bases.argtypes = [vp, vp, C.POINTER(u64)]
        L.kbbq_bloom_digest.argtypes = [vp, i32, vp]
        L.kbbq_covariates_digest.argtypes = [vp, vp]
        L.kbbq_batch_newqual_digest.argtypes = [vp, vp, vp]
        L.kbbq_synth_batch.argtypes = [vp, C.POINTER(SynthParams), u64, u32, C.POINTER(vp)]
        L.kbbq_synth_host.argtypes = [C.POINTER(SynthParams), u64, u64, vp, vp, vp, vp, i32]
        L.kbbq_batch_extent.argtypes = [vp, C.POINTER(u32), C.POINTER(u64), C.POINTER(u32)]
        L.kbbq_batch_download.argtypes = [vp, vp] + [vp] * 8
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise KbbqError(rc, lib().kbbq_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- host-only helpers -------------------------------------------------------------------------------------------
def bloom_geometry(projected: int, fpr: float, seed: int = DEFAULT_BLOOM_SEED) -> BloomGeom:
    g = BloomGeom()
    _check(lib().kbbq_bloom_geometry(projected, fpr, seed, C.byref(g)))
    return g


def bloom_patterns(g: BloomGeom) -> np.ndarray:
    out = np.zeros(65536 * 8, dtype=np.uint64)
    _check(lib().kbbq_bloom_patterns(C.byref(g), _p(out)))
    return out


def effective_fpp(table_bits: int, inserted: int, n_salts: int) -> float:
    return float(lib().kbbq_effective_fpp(table_bits, inserted, n_salts))


def thresholds_from_fpr(fpr: float, alpha, k: int) -> np.ndarray:
    out = np.zeros(k + 1, dtype=np.int32)
    _check(lib().kbbq_thresholds_from_fpr(fpr, C.c_longdouble(alpha), k, _p(out)))
    return out


def rng_state(seed: int, skip: int = 0) -> np.ndarray:
    st = np.zeros(4, dtype=np.uint64)
    _check(lib().kbbq_rng_state(seed, skip, _p(st)))
    return st


def fit_model_host(rgc, qc, cyc, dic):
    R, ml = rgc.shape[0], cyc.shape[3]
    meanq = np.zeros(R, dtype=np.int32)
    rgdq = np.zeros(R, dtype=np.int32)
    qdq = np.zeros((R, NQ), dtype=np.int32)
    cydq = np.zeros((R, NQ, 2, ml), dtype=np.int32)
    didq = np.zeros((R, NQ, 16), dtype=np.int32)
    arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in (rgc, qc, cyc, dic)]
    _check(lib().kbbq_fit_model_host(R, ml, *[_p(a) for a in arrs], _p(meanq), _p(rgdq), _p(qdq), _p(cydq), _p(didq)))
    return meanq, rgdq, qdq, cydq, didq


def digest_words(words: np.ndarray) -> tuple:
    """host restatement of the library's digest kernel (numpy): {sum, xor} of mix(word ^ mix(index)) over non-zero words"""
    w = np.ascontiguousarray(words).view(np.uint64).reshape(-1)

    def mix(x):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))

    s, x = np.uint64(0), np.uint64(0)
    with np.errstate(over="ignore"):
        for lo in range(0, w.size, 1 << 24):
            v = w[lo : lo + (1 << 24)]
            nz = np.flatnonzero(v)
            if nz.size == 0:
                continue
            h = mix(v[nz] ^ mix((nz + lo).astype(np.uint64)))
            s = s + h.sum(dtype=np.uint64)
            x = x ^ np.bitwise_xor.reduce(h)
    return int(s), int(x)


def report_tables(path: str, rgc, qc, cyc, dic, rg_names=None) -> None:
    """GATK-style report of dense count tables (host only)"""
    R, ml = rgc.shape[0], cyc.shape[3]
    arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in (rgc, qc, cyc, dic)]
    names = None
    if rg_names is not None:
        names = (C.c_char_p * R)(*[n.encode() for n in rg_names])
    _check(lib().kbbq_report_tables_host(os.fsencode(path), R, ml, *[_p(a) for a in arrs], names))


def synth_host(params: SynthParams, first_read: int, n_reads: int, n_threads: int = 8):
    """the reads kbbq_synth_batch makes, as a synth.ReadSet (ASCII arrays for the oracle)"""
    from . import synth

    L = int(params.read_len)
    seq = np.zeros(n_reads * L, dtype=np.uint8)
    qual = np.zeros(n_reads * L, dtype=np.uint8)
    rg = np.zeros(max(n_reads, 1), dtype=np.int32)
    second = np.zeros(max(n_reads, 1), dtype=np.uint8)
    _check(lib().kbbq_synth_host(C.byref(params), first_read, n_reads, _p(seq), _p(qual), _p(rg), _p(second), n_threads))
    off = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(L)).astype(np.uint64)
    return synth.ReadSet(seq=seq, off=off, qual=qual, rg=rg[:n_reads], second=second[:n_reads], n_rg=int(params.n_rg))


class PackedBatch:
    """Host-side packed read batch (the kbbq_batch layout)."""

    def _alloc_arrays(self, pinned: bool, qmask: bool):
        self._keep = []  # pinned torch tensors backing the numpy views
        n = self.n_reads

        def alloc(count, dt):
            if not pinned:
                return np.zeros(count, dtype=dt)
            import torch

            t = torch.zeros(int(count) * np.dtype(dt).itemsize, dtype=torch.uint8).pin_memory()
            self._keep.append(t)
            return t.numpy().view(dt)

        self._alloc = alloc
        self.bases = alloc(max(self.n_words, 1), np.uint64)
        self.nmask = alloc(max(self.n_words, 1), np.uint32)
        self.qual = alloc(max(self.n_words, 1) * 32, np.uint8)
        # the per-read arrays are pinned too: a pageable source turns cudaMemcpyAsync into a staged, blocking copy
        self.word_off = alloc(n + 1, np.uint32)
        self.len = alloc(max(n, 1), np.uint32)
        self.rg = alloc(max(n, 1), np.uint16)
        self.second = alloc(max(n, 1), np.uint8)
        self.qmask = alloc(max(self.n_words, 1) * 2, np.uint32) if qmask else None

    @classmethod
    def empty(cls, n_reads: int, n_words: int, max_len: int, pinned: bool = False, qmask: bool = True) -> "PackedBatch":
        """zeroed arrays of the given extent (the destination of Context.download)"""
        self = cls.__new__(cls)
        self.n_reads, self.n_words, self.max_len = n_reads, n_words, max_len
        self._alloc_arrays(pinned, qmask)
        return self

    def __init__(self, rs, n_threads: int = 8, pinned: bool = False, qmask: bool = True):
        n = rs.n
        nw = C.c_uint64(0)
        ml = C.c_uint32(0)
        off = np.ascontiguousarray(rs.off, dtype=np.uint64)
        _check(lib().kbbq_pack_size(n, _p(off), C.byref(nw), C.byref(ml)))
        self.n_reads, self.n_words, self.max_len = n, int(nw.value), int(ml.value)
        self._alloc_arrays(pinned, False)
        alloc = self._alloc
        if n:
            self.rg[:] = rs.rg
            self.second[:] = rs.second
        seq = np.ascontiguousarray(rs.seq, dtype=np.uint8)
        qual = np.ascontiguousarray(rs.qual, dtype=np.uint8)
        _check(lib().kbbq_pack_reads(n, _p(seq), _p(off), _p(qual), _p(self.bases), _p(self.nmask), _p(self.qual),
                                     _p(self.word_off), _p(self.len), n_threads))
        # qual <= 2 / qual <= 6 bit masks: with them pass 2 does not wait for the quality array (qmask=False: the
        # library derives them on the device)
        self.qmask = None
        if qmask and not os.environ.get("KBBQ_NO_HOST_QMASK"):  # (the variable: tests of the device-derived masks)
            self.qmask = alloc(max(self.n_words, 1) * 2, np.uint32)
            _check(lib().kbbq_quality_masks(self.n_words, _p(self.qual), _p(self.qmask), n_threads))

    def as_struct(self) -> _Batch:
        b = _Batch()
        b.n_reads, b.max_len, b.n_words = self.n_reads, self.max_len, self.n_words
        b.bases, b.nmask, b.qual = _p(self.bases), (_p(self.nmask) if self.has_n else None), _p(self.qual)
        b.word_off, b.len, b.rg, b.second = _p(self.word_off), _p(self.len), _p(self.rg), _p(self.second)
        b.qmask = _p(self.qmask)
        return b

    @property
    def has_n(self) -> bool:
        """any non-ACGT base in the batch?  A batch without one sends no N masks (kbbq_batch::nmask may be NULL)."""
        if getattr(self, "_has_n", None) is None:
            self._has_n = bool(self.nmask[: max(self.n_words, 1)].any())
        return self._has_n

    @property
    def h2d_bytes(self) -> int:
        return (self.n_words * (8 + (4 if self.has_n else 0) + 32 + (8 if self.qmask is not None else 0)) + (self.n_reads + 1) * 4
                + self.n_reads * (4 + 2 + 1))

    def unpack_bits(self, bits: np.ndarray, total_bases: int) -> np.ndarray:
        """bit masks in word layout -> one byte per base, reads concatenated (the oracle's layout)."""
        out = np.zeros(total_bases, dtype=np.uint8)
        expanded = np.unpackbits(bits.view(np.uint8), bitorder="little").reshape(-1, 32)
        pos = 0
        lens = self.len[: self.n_reads].astype(np.int64)
        wo = self.word_off.astype(np.int64)
        if self.n_reads and np.all(lens == lens[0]) and np.all(np.diff(wo) == wo[1] - wo[0]):
            L, W = int(lens[0]), int(wo[1] - wo[0])
            return np.ascontiguousarray(expanded.reshape(self.n_reads, W * 32)[:, :L]).reshape(-1)
        for r in range(self.n_reads):
            L = int(lens[r])
            out[pos : pos + L] = expanded[wo[r] : wo[r + 1]].reshape(-1)[:L]
            pos += L
        return out

    def pack_bits(self, flags: np.ndarray) -> np.ndarray:
        """one byte per base (oracle layout) -> bit masks in word layout."""
        tmp = np.zeros((max(self.n_words, 1), 32), dtype=np.uint8)
        lens = self.len[: self.n_reads].astype(np.int64)
        wo = self.word_off.astype(np.int64)
        pos = 0
        for r in range(self.n_reads):
            L = int(lens[r])
            tmp[wo[r] : wo[r + 1]].reshape(-1)[:L] = flags[pos : pos + L] != 0
            pos += L
        return np.packbits(tmp, axis=1, bitorder="little").view(np.uint32).reshape(-1)

    def unpack_qual(self, q: np.ndarray, total_bases: int) -> np.ndarray:
        lens = self.len[: self.n_reads].astype(np.int64)
        wo = self.word_off.astype(np.int64)
        if self.n_reads and np.all(lens == lens[0]) and np.all(np.diff(wo) == wo[1] - wo[0]):
            L, W = int(lens[0]), int(wo[1] - wo[0])
            return np.ascontiguousarray(q.reshape(self.n_reads, W * 32)[:, :L]).reshape(-1)
        out = np.zeros(total_bases, dtype=np.uint8)
        pos = 0
        for r in range(self.n_reads):
            L = int(lens[r])
            out[pos : pos + L] = q[wo[r] * 32 : wo[r] * 32 + L]
            pos += L
        return out


# ---- GPU context -------------------------------------------------------------------------------------------------
class Context:
    """One GPU: filters, count tables, model tables (kbbq_ctx)."""

    def __init__(self, k: int, n_rg: int, max_cycle: int, device: int = 0):
        self.L = lib()
        h = C.c_void_p()
        _check(self.L.kbbq_ctx_create(device, k, n_rg, max_cycle, C.byref(h)))
        self.h = h
        self.k, self.n_rg, self.max_cycle, self.device = k, n_rg, max_cycle, device

    def close(self):
        if self.h:
            self.L.kbbq_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _check(self.L.kbbq_sync(self.h))

    @property
    def stream(self) -> int:
        return int(self.L.kbbq_stream(self.h) or 0)

    def last_pass_ms(self) -> float:
        ms = C.c_float(0)
        _check(self.L.kbbq_last_pass_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def last_pass_kernels(self) -> dict:
        """{kernel name: device ms} of the most recent pass call (events between the launches)."""
        names = (C.c_char_p * 8)()
        ms = (C.c_float * 8)()
        n = self.L.kbbq_last_pass_kernels(self.h, 8, names, ms)
        if n < 0:
            _check(n)
        return {names[i].decode(): float(ms[i]) for i in range(n)}

    @property
    def launches(self) -> int:
        return int(self.L.kbbq_launch_count(self.h))

    # filters
    def bloom_create(self, which: int, projected: int, fpr: float, seed: int = DEFAULT_BLOOM_SEED) -> BloomGeom:
        _check(self.L.kbbq_bloom_create(self.h, which, projected, fpr, seed))
        return self.bloom_info(which)

    def bloom_info(self, which: int) -> BloomGeom:
        g = BloomGeom()
        _check(self.L.kbbq_bloom_info(self.h, which, C.byref(g)))
        return g

    def bloom_clear(self, which: int):
        _check(self.L.kbbq_bloom_clear(self.h, which))

    def bloom_inserted(self, which: int) -> int:
        v = C.c_uint64(0)
        _check(self.L.kbbq_bloom_inserted(self.h, which, C.byref(v)))
        return int(v.value)

    def bloom_words(self, which: int) -> np.ndarray:
        g = self.bloom_info(which)
        out = np.zeros(g.table_bits // 64, dtype=np.uint64)
        _check(self.L.kbbq_bloom_download(self.h, which, _p(out), None))
        return out

    def bloom_upload(self, which: int, words: np.ndarray, inserted: int):
        _check(self.L.kbbq_bloom_upload(self.h, which, _p(np.ascontiguousarray(words, dtype=np.uint64)), inserted))

    def bloom_insert_keys(self, which: int, keys: np.ndarray):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        _check(self.L.kbbq_bloom_insert_keys(self.h, which, _p(keys), keys.size))

    def bloom_query_keys(self, which: int, keys: np.ndarray) -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        hits = np.zeros(max(keys.size, 1), dtype=np.uint8)
        _check(self.L.kbbq_bloom_query_keys(self.h, which, _p(keys), keys.size, _p(hits)))
        return hits[: keys.size]

    def bloom_fprate(self, which: int) -> float:
        v = C.c_double(0)
        _check(self.L.kbbq_bloom_fprate(self.h, which, C.byref(v)))
        return float(v.value)

    def bloom_device_words(self, which: int) -> int:
        return int(self.L.kbbq_bloom_device_words(self.h, which) or 0)

    def bloom_device_counter(self, which: int) -> int:
        return int(self.L.kbbq_bloom_device_counter(self.h, which) or 0)

    def bloom_or_from(self, which: int, other_ptr: int, first_word: int, n_words: int):
        _check(self.L.kbbq_bloom_or_from(self.h, which, C.c_void_p(other_ptr), first_word, n_words))

    def bloom_digest(self, which: int) -> tuple:
        d = np.zeros(2, dtype=np.uint64)
        _check(self.L.kbbq_bloom_digest(self.h, which, _p(d)))
        return int(d[0]), int(d[1])

    def covariates_digest(self) -> tuple:
        d = np.zeros(2, dtype=np.uint64)
        _check(self.L.kbbq_covariates_digest(self.h, _p(d)))
        return int(d[0]), int(d[1])

    def newqual_digest(self, db: "DeviceBatch") -> tuple:
        d = np.zeros(2, dtype=np.uint64)
        _check(self.L.kbbq_batch_newqual_digest(self.h, db.h, _p(d)))
        return int(d[0]), int(d[1])

    def bloom_save(self, which: int, path: str):
        _check(self.L.kbbq_bloom_save(self.h, which, os.fsencode(path)))

    def bloom_load(self, which: int, path: str):
        _check(self.L.kbbq_bloom_load(self.h, which, os.fsencode(path)))

    def bloom_ipc_handle(self, which: int) -> np.ndarray:
        h = np.zeros(64, dtype=np.uint8)
        _check(self.L.kbbq_bloom_ipc_handle(self.h, which, _p(h)))
        return h

    def bloom_or_allreduce_p2p(self, which: int, rank: int, world: int, handles: np.ndarray):
        """handles: uint8[world, 64], every rank's kbbq_bloom_ipc_handle.  Barriers before and after are the caller's."""
        handles = np.ascontiguousarray(handles, dtype=np.uint8)
        assert handles.shape == (world, 64)
        _check(self.L.kbbq_bloom_or_allreduce_p2p(self.h, which, rank, world, _p(handles)))

    # the merge through NVSwitch multicast (kbbq_nvls.cuh); every step fails soft with KbbqError
    def bloom_mc_supported(self, which: int) -> bool:
        return bool(self.L.kbbq_bloom_mc_supported(self.h, which))

    def bloom_mc_create(self, which: int, world: int) -> int:
        fd = C.c_int(-1)
        _check(self.L.kbbq_bloom_mc_create(self.h, which, world, C.byref(fd)))
        return int(fd.value)

    def bloom_mc_import(self, which: int, world: int, fd: int):
        _check(self.L.kbbq_bloom_mc_import(self.h, which, world, fd))

    def bloom_mc_add_device(self, which: int):
        _check(self.L.kbbq_bloom_mc_add_device(self.h, which))

    def bloom_mc_bind(self, which: int):
        _check(self.L.kbbq_bloom_mc_bind(self.h, which))

    def bloom_or_allreduce_mc(self, which: int, rank: int, world: int):
        _check(self.L.kbbq_bloom_or_allreduce_mc(self.h, which, rank, world))

    # batches
    def upload(self, pb: PackedBatch) -> "DeviceBatch":
        """The qualities travel on the library's side stream while pass 1 runs (kbbq_batch_upload_async): the
        DeviceBatch keeps `pb` alive, and pb's arrays must not be modified before the next sync()."""
        s = pb.as_struct()
        h = C.c_void_p()
        _check(self.L.kbbq_batch_upload_async(self.h, C.byref(s), C.byref(h)))
        return DeviceBatch(self, h, pb)

    def upload_stream(self, pb: PackedBatch) -> "DeviceBatch":
        """Every array on the library's side stream, nothing waited for (kbbq_batch_upload_stream): the passes wait on the
        device, and the upload of the next batch overlaps the kernels of this one.  `pb` must be pinned."""
        s = pb.as_struct()
        h = C.c_void_p()
        _check(self.L.kbbq_batch_upload_stream(self.h, C.byref(s), C.byref(h)))
        return DeviceBatch(self, h, pb)

    def synth_batch(self, params: SynthParams, first_read: int, n_reads: int) -> "DeviceBatch":
        """reads [first_read, first_read + n_reads) of the synthetic stream, made on the device (measurement tooling)"""
        h = C.c_void_p()
        _check(self.L.kbbq_synth_batch(self.h, C.byref(params), first_read, n_reads, C.byref(h)))
        nr, nw, ml = C.c_uint32(0), C.c_uint64(0), C.c_uint32(0)
        _check(self.L.kbbq_batch_extent(h, C.byref(nr), C.byref(nw), C.byref(ml)))
        return DeviceBatch(self, h, BatchExtent(int(nr.value), int(nw.value), int(ml.value)))

    def download(self, db: "DeviceBatch", pinned: bool = False) -> PackedBatch:
        """a resident batch's arrays as a host PackedBatch (e.g. the host copy of a synthetic batch)"""
        e = db.pb
        pb = PackedBatch.empty(e.n_reads, e.n_words, e.max_len, pinned=pinned, qmask=True)
        _check(self.L.kbbq_batch_download(self.h, db.h, _p(pb.bases), _p(pb.nmask), _p(pb.qual), _p(pb.word_off), _p(pb.len),
                                          _p(pb.rg), _p(pb.second), _p(pb.qmask)))
        return pb

    # passes
    def sampler_seed(self, seed: int, alpha: float):
        _check(self.L.kbbq_sampler_seed(self.h, seed, alpha))

    def sampler_seek(self, ordinal: int):
        _check(self.L.kbbq_sampler_seek(self.h, ordinal))

    def pass1(self, db: "DeviceBatch"):
        _check(self.L.kbbq_pass1_subsample(self.h, db.h))

    def set_thresholds(self, th: np.ndarray):
        _check(self.L.kbbq_set_thresholds(self.h, _p(np.ascontiguousarray(th, dtype=np.int32))))

    def compute_thresholds(self, alpha) -> np.ndarray:
        out = np.zeros(self.k + 1, dtype=np.int32)
        _check(self.L.kbbq_compute_thresholds(self.h, C.c_longdouble(alpha), _p(out)))
        return out

    def pass2(self, db: "DeviceBatch", want_errors: bool = False):
        bits = np.zeros(max(db.pb.n_words, 1), dtype=np.uint32) if want_errors else None
        _check(self.L.kbbq_pass2_find_trusted(self.h, db.h, _p(bits)))
        return bits

    def pass3(self, db: "DeviceBatch", want_errors: bool = False):
        bits = np.zeros(max(db.pb.n_words, 1), dtype=np.uint32) if want_errors else None
        _check(self.L.kbbq_pass3_covariates(self.h, db.h, _p(bits)))
        return bits

    def consume_errors(self, db: "DeviceBatch", bits: np.ndarray):
        _check(self.L.kbbq_consume_errors(self.h, db.h, _p(np.ascontiguousarray(bits, dtype=np.uint32))))

    def fixed_errors(self, db: "DeviceBatch", fixed: "DeviceBatch", n_first: int | None = None, want_errors: bool = False):
        """--fixed mode on the device: flags = `db` differs from `fixed`, then consume the first n_first reads"""
        n = min(db.pb.n_reads, fixed.pb.n_reads) if n_first is None else n_first
        bits = np.zeros(max(db.pb.n_words, 1), dtype=np.uint32) if want_errors else None
        _check(self.L.kbbq_fixed_errors(self.h, db.h, fixed.h, n, _p(bits)))
        return bits

    def total_bases(self, db: "DeviceBatch") -> int:
        v = C.c_uint64(0)
        _check(self.L.kbbq_batch_total_bases(self.h, db.h, C.byref(v)))
        return int(v.value)

    def covariates_report(self, path: str, rg_names=None):
        names = None
        if rg_names is not None:
            names = (C.c_char_p * self.n_rg)(*[n.encode() for n in rg_names])
        _check(self.L.kbbq_covariates_report(self.h, os.fsencode(path), names))

    def covariates_reset(self):
        _check(self.L.kbbq_covariates_reset(self.h))

    def covariates(self):
        R, ml = self.n_rg, self.max_cycle
        rgc = np.zeros((R, 2), dtype=np.uint64)
        qc = np.zeros((R, NQ, 2), dtype=np.uint64)
        cyc = np.zeros((R, NQ, 2, ml, 2), dtype=np.uint64)
        dic = np.zeros((R, NQ, 16, 2), dtype=np.uint64)
        _check(self.L.kbbq_covariates_download(self.h, _p(rgc), _p(qc), _p(cyc), _p(dic)))
        return rgc, qc, cyc, dic

    def covariates_device(self):
        n = C.c_uint64(0)
        p = self.L.kbbq_covariates_device(self.h, C.byref(n))
        return int(p or 0), int(n.value)

    def fit_model(self):
        R, ml = self.n_rg, self.max_cycle
        meanq = np.zeros(R, dtype=np.int32)
        rgdq = np.zeros(R, dtype=np.int32)
        qdq = np.zeros((R, NQ), dtype=np.int32)
        cydq = np.zeros((R, NQ, 2, ml), dtype=np.int32)
        didq = np.zeros((R, NQ, 16), dtype=np.int32)
        _check(self.L.kbbq_fit_model(self.h, _p(meanq), _p(rgdq), _p(qdq), _p(cydq), _p(didq)))
        return meanq, rgdq, qdq, cydq, didq

    def fit_model_quiet(self):
        _check(self.L.kbbq_fit_model(self.h, None, None, None, None, None))

    def set_model(self, dqs):
        arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in dqs]
        _check(self.L.kbbq_set_model(self.h, *[_p(a) for a in arrs]))

    def pass4(self, db: "DeviceBatch", out: np.ndarray | None = None, fetch: bool = True):
        if fetch and out is None:
            out = np.zeros(max(db.pb.n_words, 1) * 32, dtype=np.uint8)
        _check(self.L.kbbq_pass4_recalibrate(self.h, db.h, _p(out) if fetch else None))
        return out


class BatchExtent:
    """what the passes need to know of a batch that has no host copy"""

    def __init__(self, n_reads: int, n_words: int, max_len: int):
        self.n_reads, self.n_words, self.max_len = n_reads, n_words, max_len


class DeviceBatch:
    def __init__(self, ctx: Context, h, pb: PackedBatch):
        self.ctx, self.h, self.pb = ctx, h, pb

    @property
    def slots(self) -> int:
        return int(self.ctx.L.kbbq_batch_slots(self.h))

    def free(self):
        if self.h and self.ctx.h:
            self.ctx.L.kbbq_batch_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
