"""ctypes loader for oracle/_ref/libkbbq_ref.so -- TEST INFRASTRUCTURE ONLY.

The library is the UNMODIFIED reference compiled by oracle/Makefile (target
`ref`) plus oracle/ref_harness.cc.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libkbbq_ref.so")

_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_new.restype = C.c_void_p
        L.ref_new.argtypes = [C.c_int, C.c_int]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_set_reads.argtypes = [C.c_void_p, C.c_uint64, _u8p, _u64p, _u8p, _i32p, _u8p]
        L.ref_bloom_new.restype = C.c_void_p
        L.ref_bloom_new.argtypes = [C.c_ulonglong, C.c_double]
        L.ref_bloom_new_seed.restype = C.c_void_p
        L.ref_bloom_new_seed.argtypes = [C.c_ulonglong, C.c_double, C.c_ulonglong]
        L.ref_bloom_free.argtypes = [C.c_void_p]
        L.ref_bloom_info.argtypes = [C.c_void_p, _u64p, _u32p]
        L.ref_bloom_words.restype = C.POINTER(C.c_uint64)
        L.ref_bloom_words.argtypes = [C.c_void_p]
        L.ref_bloom_patterns.restype = C.POINTER(C.c_uint64)
        L.ref_bloom_patterns.argtypes = [C.c_void_p]
        L.ref_bloom_insert_u64.argtypes = [C.c_void_p, C.c_uint64]
        L.ref_bloom_contains_u64.restype = C.c_int
        L.ref_bloom_contains_u64.argtypes = [C.c_void_p, C.c_uint64]
        L.ref_bloom_hashes_u64.argtypes = [C.c_void_p, C.c_uint64, _u32p]
        L.ref_bloom_fprate.restype = C.c_double
        L.ref_bloom_fprate.argtypes = [C.c_void_p]
        L.ref_kmer.argtypes = [C.c_int, C.c_char_p, C.c_uint64, _u64p]
        L.ref_rng.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, _u64p, _u64p]
        L.ref_bernoulli.argtypes = [C.c_uint32, C.c_double, C.c_uint64, _u8p]
        L.ref_subsample.restype = C.c_double
        L.ref_subsample.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.POINTER(C.c_uint64)]
        L.ref_phit.restype = C.c_double
        L.ref_phit.argtypes = [C.c_void_p, C.c_longdouble]
        L.ref_thresholds_from_filter.argtypes = [C.c_void_p, C.c_longdouble, C.c_longdouble, C.c_int, _i32p]
        L.ref_thresholds.argtypes = [C.c_int, C.c_double, _i32p]
        L.ref_find_trusted.restype = C.c_double
        L.ref_find_trusted.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, _i32p]
        L.ref_infer_errors.argtypes = [C.c_void_p, C.c_void_p, _i32p, _u8p]
        L.ref_get_errors.restype = C.c_double
        L.ref_get_errors.argtypes = [C.c_void_p, C.c_void_p, _u8p, C.c_void_p]
        L.ref_covariates.restype = C.c_double
        L.ref_covariates.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_consume_with_errors.restype = C.c_double
        L.ref_consume_with_errors.argtypes = [C.c_void_p, _u8p]
        L.ref_dump_covariates.restype = C.c_int
        L.ref_dump_covariates.argtypes = [C.c_void_p, C.c_int, C.c_int, _u64p, _u64p, _u64p, _u64p]
        L.ref_load_covariates.argtypes = [C.c_void_p, C.c_int, C.c_int, _u64p, _u64p, _u64p, _u64p]
        L.ref_get_dqs.restype = C.c_double
        L.ref_get_dqs.argtypes = [C.c_void_p]
        L.ref_dump_dqs.restype = C.c_int
        L.ref_dump_dqs.argtypes = [C.c_void_p, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p]
        L.ref_recalibrate.restype = C.c_double
        L.ref_recalibrate.argtypes = [C.c_void_p, _u8p]
        _lib = L
    return _lib


class RefBloom:
    """bloom::Bloom of the reference (bloom.hh:385-402)."""

    def __init__(self, projected: int, fpr: float, seed: int | None = None):
        L = lib()
        self.h = L.ref_bloom_new(projected, fpr) if seed is None else L.ref_bloom_new_seed(projected, fpr, seed)
        info = np.zeros(8, dtype=np.uint64)
        salts = np.zeros(128, dtype=np.uint32)
        L.ref_bloom_info(self.h, info, salts)
        self.table_bits = int(info[0])
        self.n_blocks = int(info[1])
        self.n_salts = int(info[2])
        self.random_seed = int(info[3])
        self.optimal_k = int(info[5])
        self.optimal_table_size = int(info[6])
        self.salts = salts[: self.n_salts].copy()

    @property
    def inserted(self) -> int:
        info = np.zeros(8, dtype=np.uint64)
        salts = np.zeros(128, dtype=np.uint32)
        lib().ref_bloom_info(self.h, info, salts)
        return int(info[4])

    def words(self) -> np.ndarray:
        p = lib().ref_bloom_words(self.h)
        return np.ctypeslib.as_array(p, shape=(self.table_bits // 64,)).copy()

    def patterns(self) -> np.ndarray:
        p = lib().ref_bloom_patterns(self.h)
        return np.ctypeslib.as_array(p, shape=(65536 * 8,)).copy()

    def insert(self, key: int) -> None:
        lib().ref_bloom_insert_u64(self.h, key)

    def contains(self, key: int) -> bool:
        return bool(lib().ref_bloom_contains_u64(self.h, key))

    def hashes(self, key: int):
        out = np.zeros(2, dtype=np.uint32)
        lib().ref_bloom_hashes_u64(self.h, key, out)
        return int(out[0]), int(out[1])

    def fprate(self) -> float:
        return float(lib().ref_bloom_fprate(self.h))

    def __del__(self):
        try:
            lib().ref_bloom_free(self.h)
        except Exception:
            pass


class RefPipeline:
    """The reference's pass drivers over an in-memory read set (kbbq.cc:264-457)."""

    def __init__(self, rs, k: int):
        self.rs = rs
        self.k = k
        self.L = lib()
        self.h = self.L.ref_new(k, rs.n_rg)
        self.L.ref_set_reads(self.h, rs.n, rs.seq, rs.off, rs.qual, rs.rg, rs.second)
        self.times = {}

    def subsample(self, sampled: RefBloom, alpha: float, seed: int) -> int:
        tot = C.c_uint64(0)
        self.times["pass1"] = self.L.ref_subsample(self.h, sampled.h, alpha, seed, C.byref(tot))
        return int(tot.value)

    def thresholds(self, sampled: RefBloom, alpha_num: float, alpha_den: float) -> np.ndarray:
        out = np.zeros(self.k + 1, dtype=np.int32)
        self.L.ref_thresholds_from_filter(sampled.h, alpha_num, alpha_den, self.k, out)
        return out

    def find_trusted(self, trusted: RefBloom, sampled: RefBloom, thresholds: np.ndarray) -> None:
        self.times["pass2"] = self.L.ref_find_trusted(self.h, trusted.h, sampled.h, np.ascontiguousarray(thresholds, dtype=np.int32))

    def infer_errors(self, sampled: RefBloom, thresholds: np.ndarray) -> np.ndarray:
        out = np.zeros(self.rs.total_bases, dtype=np.uint8)
        self.L.ref_infer_errors(self.h, sampled.h, np.ascontiguousarray(thresholds, dtype=np.int32), out)
        return out

    def get_errors(self, trusted: RefBloom, want_seq: bool = False):
        out = np.zeros(self.rs.total_bases, dtype=np.uint8)
        seq_after = np.zeros(self.rs.total_bases, dtype=np.uint8) if want_seq else None
        self.times["get_errors"] = self.L.ref_get_errors(
            self.h, trusted.h, out, None if seq_after is None else seq_after.ctypes.data_as(C.c_void_p)
        )
        return (out, seq_after) if want_seq else out

    def covariates(self, trusted: RefBloom):
        self.times["pass3"] = self.L.ref_covariates(self.h, trusted.h)
        return self.dump_covariates()

    def consume_with_errors(self, errors: np.ndarray):
        self.times["consume"] = self.L.ref_consume_with_errors(self.h, np.ascontiguousarray(errors, dtype=np.uint8))
        return self.dump_covariates()

    def dump_covariates(self, maxlen: int | None = None):
        R = self.rs.n_rg
        ml = self.rs.max_len if maxlen is None else maxlen
        rgc = np.zeros((R, 2), dtype=np.uint64)
        qc = np.zeros((R, 94, 2), dtype=np.uint64)
        cyc = np.zeros((R, 94, 2, ml, 2), dtype=np.uint64)
        dic = np.zeros((R, 94, 16, 2), dtype=np.uint64)
        self.L.ref_dump_covariates(self.h, R, ml, rgc.reshape(-1), qc.reshape(-1), cyc.reshape(-1), dic.reshape(-1))
        return rgc, qc, cyc, dic

    def load_covariates(self, rgc, qc, cyc, dic):
        R = self.rs.n_rg
        ml = cyc.shape[3]
        self.L.ref_load_covariates(
            self.h, R, ml,
            np.ascontiguousarray(rgc, dtype=np.uint64).reshape(-1), np.ascontiguousarray(qc, dtype=np.uint64).reshape(-1),
            np.ascontiguousarray(cyc, dtype=np.uint64).reshape(-1), np.ascontiguousarray(dic, dtype=np.uint64).reshape(-1),
        )

    def get_dqs(self, maxlen: int | None = None):
        self.times["get_dqs"] = self.L.ref_get_dqs(self.h)
        R = self.rs.n_rg
        ml = self.rs.max_len if maxlen is None else maxlen
        meanq = np.zeros(R, dtype=np.int32)
        rgdq = np.zeros(R, dtype=np.int32)
        qdq = np.zeros((R, 94), dtype=np.int32)
        cydq = np.zeros((R, 94, 2, ml), dtype=np.int32)
        didq = np.zeros((R, 94, 16), dtype=np.int32)
        self.L.ref_dump_dqs(self.h, R, ml, meanq, rgdq, qdq.reshape(-1), cydq.reshape(-1), didq.reshape(-1))
        return meanq, rgdq, qdq, cydq, didq

    def recalibrate(self) -> np.ndarray:
        out = np.zeros(self.rs.total_bases, dtype=np.uint8)
        self.times["pass4"] = self.L.ref_recalibrate(self.h, out)
        return out

    def __del__(self):
        try:
            self.L.ref_free(self.h)
        except Exception:
            pass


def ref_kmer(k: int, s: bytes):
    out = np.zeros(5, dtype=np.uint64)
    lib().ref_kmer(k, s, len(s), out)
    return {"fwd": int(out[0]), "rc": int(out[1]), "get": int(out[2]), "size": int(out[3]), "valid": bool(out[4])}


def ref_rng(seed: int, skip: int, n: int):
    st = np.zeros(4, dtype=np.uint64)
    out = np.zeros(max(n, 1), dtype=np.uint64)
    lib().ref_rng(seed, skip, n, st, out)
    return st, out[:n]


def ref_bernoulli(seed: int, p: float, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint8)
    lib().ref_bernoulli(seed, p, n, out)
    return out
