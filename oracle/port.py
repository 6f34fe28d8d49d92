"""ctypes loader for oracle/libkbbq_oracle.so (the plain-C restatement) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkbbq_oracle.so")

_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class _KoBloom(C.Structure):
    _fields_ = [
        ("table_bits", C.c_uint64), ("n_blocks", C.c_uint64), ("n_salts", C.c_uint32), ("salts", C.c_uint32 * 128),
        ("random_seed", C.c_uint64), ("optimal_k", C.c_uint32), ("optimal_table_size", C.c_uint64),
        ("inserted", C.c_uint64), ("words", C.POINTER(C.c_uint64)), ("patterns", C.POINTER(C.c_uint64)),
    ]


def build() -> None:
    subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        BP = C.POINTER(_KoBloom)
        L.ko_nt4.restype = C.c_int
        L.ko_nt4.argtypes = [C.c_ubyte]
        L.ko_hash_ap8.restype = C.c_uint32
        L.ko_hash_ap8.argtypes = [C.c_uint64, C.c_uint32]
        L.ko_kmer_of.argtypes = [C.c_int, C.c_char_p, C.c_size_t, _u64p]
        L.ko_rng_outputs.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, _u64p, _u64p]
        L.ko_bernoulli.argtypes = [C.c_uint32, C.c_double, C.c_uint64, _u8p]
        L.ko_bloom_new.restype = BP
        L.ko_bloom_new.argtypes = [C.c_uint64, C.c_double, C.c_uint64]
        L.ko_bloom_free.argtypes = [BP]
        L.ko_bloom_insert.argtypes = [BP, C.c_uint64]
        L.ko_bloom_contains.restype = C.c_int
        L.ko_bloom_contains.argtypes = [BP, C.c_uint64]
        L.ko_bloom_fprate.restype = C.c_double
        L.ko_bloom_fprate.argtypes = [BP]
        L.ko_phit.restype = C.c_longdouble
        L.ko_phit.argtypes = [BP, C.c_longdouble]
        L.ko_thresholds.argtypes = [C.c_int, C.c_longdouble, _i32p]
        L.ko_thresholds_from_filter.argtypes = [BP, C.c_longdouble, C.c_longdouble, C.c_int, _i32p]
        L.ko_subsample.restype = C.c_uint64
        L.ko_subsample.argtypes = [C.c_uint64, _u8p, _u64p, C.c_int, C.c_double, C.c_uint32, BP]
        L.ko_subsample_skip.restype = C.c_uint64
        L.ko_subsample_skip.argtypes = [C.c_uint64, _u8p, _u64p, C.c_int, C.c_double, C.c_uint32, C.c_uint64, BP]
        L.ko_infer_errors.argtypes = [C.c_uint64, _u8p, _u64p, _u8p, C.c_int, BP, _i32p, _u8p]
        L.ko_find_trusted.argtypes = [C.c_uint64, _u8p, _u64p, _u8p, C.c_int, BP, _i32p, BP]
        L.ko_get_errors.argtypes = [C.c_uint64, _u8p, _u64p, _u8p, C.c_int, BP, _u8p, C.c_void_p]
        L.ko_consume.argtypes = [C.c_uint64, _u8p, _u64p, _u8p, _i32p, _u8p, _u8p, C.c_int, C.c_int, _u64p, _u64p, _u64p, _u64p]
        L.ko_get_dqs.argtypes = [C.c_int, C.c_int, _u64p, _u64p, _u64p, _u64p, _i32p, _i32p, _i32p, _i32p, _i32p]
        L.ko_recalibrate.argtypes = [C.c_uint64, _u8p, _u64p, _u8p, _i32p, _u8p, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p, _u8p]
        _lib = L
    return _lib


DEFAULT_SEED = 0xA5A5A5A55A5A5A5A  # bloom.hh:389


class PortBloom:
    def __init__(self, projected: int, fpr: float, seed: int = DEFAULT_SEED):
        self.p = lib().ko_bloom_new(projected, fpr, seed)
        if not self.p:
            raise ValueError("invalid bloom filter parameters")
        c = self.p.contents
        self.table_bits = int(c.table_bits)
        self.n_blocks = int(c.n_blocks)
        self.n_salts = int(c.n_salts)
        self.random_seed = int(c.random_seed)
        self.optimal_k = int(c.optimal_k)
        self.optimal_table_size = int(c.optimal_table_size)
        self.salts = np.array(c.salts[: self.n_salts], dtype=np.uint32)

    @property
    def inserted(self) -> int:
        return int(self.p.contents.inserted)

    def words(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.p.contents.words, shape=(self.table_bits // 64,)).copy()

    def set_words(self, w: np.ndarray, inserted: int) -> None:
        C.memmove(self.p.contents.words, w.ctypes.data, self.table_bits // 8)
        self.p.contents.inserted = inserted

    def patterns(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.p.contents.patterns, shape=(65536 * 8,)).copy()

    def insert(self, key: int) -> None:
        lib().ko_bloom_insert(self.p, key)

    def contains(self, key: int) -> bool:
        return bool(lib().ko_bloom_contains(self.p, key))

    def fprate(self) -> float:
        return float(lib().ko_bloom_fprate(self.p))

    def __del__(self):
        try:
            lib().ko_bloom_free(self.p)
        except Exception:
            pass


class PortPipeline:
    """Same surface as oracle.refharness.RefPipeline, backed by the C restatement."""

    def __init__(self, rs, k: int):
        self.rs = rs
        self.k = k
        self.L = lib()

    def subsample(self, sampled: PortBloom, alpha: float, seed: int, skip: int = 0) -> int:
        return int(self.L.ko_subsample_skip(self.rs.n, self.rs.seq, self.rs.off, self.k, alpha, seed, skip, sampled.p))

    def thresholds(self, sampled: PortBloom, alpha_num: float, alpha_den: float) -> np.ndarray:
        out = np.zeros(self.k + 1, dtype=np.int32)
        self.L.ko_thresholds_from_filter(sampled.p, alpha_num, alpha_den, self.k, out)
        return out

    def find_trusted(self, trusted: PortBloom, sampled: PortBloom, thresholds) -> None:
        self.L.ko_find_trusted(self.rs.n, self.rs.seq, self.rs.off, self.rs.qual, self.k, sampled.p,
                               np.ascontiguousarray(thresholds, dtype=np.int32), trusted.p)

    def infer_errors(self, sampled: PortBloom, thresholds) -> np.ndarray:
        out = np.zeros(self.rs.total_bases, dtype=np.uint8)
        self.L.ko_infer_errors(self.rs.n, self.rs.seq, self.rs.off, self.rs.qual, self.k, sampled.p,
                               np.ascontiguousarray(thresholds, dtype=np.int32), out)
        return out

    def get_errors(self, trusted: PortBloom, want_seq: bool = False):
        out = np.zeros(self.rs.total_bases, dtype=np.uint8)
        seq_after = np.zeros(self.rs.total_bases, dtype=np.uint8) if want_seq else None
        self.L.ko_get_errors(self.rs.n, self.rs.seq, self.rs.off, self.rs.qual, self.k, trusted.p, out,
                             None if seq_after is None else seq_after.ctypes.data_as(C.c_void_p))
        return (out, seq_after) if want_seq else out

    def consume(self, errors: np.ndarray, seq: np.ndarray | None = None, maxlen: int | None = None):
        R = self.rs.n_rg
        ml = self.rs.max_len if maxlen is None else maxlen
        rgc = np.zeros((R, 2), dtype=np.uint64)
        qc = np.zeros((R, 94, 2), dtype=np.uint64)
        cyc = np.zeros((R, 94, 2, ml, 2), dtype=np.uint64)
        dic = np.zeros((R, 94, 16, 2), dtype=np.uint64)
        self.L.ko_consume(self.rs.n, self.rs.seq if seq is None else seq, self.rs.off, self.rs.qual, self.rs.rg,
                          self.rs.second, np.ascontiguousarray(errors, dtype=np.uint8), R, ml,
                          rgc.reshape(-1), qc.reshape(-1), cyc.reshape(-1), dic.reshape(-1))
        return rgc, qc, cyc, dic

    def covariates(self, trusted: PortBloom):
        err, seq_after = self.get_errors(trusted, want_seq=True)
        return self.consume(err, seq_after)

    def get_dqs(self, tables):
        rgc, qc, cyc, dic = tables
        R, ml = rgc.shape[0], cyc.shape[3]
        meanq = np.zeros(R, dtype=np.int32)
        rgdq = np.zeros(R, dtype=np.int32)
        qdq = np.zeros((R, 94), dtype=np.int32)
        cydq = np.zeros((R, 94, 2, ml), dtype=np.int32)
        didq = np.zeros((R, 94, 16), dtype=np.int32)
        self.L.ko_get_dqs(R, ml, np.ascontiguousarray(rgc).reshape(-1), np.ascontiguousarray(qc).reshape(-1),
                          np.ascontiguousarray(cyc).reshape(-1), np.ascontiguousarray(dic).reshape(-1),
                          meanq, rgdq, qdq.reshape(-1), cydq.reshape(-1), didq.reshape(-1))
        return meanq, rgdq, qdq, cydq, didq

    def recalibrate(self, dqs) -> np.ndarray:
        meanq, rgdq, qdq, cydq, didq = dqs
        out = np.zeros(self.rs.total_bases, dtype=np.uint8)
        self.L.ko_recalibrate(self.rs.n, self.rs.seq, self.rs.off, self.rs.qual, self.rs.rg, self.rs.second,
                              meanq.shape[0], cydq.shape[3], meanq, rgdq, np.ascontiguousarray(qdq).reshape(-1),
                              np.ascontiguousarray(cydq).reshape(-1), np.ascontiguousarray(didq).reshape(-1), out)
        return out


def kmer_of(k: int, s: bytes):
    out = np.zeros(5, dtype=np.uint64)
    lib().ko_kmer_of(k, s, len(s), out)
    return {"fwd": int(out[0]), "rc": int(out[1]), "get": int(out[2]), "size": int(out[3]), "valid": bool(out[4])}


def rng_outputs(seed: int, skip: int, n: int):
    st = np.zeros(4, dtype=np.uint64)
    out = np.zeros(max(n, 1), dtype=np.uint64)
    lib().ko_rng_outputs(seed, skip, n, st, out)
    return st, out[:n]


def bernoulli(seed: int, p: float, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint8)
    lib().ko_bernoulli(seed, p, n, out)
    return out
