"""ctypes view of oracle/_build/libfmsi_oracle.so — the CHECKER used by the tests (never by the
product path)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "libfmsi_oracle.so")
EXE = os.path.join(ROOT, "oracle", "_build", "fmsi_oracle")
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "fmsi")

MODE_OR, MODE_ALL = 0, 1


class Buf(C.Structure):
    _fields_ = [("s", C.c_void_p), ("len", C.c_size_t), ("cap", C.c_size_t)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("lf_steps", "rank_sectors", "mask_sectors", "klcp_steps", "strand_searches", "kmers")]


_L = None


def L():
    global _L
    if _L is None:
        if not os.path.exists(SO):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
        _L = C.CDLL(SO)
        vp, u64p, u8p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)
        _L.fmsi_oracle_load.restype = vp
        _L.fmsi_oracle_load.argtypes = [C.c_char_p, C.c_int]
        _L.fmsi_oracle_from_bits.restype = vp
        _L.fmsi_oracle_from_bits.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, u8p, C.c_size_t, u8p, C.c_size_t, u64p, C.c_uint64, u8p, C.c_size_t, C.c_int]
        _L.fmsi_oracle_free.argtypes = [vp]
        _L.fmsi_oracle_size.restype = C.c_uint64
        _L.fmsi_oracle_size.argtypes = [vp]
        _L.fmsi_oracle_k.argtypes = [vp]
        _L.fmsi_oracle_has_klcp.argtypes = [vp]
        _L.fmsi_oracle_count.restype = C.c_uint64
        _L.fmsi_oracle_count.argtypes = [vp, C.c_int]
        _L.fmsi_oracle_dollar.restype = C.c_uint64
        _L.fmsi_oracle_dollar.argtypes = [vp]
        _L.fmsi_oracle_reset_predictor.argtypes = [vp]
        _L.fmsi_oracle_rank.restype = C.c_uint64
        _L.fmsi_oracle_rank.argtypes = [vp, C.c_uint64, C.c_int]
        _L.fmsi_oracle_update_range.argtypes = [vp, u64p, u64p, C.c_int]
        _L.fmsi_oracle_extend_range_with_klcp.argtypes = [vp, u64p, u64p]
        _L.fmsi_oracle_get_range_with_pattern.argtypes = [vp, u64p, u64p, C.c_char_p, C.c_int]
        _L.fmsi_oracle_infer_presence.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int]
        _L.fmsi_oracle_kmer_order_if_present.restype = C.c_int64
        _L.fmsi_oracle_kmer_order_if_present.argtypes = [vp, C.c_uint64, C.c_uint64]
        _L.fmsi_oracle_mask_bit.argtypes = [vp, C.c_uint64]
        _L.fmsi_oracle_mask_rank.restype = C.c_uint64
        _L.fmsi_oracle_mask_rank.argtypes = [vp, C.c_uint64]
        _L.fmsi_oracle_klcp_bit.argtypes = [vp, C.c_uint64]
        _L.fmsi_oracle_buf_free.argtypes = [C.POINTER(Buf)]
        _L.fmsi_oracle_query_kmers.argtypes = [vp, C.c_int, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(Buf)]
        _L.fmsi_oracle_ms_query.restype = C.c_int64
        _L.fmsi_oracle_ms_query.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Buf)]
        _L.fmsi_oracle_kmer_both_strands.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _L.fmsi_oracle_query_packed.argtypes = [vp, C.c_int, C.c_int, u64p, C.c_size_t, C.c_int, C.POINTER(C.c_int64)]
        _L.fmsi_oracle_rrr_serialize.restype = C.c_void_p
        _L.fmsi_oracle_rrr_serialize.argtypes = [u8p, C.c_size_t, C.POINTER(C.c_size_t)]
        _L.fmsi_oracle_mask_bits.restype = C.c_void_p
        _L.fmsi_oracle_mask_bits.argtypes = [vp]
        _L.fmsi_oracle_counters_reset.argtypes = [vp]
        _L.fmsi_oracle_counters_get.argtypes = [vp, C.POINTER(Counters)]
        _L.free_ = C.CDLL(None).free
        _L.free_.argtypes = [C.c_void_p]
    return _L


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


class OracleIndex:
    def __init__(self, h):
        if not h:
            raise RuntimeError("oracle: index not loaded")
        self.h = h
        self.n = int(L().fmsi_oracle_size(h))
        self.k = int(L().fmsi_oracle_k(h))
        self.has_klcp = bool(L().fmsi_oracle_has_klcp(h))

    @staticmethod
    def load(prefix, use_klcp=True):
        return OracleIndex(L().fmsi_oracle_load(os.fsencode(prefix), int(use_klcp)))

    @staticmethod
    def from_bits(ac_gt, ac, gt, mask, counts, dollar, klcp=None, k=31):
        a = [_u8(x) for x in (ac_gt, ac, gt, mask)]
        kl = _u8(klcp if klcp is not None else [])
        cnt = np.ascontiguousarray(counts, dtype=np.uint64)
        p = lambda x: x.ctypes.data_as(C.POINTER(C.c_uint8))
        return OracleIndex(L().fmsi_oracle_from_bits(p(a[0]), a[0].size, p(a[1]), a[1].size, p(a[2]), a[2].size, p(a[3]), a[3].size,
                                                     cnt.ctypes.data_as(C.POINTER(C.c_uint64)), int(dollar), p(kl), kl.size, int(k)))

    def close(self):
        if self.h:
            L().fmsi_oracle_free(self.h)
            self.h = None

    def reset_predictor(self):
        L().fmsi_oracle_reset_predictor(self.h)

    def counts(self):
        return [int(L().fmsi_oracle_count(self.h, c)) for c in range(4)]

    def dollar(self):
        return int(L().fmsi_oracle_dollar(self.h))

    def rank(self, i, c):
        return int(L().fmsi_oracle_rank(self.h, int(i), int(c)))

    def update_range(self, i, j, c):
        a, b = C.c_uint64(int(i)), C.c_uint64(int(j))
        L().fmsi_oracle_update_range(self.h, C.byref(a), C.byref(b), int(c))
        return a.value, b.value

    def extend_range_with_klcp(self, i, j):
        a, b = C.c_uint64(int(i)), C.c_uint64(int(j))
        L().fmsi_oracle_extend_range_with_klcp(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def get_range_with_pattern(self, pattern: str):
        a, b = C.c_uint64(0), C.c_uint64(0)
        L().fmsi_oracle_get_range_with_pattern(self.h, C.byref(a), C.byref(b), pattern.encode(), len(pattern))
        return a.value, b.value

    def infer_presence(self, i, j, max_ones):
        return int(L().fmsi_oracle_infer_presence(self.h, int(i), int(j), int(max_ones)))

    def kmer_order_if_present(self, i, j):
        return int(L().fmsi_oracle_kmer_order_if_present(self.h, int(i), int(j)))

    def query_kmers(self, seq: str, k: int, mode=MODE_OR, has_klcp=False, output_orders=False) -> str:
        b = Buf()
        s = seq.encode()
        L().fmsi_oracle_query_kmers(self.h, mode, s, len(s), k, int(has_klcp), int(output_orders), C.byref(b))
        out = C.string_at(b.s, b.len).decode() if b.len else ""
        L().fmsi_oracle_buf_free(C.byref(b))
        return out

    def ms_query(self, text: bytes, k: int, mode=MODE_OR, has_klcp=False, output_orders=False) -> bytes:
        b = Buf()
        L().fmsi_oracle_ms_query(self.h, text, len(text), k, mode, int(has_klcp), int(output_orders), C.byref(b))
        out = C.string_at(b.s, b.len) if b.len else b""
        L().fmsi_oracle_buf_free(C.byref(b))
        return out

    def kmer_both_strands(self, kmer: str, mode=MODE_OR, output_orders=False):
        f, r = C.c_int64(0), C.c_int64(0)
        L().fmsi_oracle_kmer_both_strands(self.h, kmer.encode(), len(kmer), mode, int(output_orders), C.byref(f), C.byref(r))
        return f.value, r.value

    def query_packed(self, kmers, k, mode=MODE_OR, output_orders=False) -> np.ndarray:
        km = np.ascontiguousarray(kmers, dtype=np.uint64)
        out = np.empty(km.size, dtype=np.int64)
        L().fmsi_oracle_query_packed(self.h, mode, int(output_orders), km.ctypes.data_as(C.POINTER(C.c_uint64)), km.size, k,
                                     out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out

    def mask_bits(self) -> np.ndarray:
        p = L().fmsi_oracle_mask_bits(self.h)
        out = np.frombuffer(C.string_at(p, self.n), dtype=np.uint8).copy()
        L().free_(p)
        return out

    def mask_rank(self, i):
        return int(L().fmsi_oracle_mask_rank(self.h, int(i)))

    def klcp_bit(self, i):
        return int(L().fmsi_oracle_klcp_bit(self.h, int(i)))

    def counters_reset(self):
        L().fmsi_oracle_counters_reset(self.h)

    def counters(self) -> dict:
        c = Counters()
        L().fmsi_oracle_counters_get(self.h, C.byref(c))
        return {n: int(getattr(c, n)) for n, _ in Counters._fields_}


def rrr_serialize(bits) -> bytes:
    b = _u8(bits)
    n = C.c_size_t(0)
    p = L().fmsi_oracle_rrr_serialize(b.ctypes.data_as(C.POINTER(C.c_uint8)), b.size, C.byref(n))
    out = C.string_at(p, n.value)
    L().free_(p)
    return out


# the hand-built indexes of the reference's unit tests, tests/fms_index_test.h:10-69
FIXTURES = {
    1: dict(ac_gt=[1, 1, 0, 0, 0, 0, 1, 1], ac=[1, 0, 0, 0], gt=[0, 1, 0, 0], mask=[0, 0, 0, 1, 0, 1, 1, 1],
            counts=[1, 3, 4, 7], dollar=3, klcp=None),
    2: dict(ac_gt=[0, 1, 1, 0, 0, 0, 1, 1, 1], ac=[0, 0, 0, 0], gt=[0, 1, 0, 1, 0], mask=[0, 0, 1, 0, 0, 1, 1, 0, 0],
            counts=[1, 4, 4, 7], dollar=5, klcp=None),
    3: dict(ac_gt=[1, 0, 0, 0, 0, 0, 0, 0], ac=[1, 1, 1, 0, 0, 0, 0], gt=[1], mask=[0, 1, 0, 0, 1, 1, 1, 0],
            counts=[1, 4, 7, 7], dollar=5, klcp=[0, 1, 0, 0, 1, 1, 0, 0]),
}
