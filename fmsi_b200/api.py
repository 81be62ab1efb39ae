"""ctypes binding of include/fmsi_gpu.h. Plumbing only — every call lands in libfmsi_gpu.so.

The product path fails loudly when the CUDA library is missing: `lib()` raises, nothing here
computes anything on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

MODE_OR, MODE_ALL = 0, 1
OUT_PRESENCE, OUT_ORDERS, OUT_PRESENCE_BITS = 0, 1, 2
TEXT_ASCII, TEXT_PACKED2 = 0, 1
STRANDS_LAZY, STRANDS_BOTH = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
MAX_STREAM_KMERS = 64

# every symbol include/fmsi_gpu.h declares (tests check the .so exports exactly these)
EXPORTED_SYMBOLS = [
    "fmsi_gpu_last_error", "fmsi_gpu_abi_version", "fmsi_gpu_device_count", "fmsi_gpu_index_load",
    "fmsi_gpu_index_from_bits", "fmsi_gpu_index_build", "fmsi_gpu_index_save", "fmsi_gpu_index_free", "fmsi_gpu_index_get_info", "fmsi_gpu_rank",
    "fmsi_gpu_update_range", "fmsi_gpu_extend_range_with_klcp", "fmsi_gpu_get_range_with_pattern",
    "fmsi_gpu_infer_presence", "fmsi_gpu_kmer_order_if_present", "fmsi_gpu_query_kmers",
    "fmsi_gpu_query_chunks", "fmsi_gpu_launch_count", "fmsi_gpu_pool_create", "fmsi_gpu_pool_size", "fmsi_gpu_pool_member", "fmsi_gpu_pool_free",
    "fmsi_gpu_pool_query_kmers", "fmsi_gpu_pool_query_chunks", "fmsi_gpu_query_kmers_general", "fmsi_gpu_query_chunks_general",
    "fmsi_gpu_query_chunks_packed", "fmsi_gpu_count_probes", "fmsi_gpu_query_reads_packed",
]
F_OR, F_AND, F_XOR, F_RANGE = 0, 1, 2, 3


class FmsiGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libfmsi_gpu error {code}: {msg}")
        self.code = code


class Options(C.Structure):
    _fields_ = [("prefix_t", C.c_int32), ("sb_shift_log2", C.c_int32), ("dict", C.c_int32), ("multistep", C.c_int32),
                ("fold_ids", C.c_int32), ("locality", C.c_int32), ("reserved", C.c_int64 * 4)]


class Function(C.Structure):
    """Demasking function of the f-MS framework (reference src/functions.h)."""
    _fields_ = [("kind", C.c_int32), ("r", C.c_int32), ("s", C.c_int32), ("reserved", C.c_int32)]

    @staticmethod
    def parse(name: str) -> "Function":
        if name in ("or", "and", "xor"):
            return Function(kind={"or": F_OR, "and": F_AND, "xor": F_XOR}[name])
        r, s = name.split("-")
        return Function(kind=F_RANGE, r=int(r), s=int(s))


class IndexInfo(C.Structure):
    _fields_ = [
        ("n_bwt", C.c_uint64), ("counts", C.c_uint64 * 4), ("dollar_position", C.c_uint64),
        ("mask_ones", C.c_uint64), ("hbm_bytes", C.c_uint64), ("k", C.c_int32), ("has_klcp", C.c_int32),
        ("prefix_t", C.c_int32), ("wide", C.c_int32), ("device", C.c_int32), ("dict", C.c_int32),
        ("dict_t", C.c_int32), ("multistep", C.c_int32), ("fold_ids", C.c_int32), ("locality", C.c_int32),
        ("reserved", C.c_int32 * 2),
    ]


_lib = None


def lib_path() -> str:
    return os.environ.get("FMSI_GPU_LIB", os.path.join(HERE, "libfmsi_gpu.so"))


def lib() -> C.CDLL:
    """Load libfmsi_gpu.so (built in-tree by fmsi_b200/build.py). Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise FmsiGpuError(-3, f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp, u64p, u8p, i8p, i64p, u32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.POINTER(C.c_int8), C.POINTER(C.c_int64), C.POINTER(C.c_uint32)
    L.fmsi_gpu_last_error.restype = C.c_char_p
    L.fmsi_gpu_abi_version.restype = C.c_int
    L.fmsi_gpu_device_count.restype = C.c_int
    L.fmsi_gpu_launch_count.restype = C.c_uint64
    L.fmsi_gpu_index_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(Options), C.POINTER(vp)]
    L.fmsi_gpu_index_from_bits.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, u8p, C.c_size_t, u8p, C.c_size_t, u64p,
                                           C.c_uint64, u8p, C.c_size_t, C.c_int, C.c_int, C.POINTER(Options), C.POINTER(vp)]
    L.fmsi_gpu_index_build.argtypes = [vp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Options), C.POINTER(vp)]
    L.fmsi_gpu_index_save.argtypes = [vp, C.c_char_p]
    L.fmsi_gpu_index_free.argtypes = [vp]
    L.fmsi_gpu_index_get_info.argtypes = [vp, C.POINTER(IndexInfo)]
    L.fmsi_gpu_rank.argtypes = [vp, u64p, u8p, C.c_size_t, u64p]
    L.fmsi_gpu_update_range.argtypes = [vp, u64p, u64p, u8p, C.c_size_t]
    L.fmsi_gpu_extend_range_with_klcp.argtypes = [vp, u64p, u64p, C.c_size_t]
    L.fmsi_gpu_get_range_with_pattern.argtypes = [vp, u64p, C.c_int, C.c_size_t, C.c_int, u64p, u64p]
    L.fmsi_gpu_infer_presence.argtypes = [vp, u64p, u64p, C.c_size_t, C.c_int, i8p]
    L.fmsi_gpu_kmer_order_if_present.argtypes = [vp, u64p, u64p, C.c_size_t, i64p]
    L.fmsi_gpu_query_kmers.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, C.c_int, vp, C.c_int, vp]
    L.fmsi_gpu_query_chunks.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp, vp, vp,
                                        C.c_size_t, C.c_size_t, C.c_int, vp, C.c_int, vp]
    L.fmsi_gpu_query_chunks_packed.argtypes = L.fmsi_gpu_query_chunks.argtypes
    L.fmsi_gpu_count_probes.argtypes = [vp, C.c_int, u64p]
    L.fmsi_gpu_query_reads_packed.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t, C.c_int, vp, C.c_int, vp]
    L.fmsi_gpu_query_kmers_general.argtypes = [vp, C.POINTER(Function), vp, C.c_size_t, C.c_int, vp, C.c_int, vp]
    L.fmsi_gpu_query_chunks_general.argtypes = [vp, C.POINTER(Function), vp, C.c_size_t, vp, vp, vp, C.c_size_t, C.c_size_t, C.c_int,
                                                vp, C.c_int, vp]
    L.fmsi_gpu_pool_create.argtypes = [vp, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.fmsi_gpu_pool_size.argtypes = [vp]
    L.fmsi_gpu_pool_member.argtypes = [vp, C.c_int]
    L.fmsi_gpu_pool_member.restype = vp
    L.fmsi_gpu_pool_free.argtypes = [vp]
    L.fmsi_gpu_pool_query_kmers.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, C.c_int, vp]
    L.fmsi_gpu_pool_query_chunks.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp, vp, C.c_size_t,
                                             C.c_size_t, C.c_int, vp]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("fmsi_gpu_abi_version", "fmsi_gpu_device_count"):
            fn.restype = C.c_int
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise FmsiGpuError(rc, lib().fmsi_gpu_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(lib().fmsi_gpu_launch_count())


def device_count() -> int:
    return int(lib().fmsi_gpu_device_count())


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _ptr(a: np.ndarray, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def result_dtype_shape(output: int, strands: int, n: int):
    if output == OUT_PRESENCE:
        return np.uint8, (n,)
    if output == OUT_PRESENCE_BITS:
        return np.uint8, ((n + 7) // 8,)
    return np.int64, ((n, 2) if strands == STRANDS_BOTH else (n,))


class Index:
    """Handle to a GPU-resident FMS-index (the reference's `fms_index`, src/fms_index.h:52-66)."""

    def __init__(self, handle: C.c_void_p):
        self._h = handle
        self.refresh_info()

    def refresh_info(self) -> "Index":
        """Re-read fmsi_gpu_index_get_info (the resident tiers can change: lookup ids are built on the first lookup)."""
        info = IndexInfo()
        _check(lib().fmsi_gpu_index_get_info(self._h, C.byref(info)))
        self.info = info
        self.n = int(info.n_bwt)
        self.k = int(info.k)
        self.has_klcp = bool(info.has_klcp)
        self.counts = [int(c) for c in info.counts]
        self.dollar_position = int(info.dollar_position)
        self.prefix_t = int(info.prefix_t)
        self.wide = bool(info.wide)
        self.dict = bool(info.dict)
        self.dict_kind = int(info.dict)   # 0 none, 1 SA-ordered dictionary, 2 strand-folded dictionary
        self.dict_t = int(info.dict_t)
        self.multistep = int(info.multistep)
        self.fold_ids = bool(info.fold_ids)
        self.locality = int(info.locality)  # minimizer length of the resident minimizer-bucketed dictionary (0 = none)
        self.hbm_bytes = int(info.hbm_bytes)
        self.mask_ones = int(info.mask_ones)
        return self

    # ---- construction -------------------------------------------------------------------------
    @staticmethod
    def load(prefix: str, use_klcp: bool = True, device: int = 0, prefix_t: int = -1, sb_shift_log2: int = 0,
             dict: int = -1, multistep: int = -1, fold_ids: int = 0, locality: int = 0) -> "Index":
        """load_index(fn, use_klcp) — reference src/fms_index.h:502."""
        opts = Options(prefix_t=prefix_t, sb_shift_log2=sb_shift_log2, dict=dict, multistep=multistep, fold_ids=fold_ids, locality=locality)
        h = C.c_void_p()
        _check(lib().fmsi_gpu_index_load(os.fsencode(prefix), int(use_klcp), device, C.byref(opts), C.byref(h)))
        return Index(h)

    @staticmethod
    def from_bits(ac_gt, ac, gt, mask, counts, dollar_position, klcp=None, k=31, device=0, prefix_t=-1,
                  sb_shift_log2=0, dict=-1, multistep=-1, fold_ids=0, locality=0) -> "Index":
        """In-memory fixture, the form of tests/fms_index_test.h:10-69."""
        arrs = [np.ascontiguousarray(x, dtype=np.uint8) for x in (ac_gt, ac, gt, mask)]
        kl = np.ascontiguousarray(klcp if klcp is not None else [], dtype=np.uint8)
        cnt = _u64(counts)
        opts = Options(prefix_t=prefix_t, sb_shift_log2=sb_shift_log2, dict=dict, multistep=multistep, fold_ids=fold_ids, locality=locality)
        h = C.c_void_p()
        _check(lib().fmsi_gpu_index_from_bits(
            _ptr(arrs[0], C.c_uint8), arrs[0].size, _ptr(arrs[1], C.c_uint8), arrs[1].size,
            _ptr(arrs[2], C.c_uint8), arrs[2].size, _ptr(arrs[3], C.c_uint8), arrs[3].size,
            _ptr(cnt, C.c_uint64), int(dollar_position), _ptr(kl, C.c_uint8), kl.size, int(k), device,
            C.byref(opts), C.byref(h)))
        return Index(h)

    @staticmethod
    def build(ms, k: int, with_klcp: bool = True, device: int = 0, prefix_t: int = -1, n: int | None = None,
              mem: int = MEM_HOST, dict: int = -1, multistep: int = -1, fold_ids: int = 0, locality: int = 0) -> "Index":
        """construct(ms, k, use_klcp) on the GPU (reference src/fms_index.h:397). ms: mask-cased ASCII
        bytes (host) or a raw device pointer (int) with n and mem=MEM_DEVICE."""
        opts = Options(prefix_t=prefix_t, sb_shift_log2=0, dict=dict, multistep=multistep, fold_ids=fold_ids, locality=locality)
        h = C.c_void_p()
        if isinstance(ms, int):
            ptr, length = ms, int(n)
        else:
            buf = np.frombuffer(ms, dtype=np.uint8)
            ptr, length = buf.ctypes.data, buf.size
        _check(lib().fmsi_gpu_index_build(ptr, length, int(k), int(with_klcp), mem, device, C.byref(opts), C.byref(h)))
        return Index(h)

    def save(self, prefix: str) -> None:
        """dump_index(index, fn) (reference src/fms_index.h:484): reference-format .fmsi.* files."""
        _check(lib().fmsi_gpu_index_save(self._h, os.fsencode(prefix)))

    def close(self) -> None:
        if self._h:
            lib().fmsi_gpu_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def count_probes(self, on: bool) -> int:
        """Switch the kernels' probe accounting on / off; returns the running total of dependent requests so far."""
        total = C.c_uint64(0)
        _check(lib().fmsi_gpu_count_probes(self._h, int(on), C.byref(total)))
        return int(total.value)

    # ---- building blocks (reference names) ----------------------------------------------------
    def rank(self, i, c) -> np.ndarray:
        i = _u64(np.atleast_1d(i))
        c = np.ascontiguousarray(np.atleast_1d(c), dtype=np.uint8)
        out = np.empty(i.size, dtype=np.uint64)
        _check(lib().fmsi_gpu_rank(self._h, _ptr(i, C.c_uint64), _ptr(c, C.c_uint8), i.size, _ptr(out, C.c_uint64)))
        return out

    def update_range(self, i, j, c):
        i = _u64(np.atleast_1d(i)).copy()
        j = _u64(np.atleast_1d(j)).copy()
        c = np.ascontiguousarray(np.atleast_1d(c), dtype=np.uint8)
        _check(lib().fmsi_gpu_update_range(self._h, _ptr(i, C.c_uint64), _ptr(j, C.c_uint64), _ptr(c, C.c_uint8), i.size))
        return i, j

    def extend_range_with_klcp(self, i, j):
        i = _u64(np.atleast_1d(i)).copy()
        j = _u64(np.atleast_1d(j)).copy()
        _check(lib().fmsi_gpu_extend_range_with_klcp(self._h, _ptr(i, C.c_uint64), _ptr(j, C.c_uint64), i.size))
        return i, j

    def get_range_with_pattern(self, kmers, k: int, use_table: bool = True):
        kmers = _u64(np.atleast_1d(kmers))
        i = np.empty(kmers.size, dtype=np.uint64)
        j = np.empty(kmers.size, dtype=np.uint64)
        _check(lib().fmsi_gpu_get_range_with_pattern(self._h, _ptr(kmers, C.c_uint64), k, kmers.size, int(use_table),
                                                    _ptr(i, C.c_uint64), _ptr(j, C.c_uint64)))
        return i, j

    def infer_presence(self, sa_start, sa_end, maximized_ones: bool) -> np.ndarray:
        i = _u64(np.atleast_1d(sa_start))
        j = _u64(np.atleast_1d(sa_end))
        out = np.empty(i.size, dtype=np.int8)
        _check(lib().fmsi_gpu_infer_presence(self._h, _ptr(i, C.c_uint64), _ptr(j, C.c_uint64), i.size,
                                            int(maximized_ones), _ptr(out, C.c_int8)))
        return out

    def kmer_order_if_present(self, sa_start, sa_end) -> np.ndarray:
        i = _u64(np.atleast_1d(sa_start))
        j = _u64(np.atleast_1d(sa_end))
        out = np.empty(i.size, dtype=np.int64)
        _check(lib().fmsi_gpu_kmer_order_if_present(self._h, _ptr(i, C.c_uint64), _ptr(j, C.c_uint64), i.size, _ptr(out, C.c_int64)))
        return out

    # ---- hot path -----------------------------------------------------------------------------
    def query_kmers(self, kmers, k: int | None = None, mode: int = MODE_OR, output: int = OUT_PRESENCE,
                    strands: int = STRANDS_LAZY) -> np.ndarray:
        """query_kmers_single over packed k-mers held in host memory (numpy)."""
        kmers = _u64(kmers)
        k = self.k if k is None else k
        dt, shape = result_dtype_shape(output, strands, kmers.size)
        out = np.empty(shape, dtype=dt)
        _check(lib().fmsi_gpu_query_kmers(self._h, mode, output, strands, kmers.ctypes.data, kmers.size, k,
                                         out.ctypes.data, MEM_HOST, None))
        return out

    def query_kmers_general(self, kmers, f: "Function | str", k: int | None = None) -> np.ndarray:
        """query_kmers<general>: f(#ON occurrences, #occurrences) over both strands; uint8 0/1."""
        kmers = _u64(kmers)
        k = self.k if k is None else k
        fn = Function.parse(f) if isinstance(f, str) else f
        out = np.empty(kmers.size, dtype=np.uint8)
        _check(lib().fmsi_gpu_query_kmers_general(self._h, C.byref(fn), kmers.ctypes.data, kmers.size, k, out.ctypes.data, MEM_HOST, None))
        return out

    def query_kmers_ptr(self, kmers_ptr: int, n: int, out_ptr: int, k: int | None = None, mode: int = MODE_OR,
                        output: int = OUT_PRESENCE, strands: int = STRANDS_LAZY, mem: int = MEM_DEVICE,
                        stream: int = 0) -> None:
        """Raw-pointer form (device tensors' data_ptr() or pinned host buffers)."""
        k = self.k if k is None else k
        _check(lib().fmsi_gpu_query_kmers(self._h, mode, output, strands, kmers_ptr, n, k, out_ptr, mem, stream or None))

    def query_chunks(self, bases: bytes | np.ndarray, chunk_off, chunk_len, k: int | None = None, mode: int = MODE_OR,
                     output: int = OUT_PRESENCE, strands: int = STRANDS_LAZY, streaming: bool = False, packed: bool = False) -> np.ndarray:
        """query_kmers<mode>() over chunks of ACGT text in host memory; results concatenated chunk by chunk.
        packed=True sends the text 2-bit packed (fmsi_gpu_query_chunks_packed); the packing is done here, on the host."""
        k = self.k if k is None else k
        b = np.frombuffer(bases, dtype=np.uint8) if isinstance(bases, (bytes, bytearray)) else np.ascontiguousarray(bases, dtype=np.uint8)
        if packed:
            return self._query_chunks_packed(pack_text(b), b.size, chunk_off, chunk_len, k, mode, output, strands, streaming)
        off = _u64(chunk_off)
        ln = np.ascontiguousarray(chunk_len, dtype=np.uint32)
        cnt = ln.astype(np.int64) - k + 1
        if (cnt < 1).any():
            raise ValueError("every chunk must hold at least one k-mer")
        res_off = np.zeros(off.size, dtype=np.uint64)
        if off.size:
            res_off[1:] = np.cumsum(cnt)[:-1].astype(np.uint64)
        n_res = int(cnt.sum())
        dt, shape = result_dtype_shape(output, strands, n_res)
        out = np.empty(shape, dtype=dt)
        _check(lib().fmsi_gpu_query_chunks(self._h, mode, output, strands, int(streaming), b.ctypes.data, b.size,
                                          off.ctypes.data, ln.ctypes.data, res_off.ctypes.data, off.size, n_res, k,
                                          out.ctypes.data, MEM_HOST, None))
        return out


    def query_reads(self, reads, k: int | None = None, mode: int = MODE_OR, output: int = OUT_PRESENCE, strands: int = STRANDS_LAZY,
                    streaming: bool = False) -> np.ndarray:
        """fmsi_gpu_query_reads_packed: `reads` = list of ASCII byte strings (any lengths); results read by read."""
        k = self.k if k is None else k
        lens = np.array([len(r) for r in reads], dtype=np.int64)
        off = np.zeros(len(reads) + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens).astype(np.uint64)
        text = np.frombuffer(b"".join(reads), dtype=np.uint8)
        words = pack_text(text)
        n_res = int(np.maximum(lens - k + 1, 0).sum())
        dt, shape = result_dtype_shape(output, strands, n_res)
        out = np.empty(shape, dtype=dt)
        _check(lib().fmsi_gpu_query_reads_packed(self._h, mode, output, strands, int(streaming), words.ctypes.data, text.size, off.ctypes.data,
                                                len(reads), n_res, k, out.ctypes.data, MEM_HOST, None))
        return out

    def _query_chunks_packed(self, words: np.ndarray, n_bases: int, chunk_off, chunk_len, k, mode, output, strands, streaming) -> np.ndarray:
        off = _u64(chunk_off)
        ln = np.ascontiguousarray(chunk_len, dtype=np.uint32)
        cnt = ln.astype(np.int64) - k + 1
        if (cnt < 1).any():
            raise ValueError("every chunk must hold at least one k-mer")
        res_off = np.zeros(off.size, dtype=np.uint64)
        if off.size:
            res_off[1:] = np.cumsum(cnt)[:-1].astype(np.uint64)
        n_res = int(cnt.sum())
        dt, shape = result_dtype_shape(output, strands, n_res)
        out = np.empty(shape, dtype=dt)
        _check(lib().fmsi_gpu_query_chunks_packed(self._h, mode, output, strands, int(streaming), words.ctypes.data, n_bases,
                                                 off.ctypes.data, ln.ctypes.data, res_off.ctypes.data, off.size, n_res, k,
                                                 out.ctypes.data, MEM_HOST, None))
        return out


def pack_text(ascii_bases: np.ndarray) -> np.ndarray:
    """ASCII ACGTacgt -> FMSI_GPU_TEXT_PACKED2 words (32 bases per uint64, first base in the highest bits)."""
    b = np.ascontiguousarray(ascii_bases, dtype=np.uint8)
    x = (b >> 1) & 3
    codes = (x ^ (x >> 1)).astype(np.uint64)
    n_words = (b.size + 31) // 32
    pad = np.zeros(n_words * 32, dtype=np.uint64)
    pad[:b.size] = codes
    shifts = (62 - 2 * np.arange(32, dtype=np.uint64)).astype(np.uint64)
    return np.bitwise_or.reduce(pad.reshape(n_words, 32) << shifts[None, :], axis=1)


class Pool:
    """Multi-GPU scheduler: `primary` plus device-to-device replicas on `devices`; host batches are
    split into contiguous ranges, one host thread per member, results in query order."""

    def __init__(self, primary: Index, devices):
        self.primary = primary
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        _check(lib().fmsi_gpu_pool_create(primary._h, devs, len(devices), C.byref(h)))
        self._h = h
        self.size = int(lib().fmsi_gpu_pool_size(self._h))

    def close(self) -> None:
        if self._h:
            lib().fmsi_gpu_pool_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query_kmers(self, kmers, k: int | None = None, mode: int = MODE_OR, output: int = OUT_PRESENCE,
                    strands: int = STRANDS_LAZY, out: np.ndarray | None = None) -> np.ndarray:
        kmers = _u64(kmers)
        k = self.primary.k if k is None else k
        dt, shape = result_dtype_shape(output, strands, kmers.size)
        if out is None:
            out = np.empty(shape, dtype=dt)
        elif out.dtype != dt or out.shape != shape or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous %s array of shape %s" % (np.dtype(dt).name, shape))
        _check(lib().fmsi_gpu_pool_query_kmers(self._h, mode, output, strands, kmers.ctypes.data, kmers.size, k, out.ctypes.data))
        return out

    def query_chunks(self, bases, chunk_off, chunk_len, k: int | None = None, mode: int = MODE_OR, output: int = OUT_PRESENCE,
                     strands: int = STRANDS_LAZY, streaming: bool = False) -> np.ndarray:
        k = self.primary.k if k is None else k
        b = np.frombuffer(bases, dtype=np.uint8) if isinstance(bases, (bytes, bytearray)) else np.ascontiguousarray(bases, dtype=np.uint8)
        off = _u64(chunk_off)
        ln = np.ascontiguousarray(chunk_len, dtype=np.uint32)
        n_res = int((ln.astype(np.int64) - k + 1).sum())
        dt, shape = result_dtype_shape(output, strands, n_res)
        out = np.empty(shape, dtype=dt)
        _check(lib().fmsi_gpu_pool_query_chunks(self._h, mode, output, strands, int(streaming), b.ctypes.data, b.size,
                                               off.ctypes.data, ln.ctypes.data, off.size, n_res, k, out.ctypes.data))
        return out


def load_index(prefix: str, use_klcp: bool = True, device: int = 0, **kw) -> Index:
    return Index.load(prefix, use_klcp, device, **kw)
