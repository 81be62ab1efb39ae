"""fmsi_b200 — B200-native query engine for FMSI's FMS-index.

The product is `libfmsi_gpu.so` (hand-written sm_100a kernels behind the C-ABI of
include/fmsi_gpu.h) plus the `fmsi` command-line front end (fmsi_b200/bin/fmsi). This package is
the thin ctypes binding used by tests and bench.py; it mirrors the names of the reference's
functions (load_index, query_kmers, rank, update_range, ...). There is no CPU fallback: importing
works without a GPU, any call needs the library and a device.
"""
from .api import (  # noqa: F401
    FmsiGpuError, Function, Index, Pool, load_index, lib, lib_path, MODE_OR, MODE_ALL, OUT_PRESENCE, OUT_ORDERS,
    STRANDS_LAZY, STRANDS_BOTH, MEM_HOST, MEM_DEVICE, EXPORTED_SYMBOLS, launch_count, device_count,
    OUT_PRESENCE_BITS, TEXT_ASCII, TEXT_PACKED2, pack_text,
)

__all__ = [
    "FmsiGpuError", "Function", "Index", "Pool", "load_index", "lib", "lib_path", "MODE_OR", "MODE_ALL", "OUT_PRESENCE",
    "OUT_ORDERS", "STRANDS_LAZY", "STRANDS_BOTH", "MEM_HOST", "MEM_DEVICE", "EXPORTED_SYMBOLS",
    "launch_count", "device_count", "OUT_PRESENCE_BITS", "TEXT_ASCII", "TEXT_PACKED2", "pack_text",
]
