"""Build every native artefact in-tree (no JIT cache): libfmsi_gpu.so (sm_100a), the fmsi CLI,
the design microbenchmark, and the test-infrastructure oracle (+ the reference binaries when
/root/reference is present). Called by __graft_entry__.build()."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfmsi_gpu.so")
BIN = os.path.join(HERE, "bin")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources() -> list[str]:
    out = [os.path.join(ROOT, "include", "fmsi_gpu.h")]
    for dp, _, fns in os.walk(CSRC):
        out += [os.path.join(dp, f) for f in fns]
    return out


def _run(cmd: list[str], verbose: bool) -> None:
    if verbose:
        print("+", " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)


def build_lib(force: bool = False, verbose: bool = True) -> str:
    if force or _newer(LIB, _sources()):
        _run([_nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB, os.path.join(CSRC, "fmsi_gpu.cu")], verbose)
    return LIB


def build_variant(name: str, defines: list[str], verbose: bool = True) -> str:
    """A/B variants of the library (fmsi_b200/variants/, git-ignored; selected with $FMSI_GPU_LIB)."""
    d = os.path.join(HERE, "variants")
    os.makedirs(d, exist_ok=True)
    out = os.path.join(d, f"libfmsi_gpu_{name}.so")
    if _newer(out, _sources()):
        _run([_nvcc(), *NVCC_FLAGS, *[f"-D{x}" for x in defines], "-shared", "-o", out, os.path.join(CSRC, "fmsi_gpu.cu")], verbose)
    return out


def build_cli(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(BIN, exist_ok=True)
    exe = os.path.join(BIN, "fmsi")
    src = os.path.join(CSRC, "fmsi_cli.cpp")
    if os.path.exists(src) and (force or _newer(exe, _sources() + [LIB])):
        _run(["g++", "-std=c++17", "-O2", "-Wall", "-pthread", "-I", os.path.join(ROOT, "include"), "-o", exe, src,
              "-L", HERE, "-lfmsi_gpu", "-lz", "-Wl,-rpath,$ORIGIN/.."], verbose)
    return exe


def build_tools(force: bool = False, verbose: bool = True) -> None:
    os.makedirs(BIN, exist_ok=True)
    src = os.path.join(CSRC, "tools", "randbw.cu")
    exe = os.path.join(BIN, "randbw")
    if force or _newer(exe, [src]):
        _run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-o", exe, src], verbose)
    for tool in ("randbw2", "locbw"):  # further design microbenchmarks (profiles/README.md)
        src = os.path.join(CSRC, "tools", tool + ".cu")
        exe = os.path.join(BIN, tool)
        if force or _newer(exe, [src]):
            _run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-o", exe, src], verbose)
    src = os.path.join(CSRC, "tools", "loc_key_check.cu")  # host-only check of loc.cuh's key function (tests/test_loc_key.py)
    exe = os.path.join(BIN, "loc_key_check")
    if force or _newer(exe, [src, os.path.join(CSRC, "loc.cuh"), os.path.join(CSRC, "fold.cuh")]):
        _run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-o", exe, src], verbose)
    src = os.path.join(CSRC, "tools", "fasta_dump.cpp")  # input-parser test tool (tests/test_fasta_blocks.py)
    exe = os.path.join(BIN, "fasta_dump")
    if force or _newer(exe, [src, os.path.join(CSRC, "fasta_blocks.hpp")]):
        _run(["g++", "-std=c++17", "-O2", "-Wall", "-o", exe, src, "-lz"], verbose)
    src = os.path.join(CSRC, "tools", "layout_dump.cpp")  # record-layout test tool (tests/test_layout.py); includes fmsi_cli.cpp
    exe = os.path.join(BIN, "layout_dump")
    if os.path.exists(LIB) and (force or _newer(exe, [src, os.path.join(CSRC, "fmsi_cli.cpp"), os.path.join(CSRC, "fasta_blocks.hpp"), LIB])):
        _run(["g++", "-std=c++17", "-O2", "-Wall", "-pthread", "-I", os.path.join(ROOT, "include"), "-o", exe, src,
              "-L", HERE, "-lfmsi_gpu", "-lz", "-Wl,-rpath,$ORIGIN/.."], verbose)
    src = os.path.join(CSRC, "tools", "predictor_check.cpp")  # replay-split test tool (tests/test_predictor.py)
    exe = os.path.join(BIN, "predictor_check")
    if force or _newer(exe, [src, os.path.join(CSRC, "predictor.hpp")]):
        _run(["g++", "-std=c++17", "-O2", "-Wall", "-o", exe, src], verbose)
    src = os.path.join(CSRC, "tools", "rrr_roundtrip.cpp")  # sdsl format code test tool (tests/test_sdsl_io.py)
    exe = os.path.join(BIN, "rrr_roundtrip")
    if force or _newer(exe, [src, os.path.join(CSRC, "sdsl_io.hpp")]):
        _run(["g++", "-std=c++17", "-O2", "-Wall", "-pthread", "-o", exe, src], verbose)


def build_wide_synth(force: bool = False, verbose: bool = True) -> str:
    """Synthetic-index generator for the N >= 2^32 tests (tests/test_gpu_wide.py)."""
    src = os.path.join(CSRC, "tools", "wide_synth.cpp")
    exe = os.path.join(BIN, "wide_synth")
    if force or _newer(exe, [src, os.path.join(CSRC, "sdsl_io.hpp")]):
        _run(["g++", "-std=c++17", "-O2", "-Wall", "-pthread", "-o", exe, src], verbose)
    return exe


def build_oracle(verbose: bool = True) -> None:
    """Test infrastructure: the C restatement, and the reference binaries when their sources exist."""
    _run(["make", "-C", os.path.join(ROOT, "oracle"), "-j4", "all"] + ([] if verbose else ["-s"]), verbose)


def build_all(force: bool = False, verbose: bool = True) -> None:
    build_lib(force, verbose)
    build_cli(force, verbose)
    build_tools(force, verbose)
    build_wide_synth(force, verbose)
    build_oracle(verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
