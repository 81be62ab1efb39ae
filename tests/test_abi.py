"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports exactly what
include/fmsi_gpu.h declares, and fails loudly (no CPU fallback) when there is no device."""
import os
import re
import subprocess

import pytest

from conftest import GOLDEN, ROOT

import fmsi_b200 as fg


def header_symbols():
    text = open(os.path.join(ROOT, "include", "fmsi_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fmsi_gpu_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(fg.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    path = fg.lib_path()
    assert os.path.exists(path), "libfmsi_gpu.so not built: run __graft_entry__.build()"
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (fmsi_gpu_[a-z_]+)", out))
    assert exported == set(fg.EXPORTED_SYMBOLS)
    lib = fg.lib()
    assert lib.fmsi_gpu_abi_version() == 1
    for name in fg.EXPORTED_SYMBOLS:
        assert getattr(lib, name) is not None


def test_library_contains_sm100a_code():
    out = subprocess.run(["cuobjdump", "--list-elf", fg.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_cpu_fallback_without_device():
    if fg.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fg.FmsiGpuError) as e:
        fg.Index.load(os.path.join(GOLDEN, "syn_k5_min", "ms.fa"))
    assert e.value.code == -3 and "no CUDA device" in str(e.value)


def test_missing_index_reports_io_error():
    with pytest.raises(fg.FmsiGpuError) as e:
        fg.Index.load("/nonexistent/prefix")
    assert e.value.code == -2


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under fmsi_b200/ may reference it."""
    for dp, _, fns in os.walk(os.path.join(ROOT, "fmsi_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                text = open(os.path.join(dp, fn), errors="replace").read()
                assert "fmsi_oracle" not in text and "oracle_ffi" not in text, os.path.join(dp, fn)
