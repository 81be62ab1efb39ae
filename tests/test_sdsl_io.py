"""The product's host-side sdsl format code (fmsi_b200/csrc/sdsl_io.hpp: RRR<63> reader, decoder and the
multi-threaded encoder behind `fmsi index` / fmsi_gpu_index_save), without a GPU: decode a mask file, encode the
bits again, require the same bytes. Inputs: the reference-written .mask files under tests/golden/ and larger masks
serialised by the oracle's coder (itself pinned to those reference files in test_oracle.py), shaped to reach the
encoder's corners: complemented superblocks, uniform blocks, sizes that are multiples of 63 and of 63 * 32 (dummy
block), and enough superblocks for every worker thread to own a range."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_cases
from oracle_ffi import rrr_serialize

TOOL = os.path.join(ROOT, "fmsi_b200", "bin", "rrr_roundtrip")


def roundtrip(path, out, threads=None):
    env = dict(os.environ)
    if threads is not None:
        env["FMSI_GPU_THREADS"] = str(threads)
    r = subprocess.run([TOOL, path, out], capture_output=True, env=env)
    assert r.returncode == 0, r.stderr.decode()
    n, ones = (int(x) for x in r.stdout.split())
    return n, ones, open(out, "rb").read()


@pytest.mark.parametrize("case", golden_cases())
def test_roundtrip_of_reference_mask_files(case, tmp_path):
    src = os.path.join(GOLDEN, case, "ms.fa.fmsi.mask")
    want = open(src, "rb").read()
    for threads in (1, 5):
        _, _, got = roundtrip(src, str(tmp_path / "out.mask"), threads)
        assert got == want, (case, threads)


def synthetic_masks():
    rng = np.random.default_rng(5)
    sb = 63 * 32
    yield "dense_runs", np.repeat(rng.random(3000) < 0.8, rng.integers(1, 400, 3000))           # complemented superblocks
    yield "sparse", (rng.random(700_001) < 0.03)
    yield "half", (rng.random(64 * sb * 7 + 17) < 0.5)                                             # > 64 superblocks per thread range
    yield "all_ones_multiple_of_superblock", np.ones(sb * 130, dtype=bool)                        # dummy block + uniform blocks
    yield "multiple_of_63", (rng.random(63 * 1001) < 0.6)
    yield "zeros_then_ones", np.concatenate([np.zeros(sb * 70 + 5, dtype=bool), np.ones(sb * 70 + 11, dtype=bool)])
    yield "tiny", np.array([1, 0, 1], dtype=bool)


@pytest.mark.parametrize("name,bits", list(synthetic_masks()), ids=[n for n, _ in synthetic_masks()])
def test_threaded_encoder_matches_oracle_coder(name, bits, tmp_path):
    bits = np.asarray(bits, dtype=np.uint8)
    want = rrr_serialize(bits)
    src = tmp_path / "in.mask"
    src.write_bytes(want)
    for threads in (1, 3, 16):
        n, ones, got = roundtrip(str(src), str(tmp_path / "out.mask"), threads)
        assert (n, ones) == (bits.size, int(bits.sum())), (name, threads)
        assert got == want, (name, threads)
