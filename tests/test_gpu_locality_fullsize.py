"""The minimizer-bucketed dictionary (fmsi_b200/csrc/loc.cuh, `locality = 1`) at FULL size (BASELINE configs[3]: 3.1 Gbp,
k = 31): two exact tiers with nothing in common but the rows' states answer the same questions — reads go through the
tile kernel and its directory / rows, the k-mers cut out of the same reads through the one-probe hash table — so they
must agree k-mer for k-mer, on reads from the genome (either strand, 1 % substitutions: present and absent k-mers next
to each other, minimizers that cross a substitution) and on random reads, in both presence modes and strand policies.
The small-index versions of the same tiers face the oracle in tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest

from conftest import ROOT  # noqa: F401

import fmsi_b200 as fg
from fmsi_b200 import synth

pytestmark = pytest.mark.gpu

N_GENOME = int(os.environ.get("FMSI_TEST_FULLSIZE", 3_100_000_000))
K = 31


@pytest.fixture(scope="module")
def human_loc():
    torch = pytest.importorskip("torch")
    from bench import device_genome
    dev = torch.device("cuda", 0)
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 170e9 * (N_GENOME / 3.1e9):
        pytest.skip("not enough free device memory for the full-size index")
    codes, ascii_ = device_genome(N_GENOME, 4, K, dev)
    gi = fg.Index.build(ascii_.data_ptr(), K, with_klcp=True, device=0, n=N_GENOME, mem=fg.MEM_DEVICE, dict=2, locality=1)
    del ascii_
    torch.cuda.empty_cache()
    yield torch, dev, codes, gi
    gi.close()


def test_full_size_locality_tier_equals_one_probe_tier(human_loc):
    torch, dev, codes, gi = human_loc
    assert gi.dict_kind == 2 and gi.locality == 16 and not gi.wide
    head = codes[:200_000_000].cpu().numpy()  # reads from the first 200 Mbp (the whole genome is indexed)
    reads = synth.read_queries(head, 150, 60_000, 5)
    rng = np.random.default_rng(6)
    reads = np.concatenate([reads, rng.integers(0, 4, size=(4_000, 150), dtype=np.uint8)])
    texts = [synth.codes_to_ascii(r) for r in reads]
    texts += [synth.codes_to_ascii(head[7:7 + K]), synth.codes_to_ascii(head[1000:1000 + K + 31]), synth.codes_to_ascii(head[5000:5000 + 2000])]  # 1, 32 and 1970 k-mers
    kmers = np.concatenate([synth.pack_kmers(synth.ascii_to_codes(t), K) for t in texts])
    present = None
    for mode in (fg.MODE_ALL, fg.MODE_OR):
        for strands in (fg.STRANDS_LAZY, fg.STRANDS_BOTH):
            a = gi.query_reads(texts, K, mode, fg.OUT_PRESENCE, strands, False)  # text-derived queries: the minimizer-bucketed tier
            b = gi.query_kmers(kmers, K, mode, fg.OUT_PRESENCE, strands)        # packed k-mers: the one-probe tier
            assert a.shape == b.shape and np.array_equal(a, b), (mode, strands, int((a != b).sum()))
            if strands == fg.STRANDS_LAZY:
                present = a
    assert 0.55 < present[:60_000 * 120].mean() < 0.85   # 1 % substitutions take out about a quarter of the k-mers
    assert present[60_000 * 120:64_000 * 120].sum() <= 2  # random reads: essentially nothing
    assert present[64_000 * 120:].all()                   # unmutated stretches of the genome
    # bit-packed output and streamed calls take the same route
    bits = gi.query_reads(texts, K, fg.MODE_ALL, fg.OUT_PRESENCE_BITS, fg.STRANDS_LAZY, True)
    assert np.array_equal(bits, np.packbits(gi.query_reads(texts, K, fg.MODE_ALL, fg.OUT_PRESENCE, fg.STRANDS_LAZY, False), bitorder="little"))
