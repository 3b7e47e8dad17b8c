"""Single-process multi-GPU entry points on two real devices against one device (skipped on a box with one GPU;
the same calls over replicas on one device run in tests/test_gpu_parity.py)."""
import numpy as np
import pytest

from gcsa2_b200 import GCSA, MultiGCSA, capi, synth
from gcsa2_b200.builder import build_index


@pytest.mark.gpu
def test_two_devices_equal_one_device():
    if capi.lib().gcsa_b200_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    seq = synth.random_sequence(2_000_000, seed=61)
    graph, sites, alt = synth.snp_graph(seq, seed=61, snp_rate=0.01)
    flat, _, _ = build_index(graph, 16, 3)
    one = GCSA(flat, device=0, kmer_table_k=12)
    two = MultiGCSA(flat, [0, 1], kmer_table_k=12)
    chars, _ = synth.patterns_from_snp_graph(seq, sites, alt, 2_000_001, 32, seed=62)
    a, b = one.find_fixed_batch(chars, 32)
    c, d = two.find_fixed_batch(chars, 32)
    assert (a == c).all() and (b == d).all()
    mchars, moffsets = synth.mixed_length_patterns(seq, sites, alt, 300_001, 12, 120, seed=63, error_rate=0.01)
    e, f = one.find_batch(mchars, moffsets)
    g, h = two.find_batch(mchars, moffsets)
    assert (e == g).all() and (f == h).all()
    offs, vals = one.locate_batch(a[:1_000_000], b[:1_000_000])
    moffs, mvals = two.locate_batch(a[:1_000_000], b[:1_000_000])
    assert (offs == moffs).all() and (vals == mvals).all()
    one.close(); two.close()
