"""The device builder for linear references (gcsa2_b200/csrc/linear_builder.cu) against the host builder
(builder.cpp, itself pinned bit-for-bit to the reference's constructor in tests/test_reference.py): every array of the
index and the LCP array must be identical.  Marked `engine`: runs on the host emulation here and on the GPU there."""
import numpy as np
import pytest

from gcsa2_b200 import synth
from gcsa2_b200.builder import build_index, build_linear
from gcsa2_b200.flat import SIGMA, words_for


def assert_same_index(a, b, la, lb):
    assert (a.path_nodes, a.edge_count, a.order) == (b.path_nodes, b.edge_count, b.order)
    assert (np.asarray(a.C) == np.asarray(b.C)).all()
    def same_bits(x, y, n_bits):
        w = words_for(n_bits)
        return (np.asarray(x)[:w] == np.asarray(y)[:w]).all()
    for c in range(SIGMA):
        assert same_bits(a.bwt[c], b.bwt[c], a.path_nodes), "bwt[%d]" % c
    assert same_bits(a.edges, b.edges, a.edge_count), "edges"
    assert same_bits(a.sampled_paths, b.sampled_paths, a.path_nodes), "sampled_paths"
    assert a.sample_count == b.sample_count
    assert (a.stored_samples == b.stored_samples).all(), "stored_samples"
    assert same_bits(a.samples, b.samples, a.sample_count), "samples"
    assert same_bits(a.extra_filter, b.extra_filter, a.path_nodes), "extra_filter"
    assert a.extra_values_len == b.extra_values_len
    assert same_bits(a.extra_values, b.extra_values, a.extra_values_len), "extra_values"
    assert a.redundant_len == b.redundant_len
    assert same_bits(a.redundant, b.redundant, a.redundant_len), "redundant"
    assert (la.offsets == lb.offsets).all() and (la.data == lb.data).all(), "lcp"


def sequences():
    rng = np.random.default_rng(5)
    yield "random-3000", synth.random_sequence(3000, seed=1)
    yield "random-1", synth.random_sequence(1, seed=2)
    yield "random-17", synth.random_sequence(17, seed=3)
    # repeats longer than the order: nodes with several values, several predecessors, out-degree > 1
    unit = synth.random_sequence(37, seed=4)
    yield "tandem", np.concatenate([synth.random_sequence(50, seed=5), np.tile(unit, 12), synth.random_sequence(40, seed=6), np.tile(unit, 7)])
    yield "homopolymer", np.concatenate([np.full(700, 1, dtype=np.uint8), synth.random_sequence(30, seed=7), np.full(300, 1, dtype=np.uint8)])
    two = synth.random_sequence(400, seed=8)
    yield "copy", np.concatenate([two, synth.random_sequence(3, seed=9), two, two[:200]])
    with_n = synth.random_sequence(1500, seed=10).copy()
    with_n[rng.integers(0, 1500, size=40)] = 5
    with_n[600:800] = 5
    yield "with-N", with_n
    yield "binary", rng.integers(1, 3, size=2500).astype(np.uint8)


@pytest.mark.engine
@pytest.mark.parametrize("name,seq", list(sequences()), ids=[n for n, _ in sequences()])
def test_linear_builder_matches_host_builder(name, seq):
    for k, steps, node_len, period in ((16, 3, 32, 64), (4, 2, 32, 64), (8, 4, 7, 16), (16, 0, 1024, 64), (3, 1, 5, 3)):
        flat_h, lcp_h, _ = build_index(synth.linear_graph(seq, node_len=node_len), k, steps, sample_period=period)
        flat_d, lcp_d = build_linear(seq, k=k, doubling_steps=steps, node_len=node_len, sample_period=period)
        assert_same_index(flat_d, flat_h, lcp_d, lcp_h)


@pytest.mark.engine
def test_linear_builder_rejects_bad_input():
    from gcsa2_b200 import capi
    with pytest.raises(capi.GCSAError):
        build_linear(np.array([1, 2, 0, 3], dtype=np.uint8))          # the endmarker inside the sequence
    with pytest.raises(capi.GCSAError):
        build_linear(np.array([1, 2, 6, 3], dtype=np.uint8))
    with pytest.raises(capi.GCSAError):
        build_linear(synth.random_sequence(100, seed=1), k=16, doubling_steps=4)      # order 256
