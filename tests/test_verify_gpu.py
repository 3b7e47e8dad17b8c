"""gcsa_b200_verify_index: the reference's verifyIndex() (src/algorithms.cpp:101-295) batched on the device.
A correct index passes (as it passes the reference's own verifyIndex and the Python restatement in
verify.py); an index that does not belong to the kmers is caught at the stage the reference would catch it."""
import copy

import numpy as np
import pytest

from gcsa2_b200 import synth
from gcsa2_b200.builder import KMers, build_index
from gcsa2_b200.flat import bits_from_positions, positions_from_bits
from oracle import reference as ref

pytestmark = pytest.mark.engine


def graphs():
    yield "snp", synth.snp_graph(synth.random_sequence(30000, seed=21), seed=21, snp_rate=0.03)[0], 16, 2
    yield "linear", synth.linear_graph(synth.random_sequence(20000, seed=12)), 16, 3
    unit = np.array([1, 2, 1, 3], dtype=np.uint8)
    yield "repeats", synth.linear_graph(np.concatenate([np.tile(unit, 50), synth.random_sequence(40, 13)]), node_len=8), 4, 2


@pytest.mark.parametrize("case", list(graphs()), ids=lambda c: c[0])
def test_correct_index_passes(case):
    from gcsa2_b200 import GCSA, LCPArray
    name, graph, k, steps = case
    flat, flcp, kmers = build_index(graph, k, steps)
    gpu, glcp = GCSA(flat, kmer_table_k=4), LCPArray(flcp)
    rep = gpu.verify(kmers, glcp)
    labels = np.unique(kmers.key >> np.uint64(16)).size
    assert rep["failures"] == 0 and rep["unique"] == labels, rep
    assert gpu.verify(kmers)["failures"] == 0                          # lcp == 0: parent / depth skipped
    if ref.available():
        assert ref.ReferenceIndex.build(kmers, steps).verify()         # the reference agrees on the same input


def test_wrong_index_is_caught():
    from gcsa2_b200 import GCSA, LCPArray
    seq = synth.random_sequence(20000, seed=31)
    flat, flcp, kmers = build_index(synth.snp_graph(seq, seed=31, snp_rate=0.02)[0], 16, 2)
    gpu, glcp = GCSA(flat, kmer_table_k=4), LCPArray(flcp)

    # kmers the index does not contain: find() fails
    other = build_index(synth.linear_graph(synth.random_sequence(3000, seed=32)), 16, 2)[2]
    rep = gpu.verify(other)
    assert rep["find_failures"] > 0.9 * rep["unique"]

    # start nodes shifted: count() still right, locate() wrong for every label
    shifted = KMers(key=kmers.key.copy(), from_=kmers.from_ + np.uint64(1 << 20), to=kmers.to.copy(), k=kmers.k)
    rep = gpu.verify(shifted, glcp)
    assert rep["find_failures"] == rep["parent_failures"] == rep["depth_failures"] == rep["count_failures"] == 0
    assert rep["locate_failures"] == rep["unique"] == rep["failures"]

    # one occurrence added to every label: count() fails first
    extra = KMers(key=np.concatenate([kmers.key, kmers.key]), from_=np.concatenate([kmers.from_, kmers.from_ + np.uint64(1 << 30)]),
                  to=np.concatenate([kmers.to, kmers.to]), k=kmers.k)
    rep = gpu.verify(extra)
    assert rep["count_failures"] == rep["unique"] and rep["locate_failures"] == 0

    # an LCP array of another index: parent() / depth() fail, the rest is not reached for those labels
    wrong_lcp = copy.deepcopy(flcp)
    wrong_lcp.data = np.minimum(wrong_lcp.data, 3).astype(np.uint8)
    rep = gpu.verify(kmers, LCPArray(wrong_lcp))
    assert rep["parent_failures"] + rep["depth_failures"] > 0 and rep["find_failures"] == 0

    # samples tampered with: locate() differs for some labels, count() (Sadakane counters) does not
    bad = copy.deepcopy(flat)
    bad.stored_samples = bad.stored_samples.copy()
    bad.stored_samples[::7] += np.uint64(1)
    rep = GCSA(bad, kmer_table_k=4).verify(kmers)
    assert rep["count_failures"] == 0 and rep["locate_failures"] > 0


def test_index_with_node_mapping_verifies_with_the_mapping():
    """An index built with a NodeMapping (InputGraph::mapping: duplicated nodes of the input graph reported under the
    ids of the original graph) passes verifyIndex with that mapping and fails it without (locate() reports mapped ids)."""
    from gcsa2_b200 import GCSA, LCPArray
    from gcsa2_b200.builder import NodeMapping, build_from_kmers, enumerate_kmers
    graph = synth.snp_graph(synth.random_sequence(5000, seed=4), seed=4, snp_rate=0.05)[0]
    kmers = enumerate_kmers(graph, 16)
    ids = np.unique(graph.value >> np.uint64(11)).astype(np.int64)
    first = int(ids[len(ids) * 3 // 4])
    mapping = NodeMapping(first=first, ids=np.random.default_rng(5).integers(2, first, size=int(ids.max()) - first + 1).astype(np.uint64))
    flat, flcp = build_from_kmers(kmers, 2, mapping=mapping)
    gpu, glcp = GCSA(flat, kmer_table_k=4), LCPArray(flcp)
    rep = gpu.verify(kmers, glcp, mapping=mapping)
    assert rep["failures"] == 0, rep
    rep = gpu.verify(kmers, glcp)
    assert rep["find_failures"] == rep["parent_failures"] == 0 and rep["locate_failures"] + rep["count_failures"] > 0
    # every located value is a mapped one
    sp, ep = gpu.find_batch(["ACGT", "GGA", "T"])
    _, vals = gpu.locate_batch(sp, ep)
    assert ((vals >> np.uint64(11)).astype(np.int64) < first).all() or int((vals >> np.uint64(11)).max()) >= first + len(mapping.ids)
