"""The construction input from the reference's own files: .graph (binary) and .gcsa2 (text) kmer files and NodeMapping
files (gcsa2_b200/csrc/kmer_file.cpp), and an index built with a NodeMapping -- against the reference's own readers and
constructor (oracle/_ref) where that library is available."""
import numpy as np
import pytest

from brute import random_graph
from test_builder import flat_equal
from gcsa2_b200 import synth
from gcsa2_b200.builder import CharGraph, NodeMapping, build_from_kmers, enumerate_kmers, read_kmers
from oracle import reference as ref


def small_graphs():
    # (inputs of tests/test_reference.py: the reference's constructor is known to survive them -- on others, e.g. the
    # 8-mers of a 4 kbp graph with 3 % SNPs, it builds an index that its own verifyIndex() rejects)
    yield "snp", synth.snp_graph(synth.random_sequence(5000, seed=4), seed=4, snp_rate=0.05)[0], 16, 2
    yield "linear", synth.linear_graph(synth.random_sequence(6000, seed=5)), 16, 3
    g = random_graph(np.random.default_rng(7), 300, 8, snp_rate=0.15, node_len=4)
    yield "bubbles", CharGraph.from_lists(g.comps, g.values, g.succ, g.sources, g.sink), 2, 2


@pytest.mark.parametrize("name,graph,k,steps", list(small_graphs()), ids=[c[0] for c in small_graphs()])
def test_kmer_files_round_trip(tmp_path, name, graph, k, steps):
    kmers = enumerate_kmers(graph, k)
    binary, text = tmp_path / "input.graph", tmp_path / "input.gcsa2"
    kmers.write_binary(binary); kmers.write_text(text)
    from_binary = read_kmers(binary, binary=True)
    assert from_binary.k == k and (from_binary.key == kmers.key).all() and (from_binary.from_ == kmers.from_).all() and (from_binary.to == kmers.to).all()
    from_text = read_kmers(text, binary=False)
    assert from_text.k == k and from_text.key.size == kmers.key.size
    def canonical(km):
        order = np.lexsort((km.to, km.from_, km.key))
        return km.key[order], km.from_[order], km.to[order]
    for a, b in zip(canonical(from_text), canonical(kmers)):
        assert (a == b).all()
    # several files are concatenated (two sections in one binary file as well)
    with open(tmp_path / "twice.graph", "wb") as f:
        f.write(binary.read_bytes()); f.write(binary.read_bytes())
    twice = read_kmers([tmp_path / "twice.graph", binary], binary=True)
    assert twice.key.size == 3 * kmers.key.size and (twice.key[kmers.key.size:2 * kmers.key.size] == kmers.key).all()
    # the same index from either file
    a, la = build_from_kmers(from_binary, steps)
    b, lb = build_from_kmers(from_text, steps)
    assert flat_equal(a, b) == [] and (la.data == lb.data).all()


def test_kmer_file_errors(tmp_path):
    from gcsa2_b200 import capi
    bad = tmp_path / "bad.graph"
    bad.write_bytes(np.array([1, 0, 16], dtype=np.uint64).tobytes())               # flags != 0
    with pytest.raises(capi.GCSAError):
        read_kmers(bad)
    bad.write_bytes(np.array([0, 5, 16, 1, 2, 3], dtype=np.uint64).tobytes())      # promises 5 records, holds 1
    with pytest.raises(capi.GCSAError):
        read_kmers(bad)
    bad.write_bytes(np.array([0, 0, 17], dtype=np.uint64).tobytes())               # kmer length above Key::MAX_LENGTH
    with pytest.raises(capi.GCSAError):
        read_kmers(bad)
    with pytest.raises(capi.GCSAError):
        read_kmers(tmp_path / "missing.graph")
    text = tmp_path / "mixed.gcsa2"
    text.write_text("ACGT\t1:0\t$\tA\t1:1\nnot a kmer line\nACG\t1:1\tA\tC\t1:2\n")
    with pytest.raises(capi.GCSAError):                                            # kmer lengths 4 and 3
        read_kmers(text, binary=False)
    text.write_text("ACGT\t1:0\t$\tA,C\t1:1,2:-5\nshort line\n")
    km = read_kmers(text, binary=False)
    assert km.k == 4 and km.key.size == 2
    assert km.from_.tolist() == [1 << 11, 1 << 11] and km.to.tolist() == [(1 << 11) | 1, (2 << 11) | (1 << 10) | 5]
    assert int(km.key[0]) & 0xFFFF == ((1 << 0) << 8) | ((1 << 1) | (1 << 2))      # predecessor $, successors A and C


def duplicated_node_mapping(graph):
    """A mapping that sends the ids of the alternative alleles (the highest ids before the sink) onto low ids of the
    backbone, as if those nodes were duplicates of backbone nodes: several start positions of a path node collapse."""
    ids = (graph.value >> np.uint64(11)).astype(np.int64)
    first = int(np.sort(np.unique(ids))[len(np.unique(ids)) * 3 // 4])
    top = int(ids.max())
    rng = np.random.default_rng(5)
    return NodeMapping(first=first, ids=rng.integers(2, first, size=top - first + 1).astype(np.uint64))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libgcsa2_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("name,graph,k,steps", list(small_graphs()), ids=[c[0] for c in small_graphs()])
def test_reference_reads_our_files_and_mapped_index_matches(tmp_path, name, graph, k, steps):
    kmers = enumerate_kmers(graph, k)
    flat, flcp = build_from_kmers(kmers, steps)
    # the reference's own readText on our text file builds the same index as from the binary file
    from_text = ref.ReferenceIndex.build(kmers, steps, text=True)
    rflat, rlcp = from_text.export()
    assert flat_equal(flat, rflat) == [] and (rlcp.data == flcp.data).all()
    # NodeMapping: the reference's constructor with a mapping file == the builder with the same mapping
    mapping = duplicated_node_mapping(graph)
    mapping.write(tmp_path / "m.mapping")
    loaded = NodeMapping.load(tmp_path / "m.mapping")
    assert loaded.first == mapping.first and (loaded.ids == mapping.ids).all()
    mapped_ref = ref.ReferenceIndex.build(kmers, steps, mapping=mapping)
    mflat, mlcp = build_from_kmers(kmers, steps, mapping=mapping)
    rflat, rlcp = mapped_ref.export()
    assert flat_equal(mflat, rflat) == [] and (rlcp.data == mlcp.data).all()
    assert flat_equal(mflat, flat) != []                                            # the mapping changed the samples
    assert mapped_ref.verify()                                                     # the reference's verifyIndex with graph.mapping
