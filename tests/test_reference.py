"""Parity against the REFERENCE ITSELF: jltsiren/gcsa2's unmodified sources compiled against the SDSL
shim (oracle/_ref/libgcsa2_ref.so, built by `make -C oracle ref` where /root/reference exists; the
prebuilt library travels to the GPU box).

  * the reference's constructor, fed our kmer files, accepts them and its own verifyIndex() passes;
  * the host-side builder emits bit-identical arrays (GCSA members and the LCP tree);
  * the C restatement (oracle) answers every query exactly like the reference's own methods;
  * (gpu) so does the CUDA engine, compared directly with the reference.
"""
import numpy as np
import pytest

from brute import random_graph
from test_builder import flat_equal
from gcsa2_b200 import synth
from gcsa2_b200.builder import CharGraph, build_index
from oracle import oracle as orc
from oracle import reference as ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libgcsa2_ref.so not built (needs /root/reference)")
M64 = (1 << 64) - 1


def cases():
    seq = synth.random_sequence(20000, seed=3)
    yield "snp2", synth.snp_graph(seq, seed=3, snp_rate=0.02)[0], 16, 3
    yield "snp5", synth.snp_graph(synth.random_sequence(5000, seed=4), seed=4, snp_rate=0.05)[0], 16, 2
    yield "linear", synth.linear_graph(synth.random_sequence(30000, seed=5)), 16, 3
    unit = np.array([1, 2, 1, 3], dtype=np.uint8)
    yield "repeats", synth.linear_graph(np.concatenate([np.tile(unit, 40), synth.random_sequence(60, 6), np.tile(unit, 30)]), node_len=8), 4, 2
    rng = np.random.default_rng(7)
    g = random_graph(rng, 300, 8, snp_rate=0.15, node_len=4)
    yield "bubbles_k2", CharGraph.from_lists(g.comps, g.values, g.succ, g.sources, g.sink), 2, 2


@pytest.fixture(scope="module", params=list(cases()), ids=lambda c: c[0])
def built(request):
    name, graph, k, steps = request.param
    flat, flcp, kmers = build_index(graph, k, steps, lcp_branching=64 if name != "repeats" else 4)
    reference = ref.ReferenceIndex.build(kmers, steps, lcp_branching=64 if name != "repeats" else 4)
    return name, flat, flcp, kmers, reference


def test_reference_accepts_our_kmers_and_builder_matches_it(built):
    name, flat, flcp, kmers, reference = built
    assert reference.verify()                                  # verifyIndex(index, &lcp, graph), algorithms.cpp:85-99
    rflat, rlcp = reference.export()
    assert flat_equal(flat, rflat) == []
    assert (rlcp.size, rlcp.branching, rlcp.levels) == (flcp.size, flcp.branching, flcp.levels)
    assert list(rlcp.offsets) == list(flcp.offsets) and (rlcp.data == flcp.data).all()


def query_set(flat, kmers, seed=0):
    rng = np.random.default_rng(seed)
    from verify import kmer_table
    table = kmer_table(kmers)
    pats = [s for s, _ in table[:3000]]
    alphabet = np.frombuffer(b"ACGTACGTACGTacgtN$#x\0", dtype=np.uint8)
    for _ in range(2000):
        pats.append(bytes(alphabet[rng.integers(0, alphabet.size, size=int(rng.integers(0, 40)))]))
    chars, offsets = orc.pack_patterns(pats)
    N = flat.path_nodes
    a = rng.integers(0, N, size=3000).astype(np.uint64)
    b = np.minimum(a + rng.geometric(0.3, size=3000).astype(np.uint64) - np.uint64(1), np.uint64(N - 1))
    return chars, offsets, a, b


def test_oracle_equals_reference_on_every_operation(built):
    name, flat, flcp, kmers, reference = built
    ora, olcp = orc.OracleGCSA(flat), orc.OracleLCP(flcp)
    chars, offsets, a, b = query_set(flat, kmers)
    rsp, rep, _ = reference.find_batch(chars, offsets)
    osp, oep, _ = ora.find_batch(chars, offsets)
    assert (rsp == osp).all() and (rep == oep).all()
    comps = np.random.default_rng(1).integers(0, 7, size=a.size).astype(np.uint8)
    ra, rb = reference.lf_batch(a, b, comps)
    assert [(int(x), int(y)) for x, y in zip(ra, rb)] == [ora.LF((int(s), int(e)), int(c)) for s, e, c in zip(a, b, comps)]
    nodes = np.arange(min(flat.path_nodes, 5000), dtype=np.uint64)
    assert list(reference.lf_node_batch(nodes)) == [ora.LF(int(i)) for i in nodes]
    # count / locate on find() ranges and on arbitrary ranges
    fs, fe = rsp[:3000], rep[:3000]
    for s, e in ((fs, fe), (a, b)):
        oc, _ = ora.count_batch(s, e)
        assert (reference.count_batch(s, e) == oc).all()
        roffs, rvals, _ = reference.locate_batch(s, e)
        ooffs, ovals, _ = ora.locate_batch(s, e)
        assert (roffs == ooffs).all() and (rvals == ovals).all()
    moffs, mvals, _ = reference.locate_batch(fs[:500], fe[:500], max_positions=3)
    for i in range(500):
        assert list(mvals[int(moffs[i]):int(moffs[i + 1])]) == ora.locate((int(fs[i]), int(fe[i])), 3)
    # suffix-tree operations
    opar, _ = olcp.parent_batch(a, b)
    assert (reference.parent_batch(a, b) == opar).all()
    assert list(reference.depth_batch(a, b)) == [olcp.depth((int(s), int(e))) for s, e in zip(a, b)]
    pos = np.arange(min(flat.path_nodes, 4000), dtype=np.uint64)
    for which in ("psv", "psev", "nsv", "nsev"):
        p, v = reference.lcp_query(which, pos)
        assert [(int(x), int(y)) for x, y in zip(p, v)] == [getattr(olcp, which)(int(i)) for i in pos], which
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    p, v = reference.lcp_query("rmq", lo, hi)
    assert [(int(x), int(y)) for x, y in zip(p, v)] == [olcp.rmq(int(s), int(e)) for s, e in zip(lo, hi)]
    for k in (1, 3, 7, 12):
        assert reference.count_kmers(k) == ora.count_kmers(k)
        assert reference.count_kmers(k, include_Ns=True) == ora.count_kmers(k, include_Ns=True)


def test_reference_loaded_from_flat_arrays_answers_like_its_own_build(built):
    name, flat, flcp, kmers, reference = built
    loaded = ref.ReferenceIndex.from_flat(flat)                # what bench.py times as the CPU baseline
    chars, offsets, a, b = query_set(flat, kmers, seed=2)
    s1, e1, _ = reference.find_batch(chars, offsets)
    s2, e2, _ = loaded.find_batch(chars, offsets, threads=2)
    assert (s1 == s2).all() and (e1 == e2).all()


@pytest.mark.engine
def test_engine_equals_reference(built):
    from gcsa2_b200 import GCSA, LCPArray
    name, flat, flcp, kmers, reference = built
    chars, offsets, a, b = query_set(flat, kmers, seed=3)
    rsp, rep, _ = reference.find_batch(chars, offsets)
    for table_k, two_step in ((0, False), (4, True)):
        gpu = GCSA(flat, kmer_table_k=table_k, two_step=two_step)
        sp, ep = gpu.find_batch(chars, offsets)
        assert (sp == rsp).all() and (ep == rep).all()
    fs, fe = rsp[:3000], rep[:3000]
    for s, e in ((fs, fe), (a, b)):
        assert (gpu.count_batch(s, e) == reference.count_batch(s, e)).all()
        offs, vals = gpu.locate_batch(s, e)
        roffs, rvals, _ = reference.locate_batch(s, e)
        assert (offs == roffs).all() and (vals == rvals).all()
    comps = np.random.default_rng(1).integers(0, 7, size=a.size).astype(np.uint8)
    ga, gb = gpu.lf_batch(a, b, comps)
    ra, rb = reference.lf_batch(a, b, comps)
    assert (ga == ra).all() and (gb == rb).all()
    glcp = LCPArray(flcp)
    assert (glcp.parent_batch(a, b) == reference.parent_batch(a, b)).all()
    assert (glcp.depth_batch(a, b) == reference.depth_batch(a, b)).all()
    for k in (1, 5, 9):
        assert gpu.count_kmers(k) == reference.count_kmers(k)
