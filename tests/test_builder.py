"""The host-side builder (gcsa2_b200/csrc/builder.cpp) against the brute-force construction and
against the reference's verifyIndex predicates evaluated by the oracle (CPU only)."""
import os

import numpy as np
import pytest

from brute import BruteIndex, SimpleGraph, random_graph
from helpers import load_kat1
from verify import kmer_table, verify_index
from gcsa2_b200 import synth
from gcsa2_b200.builder import CharGraph, build_index, enumerate_kmers
from gcsa2_b200.flat import SIGMA, positions_from_bits
from oracle import oracle as orc


def flat_equal(a, b):
    diffs = [name for name in ("path_nodes", "edge_count", "order", "sample_count", "extra_values_len", "redundant_len")
             if getattr(a, name) != getattr(b, name)]
    if diffs:
        return diffs
    N = a.path_nodes
    if list(a.C) != list(b.C):
        diffs.append("C")
    for c in range(SIGMA):
        if list(positions_from_bits(a.bwt[c], N)) != list(positions_from_bits(b.bwt[c], N)):
            diffs.append("bwt%d" % c)
    for name, nb in (("edges", a.edge_count), ("sampled_paths", N), ("samples", a.sample_count), ("extra_filter", N),
                     ("extra_values", a.extra_values_len), ("redundant", a.redundant_len)):
        if list(positions_from_bits(getattr(a, name), nb)) != list(positions_from_bits(getattr(b, name), nb)):
            diffs.append(name)
    if list(a.stored_samples) != list(b.stored_samples):
        diffs.append("stored_samples")
    return diffs


def test_builder_reproduces_paper_figure():
    """Figure 2's graph in, Figure 3's GCSA out (keys, values, BWT, C, edges)."""
    kat = load_kat1()
    comp = {"$": 0, "A": 1, "C": 2, "G": 3, "T": 4, "N": 5, "#": 6}
    M1, M2 = (1 << 64) - 1, (1 << 64) - 2
    comps = [6, 6] + [comp[ch] for ch in kat["graph"]["labels"]]      # the figure's source spans 3 positions
    values = [M2, M1] + list(range(12))
    succ = [[] for _ in comps]
    succ[0], succ[1] = [1], [2]
    for a, b in kat["graph"]["edges"]:
        succ[a + 2].append(b + 2)
    g = SimpleGraph(comps=comps, values=values, succ=succ, sources=[0], sink=13)
    brute = BruteIndex(g, 3, sample_period=1 << 40)
    assert brute.consistent
    assert ["".join("$ACGTN#"[x] for x in key) for key in brute.keys] == ["$"] + kat["keys"][1:]
    assert brute.node_values == kat["values"]
    assert [sorted(s) for s in brute.bwt_sets] == [kat["bwt"][str(c)] for c in range(SIGMA)]
    assert list(brute.flat.C) == kat["C"]
    ones = set(positions_from_bits(brute.flat.edges, brute.flat.edge_count).tolist())
    assert "".join("1" if i in ones else "0" for i in range(brute.flat.edge_count)) == kat["edges"]
    flat, lcp, _ = build_index(CharGraph.from_lists(comps, values, succ, [0], 13), 3, 0, sample_period=1 << 40, lcp_branching=4)
    assert flat_equal(flat, brute.flat) == []
    assert list(lcp.data[:flat.path_nodes]) == brute.lcp


@pytest.mark.parametrize("seed", range(6))
def test_builder_matches_bruteforce(seed):
    rng = np.random.default_rng(seed)
    checked = 0
    for trial in range(40):
        length = int(rng.integers(5, 60))
        k, steps = [(1, 2), (2, 1), (2, 2), (1, 3), (3, 1), (4, 1), (2, 0)][trial % 7]
        g = random_graph(rng, length, k << steps, snp_rate=float(rng.choice([0.0, 0.1, 0.3])),
                         node_len=int(rng.integers(2, 6)), alphabet=[(1, 2, 3, 4), (1, 2)][trial % 2])
        period = int(rng.choice([4, 64, 1 << 30]))
        brute = BruteIndex(g, k << steps, sample_period=period)
        if not brute.consistent:
            continue
        flat, lcp, _ = build_index(CharGraph.from_lists(g.comps, g.values, g.succ, g.sources, g.sink), k, steps,
                                   sample_period=period, lcp_branching=4)
        assert flat_equal(flat, brute.flat) == [], (seed, trial)
        assert list(lcp.data[:flat.path_nodes]) == brute.lcp
        checked += 1
    assert checked >= 30


def _verify_with_oracle(graph, k, steps, limit=None, branching=64):
    flat, lcp, kmers = build_index(graph, k, steps, lcp_branching=branching)
    index, olcp = orc.OracleGCSA(flat), orc.OracleLCP(lcp)
    fails = verify_index(index, olcp, kmer_table(kmers), limit=limit)
    assert fails == [], fails[:5]
    return flat


def test_verify_index_random_graphs():
    rng = np.random.default_rng(11)
    for trial in range(12):
        g = random_graph(rng, int(rng.integers(20, 200)), 8, snp_rate=0.1, node_len=4)
        _verify_with_oracle(CharGraph.from_lists(g.comps, g.values, g.succ, g.sources, g.sink), 2, 2, branching=[2, 4, 64][trial % 3])


def test_verify_index_config1_linear_10kbp():
    """BASELINE.json configs[0]: 16-mers over a 10 kbp linear path (k = 16, one doubling step:
    the reference clamps doubling steps to >= 1, src/support.cpp:104-107, so the order is 32)."""
    seq = synth.random_sequence(10000, seed=1)
    flat = _verify_with_oracle(synth.linear_graph(seq, node_len=32), 16, 1, limit=1500)
    assert flat.path_nodes == 10002 and flat.order == 32


def test_verify_index_snp_graph_order128():
    seq = synth.random_sequence(30000, seed=3)
    graph, sites, alt = synth.snp_graph(seq, seed=3, snp_rate=0.02)
    flat = _verify_with_oracle(graph, 16, 3, limit=1500)
    assert flat.order == 128 and flat.path_nodes > 30002 and flat.edge_count > flat.path_nodes
    assert flat.redundant_len > flat.path_nodes - 1        # positions reachable through several path nodes


def test_low_complexity_repeats_multi_value_nodes():
    """Tandem repeats: many nodes with several values, redundant pointers, long LCPs."""
    unit = np.array([1, 2, 1, 3], dtype=np.uint8)
    seq = np.concatenate([np.tile(unit, 40), synth.random_sequence(50, 5), np.tile(unit, 30)])
    flat = _verify_with_oracle(synth.linear_graph(seq, node_len=8), 4, 2, branching=4)
    assert flat.extra_values_len > 0 and flat.path_nodes < seq.size


def test_kmers_binary_file_layout(tmp_path):
    seq = synth.random_sequence(50, seed=9)
    kmers = enumerate_kmers(synth.linear_graph(seq), 8)
    path = os.path.join(tmp_path, "x.graph")
    kmers.write_binary(path)
    raw = np.fromfile(path, dtype=np.uint64)
    assert raw[0] == 0 and raw[1] == kmers.key.size and raw[2] == 8          # GraphFileHeader, files.h:40-52
    assert raw.size == 3 + 3 * kmers.key.size
    assert (raw[3::3] == kmers.key).all() and (raw[4::3] == kmers.from_).all() and (raw[5::3] == kmers.to).all()


def test_config1_count_kmers_cpu_plumbing():
    """BASELINE.json configs[0]: countKMers (src/algorithms.cpp:387-421) on the 10 kbp linear path;
    the answer is the number of distinct k-mers of the path."""
    seq = synth.random_sequence(10000, seed=1)
    flat, _, _ = build_index(synth.linear_graph(seq, node_len=32), 16, 1)
    index = orc.OracleGCSA(flat)
    for k in (1, 4, 8, 16, 32):
        windows = np.lib.stride_tricks.sliding_window_view(seq, k)
        assert index.count_kmers(k, threads=4) == len({bytes(w) for w in windows})
    assert index.count_kmers(0) == 1
    # with N: a path containing Ns has k-mers that only LF_all follows
    seq2 = seq.copy(); seq2[100:103] = 5
    flat2, _, _ = build_index(synth.linear_graph(seq2, node_len=32), 16, 1)
    index2 = orc.OracleGCSA(flat2)
    windows = np.lib.stride_tricks.sliding_window_view(seq2, 8)
    all_kmers = {bytes(w) for w in windows}
    assert index2.count_kmers(8, include_Ns=True) == len(all_kmers)
    assert index2.count_kmers(8, include_Ns=False) == len({w for w in all_kmers if 5 not in w})


def test_builder_refuses_more_paths_than_it_can_number():
    """Path numbers are 32-bit: a kmer count at the limit is refused instead of wrapping around (the arrays are never
    touched: the check comes first)."""
    import ctypes as C
    from gcsa2_b200 import capi
    built = capi.Built()
    dummy = np.zeros(8, dtype=np.uint64)
    rc = capi.lib().gcsa_b200_build_from_kmers(dummy.ctypes.data, dummy.ctypes.data, dummy.ctypes.data, (1 << 32) - 1, 16, 3, 64, C.byref(built))
    assert rc == capi.ERR_INVALID
