"""compareKMers(left, right, k) (src/algorithms.cpp:535-616): oracle restatement against the reference's own
function and against plain set arithmetic over the k-mers of two sequences; (gpu) the device frontier
expansion against the oracle, counts and the unique-kmer records."""
import numpy as np
import pytest

from gcsa2_b200 import synth
from gcsa2_b200.builder import build_index
from oracle import oracle as orc
from oracle import reference as ref


def two_sequences(n=6000, seed=5):
    """Two sequences sharing most of their content, with a few Ns in one."""
    a = synth.random_sequence(n, seed=seed)
    b = a.copy()
    rng = np.random.default_rng(seed)
    for p in rng.integers(0, n, size=n // 50):
        b[p] = 1 + (b[p] % 4)
    b = np.concatenate([b[: n // 2], synth.random_sequence(300, seed=seed + 1), b[n // 2:]])
    a[700:702] = 5
    return a, b


def kmer_set(seq, k, with_n):
    if len(seq) < k:
        return set()
    windows = np.lib.stride_tricks.sliding_window_view(seq, k)
    return {bytes(w) for w in windows if with_n or 5 not in w}


def decode_records(records, k):
    """KMerComparisonState::kmer (algorithms.cpp:451-457): comp i of the backward walk at bits [3i, 3i+3);
    the walk goes right to left, so the k-mer in reading order is the reversed comp list."""
    out = set()
    for row in records:
        assert int(row[4]) == k
        value = int(row[5]) | (int(row[6]) << 64) | (int(row[7]) << 128)
        comps = [(value >> (3 * i)) & 7 for i in range(k)]
        out.add(bytes(reversed(comps)))
    return out


@pytest.fixture(scope="module")
def pair():
    a, b = two_sequences()
    fa, _, ka = build_index(synth.linear_graph(a, node_len=32), 16, 1)
    fb, _, kb = build_index(synth.linear_graph(b, node_len=32), 16, 1)
    return a, b, fa, fb, ka, kb


def test_oracle_against_set_arithmetic(pair):
    a, b, fa, fb, _, _ = pair
    oa, ob = orc.OracleGCSA(fa), orc.OracleGCSA(fb)
    assert oa.compare_kmers(ob, 0)[0] == (1, 0, 0)
    for k in (1, 3, 6, 11, 20, 32):
        for with_n in (False, True):
            sa, sb = kmer_set(a, k, with_n), kmer_set(b, k, with_n)
            counts, left, right = oa.compare_kmers(ob, k, include_Ns=with_n, threads=4)
            assert counts == (len(sa & sb), len(sa - sb), len(sb - sa)), (k, with_n)
            assert decode_records(left, k) == sa - sb and decode_records(right, k) == sb - sa
            assert all(r[0] <= r[1] for r in left) and all(r[2] <= r[3] for r in right)
    # a long k crossing the 64-bit word boundaries of the packed kmer
    counts, left, right = oa.compare_kmers(ob, 45, threads=4)
    sa, sb = kmer_set(a, 45, False), kmer_set(b, 45, False)
    assert decode_records(left, 45) == sa - sb and decode_records(right, 45) == sb - sa


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libgcsa2_ref.so not built (needs /root/reference)")
def test_oracle_against_the_reference(pair):
    a, b, fa, fb, ka, kb = pair
    ra, rb = ref.ReferenceIndex.build(ka, 1), ref.ReferenceIndex.build(kb, 1)
    oa, ob = orc.OracleGCSA(fa), orc.OracleGCSA(fb)
    for k in (1, 4, 9, 17, 32):
        for with_n in (False, True):
            assert ra.compare_kmers(rb, k, include_Ns=with_n) == oa.compare_kmers(ob, k, include_Ns=with_n)[0], (k, with_n)
    assert ra.compare_kmers(ra, 12) == (oa.count_kmers(12), 0, 0)


@pytest.mark.engine
def test_device_compare_kmers(pair):
    from gcsa2_b200 import GCSA
    a, b, fa, fb, _, _ = pair
    ga, gb = GCSA(fa, kmer_table_k=4), GCSA(fb, kmer_table_k=0)
    oa, ob = orc.OracleGCSA(fa), orc.OracleGCSA(fb)
    for k in (0, 1, 2, 5, 8, 16, 32, 45, 64):
        for with_n in (False, True):
            expect, eleft, eright = oa.compare_kmers(ob, k, include_Ns=with_n, threads=4)
            assert ga.compare_kmers(gb, k, include_Ns=with_n) == expect, (k, with_n)
            counts, left, right = ga.compare_kmers(gb, k, include_Ns=with_n, return_kmers=True)
            assert counts == expect
            order = lambda r: r[np.lexsort(r.T[::-1])] if len(r) else r
            assert (order(left) == order(eleft)).all() and (order(right) == order(eright)).all(), (k, with_n)
    assert ga.compare_kmers(ga, 10) == (ga.count_kmers(10), 0, 0)
    with pytest.raises(Exception):
        ga.compare_kmers(gb, 65)
    # KMerSearchParameters::output: the unique kmers as raw 64-byte records in <output>.left / <output>.right
    import os
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        base = os.path.join(tmp, "unique")
        counts, left, right = ga.compare_kmers(gb, 8, return_kmers=True)
        assert ga.compare_kmers_to_files(gb, 8, base) == counts
        order = lambda r: r[np.lexsort(r.T[::-1])] if len(r) else r
        from_left = np.fromfile(base + ".left", dtype=np.uint64).reshape(-1, 8)
        from_right = np.fromfile(base + ".right", dtype=np.uint64).reshape(-1, 8)
        assert from_left.shape[0] == counts[1] and from_right.shape[0] == counts[2]
        assert (order(from_left) == order(left)).all() and (order(from_right) == order(right)).all()
        with pytest.raises(Exception):
            ga.compare_kmers_to_files(gb, 8, os.path.join(tmp, "no", "such", "directory", "x"))
    # a graph with bubbles against its own backbone
    seq = synth.random_sequence(50_000, seed=9)
    fg, _, _ = build_index(synth.snp_graph(seq, seed=9, snp_rate=0.02)[0], 16, 2)
    fl, _, _ = build_index(synth.linear_graph(seq, node_len=32), 16, 2)
    gg, gl = GCSA(fg), GCSA(fl)
    for k in (7, 12, 20):
        shared, left, right = gg.compare_kmers(gl, k)
        assert (shared, left, right) == orc.OracleGCSA(fg).compare_kmers(orc.OracleGCSA(fl), k, threads=8)[0]
        assert right == 0 and shared == gl.count_kmers(k) and shared + left == gg.count_kmers(k)
