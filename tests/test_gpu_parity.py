"""Parity of the CUDA engine with the CPU oracle, through the C ABI (needs a GPU).

Bit-exact comparison of every operation of SURVEY.md section 8(a): find, charRange, LF(range),
LF(node), LF_fast/LF_all, count, locate (x3), parent, depth, psv/psev/nsv/nsev, rmq."""
import ctypes

import numpy as np
import pytest

from brute import random_graph
from helpers import kat1_flat, load_kat1
from verify import kmer_table, verify_index
from gcsa2_b200 import GCSA, LCPArray, capi, synth
from gcsa2_b200.builder import CharGraph, build_index
from gcsa2_b200.flat import FlatLCP
from oracle import oracle as orc

pytestmark = pytest.mark.engine
M64 = (1 << 64) - 1


def empty(sp, ep):
    return ((int(sp) + 1) & M64) > ((int(ep) + 1) & M64)


def both(flat, table_k=0, two_step=False, walk_table=None):
    return GCSA(flat, kmer_table_k=table_k, two_step=two_step, walk_table=walk_table), orc.OracleGCSA(flat)


def assert_find_equal(gpu, ora, chars, offsets, threads=4):
    sp, ep = gpu.find_batch(chars, offsets)
    osp, oep, _ = ora.find_batch(chars, offsets, threads=threads)
    bad = np.flatnonzero((sp != osp) | (ep != oep))
    assert bad.size == 0, (bad[:5], sp[bad[:5]], ep[bad[:5]], osp[bad[:5]], oep[bad[:5]])
    return sp, ep


def random_ranges(rng, n_nodes, n):
    """Valid, empty, singleton, full and out-of-range ranges."""
    a = rng.integers(0, n_nodes, size=n).astype(np.uint64)
    ln = np.minimum(rng.geometric(0.3, size=n), n_nodes).astype(np.uint64)
    b = np.minimum(a + ln - np.uint64(1), np.uint64(n_nodes - 1))
    sp = np.concatenate([a, np.array([0, 0, 1, 5, n_nodes - 1, 3], dtype=np.uint64)])
    ep = np.concatenate([b, np.array([n_nodes - 1, M64, 0, 4, n_nodes + 3, 3], dtype=np.uint64)])
    return sp, ep


# ---------------------------------------------------------------------------------------------

def test_kat1_golden_vectors():
    kat = load_kat1()
    flat, lcp = kat1_flat(kat)
    for table_k in (0, 1, 2, 3):
        gpu = GCSA(flat, kmer_table_k=table_k)
        for pattern, expected in kat["find"].items():
            got = gpu.find(pattern)
            if expected is None:
                assert empty(*got), (pattern, got)
            else:
                assert list(got) == expected, (pattern, got, table_k)
        assert gpu.find("ATT") == (5, 4)
        comp = {"$": 0, "A": 1, "C": 2, "G": 3, "T": 4, "N": 5, "#": 6}
        for case in kat["lf"]:
            assert list(gpu.LF(tuple(case["range"]), comp[case["char"]])) == case["result"]
        for pattern, expected in kat["locate"].items():
            rng = gpu.find(pattern)
            assert gpu.locate(rng) == sorted(expected) and gpu.count(rng) == len(expected)
        for i, vals in enumerate(kat["values"]):
            assert gpu.locate(i) == sorted(vals)
        assert gpu.locate(99) == [] and gpu.locate((3, 99)) == [] and gpu.count((5, 4)) == 0
        assert [gpu.charRange(c) for c in range(7)] == [orc.OracleGCSA(flat).char_range(c) for c in range(7)]
    glcp = LCPArray(FlatLCP.from_values(np.array(lcp, dtype=np.uint8), branching=4))
    assert glcp.parent((2, 4)) == (1, 4, 0, 0, 1) and glcp.depth((2, 4)) == 2
    assert glcp.parent((0, 15)) == (0, 15, 0, 0, 0)


@pytest.mark.parametrize("seed", range(4))
def test_all_operations_random_graphs(seed):
    rng = np.random.default_rng(100 + seed)
    g = random_graph(rng, int(rng.integers(200, 1500)), 8, snp_rate=0.1, node_len=4,
                     alphabet=[(1, 2, 3, 4), (1, 2)][seed % 2])
    cg = CharGraph.from_lists(g.comps, g.values, g.succ, g.sources, g.sink)
    flat, flcp, kmers = build_index(cg, 2, 2, sample_period=[4, 64][seed % 2], lcp_branching=[2, 4, 64, 3][seed])
    gpu, ora = both(flat, table_k=[0, 2, 3, 5][seed], two_step=(seed >= 2), walk_table=[True, False, 2, None][seed])
    N = flat.path_nodes

    # find: patterns of every kind (ragged lengths, empty, with $, #, N, lower case, garbage bytes)
    alphabet = np.frombuffer(b"ACGTACGTACGTacgtN$#x\0", dtype=np.uint8)
    lengths = rng.integers(0, 14, size=4000)
    offsets = np.zeros(lengths.size + 1, dtype=np.uint64); offsets[1:] = np.cumsum(lengths)
    chars = alphabet[rng.integers(0, alphabet.size, size=int(offsets[-1]) + 1)]
    assert_find_equal(gpu, ora, chars, offsets)
    table = kmer_table(kmers)
    chars2, offsets2 = orc.pack_patterns([s for s, _ in table])
    sp, ep = assert_find_equal(gpu, ora, chars2, offsets2)
    assert not any(empty(a, b) for a, b in zip(sp, ep))

    # LF(range, comp) for all comps (incl. invalid range shapes), LF(node), LF_fast / LF_all
    rsp, rep = random_ranges(rng, N, 3000)
    comps = rng.integers(0, 7, size=rsp.size).astype(np.uint8)
    valid = ~np.array([empty(a, b) or b >= N for a, b in zip(rsp, rep)])
    gsp, gep = gpu.lf_batch(rsp[valid], rep[valid], comps[valid])
    for i, (a, b, c) in enumerate(zip(rsp[valid], rep[valid], comps[valid])):
        assert (int(gsp[i]), int(gep[i])) == ora.LF((int(a), int(b)), int(c))
    nodes = np.arange(N, dtype=np.uint64)
    assert list(gpu.lf_node_batch(nodes)) == [ora.LF(int(i)) for i in nodes]
    multi_sp, multi_ep = rsp[valid][:500], rep[valid][:500]
    for all_chars in (0, 1):
        out = gpu.lf_multi_batch(multi_sp, multi_ep, all_chars)
        last = 5 if all_chars else 4
        for i, (a, b) in enumerate(zip(multi_sp, multi_ep)):
            exp = (ora.LF_all if all_chars else ora.LF_fast)((int(a), int(b)))
            for c in range(1, last + 1):
                assert (int(out[i, c, 0]), int(out[i, c, 1])) == exp[c], (a, b, c)

    # count / locate on arbitrary ranges, including empty and out-of-range ones
    cnt = gpu.count_batch(rsp, rep)
    assert list(cnt) == [ora.count((int(a), int(b))) for a, b in zip(rsp, rep)]
    offs, vals = gpu.locate_batch(rsp, rep)
    ooffs, ovals, _ = ora.locate_batch(rsp, rep, threads=2)
    assert (offs == ooffs).all() and (vals == ovals).all()
    # count() == |locate()| holds for ranges that find() produces (query_gcsa.cpp:171-179)
    foffs, _ = gpu.locate_batch(sp, ep)
    assert (np.diff(foffs) == gpu.count_batch(sp, ep)).all()
    wide = np.argsort(-(ep - sp).astype(np.int64))[:150]              # the widest find() ranges + some arbitrary ones
    msp = np.concatenate([sp[wide], rsp[:150]]); mep = np.concatenate([ep[wide], rep[:150]])
    for m in (1, 3, 10):
        moffs, mvals = gpu.locate_batch(msp, mep, max_positions=m)
        for i in range(msp.size):
            assert list(mvals[int(moffs[i]):int(moffs[i + 1])]) == ora.locate((int(msp[i]), int(mep[i])), m)

    # LCP operations
    glcp, olcp = LCPArray(flcp), orc.OracleLCP(flcp)
    psp, pep = rsp[valid], rep[valid]
    par = glcp.parent_batch(psp, pep)
    opar, _ = olcp.parent_batch(psp, pep)
    assert (par == opar).all()
    assert list(glcp.depth_batch(psp, pep)) == [olcp.depth((int(a), int(b))) for a, b in zip(psp, pep)]
    pos = np.concatenate([np.arange(N), [N, N + 5]]).astype(np.uint64)
    for which in ("psv", "psev", "nsv", "nsev"):
        a, b = glcp.sv_batch(which, pos)
        assert [(int(x), int(y)) for x, y in zip(a, b)] == [getattr(olcp, which)(int(p)) for p in pos], which
    qa, qb = rng.integers(0, N, size=2000), rng.integers(0, N + 2, size=2000)
    a, b = glcp.rmq_batch(qa, qb)
    assert [(int(x), int(y)) for x, y in zip(a, b)] == [olcp.rmq(int(s), int(e)) for s, e in zip(qa, qb)]

    # the reference's own predicates, evaluated through the engine
    assert verify_index(gpu, glcp, table, limit=150, seed=seed) == []


def test_linear_reference_order128_kmer_table_invariance():
    """1 Mbp linear reference, order 128: 32-mers sampled from it and uniform random 32-mers;
    identical answers with and without the k-mer table."""
    import helpers
    n = 60_000 if helpers.EMULATED else 200_000                       # (the emulated run is one of the slowest CPU tests)
    seq = synth.random_sequence(1_000_000, seed=2)
    flat, flcp, _ = build_index(synth.linear_graph(seq), 16, 3)
    ora = orc.OracleGCSA(flat)
    chars, offsets = synth.patterns_from_sequence(seq, n, 32, seed=5)
    rchars, roffsets = synth.random_patterns(n, 32, seed=6)
    ref = None
    for table_k, two_step in ((0, False), (4, False), (10, True), (0, True), (5, True)):
        gpu = GCSA(flat, kmer_table_k=table_k, two_step=two_step)
        assert gpu.twoStep() == two_step
        sp, ep = assert_find_equal(gpu, ora, chars, offsets)
        assert not np.any(sp > ep)                       # all sampled patterns occur
        rsp, rep = assert_find_equal(gpu, ora, rchars, roffsets)
        fsp, fep = gpu.find_fixed_batch(chars, 32)            # k-mer form of the same call
        assert (fsp == sp).all() and (fep == ep).all()
        if ref is None:
            ref = (sp, ep, rsp, rep)
        else:
            assert all((a == b).all() for a, b in zip(ref, (sp, ep, rsp, rep)))
        # round trip: locate(find(P)) contains the position P was cut from
        starts = np.random.default_rng(5).integers(0, seq.size - 32 + 1, size=n)
        offs, vals = gpu.locate_batch(sp[:5000], ep[:5000])
        graph_values = synth.linear_graph(seq).value
        for i in range(5000):
            assert graph_values[1 + starts[i]] in vals[int(offs[i]):int(offs[i + 1])]
        stats = gpu.find_batch(chars[:3200], offsets[:101], stats=True)[2]
        assert stats["queries"] == 100 and stats["found"] == 100
        assert stats["lf_steps"] == 100 * (31 if table_k == 0 else 32 - table_k)
        gpu.close()


def test_snp_graph_find_locate_parent():
    seq = synth.random_sequence(300_000, seed=3)
    graph, sites, alt = synth.snp_graph(seq, seed=3, snp_rate=0.01)
    flat, flcp, kmers = build_index(graph, 16, 3)
    gpu, ora = both(flat, table_k=8, two_step=True)
    assert gpu.twoStep()
    glcp, olcp = LCPArray(flcp), orc.OracleLCP(flcp)
    chars, offsets = synth.patterns_from_snp_graph(seq, sites, alt, 50_000, 64, seed=7)
    sp, ep = assert_find_equal(gpu, ora, chars, offsets)
    assert not np.any(sp > ep)
    offs, vals = gpu.locate_batch(sp, ep)
    ooffs, ovals, _ = ora.locate_batch(sp, ep, threads=4)
    assert (offs == ooffs).all() and (vals == ovals).all()
    assert (gpu.count_batch(sp, ep) == np.diff(offs)).all()
    par = glcp.parent_batch(sp, ep)
    opar, _ = olcp.parent_batch(sp, ep, threads=4)
    assert (par == opar).all()
    mchars, moffsets = synth.mixed_length_patterns(seq, sites, alt, 20_000, 16, 256, seed=8)
    assert_find_equal(gpu, ora, mchars, moffsets)
    assert verify_index(gpu, glcp, kmer_table(kmers), limit=100) == []


def test_device_pointer_entry_point_and_empty_batches():
    import torch
    from helpers import current_stream, device_empty, device_sync, to_device
    seq = synth.random_sequence(50_000, seed=4)
    flat, _, _ = build_index(synth.linear_graph(seq), 16, 1)
    gpu, ora = both(flat)
    chars, offsets = synth.patterns_from_sequence(seq, 10_000, 24, seed=1)
    d_chars = to_device(chars); d_off = to_device(offsets.view(np.int64))
    d_sp = device_empty(10_000, torch.int64); d_ep = device_empty(10_000, torch.int64)
    gpu.find_device(d_chars, d_off, 10_000, d_sp, d_ep, current_stream())
    device_sync()
    osp, oep, _ = ora.find_batch(chars, offsets)
    assert (d_sp.cpu().numpy().view(np.uint64) == osp).all() and (d_ep.cpu().numpy().view(np.uint64) == oep).all()
    sp, ep = gpu.find_batch([])
    assert sp.size == 0 and ep.size == 0
    assert gpu.count_batch([], []).size == 0
    offs, vals = gpu.locate_batch([], [])
    assert list(offs) == [0] and vals.size == 0
    assert gpu.find("") == (0, flat.path_nodes - 1)


def test_count_kmers_frontier_expansion():
    """countKMers on the device (breadth-first) == the reference's depth-first count (oracle), and the
    final frontier is find() of every k-mer (ordered by the reversed k-mer)."""
    seq = synth.random_sequence(10000, seed=1)
    seq[500:503] = 5                                                    # a few Ns
    flat, _, _ = build_index(synth.linear_graph(seq, node_len=32), 16, 1)
    gpu, ora = both(flat)
    for k in (0, 1, 2, 5, 8, 16, 32):
        for with_n in (False, True):
            assert gpu.count_kmers(k, include_Ns=with_n) == ora.count_kmers(k, include_Ns=with_n, threads=4), (k, with_n)
    n, sp, ep = gpu.count_kmers(6, return_ranges=True)
    # the trie is grown leftwards, so the frontier is ordered by the reversed k-mer
    kmers = sorted({bytes(w) for w in np.lib.stride_tricks.sliding_window_view(seq, 6) if 5 not in w}, key=lambda w: w[::-1])
    assert n == len(kmers)
    chars, offsets = orc.pack_patterns([bytes(synth.COMP2CHAR[np.frombuffer(k, dtype=np.uint8)]) for k in kmers])
    fsp, fep = gpu.find_batch(chars, offsets)
    assert (fsp == sp).all() and (fep == ep).all()
    # a graph with bubbles
    seq = synth.random_sequence(100_000, seed=8)
    graph, _, _ = synth.snp_graph(seq, seed=8, snp_rate=0.02)
    flat, _, _ = build_index(graph, 16, 2)
    gpu, ora = both(flat, two_step=True)
    for k in (3, 9, 12, 20):
        assert gpu.count_kmers(k) == ora.count_kmers(k, threads=8)


@pytest.mark.gpu
def test_full_size_config2_properties():
    """BASELINE.json configs[1] at full size: 10 M 32-mers sampled from a 100 Mbp linear reference,
    order 128.  Size-independent properties: every sampled pattern is found; all engine variants
    (k-mer table on/off, two-step on/off) give the same 20 M numbers; locate(find(P)) contains the
    position P was cut from; a 200 k sample equals the oracle bit for bit."""
    L, n, length = 100_000_000, 10_000_000, 32
    seq = synth.random_sequence(L, seed=2)
    graph = synth.linear_graph(seq, node_len=32)
    flat, _, _ = build_index(graph, 16, 3)
    assert flat.path_nodes == L + 2 and flat.order == 128
    chars = np.empty(n * length, dtype=np.uint8)
    starts = np.empty(n, dtype=np.int64)
    for i, q0 in enumerate(range(0, n, 1_000_000)):
        rng = np.random.default_rng(4242 + i)
        st = rng.integers(0, L - length + 1, size=1_000_000, dtype=np.int64)
        starts[q0:q0 + 1_000_000] = st
        chars[q0 * length:(q0 + 1_000_000) * length] = synth.COMP2CHAR[seq[st[:, None] + np.arange(length)[None, :]]].reshape(-1)
    ref = None
    for table_k, two_step in ((14, False), (0, False), (12, True)):
        gpu = GCSA(flat, kmer_table_k=table_k, two_step=two_step)
        sp, ep = gpu.find_fixed_batch(chars, length)
        assert not np.any(sp > ep)                                   # everything sampled from the text occurs
        if ref is None:
            ref = (sp, ep)
            m = 1_000_000
            offs, vals = gpu.locate_batch(sp[:m], ep[:m])
            want = graph.value[1 + starts[:m]]
            first = vals[offs[:-1].astype(np.int64)]
            single = np.diff(offs.astype(np.int64)) == 1
            assert (first[single] == want[single]).all() and single.mean() > 0.99
            for i in np.flatnonzero(~single)[:2000]:
                assert want[i] in vals[int(offs[i]):int(offs[i + 1])]
            assert (gpu.count_batch(sp[:m], ep[:m]) == np.diff(offs)).all()
        else:
            assert (sp == ref[0]).all() and (ep == ref[1]).all()
        gpu.close()
    ora = orc.OracleGCSA(flat)
    k = 200_000
    offsets = np.arange(k + 1, dtype=np.uint64) * np.uint64(length)
    osp, oep, _ = ora.find_batch(chars[:k * length], offsets, threads=8)
    assert (osp == ref[0][:k]).all() and (oep == ref[1][:k]).all()


def test_host_packed_find_equals_byte_path(monkeypatch):
    """gcsa_b200_find_fixed_host with host-side 2-bit packing (pack.cpp + find_kernel<.., PACKED>) gives the
    same ranges as the byte path and the oracle: lengths on and off the 32-character word boundary, mixed
    case, random (early-exit) patterns, and a batch where one chunk holds an N and must fall back."""
    seq = synth.random_sequence(300_000, seed=41)
    flat, _, _ = build_index(synth.linear_graph(seq), 16, 3)
    ora = orc.OracleGCSA(flat)
    import helpers
    for table_k in ((6,) if helpers.EMULATED else (0, 6)):           # the emulated run is the slowest test of the CPU suite
        gpu = GCSA(flat, kmer_table_k=table_k)
        for length in (20, 32, 45, 64, 100):
            n = 600_000
            chars, offsets = synth.patterns_from_sequence(seq, n, length, seed=length)
            chars = chars.copy()
            rnd, _ = synth.random_patterns(n // 4, length, seed=length + 1)
            chars[: rnd.size] = rnd                                   # the first quarter: uniform random patterns
            lower = np.random.default_rng(length).random(chars.size) < 0.3
            chars[lower] |= 0x20                                      # acgt
            monkeypatch.setenv("GCSA_B200_HOST_PACK", "0")
            bsp, bep = gpu.find_fixed_batch(chars, length)
            monkeypatch.setenv("GCSA_B200_HOST_PACK", "4")
            psp, pep = gpu.find_fixed_batch(chars, length)
            assert (psp == bsp).all() and (pep == bep).all(), (table_k, length)
            osp, oep, _ = ora.find_batch(chars[: 50_000 * length], offsets[: 50_001], threads=4)
            assert (psp[:50_000] == osp).all() and (pep[:50_000] == oep).all()
            if length == 32:
                dirty = chars.copy()
                dirty[300_000 * length + 7] = ord("N")                # second chunk of three
                dirty[599_999 * length + 31] = ord("$")
                monkeypatch.setenv("GCSA_B200_HOST_PACK", "0")
                bsp, bep = gpu.find_fixed_batch(dirty, length)
                monkeypatch.setenv("GCSA_B200_HOST_PACK", "4")
                psp, pep = gpu.find_fixed_batch(dirty, length)
                assert (psp == bsp).all() and (pep == bep).all()
                q = np.array([300_000, 599_999])
                pats = [bytes(dirty[i * length:(i + 1) * length]) for i in q]
                for i, pat in zip(q, pats):
                    assert (int(psp[i]), int(pep[i])) == ora.find(pat)


def test_host_pack_shares_the_batch_with_the_raw_path(monkeypatch):
    """The host entry point of find() with packing enabled: the caller sends raw chunks from the front of the batch while
    the rest of its OpenMP team packs chunks from the back; which chunk goes which way depends on timing, the answers do
    not -- equal to the unpacked path for every setting, with a chunk that cannot be packed (an N) in the batch; the share
    of packed chunks is reported, and without an explicit setting both ways are tried before one is chosen."""
    import ctypes
    from gcsa2_b200 import capi
    L = capi.lib()
    seq = synth.random_sequence(200_000, seed=43)
    flat, _, _ = build_index(synth.linear_graph(seq), 16, 3)
    gpu = GCSA(flat, kmer_table_k=8)
    n, length = 1_200_000, 32
    chars, offsets = synth.patterns_from_sequence(seq, n, length, seed=5)
    chars = chars.copy()
    rnd, _ = synth.random_patterns(n // 8, length, seed=6)
    chars[: rnd.size] = rnd
    chars[(n - 5) * length + 3] = ord("N")                             # in the last chunk: the packer's first claim
    monkeypatch.setenv("GCSA_B200_HOST_PACK", "0")
    bsp, bep = gpu.find_fixed_batch(chars, length)
    packed, total = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    L.gcsa_b200_internal_pack_share(ctypes.byref(packed), ctypes.byref(total))
    assert packed.value == 0 and total.value >= 1
    osp, oep, _ = orc.OracleGCSA(flat).find_batch(chars[-200_000 * length:], offsets[:200_001], threads=4)
    assert (bsp[-200_000:] == osp).all() and (bep[-200_000:] == oep).all()
    for setting, threads in (("auto", "2"), (None, "3"), ("4", None), ("1", None)):
        if setting is None:
            monkeypatch.delenv("GCSA_B200_HOST_PACK")
        else:
            monkeypatch.setenv("GCSA_B200_HOST_PACK", setting)
        if threads is not None:
            monkeypatch.setenv("GCSA_B200_HOST_PACK_THREADS", threads)
        ways = set()
        for _ in range(3):
            sp, ep = gpu.find_fixed_batch(chars, length)
            assert (sp == bsp).all() and (ep == bep).all(), (setting, threads)
            L.gcsa_b200_internal_pack_share(ctypes.byref(packed), ctypes.byref(total))
            if setting in ("auto", None) and total.value == 5:        # the automatic policy tries raw copies only as well
                assert packed.value == 0
                ways.add("raw")
            else:
                assert total.value == 10 and packed.value <= total.value - 2      # the copy engine is fed raw chunks first
                ways.add("shared")
        if setting == "auto":
            assert ways == {"raw", "shared"}                          # both ways measured before one is chosen
    small = gpu.find_fixed_batch(chars[: 1000 * length], length)      # too small for the two-thread pipeline
    assert (small[0] == bsp[:1000]).all()


def test_locate_tables_agree():
    """locate() through the locate table (one load per node), the walk table (one load per LF step) and the
    bit vectors alone: identical CSR output, equal to the oracle -- short patterns (wide ranges, many nodes per
    range, multi-valued nodes at the SNP sites), k-mers, empty and out-of-range ranges, raw and max forms."""
    seq = synth.random_sequence(200_000, seed=13)
    seq[5000:5400] = seq[1000:1400]                                   # a repeat: nodes with two start positions
    graph, sites, alt = synth.snp_graph(seq, seed=13, snp_rate=0.02)
    flat, _, _ = build_index(graph, 16, 2)
    ora = orc.OracleGCSA(flat)
    rng = np.random.default_rng(13)
    pats = []
    for length, count in ((3, 50), (6, 2000), (12, 5000), (40, 5000)):
        c, o = synth.patterns_from_snp_graph(seq, sites, alt, count, length, seed=length)
        pats += [bytes(c[int(o[i]):int(o[i + 1])]) for i in range(count)]
    pats += [bytes(seq_to_chars) for seq_to_chars in (synth.COMP2CHAR[seq[1100:1100 + L]] for L in (20, 64, 300))]
    chars, offsets = orc.pack_patterns(pats)
    sp, ep, _ = ora.find_batch(chars, offsets, threads=4)
    sp = np.concatenate([sp, [5, 1, flat.path_nodes - 2, 0]]).astype(np.uint64)
    ep = np.concatenate([ep, [4, 0, flat.path_nodes + 5, flat.path_nodes - 1]]).astype(np.uint64)   # empty, empty, out of range, everything
    sp, ep = sp[:-1], ep[:-1]                                         # (the whole index is covered by test_all_operations)
    ooffs, ovals, _ = ora.locate_batch(sp, ep, threads=4)
    results = []
    for walk_table in (1, 2, 0):
        gpu = GCSA(flat, kmer_table_k=0, walk_table=walk_table)
        offs, vals = gpu.locate_batch(sp, ep)
        assert (offs == ooffs).all() and (vals == ovals).all(), walk_table
        roffs, rvals = gpu.locate_batch(sp[:3000], ep[:3000], sort=False)
        results.append((roffs, rvals))
        for i in rng.integers(0, sp.size, size=40):
            rng_i = (int(sp[i]), int(ep[i]))
            assert list(gpu.locate(rng_i, max_positions=5)) == list(ora.locate(rng_i, max_positions=5)), (walk_table, rng_i)
    for roffs, rvals in results[1:]:
        assert (roffs == results[0][0]).all() and (rvals == results[0][1]).all()


def test_locate_short_range_path(monkeypatch):
    """locate() of short ranges in registers (one thread per range, locate table) == the general pipeline == the oracle:
    every range length around the register limit (8 nodes), duplicates inside a range (a repeat), nodes with several
    start positions (they send the range to the general pipeline), empty and out-of-range ranges mixed in, all-general
    and all-short batches, and the capacity protocol of the device entry point."""
    import torch
    from helpers import current_stream, device_empty, device_sync, to_device
    from gcsa2_b200 import capi
    seq = synth.random_sequence(120_000, seed=19)
    seq[7000:7600] = seq[2000:2600]                                   # a repeat: nodes with two start positions
    seq[9000:9300] = seq[2100:2400]
    graph, sites, alt = synth.snp_graph(seq, seed=19, snp_rate=0.02)
    flat, _, _ = build_index(graph, 16, 2)
    ora = orc.OracleGCSA(flat)
    N = flat.path_nodes
    rng = np.random.default_rng(19)
    a = rng.integers(0, N, size=60_000).astype(np.uint64)
    ln = rng.integers(1, 13, size=a.size).astype(np.uint64)           # 1..12 nodes: both sides of the limit
    ln[: 30_000] = 1
    b = np.minimum(a + ln - np.uint64(1), np.uint64(N - 1))
    sp = np.concatenate([a, np.array([5, 0, N - 3, 0, 7], dtype=np.uint64)])
    ep = np.concatenate([b, np.array([4, M64, N + 2, N - 1, 7], dtype=np.uint64)])   # empty, empty, out of range, everything, one node
    order = rng.permutation(sp.size)
    sp, ep = sp[order], ep[order]
    ooffs, ovals, _ = ora.locate_batch(sp, ep, threads=4)
    gpu = GCSA(flat, walk_table=1)
    for subset in (slice(None), np.flatnonzero(ep - sp < 8), np.flatnonzero((ep - sp >= 8) & (ep < N)), slice(0, 1)):
        s2, e2 = sp[subset], ep[subset]
        want_offs, want_vals, _ = ora.locate_batch(s2, e2, threads=4)
        monkeypatch.setenv("GCSA_B200_LOCATE_SMALL", "1")
        offs, vals = gpu.locate_batch(s2, e2)
        monkeypatch.setenv("GCSA_B200_LOCATE_SMALL", "0")
        goffs, gvals = gpu.locate_batch(s2, e2)
        assert (offs == want_offs).all() and (vals == want_vals).all()
        assert (goffs == want_offs).all() and (gvals == want_vals).all()
    monkeypatch.setenv("GCSA_B200_LOCATE_SMALL", "1")
    assert (gpu.count_batch(sp, ep)[ep < N] >= 0).all()
    # device entry point: too small a buffer reports the size and still writes the offsets
    n = sp.size
    d_sp, d_ep = to_device(sp.view(np.int64)), to_device(ep.view(np.int64))
    d_offs = device_empty(n + 1, torch.int64); small = device_empty(10, torch.int64)
    with pytest.raises(capi.GCSAError) as err:
        gpu.locate_device(d_sp, d_ep, n, d_offs, small, 10, current_stream())
    assert err.value.code == capi.ERR_CAPACITY
    device_sync()
    assert (d_offs.cpu().numpy().view(np.uint64) == ooffs).all()
    big = device_empty(int(ooffs[-1]), torch.int64)
    got = gpu.locate_device(d_sp, d_ep, n, d_offs, big, int(ooffs[-1]), current_stream())
    device_sync()
    assert got == int(ooffs[-1]) and (big.cpu().numpy().view(np.uint64) == ovals).all()


def test_locate_medium_range_path(monkeypatch):
    """locate() of ranges of tens to thousands of path nodes (what a short pattern gives): sorted and deduplicated in
    registers by a warp (up to 1024 nodes) or a block (up to 4096, up to 16384) == the general pipeline == the oracle.  Lengths
    on both sides of every limit (8 | 9, 32 | 33, 128 | 129, ..., 1024 | 1025, 4096 | 4097, 16384 | 16385), repeats (duplicates inside a range, and
    nodes with several start positions, which hand the range over to the general pipeline), mixed with short, empty
    and out-of-range ranges."""
    seq = synth.random_sequence(60_000, seed=23)
    seq[20_000:26_000] = seq[1000:7000]                               # a long repeat
    seq[40_000:40_400] = seq[1200:1600]
    graph, sites, alt = synth.snp_graph(seq, seed=23, snp_rate=0.02)
    flat, _, _ = build_index(graph, 16, 2)
    ora = orc.OracleGCSA(flat)
    N = flat.path_nodes
    rng = np.random.default_rng(23)
    edge = np.array([8, 9, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049, 4095, 4096, 4097, 5000, 8191, 8192, 8193, 16383, 16384, 16385, 20000], dtype=np.uint64)
    ln = np.concatenate([np.repeat(edge, 3), rng.integers(9, 200, size=300).astype(np.uint64), rng.integers(1, 9, size=200).astype(np.uint64),
                         rng.integers(200, 4200, size=40).astype(np.uint64)])
    a = rng.integers(0, N - 20001, size=ln.size).astype(np.uint64)
    sp = np.concatenate([a, np.array([5, 0, N - 3, 0], dtype=np.uint64)])
    ep = np.concatenate([a + ln - np.uint64(1), np.array([4, M64, N + 2, N - 1], dtype=np.uint64)])
    order = rng.permutation(sp.size)
    sp, ep = sp[order], ep[order]
    want_offs, want_vals, _ = ora.locate_batch(sp, ep, threads=4)
    gpu = GCSA(flat, walk_table=1)
    for medium in ("1", "0"):
        monkeypatch.setenv("GCSA_B200_LOCATE_MEDIUM", medium)
        offs, vals = gpu.locate_batch(sp, ep)
        assert (offs == want_offs).all() and (vals == want_vals).all(), medium
    monkeypatch.setenv("GCSA_B200_LOCATE_MEDIUM", "1")
    only = np.flatnonzero((ep - sp >= 8) & (ep - sp < 16384) & (ep < N))      # a batch without short or general ranges
    w_offs, w_vals, _ = ora.locate_batch(sp[only], ep[only], threads=4)
    offs, vals = gpu.locate_batch(sp[only], ep[only])
    assert (offs == w_offs).all() and (vals == w_vals).all()
    assert (gpu.count_batch(sp[only], ep[only]) >= 0).all()


def test_locate_into_host_buffers():
    """gcsa_b200_locate_into_host (caller-owned buffers, chunked pipeline) == gcsa_b200_locate_host; too small a
    buffer is reported with the needed size and complete offsets."""
    from gcsa2_b200 import capi
    seq = synth.random_sequence(400_000, seed=17)
    graph, sites, alt = synth.snp_graph(seq, seed=17, snp_rate=0.02)
    flat, _, _ = build_index(graph, 16, 2)
    gpu = GCSA(flat, kmer_table_k=6)
    n = 700_000                                                          # three chunks
    chars, offsets = synth.patterns_from_snp_graph(seq, sites, alt, n, 24, seed=3)
    sp, ep = gpu.find_batch(chars, offsets)
    sp[::1000] = 7; ep[::1000] = 3                                       # some empty ranges
    sp[5::5000] = 100; ep[5::5000] = 400                                 # some wide ones
    offs, vals = gpu.locate_batch(sp, ep)
    out_offs = np.zeros(n + 1, dtype=np.uint64); out_vals = np.zeros(vals.size + 8, dtype=np.uint64)
    got = gpu.locate_into_host_raw(sp.ctypes.data, ep.ctypes.data, n, out_offs.ctypes.data, out_vals.ctypes.data, out_vals.size)
    assert got == vals.size and (out_offs == offs).all() and (out_vals[:got] == vals).all()
    small = np.zeros(vals.size // 2, dtype=np.uint64); out_offs[:] = 0
    with pytest.raises(capi.GCSAError) as err:
        gpu.locate_into_host_raw(sp.ctypes.data, ep.ctypes.data, n, out_offs.ctypes.data, small.ctypes.data, small.size)
    assert err.value.code == capi.ERR_CAPACITY and (out_offs == offs).all()
    assert gpu.locate_into_host_raw(sp.ctypes.data, ep.ctypes.data, 0, out_offs.ctypes.data, small.ctypes.data, small.size) == 0


def test_jump_table_equals_single_steps():
    """find() with the jump table (one load for up to 16 steps along a unary backward path) == without it ==
    the oracle: patterns that follow the path, leave it at every possible distance (one substitution at a random
    offset: the uncanonicalised empty pair must be the one of the exact failing step), end inside a jump
    (lengths 17..70), contain N, and walks through a graph with SNP bubbles where paths branch."""
    rng = np.random.default_rng(23)
    seq = synth.random_sequence(400_000, seed=23)
    graph, sites, alt = synth.snp_graph(seq, seed=23, snp_rate=0.01)
    for name, flat, sampler in (
            ("linear", build_index(synth.linear_graph(seq), 16, 3)[0], lambda n, L, s: synth.patterns_from_sequence(seq, n, L, seed=s)),
            ("snp", build_index(graph, 16, 3)[0], lambda n, L, s: synth.patterns_from_snp_graph(seq, sites, alt, n, L, seed=s))):
        ora = orc.OracleGCSA(flat)
        pats = []
        for L in (17, 20, 31, 32, 33, 47, 48, 64, 70):
            c, o = sampler(3000, L, L)
            c = c.copy()
            for i in range(3000):
                kind = i % 4
                if kind == 1:                                        # one substitution somewhere
                    p = int(o[i]) + int(rng.integers(0, L))
                    c[p] = synth.COMP2CHAR[1 + (int(np.where(synth.COMP2CHAR == c[p])[0][0]) % 4)]
                elif kind == 2 and i % 8 == 2:                       # an N
                    c[int(o[i]) + int(rng.integers(0, L))] = ord("N")
                pats.append(bytes(c[int(o[i]):int(o[i + 1])]))
        chars, offsets = orc.pack_patterns(pats)
        osp, oep, _ = ora.find_batch(chars, offsets, threads=4)
        for table_k, two_step in ((0, False), (8, False), (6, True)):
            with_jump = GCSA(flat, kmer_table_k=table_k, two_step=two_step, jump_table=True)
            without = GCSA(flat, kmer_table_k=table_k, two_step=two_step, jump_table=False)
            assert with_jump.jumpK() == 16 and without.jumpK() == 0
            a, b = with_jump.find_batch(chars, offsets)
            c2, d2 = without.find_batch(chars, offsets)
            bad = np.flatnonzero((a != osp) | (b != oep))
            assert bad.size == 0, (name, table_k, two_step, bad[:5], [pats[i] for i in bad[:3]])
            assert (c2 == osp).all() and (d2 == oep).all()
            _, _, st = with_jump.find_batch(chars, offsets, stats=True)
            _, _, st0 = without.find_batch(chars, offsets, stats=True)
            assert st["lf_steps"] == st0["lf_steps"] and st["sector_probes"] < st0["sector_probes"]


def test_fused_table_equals_separate_loads():
    """k-mer table with fused 16-byte entries (result of the k-mer + the jump entry of its node, one load) ==
    the 8-byte table followed by a separate jump load == the oracle.  Lengths around every boundary: shorter than
    k, k exactly, k + a partial path, k + 16, beyond the packed 32-character tail; substitutions at random offsets
    (the early-exit pair of the exact failing step), an N, lower case, random patterns, and a graph with bubbles."""
    rng = np.random.default_rng(31)
    seq = synth.random_sequence(300_000, seed=31)
    graph, sites, alt = synth.snp_graph(seq, seed=31, snp_rate=0.01)
    for name, flat, sampler in (
            ("linear", build_index(synth.linear_graph(seq), 16, 3)[0], lambda n, L, s: synth.patterns_from_sequence(seq, n, L, seed=s)),
            ("snp", build_index(graph, 16, 3)[0], lambda n, L, s: synth.patterns_from_snp_graph(seq, sites, alt, n, L, seed=s))):
        ora = orc.OracleGCSA(flat)
        pats = []
        for L in (5, 8, 9, 10, 12, 17, 23, 24, 25, 31, 32, 33, 40, 64):
            c, o = sampler(1500, L, 100 + L)
            c = c.copy()
            for i in range(1500):
                kind = i % 5
                if kind == 1:                                        # one substitution somewhere
                    p = int(o[i]) + int(rng.integers(0, L))
                    c[p] = synth.COMP2CHAR[1 + (int(np.where(synth.COMP2CHAR == c[p])[0][0]) % 4)]
                elif kind == 2 and i % 10 == 2:                      # an N
                    c[int(o[i]) + int(rng.integers(0, L))] = ord("N")
                elif kind == 3:                                      # lower case
                    c[int(o[i]):int(o[i + 1])] |= 0x20
                pats.append(bytes(c[int(o[i]):int(o[i + 1])]))
        rc, ro = synth.random_patterns(2000, 32, seed=77)
        pats += [bytes(rc[int(ro[i]):int(ro[i + 1])]) for i in range(2000)]
        chars, offsets = orc.pack_patterns(pats)
        osp, oep, _ = ora.find_batch(chars, offsets, threads=4)
        for table_k, two_step in ((8, False), (4, False), (9, True), (1, False)):
            fused = GCSA(flat, kmer_table_k=table_k, two_step=two_step, jump_table=True, fused_table=True)
            plain = GCSA(flat, kmer_table_k=table_k, two_step=two_step, jump_table=True, fused_table=False)
            assert fused.fusedTable() and not plain.fusedTable()
            a, b, st = fused.find_batch(chars, offsets, stats=True)
            c2, d2, st0 = plain.find_batch(chars, offsets, stats=True)
            bad = np.flatnonzero((a != osp) | (b != oep))
            assert bad.size == 0, (name, table_k, two_step, bad[:5], [pats[i] for i in bad[:3]])
            assert (c2 == osp).all() and (d2 == oep).all()
            a, b = fused.find_batch(chars, offsets)                  # the kernel variant without statistics
            assert (a == osp).all() and (b == oep).all()
            assert st["lf_steps"] == st0["lf_steps"] and st["table_hits"] == st0["table_hits"]
            if table_k >= 8:                                         # 8-mers of a 300 kbp text are mostly unique: jumps get fused
                assert st["sector_probes"] < st0["sector_probes"], (name, table_k, st, st0)
            else:
                assert st["sector_probes"] <= st0["sector_probes"]
        assert not GCSA(flat, kmer_table_k=6, jump_table=False, fused_table=True).fusedTable()   # nothing to fuse with


def test_custom_alphabet_takes_the_general_path():
    """An index whose char2comp is not the default table (digits 1-4 are the bases, lower case is N): the SWAR
    pattern packing is off, the k-mer table and the jump table are reached through the per-character path, and
    the answers still equal the oracle's (which maps through the same table, gcsa.h:102,106)."""
    import copy
    seq = synth.random_sequence(150_000, seed=29)
    flat, _, _ = build_index(synth.linear_graph(seq), 16, 3)
    custom = copy.deepcopy(flat)
    c2c = np.full(256, 5, dtype=np.uint8)
    c2c[0] = 0; c2c[ord("$")] = 0; c2c[ord("#")] = 6
    for i, ch in enumerate("1234"):
        c2c[ord(ch)] = i + 1
    for i, ch in enumerate("ACGT"):
        c2c[ord(ch)] = i + 1
    custom.char2comp = c2c                                            # upper case and digits are bases, lower case is N
    chars, offsets = synth.patterns_from_sequence(seq, 60_000, 40, seed=9)
    chars = chars.copy()
    rng = np.random.default_rng(29)
    digits = rng.random(chars.size) < 0.5
    lut = np.arange(256, dtype=np.uint8)
    for i, ch in enumerate("ACGT"):
        lut[ord(ch)] = ord("1234"[i])
    chars[digits] = lut[chars[digits]]                               # half of the characters as digits: same comps
    lower = rng.random(60_000) < 0.1                                   # a tenth of the patterns get one lower-case letter = N here
    for i in np.flatnonzero(lower):
        p = int(offsets[i]) + int(rng.integers(0, 40))
        chars[p] = ord("acgt"[int(rng.integers(0, 4))])
    ora = orc.OracleGCSA(custom)
    osp, oep, _ = ora.find_batch(chars, offsets, threads=4)
    for table_k, jump in ((0, False), (8, True), (0, True)):
        gpu = GCSA(custom, kmer_table_k=table_k, jump_table=jump)
        sp, ep = gpu.find_batch(chars, offsets)
        assert (sp == osp).all() and (ep == oep).all(), (table_k, jump)
        fsp, fep = gpu.find_fixed_batch(chars, 40)
        assert (fsp == osp).all() and (fep == oep).all()
    found = (osp <= oep)
    assert found[~lower].all() and not found[lower].any()


def test_kmer_batches_two_kernel_form(monkeypatch):
    """Batches of k-mers (one length, at least 4096 of them) go through find_fast_kernel, find_quad_kernel, find_chain_kernel
    (patterns longer than the table plus one long jump) + the work list resumed by the general kernel; they must equal the oracle and the single general kernel (GCSA_B200_FIND_FAST=0 is read once
    per process, so the general kernel is reached through the offsets form of the same patterns).  Every table shape:
    fused and plain entries, 8- and 16-byte jump entries, no jump table, two-step blocks; lengths equal to k, between
    k and k + a path, 32; substitutions (a jump that fails on a character), N, lower case, random patterns (misses in
    the table), a repetitive text (ranges of several nodes after the table), a graph with bubbles."""
    rng = np.random.default_rng(41)
    seq = synth.random_sequence(200_000, seed=41)
    rep = np.concatenate([np.tile(synth.random_sequence(500, seed=42), 40), synth.random_sequence(30_000, seed=43)])
    graph, sites, alt = synth.snp_graph(seq, seed=41, snp_rate=0.01)
    cases = (("linear", build_index(synth.linear_graph(seq), 16, 3)[0], lambda n, L, s: synth.patterns_from_sequence(seq, n, L, seed=s)),
             ("repeats", build_index(synth.linear_graph(rep), 16, 3)[0], lambda n, L, s: synth.patterns_from_sequence(rep, n, L, seed=s)),
             ("snp", build_index(graph, 16, 3)[0], lambda n, L, s: synth.patterns_from_snp_graph(seq, sites, alt, n, L, seed=s)))
    for name, flat, sampler in cases:
        ora = orc.OracleGCSA(flat)
        for L in (8, 11, 24, 31, 32, 33, 47, 64, 100):
            n = 6000
            c, o = sampler(n, L, 200 + L)
            c = c.copy()
            for i in range(n):
                kind = i % 6
                if kind == 1:                                        # one substitution somewhere
                    p = int(o[i]) + int(rng.integers(0, L))
                    c[p] = synth.COMP2CHAR[1 + (int(np.where(synth.COMP2CHAR == c[p])[0][0]) % 4)]
                elif kind == 2 and i % 12 == 2:                      # an N
                    c[int(o[i]) + int(rng.integers(0, L))] = ord("N")
                elif kind == 3:                                      # lower case
                    c[int(o[i]):int(o[i + 1])] |= 0x20
                elif kind == 4 and i % 12 == 4:                      # a random pattern
                    c[int(o[i]):int(o[i + 1])] = synth.random_patterns(1, L, seed=i)[0]
            osp, oep, _ = ora.find_batch(c, o, threads=4)
            for options in (dict(kmer_table_k=8, jump_table=True, fused_table=True), dict(kmer_table_k=8, jump_table=True, fused_table=False),
                            dict(kmer_table_k=8, jump_table="wide"), dict(kmer_table_k=6, jump_table=False),
                            dict(kmer_table_k=7, jump_table="wide", two_step=True), dict(kmer_table_k=5, jump_table=True, fused_table=True)):
                if options["kmer_table_k"] > L:
                    continue
                gpu = GCSA(flat, **options)
                if options.get("jump_table") == "wide":
                    assert gpu.jumpK() == 16 and not gpu.fusedTable()
                fast = capi.lib().gcsa_b200_internal_fast_launches
                fast.restype = ctypes.c_ulonglong
                before = fast()
                sp, ep = gpu.find_fixed_batch(c, L)                  # the two-kernel form
                assert fast() == before + 1
                bad = np.flatnonzero((sp != osp) | (ep != oep))
                assert bad.size == 0, (name, L, options, bad[:5], [bytes(c[int(o[i]):int(o[i + 1])]) for i in bad[:3]])
                gsp, gep = gpu.find_batch(c, o)                      # the general kernel
                assert (gsp == osp).all() and (gep == oep).all(), (name, L, options)
                ssp, sep, st = gpu.find_batch(c, o, stats=True)
                assert (ssp == osp).all() and (sep == oep).all() and st["queries"] == n
                fsp, fep, fst = gpu.find_fixed_batch(c, L, stats=True)   # the counting variants of the k-mer form's kernels
                assert (fsp == osp).all() and (fep == oep).all() and fst["queries"] == n and fst["found"] == st["found"], (name, L, options, fst, st)
                assert fst["total_length"] == st["total_length"]
                gpu.close()


def test_ragged_batches_start_in_the_chain_kernel():
    """Batches in the offsets form (at least 4096 patterns of any lengths) start in find_chain_kernel's FRESH form -- k-mer table
    probe, then one jump entry or one sector per round -- and the general kernel finishes its work list: equal to the oracle
    for empty patterns, patterns shorter than the table, longer than 255 characters (left to the general kernel whole), up
    to 600 characters, with substitutions, N / $ / # / garbage bytes, lower case and random patterns, on a linear text, a
    repetitive one and a graph with bubbles, for every table shape; the counting variant agrees on what it found."""
    rng = np.random.default_rng(61)
    seq = synth.random_sequence(150_000, seed=61)
    rep = np.concatenate([np.tile(synth.random_sequence(700, seed=62), 30), synth.random_sequence(20_000, seed=63)])
    graph, sites, alt = synth.snp_graph(seq, seed=61, snp_rate=0.02)
    alphabet = np.frombuffer(b"ACGTACGTACGTACGTacgtN$#x", dtype=np.uint8)
    cases = (("linear", build_index(synth.linear_graph(seq), 16, 3)[0], lambda n, L, s: synth.patterns_from_sequence(seq, n, L, seed=s)),
             ("repeats", build_index(synth.linear_graph(rep), 16, 3)[0], lambda n, L, s: synth.patterns_from_sequence(rep, n, L, seed=s)),
             ("snp", build_index(graph, 16, 3)[0], lambda n, L, s: synth.patterns_from_snp_graph(seq, sites, alt, n, L, seed=s)))
    for name, flat, sampler in cases:
        ora = orc.OracleGCSA(flat)
        pats = [b"", b"A", b"N", b"acgt"]
        for L in (1, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 200, 254, 255, 256, 257, 400, 600):
            n = 260 if L <= 256 else 60
            c, o = sampler(n, L, 300 + L)
            c = c.copy()
            for i in range(n):
                r = i % 7
                if r == 1:
                    p = int(o[i]) + int(rng.integers(0, L)); c[p] = synth.COMP2CHAR[1 + (int(np.where(synth.COMP2CHAR == c[p])[0][0]) % 4)]
                elif r == 2 and i % 14 == 2:
                    c[int(o[i]) + int(rng.integers(0, L))] = alphabet[int(rng.integers(16, alphabet.size))]
                elif r == 3:
                    c[int(o[i]):int(o[i + 1])] |= 0x20
                pats.append(bytes(c[int(o[i]):int(o[i + 1])]))
        pats += [bytes(alphabet[rng.integers(0, 16, size=int(ln))]) for ln in rng.integers(0, 50, size=400)]
        order = rng.permutation(len(pats))
        pats = [pats[i] for i in order]
        assert len(pats) >= 4096
        chars, offsets = orc.pack_patterns(pats)
        osp, oep, _ = ora.find_batch(chars, offsets, threads=4)
        for options in (dict(kmer_table_k=8, jump_table=True, fused_table=True), dict(kmer_table_k=8, jump_table=True, fused_table=False),
                        dict(kmer_table_k=9, jump_table="wide"), dict(kmer_table_k=6, jump_table=False), dict(kmer_table_k=2, jump_table=True)):
            gpu = GCSA(flat, **options)
            sp, ep = gpu.find_batch(chars, offsets)
            bad = np.flatnonzero((sp != osp) | (ep != oep))
            assert bad.size == 0, (name, options, bad[:5], [pats[i] for i in bad[:3]])
            ssp, sep, st = gpu.find_batch(chars, offsets, stats=True)
            assert (ssp == osp).all() and (sep == oep).all() and st["queries"] == len(pats)
            hit = ~((osp + np.uint64(1)) > (oep + np.uint64(1)))       # Range::empty, include/gcsa/utils.h:93-101 ((0, -1) is empty)
            assert st["found"] == int(np.count_nonzero(hit)) and st["total_length"] == int((oep - osp + np.uint64(1))[hit].sum())
            gpu.close()


def test_single_process_multi_gpu_entry_points():
    """gcsa_b200_find_fixed_host_multi / _find_host_multi / _locate_into_host_multi: the batch cut into blocks over several
    handles (here replicas on one device; tests/test_multi_gpu.py runs them on two devices) == the single-handle calls."""
    from gcsa2_b200 import MultiGCSA
    seq = synth.random_sequence(120_000, seed=51)
    graph, sites, alt = synth.snp_graph(seq, seed=51, snp_rate=0.02)
    flat, _, _ = build_index(graph, 16, 3)
    one = GCSA(flat, kmer_table_k=8)
    for replicas in (2, 3):
        multi = MultiGCSA(flat, [0] * replicas, kmer_table_k=8)
        chars, offsets = synth.patterns_from_snp_graph(seq, sites, alt, 30_001, 32, seed=52)
        a, b = one.find_fixed_batch(chars, 32)
        c, d = multi.find_fixed_batch(chars, 32)
        assert (a == c).all() and (b == d).all()
        mchars, moffsets = synth.mixed_length_patterns(seq, sites, alt, 5003, 10, 90, seed=53, error_rate=0.02)
        e, f = one.find_batch(mchars, moffsets)
        g, h = multi.find_batch(mchars, moffsets)
        assert (e == g).all() and (f == h).all()
        # locate: short patterns have wide ranges, mixed with the 32-mers' short ones and with empty ranges
        sp = np.concatenate([e[:2000], a[:20_000]]); ep = np.concatenate([f[:2000], b[:20_000]])
        offs, vals = one.locate_batch(sp, ep)
        moffs, mvals = multi.locate_batch(sp, ep)
        assert (offs == moffs).all() and (vals == mvals).all()
        empty_offs, empty_vals = multi.locate_batch(sp[:0], ep[:0])
        assert empty_offs.tolist() == [0] and empty_vals.size == 0
        multi.close()
    one.close()
