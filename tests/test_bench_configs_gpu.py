"""The configurations bench.py times, at their full size, against the REFERENCE ITSELF (oracle/_ref: jltsiren/gcsa2's own
find() and locate() over the same index arrays) -- the parity of exactly the engine variants the benchmark lines come
from: configs[1] with the 16-mer table in its fused form and the jump tables (the two-kernel k-mer form of find()), and
configs[2] at full size through the short-range locate() path.  GPU only (the tables alone are 69 GB)."""
import numpy as np
import pytest

from gcsa2_b200 import GCSA, synth
from gcsa2_b200.builder import build_index, build_linear
from oracle import oracle as orc
from oracle import reference as ref


def cpu_checker(flat):
    """The reference's own code where its library is there, else the C restatement (pinned to it in test_reference.py)."""
    if ref.available():
        return ref.ReferenceIndex.from_flat(flat), ref.lib().ref_max_threads()
    return orc.OracleGCSA(flat), orc.lib().oracle_max_threads()


@pytest.mark.gpu
def test_config2_benchmarked_engine_against_the_reference():
    """100 Mbp linear reference, order 128, kmer_table_k = 16, fused table, jump tables: 1 M sampled 32-mers and 1 M
    uniform random 32-mers == the reference's find(); all 10 M sampled 32-mers are found; the host entry point (packed and
    raw chunks) == the device entry point."""
    import torch
    L, n, length = 100_000_000, 10_000_000, 32
    seq = synth.random_sequence(L, seed=2)
    flat, _ = build_linear(seq, k=16, doubling_steps=3, node_len=32)
    assert flat.path_nodes == L + 2 and flat.order == 128
    gpu = GCSA(flat, kmer_table_k=16, fused_table=True, jump_table=True)
    assert gpu.kmerTableK() == 16 and gpu.fusedTable() and gpu.jumpK() == 16
    chars = np.empty(n * length, dtype=np.uint8)
    for i, q0 in enumerate(range(0, n, 1_000_000)):
        c, _ = synth.patterns_from_sequence(seq, 1_000_000, length, seed=100_000 + i)        # bench.py: make_patterns(seed=100)
        chars[q0 * length:(q0 + 1_000_000) * length] = c
    d_chars = torch.from_numpy(chars).cuda()
    d_sp = torch.empty(n, dtype=torch.int64, device="cuda"); d_ep = torch.empty_like(d_sp)
    gpu.find_fixed_device(d_chars, length, n, d_sp, d_ep, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    sp = d_sp.cpu().numpy().view(np.uint64); ep = d_ep.cpu().numpy().view(np.uint64)
    assert not np.any(sp > ep)                                        # everything sampled from the text occurs
    hsp, hep = gpu.find_fixed_batch(chars, length)                    # host entry point: raw and packed chunks
    assert (hsp == sp).all() and (hep == ep).all()
    rchars, _ = synth.random_patterns(1_000_000, length, seed=900)
    rsp, rep = gpu.find_fixed_batch(rchars, length)
    checker, threads = cpu_checker(flat)
    m = 1_000_000
    offsets = np.arange(m + 1, dtype=np.uint64) * np.uint64(length)
    csp, cep, _ = checker.find_batch(chars[:m * length], offsets, threads=threads)
    assert (csp == sp[:m]).all() and (cep == ep[:m]).all()
    csp, cep, _ = checker.find_batch(rchars, offsets, threads=threads)
    assert (csp == rsp).all() and (cep == rep).all()
    assert np.count_nonzero(rsp <= rep) < 1000                        # random 32-mers miss
    gpu.close()


@pytest.mark.gpu
def test_config3_full_size_locate_against_the_reference():
    """50 Mbp backbone with 1 % SNP bubbles, order 128: 10 M 64-mers from walks through the graph, find() then
    locate(range) through the short-range path; 1 M ranges == the reference's own locate() (src/gcsa.cpp:827-842), count()
    == the sizes, and every position returned is one the pattern was sampled from or an equally valid occurrence."""
    L, n, length = 50_000_000, 10_000_000, 64
    seq = synth.random_sequence(L, seed=3)
    graph, sites, alt = synth.snp_graph(seq, seed=3, snp_rate=0.01)
    flat, _, _ = build_index(graph, 16, 3)
    gpu = GCSA(flat, kmer_table_k=12)
    chars = np.empty(n * length, dtype=np.uint8)
    for i, q0 in enumerate(range(0, n, 1_000_000)):
        c, _ = synth.patterns_from_snp_graph(seq, sites, alt, 1_000_000, length, seed=7000 + i)
        chars[q0 * length:(q0 + 1_000_000) * length] = c
    sp, ep = gpu.find_fixed_batch(chars, length)
    assert not np.any(sp > ep)
    offs, vals = gpu.locate_batch(sp, ep)
    assert vals.size >= n and (gpu.count_batch(sp, ep) == np.diff(offs)).all()
    checker, threads = cpu_checker(flat)
    m = 1_000_000
    roffs, rvals, _ = checker.locate_batch(sp[:m], ep[:m], threads=threads)
    k = int(roffs[m])
    assert (offs[:m + 1] == roffs).all() and (vals[:k] == rvals).all()
    offsets = np.arange(200_001, dtype=np.uint64) * np.uint64(length)
    csp, cep, _ = checker.find_batch(chars[:200_000 * length], offsets, threads=threads)
    assert (csp == sp[:200_000]).all() and (cep == ep[:200_000]).all()
    # the ranges of short patterns on the same index (tens to thousands of path nodes each): the register sort by a warp
    # or a block == the reference's locate() on a sample, count() == the sizes, every range sorted and free of duplicates
    for plen, nq, sample in ((10, 300_000, 20_000), (8, 60_000, 3_000), (7, 20_000, 1_000), (6, 3_000, 300)):
        pchars, poffsets = synth.patterns_from_snp_graph(seq, sites, alt, nq, plen, seed=600 + plen)
        psp, pep = gpu.find_batch(pchars, poffsets)
        poffs, pvals = gpu.locate_batch(psp, pep)
        assert (gpu.count_batch(psp, pep) == np.diff(poffs)).all()
        inner = np.ones(pvals.size, dtype=bool); inner[poffs[:-1][np.diff(poffs) > 0].astype(np.int64)] = False      # not the first of its range
        assert (pvals[1:][inner[1:]] > pvals[:-1][inner[1:]]).all()
        roffs, rvals, _ = checker.locate_batch(psp[:sample], pep[:sample], threads=threads)
        assert (poffs[:sample + 1] == roffs).all() and (pvals[:int(roffs[sample])] == rvals).all(), plen
    gpu.close()
