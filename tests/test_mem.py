"""The MEM-style scan of config 5 (LF + parent driver): properties of the CPU definition, and
bit-exact agreement of the device kernel with it."""
import random

import numpy as np
import pytest

from gcsa2_b200 import synth
from gcsa2_b200.builder import build_index
from oracle import oracle as orc


def make(n_patterns, L=50_000, seed=1):
    seq = synth.random_sequence(L, seed=seed)
    graph, sites, alt = synth.snp_graph(seq, seed=seed, snp_rate=0.01)
    flat, flcp, _ = build_index(graph, 16, 3)
    chars, offsets = synth.mixed_length_patterns(seq, sites, alt, n_patterns, 16, 256, seed=seed + 7, error_rate=0.01)
    return flat, flcp, chars, offsets


def test_mem_scan_definition_properties():
    flat, flcp, chars, offsets = make(1500)
    index, lcp = orc.OracleGCSA(flat), orc.OracleLCP(flcp)
    offs, vals, _ = orc.mem_batch(index, lcp, chars, offsets, threads=4)
    assert int(offs[-1]) == len(vals) and (np.diff(offs.astype(np.int64)) >= 1).all()
    for i in random.Random(1).sample(range(1500), 150):
        P = bytes(chars[int(offsets[i]):int(offsets[i + 1])])
        ms = vals[int(offs[i]):int(offs[i + 1])]
        starts = [int(m[0]) for m in ms]
        assert starts == sorted(starts, reverse=True)                     # reported right to left
        for st, ln, sp, ep in ((int(a), int(b), int(c), int(d)) for a, b, c, d in ms):
            assert ln >= 1 and st + ln <= len(P)
            assert index.find(P[st:st + ln]) == (sp, ep)                  # it is find() of that substring
            if st > 0:                                                    # and it is left-maximal
                r = index.find(P[st - 1:st + ln])
                assert ((r[0] + 1) % 2**64) > ((r[1] + 1) % 2**64)
    # an exact substring of the graph is one match covering the whole pattern
    seq = synth.random_sequence(50_000, seed=1)
    P = bytes(synth.COMP2CHAR[seq[1000:1100]])
    graph, _, _ = synth.snp_graph(seq, seed=1, snp_rate=0.01)
    c, o = orc.pack_patterns([P, b"", b"NNNN"])
    offs, vals, _ = orc.mem_batch(index, lcp, c, o)
    assert list(offs) == [0, 1, 1, 1] and list(vals[0][:2]) == [0, 100]


@pytest.mark.engine
def test_mem_scan_device_matches_definition(monkeypatch):
    from gcsa2_b200 import GCSA, LCPArray, mem_batch
    flat, flcp, chars, offsets = make(30_000, L=200_000, seed=3)
    index, lcp = orc.OracleGCSA(flat), orc.OracleLCP(flcp)
    ooffs, ovals, _ = orc.mem_batch(index, lcp, chars, offsets, threads=8)
    for two_step, jump in ((False, "0"), (True, "0"), (False, "1")):
        monkeypatch.setenv("GCSA_B200_MEM_JUMP", jump)               # "1": singleton ranges follow the jump tables
        gpu, glcp = GCSA(flat, kmer_table_k=6, two_step=two_step), LCPArray(flcp)
        offs, vals = mem_batch(gpu, glcp, chars, offsets)
        assert (offs == ooffs).all() and vals.shape == ovals.shape and (vals == ovals).all(), (two_step, jump)
    c, o = orc.pack_patterns([b"", b"ACGT", b"NNNN", b"$", bytes(chars[:300])])
    offs, vals = mem_batch(gpu, glcp, c, o)
    eoffs, evals, _ = orc.mem_batch(index, lcp, c, o)
    assert (offs == eoffs).all() and (vals == evals).all()
    offs, vals = mem_batch(gpu, glcp, [])
    assert list(offs) == [0] and vals.shape == (0, 4)


@pytest.mark.engine
def test_mem_scan_scratch_paths(monkeypatch):
    """The one-pass scan (matches staged in a per-pattern scratch slot, overflowing patterns redone), with
    slots that grow with the pattern (the default), slots of 16, 4 and 1 matches, and the two-pass fallback: all equal to the definition.  Noisy patterns
    (10 % substitutions) so that many patterns have more matches than a slot holds."""
    from gcsa2_b200 import GCSA, LCPArray, mem_batch, mem_device
    import torch
    from helpers import current_stream, device_empty, device_sync, to_device
    seq = synth.random_sequence(100_000, seed=5)
    graph, sites, alt = synth.snp_graph(seq, seed=5, snp_rate=0.01)
    flat, flcp, _ = build_index(graph, 16, 3)
    import helpers
    n_patterns = 8_000 if helpers.EMULATED else 20_000                # the emulated run is the slowest test of the CPU suite
    chars, offsets = synth.mixed_length_patterns(seq, sites, alt, n_patterns, 16, 256, seed=11, error_rate=0.10)
    index, lcp = orc.OracleGCSA(flat), orc.OracleLCP(flcp)
    ooffs, ovals, _ = orc.mem_batch(index, lcp, chars, offsets, threads=8)
    counts = np.diff(ooffs.astype(np.int64))
    assert (counts > 16).sum() > 40 and (counts <= 4).sum() > 40
    gpu, glcp = GCSA(flat, kmer_table_k=0), LCPArray(flcp)
    assert (counts > np.diff(offsets.astype(np.int64)) // 4 + 4).sum() > 10      # ... and more than a slot that grows with the pattern
    for stride, shift, jump in ((None, None, "0"), (None, "2", "0"), (None, "1", "0"), ("16", None, "0"), ("4", None, "0"), ("1", None, "0"), ("0", None, "0"),
                                ("4", None, "1"), ("0", None, "1"), (None, "2", "1")):
        # no stride: slots of (len >> shift) + 4 matches; the default shift is 0 where the memory allows it (no overflow possible)
        for name, value in (("GCSA_B200_MEM_STRIDE", stride), ("GCSA_B200_MEM_SHIFT", shift)):
            if value is None: monkeypatch.delenv(name, raising=False)
            else: monkeypatch.setenv(name, value)
        monkeypatch.setenv("GCSA_B200_MEM_JUMP", jump)
        offs, vals = mem_batch(gpu, glcp, chars, offsets)
        assert (offs == ooffs).all() and vals.shape == ovals.shape and (vals == ovals).all(), (stride, shift, jump)
    monkeypatch.delenv("GCSA_B200_MEM_STRIDE", raising=False)
    monkeypatch.delenv("GCSA_B200_MEM_SHIFT", raising=False)
    monkeypatch.delenv("GCSA_B200_MEM_JUMP")
    # device entry point with a caller buffer: too small -> the needed size, then the same answer
    from gcsa2_b200 import capi
    n = len(offsets) - 1
    d_chars = to_device(chars); d_off = to_device(offsets.view(np.int64))
    d_out = device_empty(n + 1, torch.int64)
    small = device_empty((10, 4), torch.int64)
    with pytest.raises(capi.GCSAError) as err:
        mem_device(gpu, glcp, d_chars, d_off, n, d_out, small, 10, current_stream())
    assert err.value.code == capi.ERR_CAPACITY
    total = int(ooffs[-1])
    big = device_empty((total, 4), torch.int64)
    got = mem_device(gpu, glcp, d_chars, d_off, n, d_out, big, total, current_stream())
    device_sync()
    assert got == total and (d_out.cpu().numpy().view(np.uint64) == ooffs).all() and (big.cpu().numpy().view(np.uint64) == ovals).all()
