"""Index files (.gcsa / .lcp): gcsa_b200_load_gcsa_file / _write_gcsa_file / _load_lcp_file / _write_lcp_file
against GCSA::serialize / GCSA::load and LCPArray::serialize / load of the reference
(src/gcsa.cpp:140-216, src/lcp.cpp:116-143).

  * a file written by the reference's own serialize() (its sources over the SDSL shim, whose member
    encodings follow the sdsl-lite on-disk layout) loads into arrays identical to the builder's;
  * a file written by us is accepted by the reference's own load(), which then answers queries like
    the oracle, and is byte-identical to the reference's own file;
  * write -> load round trip; corrupted and truncated files are rejected with a message.
"""
import os

import numpy as np
import pytest

from test_builder import flat_equal
from gcsa2_b200 import capi, synth
from gcsa2_b200.builder import build_index
from gcsa2_b200.flat import FlatGCSA, FlatLCP
from oracle import oracle as orc
from oracle import reference as ref

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libgcsa2_ref.so not built (needs /root/reference)")


def graphs():
    yield "snp", synth.snp_graph(synth.random_sequence(20000, seed=21), seed=21, snp_rate=0.03)[0], 16, 2, 64
    yield "linear", synth.linear_graph(synth.random_sequence(9000, seed=12)), 16, 3, 64
    unit = np.array([1, 2, 1, 3], dtype=np.uint8)
    yield "repeats", synth.linear_graph(np.concatenate([np.tile(unit, 50), synth.random_sequence(40, 13)]), node_len=8), 4, 2, 4


@pytest.fixture(scope="module", params=list(graphs()), ids=lambda c: c[0])
def built(request):
    name, graph, k, steps, branching = request.param
    flat, flcp, kmers = build_index(graph, k, steps, lcp_branching=branching)
    return name, flat, flcp, kmers, steps, branching


def lcp_equal(a, b):
    return (a.size, a.branching, a.levels) == (b.size, b.branching, b.levels) and \
        list(a.offsets) == list(b.offsets) and (a.data == b.data).all()


def test_round_trip(built, tmp_path):
    name, flat, flcp, kmers, steps, branching = built
    flat.to_gcsa_file(tmp_path / "x.gcsa"); flcp.to_lcp_file(tmp_path / "x.lcp")
    assert flat_equal(flat, FlatGCSA.from_gcsa_file(tmp_path / "x.gcsa")) == []
    assert lcp_equal(flcp, FlatLCP.from_lcp_file(tmp_path / "x.lcp"))


def test_header_layout(built, tmp_path):
    """GCSAHeader / LCPHeader as the reference writes them (src/files.cpp:513-538, 581-604)."""
    name, flat, flcp, kmers, steps, branching = built
    flat.to_gcsa_file(tmp_path / "x.gcsa"); flcp.to_lcp_file(tmp_path / "x.lcp")
    raw = open(tmp_path / "x.gcsa", "rb").read()
    assert np.frombuffer(raw[:8], dtype=np.uint32).tolist() == [0x6C5A6C5A, 3]
    assert np.frombuffer(raw[8:40], dtype=np.uint64).tolist() == [flat.path_nodes, flat.edge_count, flat.order, 0]
    # Alphabet: int_vector<8> char2comp = bit length 2048 followed by the 256 bytes
    assert int(np.frombuffer(raw[40:48], dtype=np.uint64)[0]) == 2048
    assert (np.frombuffer(raw[48:48 + 256], dtype=np.uint8) == flat.char2comp).all()
    raw = open(tmp_path / "x.lcp", "rb").read()
    assert np.frombuffer(raw[:8], dtype=np.uint32).tolist() == [0x6C5A7C94, 1]
    assert np.frombuffer(raw[8:32], dtype=np.uint64).tolist() == [flcp.size, flcp.branching, 0]


def test_rejects_damaged_files(built, tmp_path):
    name, flat, flcp, kmers, steps, branching = built
    good = tmp_path / "x.gcsa"
    flat.to_gcsa_file(good)
    raw = bytearray(open(good, "rb").read())
    def expect_failure(data, what):
        path = tmp_path / "bad.gcsa"
        open(path, "wb").write(bytes(data))
        with pytest.raises(capi.GCSAError) as err:
            FlatGCSA.from_gcsa_file(path)
        assert what in str(err.value), str(err.value)
    expect_failure(raw[:len(raw) // 2], "end of file")
    expect_failure(raw + b"\0", "trailing")
    bad = bytearray(raw); bad[0] ^= 1
    expect_failure(bad, "tag")
    bad = bytearray(raw); bad[4] = 2
    expect_failure(bad, "version")
    bad = bytearray(raw); bad[8] ^= 1                                   # path_nodes
    expect_failure(bad, "")
    # flip one data bit of the first fast BWT vector: its cumulative counts no longer match
    alphabet_end = 40 + (8 + 256) + (8 + 8) + (8 + 64) + 16
    first_fast = alphabet_end + 6 * 8                                   # the empty vector of comp 0
    bad = bytearray(raw); bad[first_fast + 4 * 8 + 8 + 8] ^= 1         # 4 members, data length, first count word
    expect_failure(bad, "")
    with pytest.raises(capi.GCSAError):
        FlatGCSA.from_gcsa_file(tmp_path / "does_not_exist.gcsa")
    with pytest.raises(capi.GCSAError):
        FlatLCP.from_lcp_file(good)                                      # wrong tag


@needs_ref
def test_reads_files_written_by_the_reference(built, tmp_path):
    name, flat, flcp, kmers, steps, branching = built
    reference = ref.ReferenceIndex.build(kmers, steps, lcp_branching=branching)
    reference.store(tmp_path / "ref.gcsa", tmp_path / "ref.lcp")
    assert flat_equal(flat, FlatGCSA.from_gcsa_file(tmp_path / "ref.gcsa")) == []
    assert lcp_equal(flcp, FlatLCP.from_lcp_file(tmp_path / "ref.lcp"))
    # and our writer produces the same bytes as the reference's serialize()
    flat.to_gcsa_file(tmp_path / "ours.gcsa"); flcp.to_lcp_file(tmp_path / "ours.lcp")
    assert open(tmp_path / "ours.lcp", "rb").read() == open(tmp_path / "ref.lcp", "rb").read()
    assert open(tmp_path / "ours.gcsa", "rb").read() == open(tmp_path / "ref.gcsa", "rb").read()


@needs_ref
def test_reference_loads_our_files(built, tmp_path):
    name, flat, flcp, kmers, steps, branching = built
    flat.to_gcsa_file(tmp_path / "ours.gcsa"); flcp.to_lcp_file(tmp_path / "ours.lcp")
    loaded = ref.ReferenceIndex.load(tmp_path / "ours.gcsa", tmp_path / "ours.lcp")
    assert loaded is not None
    rflat, rlcp = loaded.export()
    assert flat_equal(flat, rflat) == [] and lcp_equal(flcp, rlcp)
    from verify import kmer_table
    pats = [s.encode() if isinstance(s, str) else bytes(s) for s, _ in kmer_table(kmers)[:2000]]
    chars = np.frombuffer(b"".join(pats), dtype=np.uint8)
    offsets = np.zeros(len(pats) + 1, dtype=np.uint64); offsets[1:] = np.cumsum([len(p) for p in pats])
    ora = orc.OracleGCSA(flat)
    sp, ep, _ = ora.find_batch(chars, offsets)
    rsp, rep, _ = loaded.find_batch(chars, offsets)
    assert (sp == rsp).all() and (ep == rep).all()
    assert (loaded.count_batch(sp, ep) == ora.count_batch(sp, ep)[0]).all()
    o1, v1, _ = loaded.locate_batch(sp[:500], ep[:500]); o2, v2, _ = ora.locate_batch(sp[:500], ep[:500])
    assert (o1 == o2).all() and (v1 == v2).all()
    corrupt = bytearray(open(tmp_path / "ours.gcsa", "rb").read()); corrupt[4] = 9
    open(tmp_path / "bad.gcsa", "wb").write(bytes(corrupt))
    assert ref.ReferenceIndex.load(tmp_path / "bad.gcsa") is None        # GCSA::load throws on a bad header (gcsa.cpp:188-193)


@pytest.mark.engine
def test_engine_on_loaded_files(built, tmp_path):
    """The CUDA engine over an index that went through the file formats answers like the oracle."""
    from gcsa2_b200 import GCSA, LCPArray
    from verify import kmer_table
    name, flat, flcp, kmers, steps, branching = built
    flat.to_gcsa_file(tmp_path / "x.gcsa"); flcp.to_lcp_file(tmp_path / "x.lcp")
    gpu, glcp = GCSA.load(tmp_path / "x.gcsa", kmer_table_k=4), LCPArray.load(tmp_path / "x.lcp")
    ora, olcp = orc.OracleGCSA(flat), orc.OracleLCP(flcp)
    pats = [s.encode() if isinstance(s, str) else bytes(s) for s, _ in kmer_table(kmers)[:3000]]
    chars = np.frombuffer(b"".join(pats), dtype=np.uint8)
    offsets = np.zeros(len(pats) + 1, dtype=np.uint64); offsets[1:] = np.cumsum([len(p) for p in pats])
    sp, ep = gpu.find_batch(chars, offsets)
    osp, oep, _ = ora.find_batch(chars, offsets)
    assert (sp == osp).all() and (ep == oep).all()
    offs, vals = gpu.locate_batch(sp, ep); ooffs, ovals, _ = ora.locate_batch(sp, ep)
    assert (offs == ooffs).all() and (vals == ovals).all()
    assert (gpu.count_batch(sp, ep) == ora.count_batch(sp, ep)[0]).all()
    assert (glcp.parent_batch(sp, ep) == olcp.parent_batch(sp, ep)[0]).all()
