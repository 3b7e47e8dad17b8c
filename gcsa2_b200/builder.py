"""Host-side index construction: graph -> kmers -> FlatGCSA / FlatLCP.

Thin wrapper over the C++ builder (gcsa2_b200/csrc/builder.cpp) through the C ABI.  Stands in
for build_gcsa / GCSA::GCSA(InputGraph&, ...) of the reference (src/gcsa.cpp:447-724) for
in-memory inputs; construction stays on the CPU.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .flat import FlatGCSA, FlatLCP, SIGMA, words_for


@dataclass
class CharGraph:
    """A graph of single-character nodes in CSR form (struct gcsa_b200_graph)."""
    comp: np.ndarray            # uint8[n]
    value: np.ndarray           # uint64[n] node_type
    succ_offsets: np.ndarray    # uint64[n + 1]
    succ: np.ndarray            # uint64[edges]
    sink: int
    sources: np.ndarray         # uint64[], nodes that get predecessor '$' (technical edge sink -> source)

    @property
    def nodes(self):
        return int(self.comp.size)

    @staticmethod
    def from_lists(comps, values, succ, sources, sink):
        offsets = np.zeros(len(comps) + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum([len(s) for s in succ])
        flat = np.array([w for s in succ for w in s], dtype=np.uint64)
        return CharGraph(comp=np.array(comps, dtype=np.uint8), value=np.array(values, dtype=np.uint64),
                         succ_offsets=offsets, succ=flat, sink=int(sink), sources=np.array(sources, dtype=np.uint64))


@dataclass
class KMers:
    """KMer records of the reference (include/gcsa/support.h:475-497): key = label << 16 |
    predecessor mask << 8 | successor mask; one record per (kmer, successor position)."""
    key: np.ndarray
    from_: np.ndarray
    to: np.ndarray
    k: int

    def labels(self):
        """Decoded labels (comp tuples), for tests."""
        lab = self.key >> np.uint64(16)
        return [tuple(int((x >> np.uint64(3 * (self.k - 1 - i))) & np.uint64(7)) for i in range(self.k)) for x in lab]

    def write_text(self, path):
        """Text .gcsa2 format of the reference (src/files.cpp:86-124): kmer, start position, predecessor characters,
        successor characters, successor positions; one line per (kmer, start position)."""
        comp2char = "$ACGTN#"
        def node(v):
            v = int(v)
            return "%d:%s%d" % (v >> 11, "-" if (v >> 10) & 1 else "", v & 0x3FF)
        def chars(mask):
            return ",".join(comp2char[c] for c in range(7) if (mask >> c) & 1)
        lines = {}
        for key, frm, to in zip(self.key.tolist(), self.from_.tolist(), self.to.tolist()):
            lines.setdefault((key, frm), []).append(to)
        with open(path, "w") as f:
            for (key, frm), tos in lines.items():
                label = "".join(comp2char[(key >> (16 + 3 * (self.k - 1 - i))) & 7] for i in range(self.k))
                f.write("%s\t%s\t%s\t%s\t%s\n" % (label, node(frm), chars((key >> 8) & 0xFF), chars(key & 0xFF), ",".join(node(t) for t in tos)))

    def write_binary(self, path):
        """Binary .graph format of the reference: 24-byte GraphFileHeader {flags, kmer_count,
        kmer_length} followed by 24-byte KMer {key, from, to} records (include/gcsa/files.h:40-52,
        src/files.cpp:127-187)."""
        with open(path, "wb") as f:
            np.array([0, self.key.size, self.k], dtype=np.uint64).tofile(f)
            rec = np.empty((self.key.size, 3), dtype=np.uint64)
            rec[:, 0], rec[:, 1], rec[:, 2] = self.key, self.from_, self.to
            rec.tofile(f)


@dataclass
class NodeMapping:
    """NodeMapping of the reference (include/gcsa/support.h:167-222): node ids first .. first + len(ids) - 1 of the input
    graph are reported as ids[id - first]."""
    first: int
    ids: np.ndarray

    def __call__(self, values):
        """Node::map (src/support.cpp:604-612) on an array of node_type values."""
        values = np.asarray(values, dtype=np.uint64)
        node = (values >> np.uint64(11)).astype(np.int64)
        inside = (node >= self.first) & (node < self.first + len(self.ids))
        mapped = np.where(inside, np.asarray(self.ids, dtype=np.uint64)[np.clip(node - self.first, 0, max(0, len(self.ids) - 1))], node.astype(np.uint64))
        return (mapped << np.uint64(11)) | (values & np.uint64(0x7FF))

    def write(self, path):
        """NodeMapping::serialize (src/support.cpp:316-333): first_node, next_node, the ids."""
        with open(path, "wb") as f:
            np.array([self.first, self.first + len(self.ids)], dtype=np.uint64).tofile(f)
            np.asarray(self.ids, dtype=np.uint64).tofile(f)

    @staticmethod
    def load(path):
        first, p, size = C.c_uint64(), C.c_void_p(), C.c_uint64()
        capi.check(capi.lib().gcsa_b200_load_node_mapping(str(path).encode(), C.byref(first), C.byref(p), C.byref(size)))
        n = int(size.value)
        ids = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(max(n, 1),))[:n].copy()
        capi.lib().gcsa_b200_free(p)
        return NodeMapping(first=int(first.value), ids=ids)


def read_kmers(paths, binary=True, char2comp=None):
    """The kmer files of the reference (vg's output): binary .graph or text .gcsa2 (src/files.cpp:86-167); several files
    are concatenated like InputGraph does.  -> KMers."""
    if isinstance(paths, (str, bytes)) or hasattr(paths, "__fspath__"):
        paths = [paths]
    arr = (C.c_char_p * len(paths))(*[str(p).encode() for p in paths])
    out = capi.Kmers(); k = C.c_int()
    table = None if char2comp is None else np.ascontiguousarray(char2comp, dtype=np.uint8)
    capi.check(capi.lib().gcsa_b200_read_kmer_files(arr, len(paths), int(bool(binary)), None if table is None else table.ctypes.data,
                                                    C.byref(out), C.byref(k)))
    n = int(out.n)
    def take(p):
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(max(n, 1),))[:n].copy()
    res = KMers(key=take(out.key), from_=take(out.from_), to=take(out.to), k=int(k.value))
    capi.lib().gcsa_b200_kmers_free(C.byref(out))
    return res


def enumerate_kmers(graph, k):
    g = capi.Graph()
    comp = np.ascontiguousarray(graph.comp, dtype=np.uint8)
    value = capi.as_u64(graph.value); offs = capi.as_u64(graph.succ_offsets)
    succ = capi.as_u64(graph.succ); sources = capi.as_u64(graph.sources)
    g.nodes, g.comp, g.value = graph.nodes, comp.ctypes.data, value.ctypes.data
    g.succ_offsets, g.succ, g.sink = offs.ctypes.data, succ.ctypes.data, int(graph.sink)
    g.n_sources, g.sources = int(np.asarray(graph.sources).size), sources.ctypes.data
    out = capi.Kmers()
    capi.check(capi.lib().gcsa_b200_enumerate_kmers(C.byref(g), int(k), C.byref(out)))
    n = int(out.n)
    def take(p):
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(max(n, 1),))[:n].copy()
    res = KMers(key=take(out.key), from_=take(out.from_), to=take(out.to), k=int(k))
    capi.lib().gcsa_b200_kmers_free(C.byref(out))
    return res


def build_from_kmers(kmers, doubling_steps, sample_period=64, lcp_branching=64, allow_inconsistent=False, mapping=None):
    """-> (FlatGCSA, FlatLCP).  Raises GCSAError(ERR_INCONSISTENT) if the kmers do not describe a
    graph whose order-K pruned de Bruijn graph satisfies the GCSA invariants.  mapping: a NodeMapping (the index
    reports mapped node ids, InputGraph::mapping of the reference)."""
    key, frm, to = capi.as_u64(kmers.key), capi.as_u64(kmers.from_), capi.as_u64(kmers.to)
    built = capi.Built()
    ids = capi.as_u64(mapping.ids) if mapping is not None else None
    rc = capi.lib().gcsa_b200_build_from_kmers_mapped(key.ctypes.data, frm.ctypes.data, to.ctypes.data, int(kmers.key.size),
                                                      int(kmers.k), int(doubling_steps), int(sample_period),
                                                      int(mapping.first) if mapping is not None else 0,
                                                      ids.ctypes.data if mapping is not None else None,
                                                      len(mapping.ids) if mapping is not None else 0, C.byref(built))
    capi.check(rc, allow=(capi.ERR_INCONSISTENT,) if allow_inconsistent else ())
    flat = capi.flat_from_struct(built.index)
    lcp = np.ctypeslib.as_array(C.cast(built.lcp, C.POINTER(C.c_uint8)), shape=(max(1, int(built.lcp_size)),))[:int(built.lcp_size)].copy()
    capi.lib().gcsa_b200_built_free(C.byref(built))
    flat.consistent = (rc == 0)
    return flat, FlatLCP.from_values(lcp, branching=lcp_branching)


def build_index(graph, k, doubling_steps, sample_period=64, lcp_branching=64, allow_inconsistent=False):
    """graph -> order k * 2^steps index; returns (FlatGCSA, FlatLCP, KMers)."""
    kmers = enumerate_kmers(graph, k)
    flat, lcp = build_from_kmers(kmers, doubling_steps, sample_period, lcp_branching, allow_inconsistent)
    return flat, lcp, kmers


class BuiltIndex:
    """The arrays of a freshly built index, still owned by the library (struct gcsa_b200_built): hand it to GCSA(...)
    as it is -- no numpy copies, which matters at 3 Gbp -- or take copies with flat() / lcp().  free() releases it."""

    def __init__(self, built):
        self._built = built

    @property
    def struct(self):
        if self._built is None:
            raise ValueError("BuiltIndex: already freed")
        return self._built.index

    @property
    def char2comp(self):
        return np.frombuffer(bytes(self.struct.char2comp), dtype=np.uint8).copy()

    @property
    def C(self):
        return np.array([self.struct.C[i] for i in range(SIGMA + 1)], dtype=np.uint64)

    @property
    def path_nodes(self):
        return int(self.struct.path_nodes)

    def flat(self):
        flat = capi.flat_from_struct(self.struct)
        flat.consistent = True
        return flat

    def lcp_values(self):
        n = int(self._built.lcp_size)
        return np.ctypeslib.as_array(C.cast(self._built.lcp, C.POINTER(C.c_uint8)), shape=(max(1, n),))[:n].copy()

    def lcp(self, branching=64):
        return FlatLCP.from_values(self.lcp_values(), branching=branching)

    def free(self):
        if self._built is not None:
            capi.lib().gcsa_b200_built_free(C.byref(self._built))
            self._built = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def build_linear(seq, k=16, doubling_steps=3, node_len=32, sample_period=64, lcp_branching=64, device=0, raw=False):
    """Index of the linear reference `seq` (comp values 1..5; a numpy array or a CUDA uint8 tensor) built on the
    device (csrc/linear_builder.cu): the same arrays build_index(synth.linear_graph(seq, node_len), k, doubling_steps)
    returns, without enumerating kmers.  -> (FlatGCSA, FlatLCP), or with raw=True a BuiltIndex (the arrays stay in the
    library's buffers)."""
    on_device = not isinstance(seq, np.ndarray)
    if not on_device:
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
    built = capi.Built()
    capi.check(capi.lib().gcsa_b200_build_linear(capi.ptr(seq), int(seq.numel() if on_device else seq.size), int(on_device),
                                                 int(node_len), int(k), int(doubling_steps), int(sample_period), int(device),
                                                 C.byref(built)))
    res = BuiltIndex(built)
    if raw:
        return res
    flat, lcp = res.flat(), res.lcp(lcp_branching)
    res.free()
    return flat, lcp
