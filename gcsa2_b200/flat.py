"""FlatGCSA / FlatLCP: the index as a handful of plain host arrays.

This is the interchange format at the C-ABI boundary (struct gcsa_flat_index /
gcsa_flat_lcp in include/gcsa2_b200.h).  It names the members of gcsa::GCSA
(reference include/gcsa/gcsa.h:214-240) and gcsa::LCPArray (include/gcsa/lcp.h:182-190)
one by one, with every bit vector stored as plain little-endian 64-bit words (bit i of
the vector is bit (i & 63) of word i >> 6) instead of an SDSL encoding.  The engine
builds its own device layout from it ("Data layout in HBM").
"""
from dataclasses import dataclass, field

import numpy as np

SIGMA = 7
FAST_CHARS = 4
COMP2CHAR = b"$ACGTN#"

# src/support.cpp:69-92: \0 and $ -> 0, ACGT/acgt -> 1..4, # -> 6, everything else -> N (5).
DEFAULT_CHAR2COMP = np.full(256, 5, dtype=np.uint8)
DEFAULT_CHAR2COMP[0] = 0
DEFAULT_CHAR2COMP[ord("$")] = 0
DEFAULT_CHAR2COMP[ord("#")] = 6
for _i, _ch in enumerate("ACGT"):
    DEFAULT_CHAR2COMP[ord(_ch)] = _i + 1
    DEFAULT_CHAR2COMP[ord(_ch.lower())] = _i + 1

# include/gcsa/support.h:443-471 (struct Node)
NODE_OFFSET_BITS = 10
NODE_ID_OFFSET = NODE_OFFSET_BITS + 1


def node_encode(node_id, offset, rc=False):
    return (int(node_id) << NODE_ID_OFFSET) | (int(rc) << NODE_OFFSET_BITS) | int(offset)


def node_id(node): return int(node) >> NODE_ID_OFFSET
def node_rc(node): return bool((int(node) >> NODE_OFFSET_BITS) & 1)
def node_offset(node): return int(node) & ((1 << NODE_OFFSET_BITS) - 1)


def words_for(n_bits):
    return (int(n_bits) + 63) // 64


def bits_from_positions(positions, n_bits):
    """Plain bit vector (uint64 words) with the given positions set."""
    words = np.zeros(max(1, words_for(n_bits)), dtype=np.uint64)
    pos = np.asarray(positions, dtype=np.uint64)
    if pos.size:
        np.bitwise_or.at(words, (pos >> np.uint64(6)).astype(np.int64),
                         np.uint64(1) << (pos & np.uint64(63)))
    return words


def positions_from_bits(words, n_bits):
    bits = np.unpackbits(np.ascontiguousarray(words, dtype=np.uint64).view(np.uint8), bitorder="little")[:int(n_bits)]
    return np.flatnonzero(bits).astype(np.uint64)


@dataclass
class FlatGCSA:
    path_nodes: int
    edge_count: int
    order: int
    C: np.ndarray                      # uint64[SIGMA + 1]
    bwt: list                          # SIGMA plain bit vectors, path_nodes bits each
    edges: np.ndarray                  # edge_count bits; 1 = last outgoing edge of a node
    sampled_paths: np.ndarray          # path_nodes bits
    sample_count: int
    stored_samples: np.ndarray         # uint64[sample_count] (unpacked int_vector<0>)
    samples: np.ndarray                # sample_count bits; 1 = last sample of a node
    extra_filter: np.ndarray           # SadaSparse::filter, path_nodes bits
    extra_values_len: int
    extra_values: np.ndarray           # SadaSparse::values (k -> 0^(k-1) 1)
    redundant_len: int
    redundant: np.ndarray              # SadaCount::data (k -> 0^k 1), path_nodes - 1 ones
    sigma: int = SIGMA
    fast_chars: int = FAST_CHARS
    char2comp: np.ndarray = field(default_factory=lambda: DEFAULT_CHAR2COMP.copy())

    def size(self):
        return self.path_nodes

    def sample_bits(self):
        m = int(self.stored_samples.max()) if self.sample_count else 0
        return max(1, m.bit_length())

    def save(self, path):
        np.savez(path, header=np.array([self.path_nodes, self.edge_count, self.order, self.sample_count,
                                        self.extra_values_len, self.redundant_len, self.sigma, self.fast_chars],
                                       dtype=np.uint64),
                 C=self.C, char2comp=self.char2comp, edges=self.edges, sampled_paths=self.sampled_paths,
                 stored_samples=self.stored_samples, samples=self.samples, extra_filter=self.extra_filter,
                 extra_values=self.extra_values, redundant=self.redundant,
                 **{"bwt%d" % c: self.bwt[c] for c in range(SIGMA)})

    @staticmethod
    def from_gcsa_file(path):
        """Reads a .gcsa file of the reference (GCSA::load, src/gcsa.cpp:182-216) through the C ABI."""
        import ctypes as C
        from . import capi
        built = capi.Built()
        capi.check(capi.lib().gcsa_b200_load_gcsa_file(str(path).encode(), C.byref(built)))
        flat = capi.flat_from_struct(built.index)
        capi.lib().gcsa_b200_built_free(C.byref(built))
        return flat

    def to_gcsa_file(self, path):
        """Writes the index in the reference's file format (GCSA::serialize, src/gcsa.cpp:140-180)."""
        import ctypes as C
        from . import capi
        keep = []
        f = capi.flat_struct(self, keep)
        capi.check(capi.lib().gcsa_b200_write_gcsa_file(C.byref(f), str(path).encode()))

    @staticmethod
    def load(path):
        z = np.load(path)
        h = [int(x) for x in z["header"]]
        return FlatGCSA(path_nodes=h[0], edge_count=h[1], order=h[2], C=z["C"],
                        bwt=[z["bwt%d" % c] for c in range(SIGMA)], edges=z["edges"],
                        sampled_paths=z["sampled_paths"], sample_count=h[3], stored_samples=z["stored_samples"],
                        samples=z["samples"], extra_filter=z["extra_filter"], extra_values_len=h[4],
                        extra_values=z["extra_values"], redundant_len=h[5], redundant=z["redundant"],
                        sigma=h[6], fast_chars=h[7], char2comp=z["char2comp"])


@dataclass
class FlatLCP:
    size: int                          # number of LCP values = path nodes
    branching: int
    levels: int
    offsets: np.ndarray                # uint64[levels + 1]
    data: np.ndarray                   # uint8[offsets[levels]], levels concatenated (lcp.h:182-190)

    @staticmethod
    def from_lcp_file(path):
        """Reads a .lcp file of the reference (LCPArray::load, src/lcp.cpp:128-143) through the C ABI."""
        import ctypes as C
        from . import capi
        f = capi.FlatLcp()
        capi.check(capi.lib().gcsa_b200_load_lcp_file(str(path).encode(), C.byref(f)))
        levels, total = int(f.levels), 0
        offsets = np.ctypeslib.as_array(C.cast(f.offsets, C.POINTER(C.c_uint64)), shape=(levels + 1,)).copy()
        total = int(offsets[levels])
        data = np.ctypeslib.as_array(C.cast(f.data, C.POINTER(C.c_uint8)), shape=(max(1, total),))[:total].copy()
        res = FlatLCP(size=int(f.size), branching=int(f.branching), levels=levels, offsets=offsets, data=data)
        capi.lib().gcsa_b200_flat_lcp_free(C.byref(f))
        return res

    def to_lcp_file(self, path):
        """Writes the array in the reference's file format (LCPArray::serialize, src/lcp.cpp:116-126)."""
        import ctypes as C
        from . import capi
        keep = []
        f = capi.lcp_struct(self, keep)
        capi.check(capi.lib().gcsa_b200_write_lcp_file(C.byref(f), str(path).encode()))

    @staticmethod
    def from_values(lcp_values, branching=64):
        """Range-minimum tree over an LCP array, laid out like src/lcp.cpp:224-258."""
        lcp_values = np.ascontiguousarray(lcp_values, dtype=np.uint8)
        size = int(lcp_values.size)
        level_sizes = [size]
        while level_sizes[-1] > 1:
            level_sizes.append((level_sizes[-1] + branching - 1) // branching)
        offsets = np.zeros(len(level_sizes) + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum(level_sizes)
        parts = [lcp_values]
        for lv in range(1, len(level_sizes)):
            prev = parts[-1]
            pad = (-prev.size) % branching
            padded = np.concatenate([prev, np.full(pad, 255, dtype=np.uint8)]) if pad else prev
            parts.append(padded.reshape(-1, branching).min(axis=1))
        return FlatLCP(size=size, branching=branching, levels=len(level_sizes),
                       offsets=offsets, data=np.concatenate(parts) if size else np.zeros(0, dtype=np.uint8))
