"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this module.  The product package gcsa2_b200 never does.

The index is handed over as any object with the attributes of gcsa2_b200.flat.FlatGCSA
(plain numpy bit vectors and arrays); the oracle builds its own rank directories.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
SIGMA = 7
UNKNOWN = (1 << 64) - 1


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc, no external dependencies)."""
    src_time = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("gcsa_oracle.c", "gcsa_oracle.h", "Makefile"))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_time:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Flat(C.Structure):
    _fields_ = [
        ("path_nodes", C.c_uint64), ("edge_count", C.c_uint64), ("order", C.c_uint64),
        ("sigma", C.c_uint64), ("fast_chars", C.c_uint64),
        ("C", C.c_uint64 * (SIGMA + 1)),
        ("char2comp", C.c_uint8 * 256),
        ("bwt", C.c_void_p * SIGMA),
        ("edges", C.c_void_p),
        ("sampled_paths", C.c_void_p),
        ("sample_count", C.c_uint64),
        ("stored_samples", C.c_void_p),
        ("samples", C.c_void_p),
        ("extra_filter", C.c_void_p),
        ("extra_values_len", C.c_uint64),
        ("extra_values", C.c_void_p),
        ("redundant_len", C.c_uint64),
        ("redundant", C.c_void_p),
    ]


class STNode(C.Structure):
    _fields_ = [("sp", C.c_uint64), ("ep", C.c_uint64), ("left_lcp", C.c_uint64),
                ("right_lcp", C.c_uint64), ("node_lcp", C.c_uint64)]

    def astuple(self):
        return (self.sp, self.ep, self.left_lcp, self.right_lcp, self.node_lcp)


class _MT64(C.Structure):
    _fields_ = [("mt", C.c_uint64 * 312), ("idx", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    u64, vp, u64p = C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)
    L.oracle_gcsa_create.restype = vp; L.oracle_gcsa_create.argtypes = [C.POINTER(_Flat)]
    L.oracle_gcsa_destroy.argtypes = [vp]
    L.oracle_lcp_create.restype = vp; L.oracle_lcp_create.argtypes = [u64, u64, u64, vp, vp]
    L.oracle_lcp_destroy.argtypes = [vp]
    L.oracle_bv_create.restype = vp; L.oracle_bv_create.argtypes = [vp, u64]
    L.oracle_bv_destroy.argtypes = [vp]
    L.oracle_bv_rank.restype = u64; L.oracle_bv_rank.argtypes = [vp, u64]
    L.oracle_bv_select.restype = u64; L.oracle_bv_select.argtypes = [vp, u64]
    L.oracle_bv_get.restype = C.c_int; L.oracle_bv_get.argtypes = [vp, u64]
    L.oracle_find.argtypes = [vp, vp, u64, u64p, u64p]
    L.oracle_find_stats.argtypes = [vp, vp, u64, u64p, u64p, u64p, u64p]
    L.oracle_char_range.argtypes = [vp, u64, u64p, u64p]
    L.oracle_lf_range.argtypes = [vp, u64, u64, u64, u64p, u64p]
    L.oracle_lf_node.restype = u64; L.oracle_lf_node.argtypes = [vp, u64]
    L.oracle_lf_fast.argtypes = [vp, u64, u64, vp]
    L.oracle_lf_all.argtypes = [vp, u64, u64, vp]
    L.oracle_count.restype = u64; L.oracle_count.argtypes = [vp, u64, u64]
    L.oracle_locate_node.restype = u64; L.oracle_locate_node.argtypes = [vp, u64, C.POINTER(u64p)]
    L.oracle_locate_range.restype = u64; L.oracle_locate_range.argtypes = [vp, u64, u64, C.POINTER(u64p)]
    L.oracle_locate_max.restype = u64; L.oracle_locate_max.argtypes = [vp, u64, u64, u64, C.POINTER(u64p)]
    L.oracle_free.argtypes = [vp]
    L.oracle_find_batch.restype = C.c_double
    L.oracle_find_batch.argtypes = [vp, vp, vp, u64, vp, vp, C.c_int]
    L.oracle_find_batch_stats.restype = C.c_double
    L.oracle_find_batch_stats.argtypes = [vp, vp, vp, u64, vp, vp, C.c_int, u64p, u64p]
    L.oracle_count_batch.restype = C.c_double
    L.oracle_count_batch.argtypes = [vp, vp, vp, u64, vp, C.c_int]
    L.oracle_locate_batch.restype = C.c_double
    L.oracle_locate_batch.argtypes = [vp, vp, vp, u64, vp, C.POINTER(u64p), C.c_int]
    L.oracle_max_threads.restype = C.c_int
    L.oracle_count_kmers.restype = u64; L.oracle_count_kmers.argtypes = [vp, u64, C.c_int, C.c_int]
    L.oracle_compare_kmers.restype = None
    L.oracle_compare_kmers.argtypes = [vp, vp, u64, C.c_int, C.c_int, vp, C.POINTER(vp), C.POINTER(vp)]
    L.oracle_lcp_parent.argtypes = [vp, u64, u64, C.POINTER(STNode)]
    L.oracle_lcp_depth.restype = u64; L.oracle_lcp_depth.argtypes = [vp, u64, u64]
    for name in ("psv", "psev", "nsv", "nsev"):
        getattr(L, "oracle_lcp_" + name).argtypes = [vp, u64, u64p, u64p]
    L.oracle_lcp_rmq.argtypes = [vp, u64, u64, u64p, u64p]
    L.oracle_parent_batch.restype = C.c_double
    L.oracle_parent_batch.argtypes = [vp, vp, vp, u64, vp, C.c_int]
    L.oracle_mem_batch.restype = C.c_double
    L.oracle_mem_batch.argtypes = [vp, vp, vp, vp, u64, vp, C.POINTER(u64p), C.c_int]
    L.oracle_mt64_seed.argtypes = [C.POINTER(_MT64), u64]
    L.oracle_mt64_next.restype = u64; L.oracle_mt64_next.argtypes = [C.POINTER(_MT64)]
    L.oracle_wang_hash_64.restype = u64; L.oracle_wang_hash_64.argtypes = [u64]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _bits(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.size == 0:
        a = np.zeros(1, dtype=np.uint64)
    return a


def pack_patterns(patterns):
    """list of bytes/str -> (chars uint8[], offsets uint64[n+1])"""
    bs = [p.encode() if isinstance(p, str) else bytes(p) for p in patterns]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    chars = np.frombuffer(b"".join(bs), dtype=np.uint8).copy()
    if chars.size == 0:
        chars = np.zeros(1, dtype=np.uint8)
    return chars, offsets


class BitVector:
    """Rank/select/access on a plain bit vector (for primitive-level tests)."""

    def __init__(self, words, n_bits):
        self._words = _bits(words)
        self._h = lib().oracle_bv_create(_ptr(self._words), n_bits)

    def rank(self, i): return lib().oracle_bv_rank(self._h, i)
    def select(self, k): return lib().oracle_bv_select(self._h, k)
    def get(self, i): return lib().oracle_bv_get(self._h, i)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_bv_destroy(self._h); self._h = None


class OracleGCSA:
    """CPU restatement of gcsa::GCSA's query interface (include/gcsa/gcsa.h:96-210)."""

    def __init__(self, flat):
        L = lib()
        f = _Flat()
        f.path_nodes, f.edge_count, f.order = int(flat.path_nodes), int(flat.edge_count), int(flat.order)
        f.sigma, f.fast_chars = int(flat.sigma), int(flat.fast_chars)
        for i in range(SIGMA + 1):
            f.C[i] = int(flat.C[i])
        C.memmove(f.char2comp, np.ascontiguousarray(flat.char2comp, dtype=np.uint8).ctypes.data, 256)
        self._keep = []
        for c in range(SIGMA):
            a = _bits(flat.bwt[c]); self._keep.append(a); f.bwt[c] = a.ctypes.data
        def keep(a, dtype=np.uint64):
            a = np.ascontiguousarray(a, dtype=dtype)
            if a.size == 0:
                a = np.zeros(1, dtype=dtype)
            self._keep.append(a)
            return a.ctypes.data
        f.edges = keep(flat.edges)
        f.sampled_paths = keep(flat.sampled_paths)
        f.sample_count = int(flat.sample_count)
        f.stored_samples = keep(flat.stored_samples)
        f.samples = keep(flat.samples)
        f.extra_filter = keep(flat.extra_filter)
        f.extra_values_len = int(flat.extra_values_len)
        f.extra_values = keep(flat.extra_values)
        f.redundant_len = int(flat.redundant_len)
        f.redundant = keep(flat.redundant)
        self._h = L.oracle_gcsa_create(C.byref(f))
        self.path_nodes = f.path_nodes
        self.char2comp = np.array(flat.char2comp, dtype=np.uint8)
        self._keep = None   # the oracle copied everything

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_gcsa_destroy(self._h); self._h = None

    def size(self): return self.path_nodes

    def find(self, pattern):
        b = pattern.encode() if isinstance(pattern, str) else bytes(pattern)
        buf = (C.c_uint8 * max(1, len(b))).from_buffer_copy(b if b else b"\0")
        sp, ep = C.c_uint64(), C.c_uint64()
        lib().oracle_find(self._h, buf, len(b), C.byref(sp), C.byref(ep))
        return (sp.value, ep.value)

    def find_stats(self, pattern):
        b = pattern.encode() if isinstance(pattern, str) else bytes(pattern)
        buf = (C.c_uint8 * max(1, len(b))).from_buffer_copy(b if b else b"\0")
        sp, ep, st, pr = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib().oracle_find_stats(self._h, buf, len(b), C.byref(sp), C.byref(ep), C.byref(st), C.byref(pr))
        return (sp.value, ep.value), st.value, pr.value

    def char_range(self, comp):
        sp, ep = C.c_uint64(), C.c_uint64()
        lib().oracle_char_range(self._h, comp, C.byref(sp), C.byref(ep))
        return (sp.value, ep.value)

    def LF(self, rng_or_node, comp=None):
        if comp is None:
            return lib().oracle_lf_node(self._h, int(rng_or_node))
        sp, ep = C.c_uint64(), C.c_uint64()
        lib().oracle_lf_range(self._h, int(rng_or_node[0]), int(rng_or_node[1]), int(comp), C.byref(sp), C.byref(ep))
        return (sp.value, ep.value)

    def _lf_multi(self, fn, rng):
        out = np.zeros(2 * SIGMA, dtype=np.uint64)
        fn(self._h, int(rng[0]), int(rng[1]), _ptr(out))
        return [(int(out[2 * c]), int(out[2 * c + 1])) for c in range(SIGMA)]

    def LF_fast(self, rng): return self._lf_multi(lib().oracle_lf_fast, rng)
    def LF_all(self, rng): return self._lf_multi(lib().oracle_lf_all, rng)

    def count(self, rng):
        return lib().oracle_count(self._h, int(rng[0]), int(rng[1]))

    def _take(self, n, p):
        res = [p[i] for i in range(n)]
        lib().oracle_free(p)
        return res

    def locate(self, rng_or_node, max_positions=None):
        p = C.POINTER(C.c_uint64)()
        if isinstance(rng_or_node, tuple):
            if max_positions is None:
                n = lib().oracle_locate_range(self._h, int(rng_or_node[0]), int(rng_or_node[1]), C.byref(p))
            else:
                n = lib().oracle_locate_max(self._h, int(rng_or_node[0]), int(rng_or_node[1]), int(max_positions), C.byref(p))
        else:
            n = lib().oracle_locate_node(self._h, int(rng_or_node), C.byref(p))
        return self._take(n, p)

    def count_kmers(self, k, include_Ns=False, threads=1):
        """countKMers(index, k), src/algorithms.cpp:387-421."""
        return lib().oracle_count_kmers(self._h, int(k), int(bool(include_Ns)), int(threads))

    def compare_kmers(self, other, k, include_Ns=False, threads=1):
        """compareKMers(left, right, k), src/algorithms.cpp:535-616 -> ((shared, left, right), left_kmers, right_kmers);
        the kmer arrays hold KMerComparisonState records as rows of 8 uint64 (left range, right range, k, kmer[3])."""
        res = np.zeros(3, dtype=np.uint64)
        pl, pr = C.c_void_p(), C.c_void_p()
        lib().oracle_compare_kmers(self._h, other._h, int(k), int(bool(include_Ns)), int(threads), res.ctypes.data, C.byref(pl), C.byref(pr))
        def take(p, n):
            if not p or n == 0:
                return np.zeros((0, 8), dtype=np.uint64)
            a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(n * 8,)).copy().reshape(n, 8)
            return a
        left, right = take(pl, int(res[1])), take(pr, int(res[2]))
        libc = C.CDLL(None); libc.free.argtypes = [C.c_void_p]
        libc.free(pl); libc.free(pr)
        return tuple(int(x) for x in res), left, right

    # ---- batch drivers (timed CPU baseline) ----
    def find_batch(self, chars, offsets, threads=1, stats=False):
        n = len(offsets) - 1
        chars = np.ascontiguousarray(chars, dtype=np.uint8); offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        sp = np.zeros(max(1, n), dtype=np.uint64); ep = np.zeros(max(1, n), dtype=np.uint64)
        if stats:
            ts, tp = C.c_uint64(), C.c_uint64()
            secs = lib().oracle_find_batch_stats(self._h, _ptr(chars), _ptr(offsets), n, _ptr(sp), _ptr(ep), threads,
                                                 C.byref(ts), C.byref(tp))
            return sp[:n], ep[:n], secs, ts.value, tp.value
        secs = lib().oracle_find_batch(self._h, _ptr(chars), _ptr(offsets), n, _ptr(sp), _ptr(ep), threads)
        return sp[:n], ep[:n], secs

    def count_batch(self, sp, ep, threads=1):
        sp = np.ascontiguousarray(sp, dtype=np.uint64); ep = np.ascontiguousarray(ep, dtype=np.uint64)
        n = len(sp); out = np.zeros(max(1, n), dtype=np.uint64)
        secs = lib().oracle_count_batch(self._h, _ptr(sp), _ptr(ep), n, _ptr(out), threads)
        return out[:n], secs

    def locate_batch(self, sp, ep, threads=1):
        sp = np.ascontiguousarray(sp, dtype=np.uint64); ep = np.ascontiguousarray(ep, dtype=np.uint64)
        n = len(sp); offs = np.zeros(n + 1, dtype=np.uint64)
        p = C.POINTER(C.c_uint64)()
        secs = lib().oracle_locate_batch(self._h, _ptr(sp), _ptr(ep), n, _ptr(offs), C.byref(p), threads)
        total = int(offs[n])
        vals = np.ctypeslib.as_array(p, shape=(max(1, total),))[:total].copy()
        lib().oracle_free(p)
        return offs, vals, secs


class OracleLCP:
    """CPU restatement of gcsa::LCPArray's query interface (include/gcsa/lcp.h:137-178)."""

    def __init__(self, flat_lcp):
        offs = np.ascontiguousarray(flat_lcp.offsets, dtype=np.uint64)
        data = np.ascontiguousarray(flat_lcp.data, dtype=np.uint8)
        self.size_, self.values_ = int(flat_lcp.size), int(offs[-1])
        self._h = lib().oracle_lcp_create(int(flat_lcp.size), int(flat_lcp.branching), int(flat_lcp.levels),
                                          _ptr(offs), _ptr(data))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_lcp_destroy(self._h); self._h = None

    def not_found(self): return (self.values_, self.values_)

    def parent(self, rng):
        out = STNode()
        lib().oracle_lcp_parent(self._h, int(rng[0]), int(rng[1]), C.byref(out))
        return out.astuple()

    def depth(self, rng):
        return lib().oracle_lcp_depth(self._h, int(rng[0]), int(rng[1]))

    def _pv(self, name, *args):
        a, b = C.c_uint64(), C.c_uint64()
        getattr(lib(), "oracle_lcp_" + name)(self._h, *[int(x) for x in args], C.byref(a), C.byref(b))
        return (a.value, b.value)

    def psv(self, pos): return self._pv("psv", pos)
    def psev(self, pos): return self._pv("psev", pos)
    def nsv(self, pos): return self._pv("nsv", pos)
    def nsev(self, pos): return self._pv("nsev", pos)
    def rmq(self, sp, ep): return self._pv("rmq", sp, ep)

    def parent_batch(self, sp, ep, threads=1):
        sp = np.ascontiguousarray(sp, dtype=np.uint64); ep = np.ascontiguousarray(ep, dtype=np.uint64)
        n = len(sp); out = np.zeros((max(1, n), 5), dtype=np.uint64)
        secs = lib().oracle_parent_batch(self._h, _ptr(sp), _ptr(ep), n, _ptr(out), threads)
        return out[:n], secs


def mem_batch(index, lcp, chars, offsets, threads=1):
    """MEM-style scan of every pattern: (out_offsets, matches[k, 4] = (start, length, sp, ep), seconds)."""
    chars = np.ascontiguousarray(chars, dtype=np.uint8); offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    offs = np.zeros(n + 1, dtype=np.uint64)
    p = C.POINTER(C.c_uint64)()
    secs = lib().oracle_mem_batch(index._h, lcp._h, _ptr(chars), _ptr(offsets), n, _ptr(offs), C.byref(p), threads)
    total = int(offs[n])
    vals = np.ctypeslib.as_array(p, shape=(max(1, 4 * total),))[:4 * total].copy().reshape(-1, 4)
    lib().oracle_free(p)
    return offs, vals, secs


def mt19937_64(seed, n):
    r = _MT64(); lib().oracle_mt64_seed(C.byref(r), seed)
    return [lib().oracle_mt64_next(C.byref(r)) for _ in range(n)]


def wang_hash_64(key):
    return lib().oracle_wang_hash_64(key)
