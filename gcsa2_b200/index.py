"""Host-side mirror of the reference's query interface over the C ABI.

GCSA     <-> gcsa::GCSA      (reference include/gcsa/gcsa.h:40-275): find, charRange, LF,
                              LF_fast, LF_all, count, locate and the size accessors, with the
                              reference's names, argument meaning and error behaviour (queries
                              never raise for odd ranges; count/locate treat ranges past the
                              index as empty, src/gcsa.cpp:805, 817, 831), plus *_batch forms.
LCPArray <-> gcsa::LCPArray  (include/gcsa/lcp.h:90-194): parent, depth, psv/psev/nsv/nsev, rmq.

Everything here is plumbing: the work happens in libgcsa2_b200.so (CUDA, sm_100a).  There is no
CPU implementation behind these classes; constructing one without a CUDA device raises.
"""
import atexit
import ctypes as C
import weakref

import numpy as np

from . import capi
from .flat import FlatGCSA, FlatLCP, SIGMA

UNKNOWN = (1 << 64) - 1


def pack_patterns(patterns):
    """list of str/bytes -> (chars uint8[], offsets uint64[n + 1])"""
    bs = [p.encode() if isinstance(p, str) else bytes(p) for p in patterns]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    chars = np.frombuffer(b"".join(bs), dtype=np.uint8).copy()
    if chars.size == 0:
        chars = np.zeros(1, dtype=np.uint8)
    return chars, offsets


def range_empty(rng):
    """Range::empty, include/gcsa/utils.h:93-101."""
    return ((int(rng[0]) + 1) & UNKNOWN) > ((int(rng[1]) + 1) & UNKNOWN)


def range_length(rng):
    return (int(rng[1]) + 1 - int(rng[0])) & UNKNOWN


# Handles still open when the interpreter exits are destroyed before the CUDA context goes away (objects alive at exit are
# not guaranteed their __del__): device memory is released by the library that allocated it, and leak checkers stay quiet.
_open_handles = weakref.WeakSet()


@atexit.register
def _close_open_handles():
    for handle in list(_open_handles):
        try:
            handle.close()
        except Exception:
            pass


class GCSA:
    def __init__(self, flat, device=0, kmer_table_k=0, two_step=None, walk_table=None, jump_table=None, fused_table=None):
        self._h = None
        L = self._L = capi.lib()            # the handle is destroyed by the library that made it
        keep = []
        # a FlatGCSA (numpy arrays) or a builder.BuiltIndex (the library's own buffers, no copies)
        f = flat.struct if hasattr(flat, "struct") else capi.flat_struct(flat, keep)
        opt = capi.Options(); opt.kmer_table_k = int(kmer_table_k); opt.two_step = (-1 if two_step is None else int(bool(two_step)))
        opt.walk_table = (-1 if walk_table is None else int(walk_table))
        opt.jump_table = (0 if jump_table is None else (2 if jump_table == "wide" else (1 if jump_table else -1)))
        opt.fused_table = (0 if fused_table is None else (1 if fused_table else -1))
        h = C.c_void_p()
        capi.check(L.gcsa_b200_index_create(C.byref(f), int(device), C.byref(opt), C.byref(h)))
        self._h = h
        _open_handles.add(self)
        self.device = int(device)
        self.char2comp = np.array(flat.char2comp, dtype=np.uint8)
        self.C = np.array(flat.C, dtype=np.uint64)
        info = capi.Info()
        capi.check(L.gcsa_b200_index_info(self._h, C.byref(info)))
        self._info = info

    @classmethod
    def load(cls, gcsa_file, **options):
        """sdsl::load_from_file(index, name) -> GCSA::load() of the reference (src/gcsa.cpp:182-216)."""
        from .flat import FlatGCSA
        return cls(FlatGCSA.from_gcsa_file(gcsa_file), **options)

    def close(self):
        if self._h is not None:
            self._L.gcsa_b200_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- accessors (gcsa.h:137-148) ----
    def size(self): return int(self._info.path_nodes)
    def empty(self): return self.size() == 0
    def edgeCount(self): return int(self._info.edge_count)
    def order(self): return int(self._info.order)
    def sampleCount(self): return int(self._info.sample_count)
    def deviceBytes(self): return int(self._info.device_bytes)
    def kmerTableK(self): return int(self._info.kmer_table_k)
    def twoStep(self): return bool(self._info.two_step)
    def jumpK(self): return int(self._info.jump_k)
    def fusedTable(self): return bool(self._info.fused_table)
    def smCount(self): return int(self._info.sm_count)
    @property
    def handle(self): return self._h

    # ---- find (gcsa.h:96-122) ----
    def find(self, pattern):
        sp, ep = self.find_batch([pattern])
        return (int(sp[0]), int(ep[0]))

    def find_batch(self, patterns, offsets=None, stats=False):
        """patterns: list of str/bytes, or a uint8 array together with `offsets` (n + 1)."""
        if offsets is None:
            chars, offsets = pack_patterns(patterns)
        else:
            chars = np.ascontiguousarray(patterns, dtype=np.uint8)
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        sp = np.zeros(max(n, 1), dtype=np.uint64); ep = np.zeros(max(n, 1), dtype=np.uint64)
        if stats:
            st = capi.FindStats()
            capi.check(capi.lib().gcsa_b200_find_stats_host(self._h, chars.ctypes.data, offsets.ctypes.data, n,
                                                            sp.ctypes.data, ep.ctypes.data, C.byref(st)))
            return sp[:n], ep[:n], {k: int(getattr(st, k)) for k, _ in capi.FindStats._fields_}
        capi.check(capi.lib().gcsa_b200_find_host(self._h, chars.ctypes.data, offsets.ctypes.data, n,
                                                  sp.ctypes.data, ep.ctypes.data))
        return sp[:n], ep[:n]

    def find_host_raw(self, chars_ptr, offsets_ptr, n, sp_ptr, ep_ptr):
        """Host pointers (e.g. pinned torch tensors' data_ptr()); the C ABI call a user makes."""
        capi.check(capi.lib().gcsa_b200_find_host(self._h, chars_ptr, offsets_ptr, int(n), sp_ptr, ep_ptr))

    def find_fixed_host_raw(self, chars_ptr, pattern_length, n, sp_ptr, ep_ptr):
        """k-mer form: n patterns of one length back to back, host pointers."""
        capi.check(capi.lib().gcsa_b200_find_fixed_host(self._h, chars_ptr, int(pattern_length), int(n), sp_ptr, ep_ptr))

    def find_fixed_batch(self, chars, pattern_length, stats=False):
        chars = np.ascontiguousarray(chars, dtype=np.uint8)
        n = chars.size // int(pattern_length) if pattern_length else 0
        sp = np.zeros(max(n, 1), dtype=np.uint64); ep = np.zeros(max(n, 1), dtype=np.uint64)
        if stats:
            st = capi.FindStats()
            capi.check(capi.lib().gcsa_b200_find_fixed_stats_host(self._h, chars.ctypes.data, int(pattern_length), n,
                                                                  sp.ctypes.data, ep.ctypes.data, C.byref(st)))
            return sp[:n], ep[:n], {k: int(getattr(st, k)) for k, _ in capi.FindStats._fields_}
        self.find_fixed_host_raw(chars.ctypes.data, pattern_length, n, sp.ctypes.data, ep.ctypes.data)
        return sp[:n], ep[:n]

    def find_fixed_device(self, d_chars, pattern_length, n, d_sp, d_ep, stream=0):
        capi.check(capi.lib().gcsa_b200_find_fixed_batch(self._h, capi.ptr(d_chars), int(pattern_length), int(n),
                                                         capi.ptr(d_sp), capi.ptr(d_ep), stream or None))

    def find_device(self, d_chars, d_offsets, n, d_sp, d_ep, stream=0):
        """Device pointers / tensors, stream-ordered, no synchronisation."""
        capi.check(capi.lib().gcsa_b200_find_batch(self._h, capi.ptr(d_chars), capi.ptr(d_offsets), int(n),
                                                   capi.ptr(d_sp), capi.ptr(d_ep), stream or None))

    # ---- device-pointer forms (stream-ordered; tensors or raw addresses) ----
    def lf_device(self, d_sp, d_ep, d_comp, n, d_sp_out, d_ep_out, stream=0):
        capi.check(capi.lib().gcsa_b200_lf_batch(self._h, capi.ptr(d_sp), capi.ptr(d_ep), capi.ptr(d_comp), int(n),
                                                 capi.ptr(d_sp_out), capi.ptr(d_ep_out), stream or None))

    def count_device(self, d_sp, d_ep, n, d_out, stream=0):
        capi.check(capi.lib().gcsa_b200_count_batch(self._h, capi.ptr(d_sp), capi.ptr(d_ep), int(n), capi.ptr(d_out), stream or None))

    def locate_device(self, d_sp, d_ep, n, d_offsets, d_values, capacity, stream=0):
        """Returns the number of values written (raises GCSAError(ERR_CAPACITY) if capacity is too small)."""
        needed = C.c_uint64()
        capi.check(capi.lib().gcsa_b200_locate_batch(self._h, capi.ptr(d_sp), capi.ptr(d_ep), int(n), capi.ptr(d_offsets),
                                                     capi.ptr(d_values), int(capacity), C.byref(needed), stream or None))
        return int(needed.value)

    # ---- low-level interface (gcsa.h:150-183, gcsa.cpp:742-798) ----
    def charRange(self, comp):
        sp, ep = C.c_uint64(), C.c_uint64()
        capi.check(capi.lib().gcsa_b200_char_range(self._h, int(comp), C.byref(sp), C.byref(ep)))
        return (sp.value, ep.value)

    def LF(self, range_or_node, comp=None):
        if comp is None:
            return int(self.lf_node_batch([range_or_node])[0])
        sp, ep = self.lf_batch([range_or_node[0]], [range_or_node[1]], [comp])
        return (int(sp[0]), int(ep[0]))

    def lf_batch(self, sp, ep, comp):
        sp, ep = capi.as_u64(sp), capi.as_u64(ep)
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        n = int(comp.size)
        osp = np.zeros(max(n, 1), dtype=np.uint64); oep = np.zeros(max(n, 1), dtype=np.uint64)
        capi.check(capi.lib().gcsa_b200_lf_host(self._h, sp.ctypes.data, ep.ctypes.data, comp.ctypes.data, n,
                                                osp.ctypes.data, oep.ctypes.data))
        return osp[:n], oep[:n]

    def lf_node_batch(self, nodes):
        n = len(nodes)
        nodes = capi.as_u64(nodes)
        out = np.zeros(max(n, 1), dtype=np.uint64)
        capi.check(capi.lib().gcsa_b200_lf_node_host(self._h, nodes.ctypes.data, n, out.ctypes.data))
        return out[:n]

    def lf_multi_batch(self, sp, ep, all_chars):
        n = len(sp)
        sp, ep = capi.as_u64(sp), capi.as_u64(ep)
        out = np.zeros((max(n, 1), SIGMA, 2), dtype=np.uint64)
        capi.check(capi.lib().gcsa_b200_lf_multi_host(self._h, sp.ctypes.data, ep.ctypes.data, n, int(all_chars), out.ctypes.data))
        return out[:n]

    def LF_fast(self, rng):
        """results[comp] for 1 <= comp <= fast_chars (gcsa.cpp:742-767); other slots are (1, 0)."""
        out = self.lf_multi_batch([rng[0]], [rng[1]], 0)[0]
        return [(int(a), int(b)) for a, b in out]

    def LF_all(self, rng):
        out = self.lf_multi_batch([rng[0]], [rng[1]], 1)[0]
        return [(int(a), int(b)) for a, b in out]

    # ---- countKMers (algorithms.cpp:387-421) ----
    def count_kmers(self, k, include_Ns=False, return_ranges=False):
        res = C.c_uint64(); p = C.c_void_p()
        capi.check(capi.lib().gcsa_b200_count_kmers(self._h, int(k), int(bool(include_Ns)), C.byref(res),
                                                    C.byref(p) if return_ranges else None))
        if not return_ranges:
            return int(res.value)
        n = int(res.value)
        if not p.value:
            return n, np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.uint64)
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(2 * n,)).copy()
        capi.lib().gcsa_b200_free(p)
        return n, arr[:n], arr[n:]

    def verify(self, kmers, lcp=None, mapping=None):
        """verifyIndex(index, lcp, kmers, kmer_length, mapping), src/algorithms.cpp:101-295, batched on the device.
        kmers: builder.KMers (the construction input).  -> dict of the report; ok iff report["failures"] == 0."""
        rep = capi.VerifyReport()
        key, frm = capi.as_u64(kmers.key), capi.as_u64(kmers.from_)
        ids = capi.as_u64(mapping.ids) if mapping is not None else None
        capi.check(capi.lib().gcsa_b200_verify_index_mapped(self._h, lcp._h if lcp is not None else None, key.ctypes.data, frm.ctypes.data,
                                                            int(kmers.key.size), int(kmers.k),
                                                            int(mapping.first) if mapping is not None else 0,
                                                            ids.ctypes.data if mapping is not None else None,
                                                            len(mapping.ids) if mapping is not None else 0, C.byref(rep)))
        return {name: getattr(rep, name) for name, _ in capi.VerifyReport._fields_}

    def compare_kmers_to_files(self, other, k, output, include_Ns=False):
        """compareKMers with parameters.output = `output`: counts, and the unique kmers in <output>.left / <output>.right."""
        res = np.zeros(3, dtype=np.uint64)
        capi.check(capi.lib().gcsa_b200_compare_kmers_to_files(self._h, other._h, int(k), int(bool(include_Ns)), str(output).encode(), res.ctypes.data))
        return tuple(int(x) for x in res)

    def compare_kmers(self, other, k, include_Ns=False, return_kmers=False):
        """compareKMers(left, right, k), src/algorithms.cpp:535-616 -> (shared, left only, right only)
        [, left_kmers, right_kmers as rows of 8 uint64: KMerComparisonState records]."""
        res = np.zeros(3, dtype=np.uint64)
        pl, pr = C.c_void_p(), C.c_void_p()
        capi.check(capi.lib().gcsa_b200_compare_kmers(self._h, other._h, int(k), int(bool(include_Ns)), res.ctypes.data,
                                                      C.byref(pl) if return_kmers else None, C.byref(pr) if return_kmers else None))
        counts = tuple(int(x) for x in res)
        if not return_kmers:
            return counts
        def take(p, n):
            if not p.value or n == 0:
                return np.zeros((0, 8), dtype=np.uint64)
            a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(n * 8,)).copy().reshape(n, 8)
            capi.lib().gcsa_b200_free(p)
            return a
        return counts, take(pl, counts[1]), take(pr, counts[2])

    # ---- count / locate (gcsa.cpp:802-878) ----
    def count(self, rng):
        return int(self.count_batch([rng[0]], [rng[1]])[0])

    def count_batch(self, sp, ep):
        n = len(sp)
        sp, ep = capi.as_u64(sp), capi.as_u64(ep)
        out = np.zeros(max(n, 1), dtype=np.uint64)
        capi.check(capi.lib().gcsa_b200_count_host(self._h, sp.ctypes.data, ep.ctypes.data, n, out.ctypes.data))
        return out[:n]

    def locate(self, range_or_node, max_positions=None):
        """locate(path_node) / locate(range) / locate(range, max_positions): sorted distinct values."""
        if isinstance(range_or_node, tuple):
            sp, ep = [range_or_node[0]], [range_or_node[1]]
        else:
            node = int(range_or_node)
            if node >= self.size():                       # gcsa.cpp:817
                return []
            sp, ep = [node], [node]
        offs, vals = self.locate_batch(sp, ep, max_positions=max_positions)
        return [int(x) for x in vals[int(offs[0]):int(offs[1])]]

    def locate_into_host_raw(self, sp_ptr, ep_ptr, n, offsets_ptr, values_ptr, capacity):
        """gcsa_b200_locate_into_host on raw host addresses (pinned buffers); returns the number of values,
        raises GCSAError(ERR_CAPACITY) if capacity was too small."""
        needed = C.c_uint64()
        capi.check(capi.lib().gcsa_b200_locate_into_host(self._h, sp_ptr, ep_ptr, int(n), offsets_ptr, values_ptr, int(capacity), C.byref(needed)))
        return int(needed.value)

    def locate_batch(self, sp, ep, max_positions=None, sort=True):
        """CSR result: values[offsets[i]:offsets[i+1]] are the sorted distinct positions of range i
        (sort=False: every emitted value in the reference's order, gcsa.cpp:840)."""
        n = len(sp)
        sp, ep = capi.as_u64(sp), capi.as_u64(ep)
        offs = np.zeros(n + 1, dtype=np.uint64)
        p = C.c_void_p()
        if max_positions is None and not sort:
            capi.check(capi.lib().gcsa_b200_locate_raw_host(self._h, sp.ctypes.data, ep.ctypes.data, n, offs.ctypes.data, C.byref(p)))
        elif max_positions is None:
            capi.check(capi.lib().gcsa_b200_locate_host(self._h, sp.ctypes.data, ep.ctypes.data, n, offs.ctypes.data, C.byref(p)))
        else:
            capi.check(capi.lib().gcsa_b200_locate_max_host(self._h, sp.ctypes.data, ep.ctypes.data, n, int(max_positions),
                                                            offs.ctypes.data, C.byref(p)))
        total = int(offs[n])
        if not p.value:
            return offs, np.zeros(0, dtype=np.uint64)
        vals = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(max(total, 1),))[:total].copy()
        capi.lib().gcsa_b200_free(p)
        return offs, vals


class MultiGCSA:
    """Replicas of one index on several devices behind the single-process multi-GPU entry points
    (gcsa_b200_find_fixed_host_multi / _find_host_multi / _locate_into_host_multi): the batch is cut into contiguous
    blocks, one per replica, each driven by its own host thread.  `devices` may name a device more than once."""

    def __init__(self, flat, devices, **options):
        self.replicas = [GCSA(flat, device=d, **options) for d in devices]
        self._handles = (C.c_void_p * len(self.replicas))(*[r.handle for r in self.replicas])

    def close(self):
        for r in self.replicas:
            r.close()

    def size(self): return self.replicas[0].size()

    def find_fixed_host_raw(self, chars_ptr, pattern_length, n, sp_ptr, ep_ptr):
        capi.check(capi.lib().gcsa_b200_find_fixed_host_multi(self._handles, len(self.replicas), chars_ptr, int(pattern_length), int(n), sp_ptr, ep_ptr))

    def find_fixed_batch(self, chars, pattern_length):
        chars = np.ascontiguousarray(chars, dtype=np.uint8)
        n = chars.size // int(pattern_length) if pattern_length else 0
        sp = np.zeros(max(n, 1), dtype=np.uint64); ep = np.zeros(max(n, 1), dtype=np.uint64)
        self.find_fixed_host_raw(chars.ctypes.data, pattern_length, n, sp.ctypes.data, ep.ctypes.data)
        return sp[:n], ep[:n]

    def find_batch(self, patterns, offsets=None):
        if offsets is None:
            chars, offsets = pack_patterns(patterns)
        else:
            chars = np.ascontiguousarray(patterns, dtype=np.uint8); offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        sp = np.zeros(max(n, 1), dtype=np.uint64); ep = np.zeros(max(n, 1), dtype=np.uint64)
        capi.check(capi.lib().gcsa_b200_find_host_multi(self._handles, len(self.replicas), chars.ctypes.data, offsets.ctypes.data, n,
                                                        sp.ctypes.data, ep.ctypes.data))
        return sp[:n], ep[:n]

    def locate_into_host_raw(self, sp_ptr, ep_ptr, n, offsets_ptr, values_ptr, capacity):
        needed = C.c_uint64()
        capi.check(capi.lib().gcsa_b200_locate_into_host_multi(self._handles, len(self.replicas), sp_ptr, ep_ptr, int(n), offsets_ptr,
                                                               values_ptr, int(capacity), C.byref(needed)))
        return int(needed.value)

    def locate_batch(self, sp, ep):
        """CSR of sorted distinct positions, like GCSA.locate_batch."""
        n = len(sp)
        sp, ep = capi.as_u64(sp), capi.as_u64(ep)
        offs = np.zeros(n + 1, dtype=np.uint64)
        needed = C.c_uint64()
        rc = capi.lib().gcsa_b200_locate_into_host_multi(self._handles, len(self.replicas), sp.ctypes.data, ep.ctypes.data, n,
                                                         offs.ctypes.data, None, 0, C.byref(needed))
        capi.check(rc, allow=(capi.ERR_CAPACITY,))
        vals = np.zeros(max(1, int(needed.value)), dtype=np.uint64)
        got = self.locate_into_host_raw(sp.ctypes.data, ep.ctypes.data, n, offs.ctypes.data, vals.ctypes.data, vals.size)
        return offs, vals[:got]


class LCPArray:
    @classmethod
    def load(cls, lcp_file, device=0):
        """LCPArray::load() of the reference (src/lcp.cpp:128-143)."""
        from .flat import FlatLCP
        return cls(FlatLCP.from_lcp_file(lcp_file), device=device)

    def __init__(self, flat_lcp, device=0):
        self._h = None
        self._offsets = capi.as_u64(flat_lcp.offsets)
        data = np.ascontiguousarray(flat_lcp.data, dtype=np.uint8)
        if data.size == 0:
            data = np.zeros(1, dtype=np.uint8)
        f = capi.FlatLcp()
        f.size, f.branching, f.levels = int(flat_lcp.size), int(flat_lcp.branching), int(flat_lcp.levels)
        f.offsets, f.data = self._offsets.ctypes.data, data.ctypes.data
        h = C.c_void_p()
        self._L = capi.lib()
        capi.check(self._L.gcsa_b200_lcp_create(C.byref(f), int(device), C.byref(h)))
        self._h = h
        _open_handles.add(self)
        self._size, self._values = int(flat_lcp.size), int(self._offsets[int(flat_lcp.levels)])
        self._branching, self._levels = int(flat_lcp.branching), int(flat_lcp.levels)

    def close(self):
        if self._h is not None:
            self._L.gcsa_b200_lcp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self): return self._size
    def values(self): return self._values
    def levels(self): return self._levels
    def branching(self): return self._branching
    def root(self): return (0, self._size - 1, 0, 0, 0)            # lcp.h:137
    def notFound(self): return (self._values, self._values)        # lcp.h:178
    @property
    def handle(self): return self._h

    def parent(self, rng):
        return tuple(int(x) for x in self.parent_batch([rng[0]], [rng[1]])[0])

    def parent_batch(self, sp, ep):
        n = len(sp)
        sp, ep = capi.as_u64(sp), capi.as_u64(ep)
        out = np.zeros((max(n, 1), 5), dtype=np.uint64)
        capi.check(capi.lib().gcsa_b200_parent_host(self._h, sp.ctypes.data, ep.ctypes.data, n, out.ctypes.data))
        return out[:n]

    def parent_device(self, d_sp, d_ep, n, d_out, stream=0):
        capi.check(capi.lib().gcsa_b200_parent_batch(self._h, capi.ptr(d_sp), capi.ptr(d_ep), int(n), capi.ptr(d_out), stream or None))

    def depth_device(self, d_sp, d_ep, n, d_out, stream=0):
        capi.check(capi.lib().gcsa_b200_depth_batch(self._h, capi.ptr(d_sp), capi.ptr(d_ep), int(n), capi.ptr(d_out), stream or None))

    def depth(self, rng):
        return int(self.depth_batch([rng[0]], [rng[1]])[0])

    def depth_batch(self, sp, ep):
        n = len(sp)
        sp, ep = capi.as_u64(sp), capi.as_u64(ep)
        out = np.zeros(max(n, 1), dtype=np.uint64)
        capi.check(capi.lib().gcsa_b200_depth_host(self._h, sp.ctypes.data, ep.ctypes.data, n, out.ctypes.data))
        return out[:n]

    def _sv(self, which, pos):
        n = len(pos)
        pos = capi.as_u64(pos)
        a = np.zeros(max(n, 1), dtype=np.uint64); b = np.zeros(max(n, 1), dtype=np.uint64)
        capi.check(capi.lib().gcsa_b200_lcp_sv_host(self._h, which, pos.ctypes.data, n, a.ctypes.data, b.ctypes.data))
        return a[:n], b[:n]

    def psv(self, pos): a, b = self._sv(0, [pos]); return (int(a[0]), int(b[0]))
    def psev(self, pos): a, b = self._sv(1, [pos]); return (int(a[0]), int(b[0]))
    def nsv(self, pos): a, b = self._sv(2, [pos]); return (int(a[0]), int(b[0]))
    def nsev(self, pos): a, b = self._sv(3, [pos]); return (int(a[0]), int(b[0]))
    def sv_batch(self, which, pos): return self._sv({"psv": 0, "psev": 1, "nsv": 2, "nsev": 3}[which], pos)

    def rmq(self, sp, ep):
        a, b = self.rmq_batch([sp], [ep])
        return (int(a[0]), int(b[0]))

    def rmq_batch(self, sp, ep):
        n = len(sp)
        sp, ep = capi.as_u64(sp), capi.as_u64(ep)
        a = np.zeros(max(n, 1), dtype=np.uint64); b = np.zeros(max(n, 1), dtype=np.uint64)
        capi.check(capi.lib().gcsa_b200_lcp_rmq_host(self._h, sp.ctypes.data, ep.ctypes.data, n, a.ctypes.data, b.ctypes.data))
        return a[:n], b[:n]


def mem_batch(index, lcp, patterns, offsets=None):
    """MEM-style scan (config 5) of every pattern through the engine:
    -> (out_offsets uint64[n + 1], matches uint64[k, 4] = (start, length, sp, ep))."""
    if offsets is None:
        chars, offsets = pack_patterns(patterns)
    else:
        chars = np.ascontiguousarray(patterns, dtype=np.uint8); offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    offs = np.zeros(n + 1, dtype=np.uint64)
    p = C.c_void_p()
    capi.check(capi.lib().gcsa_b200_mem_host(index.handle, lcp.handle, chars.ctypes.data, offsets.ctypes.data, n,
                                             offs.ctypes.data, C.byref(p)))
    total = int(offs[n])
    if not p.value:
        return offs, np.zeros((0, 4), dtype=np.uint64)
    vals = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(max(4 * total, 1),))[:4 * total].copy().reshape(-1, 4)
    capi.lib().gcsa_b200_free(p)
    return offs, vals


def mem_device(index, lcp, d_chars, d_offsets, n, d_out_offsets, d_matches, capacity, stream=0):
    needed = C.c_uint64()
    capi.check(capi.lib().gcsa_b200_mem_batch(index.handle, lcp.handle, capi.ptr(d_chars), capi.ptr(d_offsets), int(n),
                                              capi.ptr(d_out_offsets), capi.ptr(d_matches), int(capacity), C.byref(needed), stream or None))
    return int(needed.value)
