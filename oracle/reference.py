"""ctypes binding of oracle/_ref/libgcsa2_ref.so: the REFERENCE's own, unmodified sources
(jltsiren/gcsa2 src/*.cpp) compiled against the SDSL shim of oracle/sdsl_shim/.
TEST INFRASTRUCTURE ONLY -- used to pin the C restatement (oracle/gcsa_oracle.c), the index builder
and the CUDA engine to what the reference itself computes.

The library is built in this container by `make -C oracle ref` (needs /root/reference); it travels
to the GPU box as a prebuilt file.  available() says whether it can be loaded.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libgcsa2_ref.so")
REFERENCE_TREE = "/root/reference"
SIGMA = 7
_lib = None


def build(force=False):
    """Compile the reference from where it lies; no-op (returns None) where /root/reference is absent."""
    if not os.path.isdir(os.path.join(REFERENCE_TREE, "src")):
        return _LIB_PATH if os.path.exists(_LIB_PATH) else None
    deps = [os.path.join(_HERE, "ref_driver.cpp"), os.path.join(_HERE, "sdsl_shim", "sdsl", "wavelet_trees.hpp"), os.path.join(_HERE, "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(d) > os.path.getmtime(_LIB_PATH) for d in deps)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "ref"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def available():
    try:
        return lib() is not None
    except OSError:
        return False


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = build()
    if path is None or not os.path.exists(path):
        return None
    L = C.CDLL(path)
    vp, u64 = C.c_void_p, C.c_uint64
    L.ref_build.restype = vp; L.ref_build.argtypes = [C.c_char_p, C.c_int, u64, u64, C.c_char_p]
    L.ref_build_from.restype = vp; L.ref_build_from.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, u64, u64, C.c_char_p]
    L.ref_destroy.argtypes = [vp]
    L.ref_verify.restype = C.c_int; L.ref_verify.argtypes = [vp]
    L.ref_sizes.argtypes = [vp, vp]
    L.ref_export.argtypes = [vp] + [vp] * 11
    L.ref_from_flat.restype = vp; L.ref_from_flat.argtypes = [u64, u64, u64, vp, vp, vp, vp, u64, vp, vp]
    L.ref_store.restype = C.c_int; L.ref_store.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.ref_load.restype = vp; L.ref_load.argtypes = [C.c_char_p, C.c_char_p]
    L.ref_max_threads.restype = C.c_int
    L.ref_find_batch.restype = C.c_double; L.ref_find_batch.argtypes = [vp, vp, vp, u64, vp, vp, C.c_int]
    L.ref_lf_batch.argtypes = [vp, vp, vp, vp, u64, vp, vp]
    L.ref_lf_node_batch.argtypes = [vp, vp, u64, vp]
    L.ref_count_batch.argtypes = [vp, vp, vp, u64, vp]
    L.ref_locate_batch.restype = C.c_double; L.ref_locate_batch.argtypes = [vp, vp, vp, u64, u64, vp, C.POINTER(vp), C.c_int]
    L.ref_free.argtypes = [vp]
    L.ref_parent_batch.argtypes = [vp, vp, vp, u64, vp]
    L.ref_depth_batch.argtypes = [vp, vp, vp, u64, vp]
    L.ref_lcp_query.argtypes = [vp, C.c_int, vp, vp, u64, vp, vp]
    L.ref_count_kmers.restype = u64; L.ref_count_kmers.argtypes = [vp, u64, C.c_int]
    L.ref_compare_kmers.restype = None; L.ref_compare_kmers.argtypes = [vp, vp, u64, C.c_int, vp]
    _lib = L
    return L


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a if a.size else np.zeros(1, dtype=np.uint64)


class ReferenceIndex:
    """A gcsa::GCSA (+ gcsa::LCPArray) object living inside the reference's own code."""

    def __init__(self, handle, has_lcp, tmp_dir=None):
        self._h, self.has_lcp, self._tmp = handle, has_lcp, tmp_dir

    @staticmethod
    def build(kmers, doubling_steps, sample_period=64, lcp_branching=64, text=False, mapping=None):
        """GCSA::GCSA(InputGraph&, ConstructionParameters) on the kmers, written as a binary .graph file (text=True: as a
        text .gcsa2 file, read back by the reference's own readText); mapping: a builder.NodeMapping, written as the
        mapping file build_gcsa takes."""
        tmp = tempfile.mkdtemp(prefix="gcsa_ref_")
        path = os.path.join(tmp, "input.gcsa2" if text else "input.graph")
        if text:
            kmers.write_text(path)
        else:
            kmers.write_binary(path)
        mapping_path = None
        if mapping is not None:
            mapping_path = os.path.join(tmp, "input.mapping")
            mapping.write(mapping_path)
        h = lib().ref_build_from(path.encode(), 0 if text else 1, mapping_path.encode() if mapping_path else None,
                                 int(doubling_steps), int(sample_period), int(lcp_branching), tmp.encode())
        return ReferenceIndex(h, True, tmp)      # the InputGraph re-reads its file in verify()

    @staticmethod
    def from_flat(flat):
        """Reference GCSA object over arrays built elsewhere (for timing the reference's find())."""
        keep = [_u64(flat.bwt[c]) for c in range(SIGMA)]
        ptrs = (C.c_void_p * SIGMA)(*[a.ctypes.data for a in keep])
        Cc, edges, sampled = _u64(flat.C), _u64(flat.edges), _u64(flat.sampled_paths)
        stored, samples = _u64(flat.stored_samples), _u64(flat.samples)
        h = lib().ref_from_flat(int(flat.path_nodes), int(flat.edge_count), int(flat.order), Cc.ctypes.data, ptrs,
                                edges.ctypes.data, sampled.ctypes.data, int(flat.sample_count), stored.ctypes.data, samples.ctypes.data)
        return ReferenceIndex(h, False)

    def store(self, gcsa_path, lcp_path=None):
        """sdsl::store_to_file -> GCSA::serialize / LCPArray::serialize of the reference (through the shim)."""
        ok = lib().ref_store(self._h, str(gcsa_path).encode(), str(lcp_path).encode() if lcp_path else None)
        if not ok:
            raise IOError("reference could not write %s" % gcsa_path)

    @staticmethod
    def load(gcsa_path, lcp_path=None):
        """sdsl::load_from_file -> GCSA::load / LCPArray::load of the reference; None if it rejects the file."""
        h = lib().ref_load(str(gcsa_path).encode(), str(lcp_path).encode() if lcp_path else None)
        return ReferenceIndex(h, lcp_path is not None) if h else None

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_destroy(self._h); self._h = None
        if getattr(self, "_tmp", None):
            import shutil
            shutil.rmtree(self._tmp, ignore_errors=True); self._tmp = None

    def verify(self):
        """The reference's verifyIndex() on its own index and input (src/algorithms.cpp:85-99)."""
        return bool(lib().ref_verify(self._h))

    def sizes(self):
        out = np.zeros(10, dtype=np.uint64)
        lib().ref_sizes(self._h, out.ctypes.data)
        return [int(x) for x in out]

    def export(self):
        """-> (FlatGCSA, FlatLCP) with the reference's own arrays."""
        from gcsa2_b200.flat import FlatGCSA, FlatLCP, words_for
        N, E, order, ns, evl, rdl, lsize, lbranch, llevels, lvalues = self.sizes()
        Cc = np.zeros(SIGMA + 1, dtype=np.uint64)
        bwt = [np.zeros(words_for(N) + 1, dtype=np.uint64) for _ in range(SIGMA)]
        ptrs = (C.c_void_p * SIGMA)(*[a.ctypes.data for a in bwt])
        edges = np.zeros(words_for(E) + 1, dtype=np.uint64); sampled = np.zeros(words_for(N) + 1, dtype=np.uint64)
        stored = np.zeros(max(ns, 1), dtype=np.uint64); samples = np.zeros(words_for(ns) + 1, dtype=np.uint64)
        filt = np.zeros(words_for(N) + 1, dtype=np.uint64); vals = np.zeros(words_for(evl) + 1, dtype=np.uint64)
        red = np.zeros(words_for(rdl) + 1, dtype=np.uint64)
        offs = np.zeros(llevels + 1, dtype=np.uint64); data = np.zeros(max(lvalues, 1), dtype=np.uint8)
        lib().ref_export(self._h, Cc.ctypes.data, ptrs, edges.ctypes.data, sampled.ctypes.data, stored.ctypes.data, samples.ctypes.data,
                         filt.ctypes.data, vals.ctypes.data, red.ctypes.data, offs.ctypes.data, data.ctypes.data)
        flat = FlatGCSA(path_nodes=N, edge_count=E, order=order, C=Cc, bwt=bwt, edges=edges, sampled_paths=sampled, sample_count=ns,
                        stored_samples=stored[:ns], samples=samples, extra_filter=filt, extra_values_len=evl, extra_values=vals,
                        redundant_len=rdl, redundant=red)
        lcp = FlatLCP(size=lsize, branching=lbranch, levels=llevels, offsets=offs, data=data[:lvalues])
        return flat, lcp

    def find_batch(self, chars, offsets, threads=1):
        chars = np.ascontiguousarray(chars, dtype=np.uint8); offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        sp = np.zeros(max(n, 1), dtype=np.uint64); ep = np.zeros(max(n, 1), dtype=np.uint64)
        secs = lib().ref_find_batch(self._h, chars.ctypes.data, offsets.ctypes.data, n, sp.ctypes.data, ep.ctypes.data, threads)
        return sp[:n], ep[:n], secs

    def lf_batch(self, sp, ep, comp):
        sp, ep = _u64(sp), _u64(ep); comp = np.ascontiguousarray(comp, dtype=np.uint8); n = comp.size
        a = np.zeros(max(n, 1), dtype=np.uint64); b = np.zeros(max(n, 1), dtype=np.uint64)
        lib().ref_lf_batch(self._h, sp.ctypes.data, ep.ctypes.data, comp.ctypes.data, n, a.ctypes.data, b.ctypes.data)
        return a[:n], b[:n]

    def lf_node_batch(self, nodes):
        n = len(nodes); nodes = _u64(nodes); out = np.zeros(max(n, 1), dtype=np.uint64)
        lib().ref_lf_node_batch(self._h, nodes.ctypes.data, n, out.ctypes.data)
        return out[:n]

    def count_batch(self, sp, ep):
        n = len(sp); sp, ep = _u64(sp), _u64(ep); out = np.zeros(max(n, 1), dtype=np.uint64)
        lib().ref_count_batch(self._h, sp.ctypes.data, ep.ctypes.data, n, out.ctypes.data)
        return out[:n]

    def locate_batch(self, sp, ep, max_positions=0, threads=1):
        n = len(sp); sp, ep = _u64(sp), _u64(ep)
        offs = np.zeros(n + 1, dtype=np.uint64); p = C.c_void_p()
        secs = lib().ref_locate_batch(self._h, sp.ctypes.data, ep.ctypes.data, n, int(max_positions), offs.ctypes.data, C.byref(p), threads)
        total = int(offs[n])
        vals = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(max(total, 1),))[:total].copy()
        lib().ref_free(p)
        return offs, vals, secs

    def parent_batch(self, sp, ep):
        n = len(sp); sp, ep = _u64(sp), _u64(ep); out = np.zeros((max(n, 1), 5), dtype=np.uint64)
        lib().ref_parent_batch(self._h, sp.ctypes.data, ep.ctypes.data, n, out.ctypes.data)
        return out[:n]

    def depth_batch(self, sp, ep):
        n = len(sp); sp, ep = _u64(sp), _u64(ep); out = np.zeros(max(n, 1), dtype=np.uint64)
        lib().ref_depth_batch(self._h, sp.ctypes.data, ep.ctypes.data, n, out.ctypes.data)
        return out[:n]

    def lcp_query(self, which, a, b=None):
        code = {"psv": 0, "psev": 1, "nsv": 2, "nsev": 3, "rmq": 4}[which]
        n = len(a); a = _u64(a); b = _u64(b if b is not None else a)
        op = np.zeros(max(n, 1), dtype=np.uint64); ov = np.zeros(max(n, 1), dtype=np.uint64)
        lib().ref_lcp_query(self._h, code, a.ctypes.data, b.ctypes.data, n, op.ctypes.data, ov.ctypes.data)
        return op[:n], ov[:n]

    def count_kmers(self, k, include_Ns=False):
        return int(lib().ref_count_kmers(self._h, int(k), int(bool(include_Ns))))

    def compare_kmers(self, other, k, include_Ns=False):
        """compareKMers(left, right, k) of the reference -> (shared, left only, right only)."""
        res = np.zeros(3, dtype=np.uint64)
        lib().ref_compare_kmers(self._h, other._h, int(k), int(bool(include_Ns)), res.ctypes.data)
        return tuple(int(x) for x in res)
