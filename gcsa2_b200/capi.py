"""ctypes binding of include/gcsa2_b200.h (the C ABI of libgcsa2_b200.so)."""
import ctypes as C
import os

import numpy as np

from . import build as _build

SIGMA = 7

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_CAPACITY, ERR_INCONSISTENT = 0, -1, -2, -3, -4, -5


class GCSAError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("gcsa2_b200 error %d: %s" % (code, message))
        self.code = code


class FlatIndex(C.Structure):
    _fields_ = [
        ("path_nodes", C.c_uint64), ("edge_count", C.c_uint64), ("order", C.c_uint64),
        ("sigma", C.c_uint64), ("fast_chars", C.c_uint64),
        ("C", C.c_uint64 * (SIGMA + 1)),
        ("char2comp", C.c_uint8 * 256),
        ("bwt", C.c_void_p * SIGMA),
        ("edges", C.c_void_p), ("sampled_paths", C.c_void_p),
        ("sample_count", C.c_uint64), ("stored_samples", C.c_void_p), ("samples", C.c_void_p),
        ("extra_filter", C.c_void_p), ("extra_values_len", C.c_uint64), ("extra_values", C.c_void_p),
        ("redundant_len", C.c_uint64), ("redundant", C.c_void_p),
    ]


class FlatLcp(C.Structure):
    _fields_ = [("size", C.c_uint64), ("branching", C.c_uint64), ("levels", C.c_uint64),
                ("offsets", C.c_void_p), ("data", C.c_void_p)]


class Options(C.Structure):
    _fields_ = [("kmer_table_k", C.c_int), ("two_step", C.c_int), ("walk_table", C.c_int), ("jump_table", C.c_int), ("fused_table", C.c_int), ("reserved", C.c_int * 3)]


class Info(C.Structure):
    _fields_ = [("path_nodes", C.c_uint64), ("edge_count", C.c_uint64), ("order", C.c_uint64),
                ("sample_count", C.c_uint64), ("device_bytes", C.c_uint64),
                ("kmer_table_k", C.c_int), ("device", C.c_int), ("sm_count", C.c_int), ("two_step", C.c_int), ("jump_k", C.c_int), ("fused_table", C.c_int)]


class FindStats(C.Structure):
    _fields_ = [("queries", C.c_uint64), ("found", C.c_uint64), ("total_length", C.c_uint64),
                ("lf_steps", C.c_uint64), ("sector_probes", C.c_uint64), ("table_hits", C.c_uint64)]


class VerifyReport(C.Structure):
    _fields_ = [("unique", C.c_uint64), ("failures", C.c_uint64), ("find_failures", C.c_uint64),
                ("parent_failures", C.c_uint64), ("depth_failures", C.c_uint64), ("count_failures", C.c_uint64),
                ("locate_failures", C.c_uint64), ("random_locate_failures", C.c_uint64), ("seconds", C.c_double),
                ("engine_seconds", C.c_double)]


class Built(C.Structure):
    _fields_ = [("index", FlatIndex), ("lcp_size", C.c_uint64), ("lcp", C.c_void_p)]


class Graph(C.Structure):
    _fields_ = [("nodes", C.c_uint64), ("comp", C.c_void_p), ("value", C.c_void_p),
                ("succ_offsets", C.c_void_p), ("succ", C.c_void_p), ("sink", C.c_uint64),
                ("n_sources", C.c_uint64), ("sources", C.c_void_p)]


class Kmers(C.Structure):
    _fields_ = [("n", C.c_uint64), ("key", C.c_void_p), ("from_", C.c_void_p), ("to", C.c_void_p)]


# every symbol include/gcsa2_b200.h declares
SYMBOLS = [
    "gcsa_b200_last_error", "gcsa_b200_version", "gcsa_b200_device_count",
    "gcsa_b200_index_create", "gcsa_b200_index_destroy", "gcsa_b200_index_info",
    "gcsa_b200_find_batch", "gcsa_b200_find_host", "gcsa_b200_find_fixed_batch", "gcsa_b200_find_fixed_host",
    "gcsa_b200_find_stats_host", "gcsa_b200_find_fixed_stats_host", "gcsa_b200_char_range",
    "gcsa_b200_find_fixed_host_multi", "gcsa_b200_find_host_multi", "gcsa_b200_locate_into_host_multi",
    "gcsa_b200_lf_batch", "gcsa_b200_lf_host", "gcsa_b200_lf_node_batch", "gcsa_b200_lf_node_host",
    "gcsa_b200_lf_multi_batch", "gcsa_b200_lf_multi_host",
    "gcsa_b200_count_batch", "gcsa_b200_count_host",
    "gcsa_b200_locate_host", "gcsa_b200_locate_into_host", "gcsa_b200_locate_raw_host", "gcsa_b200_locate_batch", "gcsa_b200_locate_max_host", "gcsa_b200_free", "gcsa_b200_count_kmers", "gcsa_b200_compare_kmers", "gcsa_b200_compare_kmers_to_files", "gcsa_b200_verify_index",
    "gcsa_b200_lcp_create", "gcsa_b200_lcp_destroy",
    "gcsa_b200_parent_batch", "gcsa_b200_parent_host", "gcsa_b200_depth_batch", "gcsa_b200_depth_host",
    "gcsa_b200_lcp_sv_host", "gcsa_b200_lcp_rmq_host", "gcsa_b200_mem_batch", "gcsa_b200_mem_host",
    "gcsa_b200_build_from_kmers", "gcsa_b200_built_free", "gcsa_b200_build_linear", "gcsa_b200_build_from_kmers_mapped",
    "gcsa_b200_verify_index_mapped", "gcsa_b200_read_kmer_files", "gcsa_b200_load_node_mapping",
    "gcsa_b200_enumerate_kmers", "gcsa_b200_kmers_free", "gcsa_b200_default_char2comp",
    "gcsa_b200_load_gcsa_file", "gcsa_b200_write_gcsa_file", "gcsa_b200_load_lcp_file", "gcsa_b200_write_lcp_file",
    "gcsa_b200_flat_lcp_free",
]

_lib = None


def lib():
    """Loads libgcsa2_b200.so, building it first if the sources are newer.  Raises if it cannot be
    loaded: there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    try:
        path = _build.build()
    except Exception:
        if not os.path.exists(path):
            raise
    _lib = _bind(C.CDLL(path))
    return _lib


def _bind(L):
    """Declares the prototypes of include/gcsa2_b200.h on a loaded library."""
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.gcsa_b200_last_error.restype = C.c_char_p
    L.gcsa_b200_version.restype = C.c_char_p
    L.gcsa_b200_device_count.restype = i32
    L.gcsa_b200_index_create.argtypes = [C.POINTER(FlatIndex), i32, C.POINTER(Options), C.POINTER(vp)]
    L.gcsa_b200_index_destroy.argtypes = [vp]; L.gcsa_b200_index_destroy.restype = None
    L.gcsa_b200_index_info.argtypes = [vp, C.POINTER(Info)]
    L.gcsa_b200_find_batch.argtypes = [vp, vp, vp, u64, vp, vp, vp]
    L.gcsa_b200_find_host.argtypes = [vp, vp, vp, u64, vp, vp]
    L.gcsa_b200_find_fixed_batch.argtypes = [vp, vp, u64, u64, vp, vp, vp]
    L.gcsa_b200_find_fixed_host.argtypes = [vp, vp, u64, u64, vp, vp]
    L.gcsa_b200_find_stats_host.argtypes = [vp, vp, vp, u64, vp, vp, C.POINTER(FindStats)]
    L.gcsa_b200_find_fixed_host_multi.argtypes = [vp, i32, vp, u64, u64, vp, vp]
    L.gcsa_b200_find_host_multi.argtypes = [vp, i32, vp, vp, u64, vp, vp]
    L.gcsa_b200_locate_into_host_multi.argtypes = [vp, i32, vp, vp, u64, vp, vp, u64, C.POINTER(u64)]
    L.gcsa_b200_find_fixed_stats_host.argtypes = [vp, vp, u64, u64, vp, vp, C.POINTER(FindStats)]
    L.gcsa_b200_char_range.argtypes = [vp, u64, C.POINTER(u64), C.POINTER(u64)]
    L.gcsa_b200_lf_batch.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp]
    L.gcsa_b200_lf_host.argtypes = [vp, vp, vp, vp, u64, vp, vp]
    L.gcsa_b200_lf_node_batch.argtypes = [vp, vp, u64, vp, vp]
    L.gcsa_b200_lf_node_host.argtypes = [vp, vp, u64, vp]
    L.gcsa_b200_lf_multi_batch.argtypes = [vp, vp, vp, u64, i32, vp, vp]
    L.gcsa_b200_lf_multi_host.argtypes = [vp, vp, vp, u64, i32, vp]
    L.gcsa_b200_count_batch.argtypes = [vp, vp, vp, u64, vp, vp]
    L.gcsa_b200_count_host.argtypes = [vp, vp, vp, u64, vp]
    L.gcsa_b200_locate_host.argtypes = [vp, vp, vp, u64, vp, C.POINTER(vp)]
    L.gcsa_b200_locate_into_host.argtypes = [vp, vp, vp, u64, vp, vp, u64, C.POINTER(u64)]
    L.gcsa_b200_locate_raw_host.argtypes = [vp, vp, vp, u64, vp, C.POINTER(vp)]
    L.gcsa_b200_locate_batch.argtypes = [vp, vp, vp, u64, vp, vp, u64, C.POINTER(u64), vp]
    L.gcsa_b200_locate_max_host.argtypes = [vp, vp, vp, u64, u64, vp, C.POINTER(vp)]
    L.gcsa_b200_free.argtypes = [vp]; L.gcsa_b200_free.restype = None
    L.gcsa_b200_count_kmers.argtypes = [vp, u64, i32, C.POINTER(u64), C.POINTER(vp)]
    L.gcsa_b200_compare_kmers.argtypes = [vp, vp, u64, i32, vp, C.POINTER(vp), C.POINTER(vp)]
    L.gcsa_b200_compare_kmers_to_files.argtypes = [vp, vp, u64, i32, C.c_char_p, vp]
    L.gcsa_b200_verify_index.argtypes = [vp, vp, vp, vp, u64, i32, C.POINTER(VerifyReport)]
    L.gcsa_b200_lcp_create.argtypes = [C.POINTER(FlatLcp), i32, C.POINTER(vp)]
    L.gcsa_b200_lcp_destroy.argtypes = [vp]; L.gcsa_b200_lcp_destroy.restype = None
    L.gcsa_b200_parent_batch.argtypes = [vp, vp, vp, u64, vp, vp]
    L.gcsa_b200_parent_host.argtypes = [vp, vp, vp, u64, vp]
    L.gcsa_b200_depth_batch.argtypes = [vp, vp, vp, u64, vp, vp]
    L.gcsa_b200_depth_host.argtypes = [vp, vp, vp, u64, vp]
    L.gcsa_b200_lcp_sv_host.argtypes = [vp, i32, vp, u64, vp, vp]
    L.gcsa_b200_lcp_rmq_host.argtypes = [vp, vp, vp, u64, vp, vp]
    L.gcsa_b200_mem_batch.argtypes = [vp, vp, vp, vp, u64, vp, vp, u64, C.POINTER(u64), vp]
    L.gcsa_b200_mem_host.argtypes = [vp, vp, vp, vp, u64, vp, C.POINTER(vp)]
    L.gcsa_b200_build_from_kmers.argtypes = [vp, vp, vp, u64, i32, i32, u64, C.POINTER(Built)]
    L.gcsa_b200_build_from_kmers_mapped.argtypes = [vp, vp, vp, u64, i32, i32, u64, u64, vp, u64, C.POINTER(Built)]
    L.gcsa_b200_verify_index_mapped.argtypes = [vp, vp, vp, vp, u64, i32, u64, vp, u64, C.POINTER(VerifyReport)]
    L.gcsa_b200_read_kmer_files.argtypes = [C.POINTER(C.c_char_p), i32, i32, vp, C.POINTER(Kmers), C.POINTER(i32)]
    L.gcsa_b200_load_node_mapping.argtypes = [C.c_char_p, C.POINTER(u64), C.POINTER(vp), C.POINTER(u64)]
    L.gcsa_b200_built_free.argtypes = [C.POINTER(Built)]; L.gcsa_b200_built_free.restype = None
    L.gcsa_b200_build_linear.argtypes = [vp, u64, i32, u64, i32, i32, u64, i32, C.POINTER(Built)]
    L.gcsa_b200_enumerate_kmers.argtypes = [C.POINTER(Graph), i32, C.POINTER(Kmers)]
    L.gcsa_b200_kmers_free.argtypes = [C.POINTER(Kmers)]; L.gcsa_b200_kmers_free.restype = None
    L.gcsa_b200_default_char2comp.argtypes = [vp]; L.gcsa_b200_default_char2comp.restype = None
    L.gcsa_b200_load_gcsa_file.argtypes = [C.c_char_p, C.POINTER(Built)]
    L.gcsa_b200_write_gcsa_file.argtypes = [C.POINTER(FlatIndex), C.c_char_p]
    L.gcsa_b200_load_lcp_file.argtypes = [C.c_char_p, C.POINTER(FlatLcp)]
    L.gcsa_b200_write_lcp_file.argtypes = [C.POINTER(FlatLcp), C.c_char_p]
    L.gcsa_b200_flat_lcp_free.argtypes = [C.POINTER(FlatLcp)]; L.gcsa_b200_flat_lcp_free.restype = None
    return L


def check(rc, allow=()):
    if rc != 0 and rc not in allow:
        raise GCSAError(rc, lib().gcsa_b200_last_error().decode(errors="replace"))
    return rc


def ptr(a):
    """Raw address of a numpy array, a torch tensor (host or device) or an int."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def as_u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a if a.size else np.zeros(1, dtype=np.uint64)


def flat_from_struct(f):
    """C struct gcsa_flat_index (arrays owned by the library) -> FlatGCSA with its own copies."""
    from .flat import FlatGCSA, words_for
    def bits(p, n_bits):
        n = words_for(n_bits) + 1
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(n,)).copy()
    N = int(f.path_nodes)
    return FlatGCSA(
        path_nodes=N, edge_count=int(f.edge_count), order=int(f.order),
        C=np.array([f.C[i] for i in range(SIGMA + 1)], dtype=np.uint64),
        bwt=[bits(f.bwt[c], N) for c in range(SIGMA)],
        edges=bits(f.edges, f.edge_count), sampled_paths=bits(f.sampled_paths, N),
        sample_count=int(f.sample_count),
        stored_samples=np.ctypeslib.as_array(C.cast(f.stored_samples, C.POINTER(C.c_uint64)),
                                             shape=(max(1, int(f.sample_count)),))[:int(f.sample_count)].copy(),
        samples=bits(f.samples, f.sample_count), extra_filter=bits(f.extra_filter, N),
        extra_values_len=int(f.extra_values_len), extra_values=bits(f.extra_values, f.extra_values_len),
        redundant_len=int(f.redundant_len), redundant=bits(f.redundant, f.redundant_len),
        sigma=int(f.sigma), fast_chars=int(f.fast_chars),
        char2comp=np.frombuffer(bytes(f.char2comp), dtype=np.uint8).copy())


def lcp_struct(lcp, keep):
    """FlatLCP (numpy) -> C struct gcsa_flat_lcp."""
    f = FlatLcp()
    offsets = as_u64(lcp.offsets)
    data = np.ascontiguousarray(lcp.data, dtype=np.uint8)
    if data.size == 0:
        data = np.zeros(1, dtype=np.uint8)
    keep.extend([offsets, data])
    f.size, f.branching, f.levels = int(lcp.size), int(lcp.branching), int(lcp.levels)
    f.offsets, f.data = offsets.ctypes.data, data.ctypes.data
    return f


def flat_struct(flat, keep):
    """FlatGCSA (numpy) -> C struct; `keep` receives the arrays that must outlive the call."""
    f = FlatIndex()
    f.path_nodes, f.edge_count, f.order = int(flat.path_nodes), int(flat.edge_count), int(flat.order)
    f.sigma, f.fast_chars = int(flat.sigma), int(flat.fast_chars)
    for i in range(SIGMA + 1):
        f.C[i] = int(flat.C[i])
    c2c = np.ascontiguousarray(flat.char2comp, dtype=np.uint8)
    C.memmove(f.char2comp, c2c.ctypes.data, 256)
    def hold(a, dtype=np.uint64):
        a = np.ascontiguousarray(a, dtype=dtype)
        if a.size == 0:
            a = np.zeros(1, dtype=dtype)
        keep.append(a)
        return a.ctypes.data
    for c in range(SIGMA):
        f.bwt[c] = hold(flat.bwt[c])
    f.edges = hold(flat.edges); f.sampled_paths = hold(flat.sampled_paths)
    f.sample_count = int(flat.sample_count)
    f.stored_samples = hold(flat.stored_samples); f.samples = hold(flat.samples)
    f.extra_filter = hold(flat.extra_filter)
    f.extra_values_len = int(flat.extra_values_len); f.extra_values = hold(flat.extra_values)
    f.redundant_len = int(flat.redundant_len); f.redundant = hold(flat.redundant)
    return f
