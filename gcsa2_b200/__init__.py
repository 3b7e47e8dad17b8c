"""gcsa2_b200: B200-native batched backward-search engine behind GCSA2's query interface.

The product is libgcsa2_b200.so (hand-written CUDA for sm_100a + C++ host code) with the C ABI
of include/gcsa2_b200.h.  This package is the Python host-side mirror of the reference's query
interface (GCSA, LCPArray), the builder wrapper and the synthetic inputs of the benchmarks.
"""
from .flat import FlatGCSA, FlatLCP, node_encode, node_id, node_offset, node_rc  # noqa: F401
from .index import GCSA, LCPArray, MultiGCSA, pack_patterns, range_empty, range_length, mem_batch, mem_device, UNKNOWN  # noqa: F401
from .capi import GCSAError  # noqa: F401
