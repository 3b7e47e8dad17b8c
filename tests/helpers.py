"""Shared fixtures for the test-suite (test infrastructure)."""
import json
import os

import numpy as np

from gcsa2_b200.flat import FlatGCSA, SIGMA, bits_from_positions

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_kat1():
    with open(os.path.join(GOLDEN, "kat1_fig3.json")) as f:
        return json.load(f)


def bits_from_string(s):
    return bits_from_positions([i for i, ch in enumerate(s) if ch == "1"], len(s))


def kat1_flat(kat=None):
    """FlatGCSA for the paper's Figure 3.  The figure has no counting structures; they are
    derived from its value sets (A[i] = |values| - 1) and from a brute-force distinct count."""
    kat = kat or load_kat1()
    n = len(kat["keys"])
    occ = [len(v) - 1 for v in kat["values"]]
    filt = [i for i in range(n) if occ[i] > 0]
    ev_ones, tail = [], 0
    for i in filt:
        tail += occ[i]; ev_ones.append(tail - 1)
    lcp = [0] * n
    for i in range(1, n):
        a, b = kat["keys"][i - 1], kat["keys"][i]
        k = 0
        while k < min(len(a), len(b)) and a[k] == b[k]:
            k += 1
        lcp[i] = k
    from brute import redundant_counts
    red = redundant_counts(kat["values"], lcp)      # the walk of src/gcsa.cpp:590-619
    rd_ones, tail = [], 0
    for i in range(n - 1):
        tail += red[i] + 1; rd_ones.append(tail - 1)
    flat = FlatGCSA(
        path_nodes=n, edge_count=len(kat["edges"]), order=kat["order"], C=np.array(kat["C"], dtype=np.uint64),
        bwt=[bits_from_positions(kat["bwt"][str(c)], n) for c in range(SIGMA)],
        edges=bits_from_string(kat["edges"]),
        sampled_paths=bits_from_positions(kat["sampled_paths"], n),
        sample_count=len(kat["stored_samples"]), stored_samples=np.array(kat["stored_samples"], dtype=np.uint64),
        samples=bits_from_string(kat["samples"]),
        extra_filter=bits_from_positions(filt, n), extra_values_len=sum(occ),
        extra_values=bits_from_positions(ev_ones, sum(occ)),
        redundant_len=(n - 1) + sum(red), redundant=bits_from_positions(rd_ones, (n - 1) + sum(red)))
    return flat, lcp


# ---- "device" buffers for the tests that call the device-pointer entry points -------------------------------
# Under the host emulation (tests/emu) device memory is host memory and there is one implicit stream.
EMULATED = False


def as_emulated_device_memory(tensor):
    """Under the emulation a host tensor stands for device memory: tell the emulated runtime, which refuses kernel
    arguments that point anywhere else (as the device would fault on them)."""
    import ctypes
    import weakref
    from gcsa2_b200 import capi
    lib, address = capi.lib(), tensor.data_ptr()
    lib.emu_register_device_range(ctypes.c_void_p(address), ctypes.c_size_t(max(1, tensor.numel() * tensor.element_size())))
    weakref.finalize(tensor, lib.emu_unregister_device_range, ctypes.c_void_p(address))
    return tensor


def to_device(array):
    import torch
    t = torch.from_numpy(array)
    return as_emulated_device_memory(t.clone()) if EMULATED else t.cuda()


def device_empty(shape, dtype):
    import torch
    if EMULATED:
        return as_emulated_device_memory(torch.empty(shape, dtype=dtype))
    return torch.empty(shape, dtype=dtype, device="cuda")


def current_stream():
    import torch
    return 0 if EMULATED else torch.cuda.current_stream().cuda_stream


def device_sync():
    import torch
    if not EMULATED:
        torch.cuda.synchronize()
