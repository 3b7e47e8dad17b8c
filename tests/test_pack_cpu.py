"""Host-side 2-bit packing of fixed-length patterns (gcsa2_b200/csrc/pack.cpp), the step that precedes the
H2D copy in gcsa_b200_find_fixed_host: AVX2 and table paths against a numpy restatement of the layout."""
import ctypes as C

import numpy as np
import pytest

from gcsa2_b200 import capi
from gcsa2_b200.flat import DEFAULT_CHAR2COMP


def pack_numpy(chars, n, length, code):
    words = (length + 31) // 32
    out = np.zeros((n, words), dtype=np.uint64)
    c = code[chars.reshape(n, length)].astype(np.uint64)
    for p in range(length):
        out[:, p // 32] |= (c[:, p] & np.uint64(3)) << np.uint64(2 * (p % 32))
    return out, bool((code[chars] != 0xFF).all())


def run(chars, n, length, code, default_alphabet, threads):
    L = capi.lib()
    fn = L.gcsa_b200_internal_pack_patterns
    fn.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    fn.restype = C.c_int
    words = (length + 31) // 32
    out = np.zeros(max(1, n * words), dtype=np.uint64)
    buf = np.concatenate([chars, np.zeros(64, dtype=np.uint8)])          # the caller's buffer continues after the batch
    ok = fn(buf.ctypes.data, n, length, code.ctypes.data, default_alphabet, out.ctypes.data, threads)
    return out[:n * words].reshape(n, words), bool(ok)


def default_code():
    code = np.full(256, 0xFF, dtype=np.uint8)
    fast = (DEFAULT_CHAR2COMP >= 1) & (DEFAULT_CHAR2COMP <= 4)
    code[fast] = DEFAULT_CHAR2COMP[fast] - 1
    return code


@pytest.mark.parametrize("length", [1, 5, 16, 31, 32, 33, 64, 70, 128])
def test_pack_matches_layout(length):
    rng = np.random.default_rng(length)
    n = 20_011
    code = default_code()
    chars = np.frombuffer(b"ACGTacgt", dtype=np.uint8)[rng.integers(0, 8, size=n * length)]
    expect, ok = pack_numpy(chars, n, length, code)
    assert ok
    for simd in (0, 1):
        for threads in (1, 4):
            got, good = run(chars, n, length, code, simd, threads)
            assert good and (got == expect).all(), (simd, threads)


def test_pack_reports_other_characters():
    code = default_code()
    rng = np.random.default_rng(3)
    n, length = 5000, 32
    base = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n * length)]
    for bad in (ord("N"), ord("$"), ord("#"), 0, ord("B"), ord("U"), 0xC1, ord("@"), ord("t") + 1):
        for where in (0, n * length - 1, 77 * length + 13):
            chars = base.copy(); chars[where] = bad
            for simd in (0, 1):
                assert run(chars, n, length, code, simd, 3)[1] is False, (bad, where, simd)
    assert run(base, n, length, code, 1, 3)[1] is True
    # a remapped alphabet goes through the table: lower-case letters are not bases here
    remapped = code.copy(); remapped[[ord(c) for c in "acgt"]] = 0xFF
    lower = np.frombuffer(b"acgt", dtype=np.uint8)[rng.integers(0, 4, size=n * length)]
    assert run(lower, n, length, remapped, 0, 2)[1] is False
