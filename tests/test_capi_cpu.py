"""The C-ABI library on a machine without a GPU: it loads, exports every symbol the header
declares, and refuses to create an index instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import pytest

from helpers import kat1_flat
from gcsa2_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    with open(os.path.join(ROOT, "include", "gcsa2_b200.h")) as f:
        header = f.read()
    declared = sorted(set(re.findall(r"\b(gcsa_b200_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(L, name), "missing symbol " + name
    assert sorted(capi.SYMBOLS) == declared
    assert b"sm_100a" in L.gcsa_b200_version()


def test_struct_sizes_match_header():
    assert C.sizeof(capi.FlatIndex) == 8 * 5 + 8 * 8 + 256 + 8 * 7 + 8 * 10
    assert C.sizeof(capi.FlatLcp) == 40 and C.sizeof(capi.Info) == 64 and C.sizeof(capi.FindStats) == 48


def test_ctypes_structs_have_the_sizes_the_compiler_gives_the_header(tmp_path):
    """The ctypes mirrors of capi.py against sizeof() of the structs of include/gcsa2_b200.h, from the C compiler."""
    import subprocess
    pairs = [("gcsa_flat_index", capi.FlatIndex), ("gcsa_flat_lcp", capi.FlatLcp), ("gcsa_b200_options", capi.Options),
             ("gcsa_b200_info", capi.Info), ("gcsa_b200_find_stats", capi.FindStats), ("gcsa_b200_verify_report", capi.VerifyReport),
             ("gcsa_b200_built", capi.Built), ("gcsa_b200_graph", capi.Graph), ("gcsa_b200_kmers", capi.Kmers)]
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "gcsa2_b200.h"\nint main(void) {\n' +
                   "".join('  printf("%s %%zu\\n", sizeof(%s));\n' % (name, name) for name, _ in pairs) + "  return 0;\n}\n")
    exe = str(tmp_path / "sizes")
    subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe])
    sizes = dict(line.split() for line in subprocess.check_output([exe], text=True).splitlines())
    for name, mirror in pairs:
        assert int(sizes[name]) == C.sizeof(mirror), (name, sizes[name], C.sizeof(mirror))


def test_product_does_not_know_the_emulation():
    """tests/emu is test infrastructure: nothing under gcsa2_b200/ or include/ mentions it."""
    for top in ("gcsa2_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for name in files:
                if name.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    with open(os.path.join(dirpath, name), errors="replace") as f:
                        text = f.read()
                    assert "cuda_emu" not in text and "build_emu" not in text and "GCSA_EMU" not in text, os.path.join(dirpath, name)


def test_no_cuda_device_fails_loudly():
    L = capi.lib()
    if L.gcsa_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    flat, _ = kat1_flat()
    keep = []
    f = capi.flat_struct(flat, keep)
    h = C.c_void_p()
    rc = L.gcsa_b200_index_create(C.byref(f), 0, None, C.byref(h))
    assert rc == capi.ERR_CUDA and not h.value
    assert b"no CPU fallback" in L.gcsa_b200_last_error()
    from gcsa2_b200 import GCSA, GCSAError
    with pytest.raises(GCSAError):
        GCSA(flat)


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "gcsa2_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cpp", ".h")):
                with open(os.path.join(dirpath, name), errors="replace") as f:
                    text = f.read()
                assert "oracle" not in text.lower().replace("test-suite", ""), os.path.join(dirpath, name)
