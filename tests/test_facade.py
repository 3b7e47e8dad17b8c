"""The C++ facade (include/gcsa2_b200.hpp): it compiles against the C ABI on any machine; on a GPU
it reproduces the reference's query_gcsa / verifyIndex checks against a naive scan of the text."""
import os
import subprocess

import pytest

from gcsa2_b200 import build as _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def compile_facade_test(tmp_path, emulated=False):
    """Links tests/cpp/facade_test.cpp against the product library, or (emulated) against the host emulation of
    the engine built by tests/emu -- the same program, the same header."""
    _build.build()
    exe = os.path.join(str(tmp_path), "facade_test")
    lib_dir, lib = os.path.join(ROOT, "gcsa2_b200"), "gcsa2_b200"
    if emulated:
        from emu import build_emu
        lib_dir, lib = os.path.dirname(build_emu.build()), "gcsa2_b200_emu"
    subprocess.check_call([_build.CXX, "-std=c++17", "-O1", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"), "-L" + lib_dir, "-l" + lib,
                           "-Wl,-rpath," + lib_dir, "-o", exe])
    return exe


def test_facade_compiles_and_refuses_to_run_without_gpu(tmp_path):
    exe = compile_facade_test(tmp_path)
    from gcsa2_b200 import capi
    if capi.lib().gcsa_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode != 0 and "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_facade_query_loop_on_gpu(tmp_path):
    exe = compile_facade_test(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "facade_test OK" in res.stdout, res.stdout + res.stderr


def test_facade_query_loop_on_the_emulated_engine(tmp_path):
    exe = compile_facade_test(tmp_path, emulated=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "facade_test OK" in res.stdout, res.stdout + res.stderr
