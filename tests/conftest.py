import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "engine: exercises the CUDA engine through the C ABI; collected twice -- [cuda] on the real "
                                       "library (carries the gpu marker) and [emu] on the same source compiled against the host "
                                       "emulation of tests/emu (runs anywhere)")


def pytest_generate_tests(metafunc):
    if metafunc.definition.get_closest_marker("engine") is not None:
        kinds = [pytest.param("cuda", marks=pytest.mark.gpu)]
        if metafunc.definition.get_closest_marker("gpu") is None:      # `gpu` on top of `engine`: too large to emulate
            kinds.append(pytest.param("emu"))
        metafunc.parametrize("engine_kind", kinds, indirect=True)


_emulated = {}


@pytest.fixture(autouse=True)
def engine_kind(request):
    """For tests marked `engine`: which build of gcsa2_b200/csrc/engine.cu answers capi.lib().  The emulated build
    is test infrastructure (tests/emu); it is swapped in for the duration of one test only."""
    kind = getattr(request, "param", None)
    if kind != "emu":
        yield kind
        return
    import helpers
    from gcsa2_b200 import capi
    if "lib" not in _emulated:
        from emu import build_emu
        _emulated["lib"] = capi._bind(ctypes.CDLL(build_emu.build()))
    saved = capi._lib
    capi._lib = _emulated["lib"]
    helpers.EMULATED = True
    try:
        yield kind
    finally:
        capi._lib = saved
        helpers.EMULATED = False
