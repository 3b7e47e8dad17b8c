"""The host emulation the [emu] tests rest on (tests/emu/cuda_emu.h), checked on hand-verifiable kernels: ballots
with exited lanes, shuffles, barriers, atomics, intrinsics, and the guard page behind every device buffer."""
import os
import subprocess

import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


@pytest.fixture(scope="module")
def selfcheck(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("emu") / "selfcheck")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-I", HERE, os.path.join(HERE, "selfcheck.cpp"), "-o", exe])
    return exe


@pytest.mark.parametrize("shuffle", ["", "3", "99"])
def test_emulated_constructs(selfcheck, shuffle):
    env = dict(os.environ)
    env.pop("GCSA_EMU_SHUFFLE", None)
    if shuffle:
        env["GCSA_EMU_SHUFFLE"] = shuffle
    res = subprocess.run([selfcheck], capture_output=True, text=True, env=env, timeout=120)
    assert res.returncode == 0 and "selfcheck OK" in res.stdout, res.stdout + res.stderr
    assert ("shuffled" in res.stdout) == bool(shuffle)


def test_guard_page_catches_an_overrun(tmp_path):
    src = tmp_path / "overrun.cpp"
    src.write_text('#define EMU_DEFINE_SWITCH\n#include "cuda_emu.h"\n'
                   '__global__ void k(unsigned long long* p) { p[threadIdx.x] = 1; }\n'
                   'int main(int argc, char**) { unsigned long long* p = nullptr; cudaMalloc(&p, 64 * sizeof(*p));\n'
                   '  emu::launch(emu::Cfg(1, argc > 1 ? 65 : 64), [&] { k(p); }, false); std::puts("done"); return 0; }\n')
    exe = str(tmp_path / "overrun")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", HERE, str(src), "-o", exe])
    assert subprocess.run([exe], capture_output=True, text=True).stdout.strip() == "done"       # 64 threads: inside the buffer
    assert subprocess.run([exe, "x"], capture_output=True, text=True).returncode < 0            # 65: one element past the end -> SIGSEGV


def test_kernel_arguments_must_be_device_pointers(tmp_path):
    """A pointer to ordinary host memory works on the host and faults on the device: the emulated launch refuses it."""
    src = tmp_path / "hostptr.cpp"
    src.write_text('#define EMU_DEFINE_SWITCH\n#include "cuda_emu.h"\n'
                   '__global__ void k(unsigned* p, int n) { if((int)threadIdx.x < n) { p[threadIdx.x] = 1; } }\n'
                   'int main(int argc, char**) { unsigned host[64]; unsigned* dev = nullptr; cudaMalloc(&dev, sizeof(host));\n'
                   '  unsigned* p = (argc > 1 ? host : dev);\n'
                   '  emu::launchChecked(emu::Cfg(1, 64), false, [&](auto&&... a) { k(a...); }, p + 3, 8); std::puts("done"); return 0; }\n')
    exe = str(tmp_path / "hostptr")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", HERE, str(src), "-o", exe])
    assert subprocess.run([exe], capture_output=True, text=True).stdout.strip() == "done"
    res = subprocess.run([exe, "x"], capture_output=True, text=True)
    assert res.returncode < 0 and "not a device pointer" in res.stderr
