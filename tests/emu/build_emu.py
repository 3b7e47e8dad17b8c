"""TEST INFRASTRUCTURE: builds tests/emu/_build/libgcsa2_b200_emu.so, the CUDA engine's own source
(gcsa2_b200/csrc/engine.cu) compiled for the host against tests/emu/cuda_emu.h, so that the kernels can be
executed and compared with the oracle on a machine without a GPU.

The translation is textual and small:
  * `#include <cuda_runtime.h>` / `<cub/cub.cuh>`  ->  `#include "cuda_emu.h"`;
  * `kernel<T...><<<cfg>>>(args);`                 ->  `emu::launchChecked(emu::Cfg(cfg), collectives, [&](auto&&... a) { kernel<T...>(a...); }, args);`
    where `collectives` says whether the kernel (or a device function it names) uses a warp / block collective
    and therefore has to run on fibers; pointer arguments must be device pointers;
  * the one inline-PTX statement (`ld.global.nc.v4.u64`, a 32-byte load) -> a plain load.
Nothing under gcsa2_b200/ knows about this; the product library has no CPU path."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "gcsa2_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libgcsa2_b200_emu.so")
DEVICE_SOURCES = ["engine.cu", "find.cu", "ops.cu", "locate.cu", "lcp.cu", "kmers.cu", "verify.cu", "linear_builder.cu"]   # the first one carries the emulation's out-of-line definitions
HOST_SOURCES = ["builder.cpp", "gcsa_file.cpp", "kmer_file.cpp", "pack.cpp"]
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
COLLECTIVES = ("__ballot_sync", "__any_sync", "__all_sync", "__syncwarp", "__syncthreads", "__shfl_sync",
               "__shfl_down_sync", "__shfl_up_sync", "__shfl_xor_sync")


def _matching(text, start, open_ch, close_ch):
    """Index just past the bracket that closes the one at text[start]."""
    depth = 0
    for i in range(start, len(text)):
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced %s at %d" % (open_ch, start))


def _functions(text, qualifier):
    """(name, body) of every function definition introduced by `qualifier` (__global__ / __device__)."""
    out = []
    for m in re.finditer(re.escape(qualifier), text):
        paren = text.find("(", m.end())
        if text[m.end():paren].rstrip().endswith("__launch_bounds__"):
            after = _matching(text, paren, "(", ")")
            paren = text.find("(", after)
            name = re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", text[after:paren])
        else:
            name = re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", text[m.end():paren])
        close = _matching(text, paren, "(", ")")
        rest = text[close:close + 200].lstrip()
        if not name or not (rest.startswith("{") or rest.startswith("const") or rest.startswith(":")):
            continue
        brace = text.find("{", close)
        if rest.startswith(":"):            # constructor with an initialiser list
            brace = text.find("{", brace)
        out.append((name[0], text[brace:_matching(text, brace, "{", "}")]))
    return out


def inline_device_headers(text):
    """A translation unit includes engine.h and the kernels it launches from csrc/device/*.cuh: paste them in (engine.h
    first, it includes device/layout.cuh itself)."""
    def paste(m):
        with open(os.path.join(CSRC, m.group(1))) as f:
            return "// ---- %s ----\n%s" % (m.group(1), f.read())
    text = re.sub(r'#include "(engine\.h)"', paste, text)
    return re.sub(r'#include "(device/[a-z_]+\.cuh)"', paste, text)


def translate(text, main=True):
    """main: the translation unit that defines the emulation's out-of-line symbols (and holds the one inline-PTX load)."""
    text = inline_device_headers(text)
    if main:
        text = text.replace("#include <cuda_runtime.h>", '#define EMU_DEFINE_SWITCH\n#include "cuda_emu.h"', 1)
    text = text.replace("#include <cuda_runtime.h>", '#include "cuda_emu.h"')
    text = text.replace("#include <cub/cub.cuh>", "")
    text, n_asm = re.subn(r'asm volatile\("ld\.global\.nc\.v4\.u64[^;]*;[^;]*;', "emu::checkAligned(p, 32); r = *p;", text)
    assert n_asm <= 1, "expected at most one inline-PTX load (ld256 of device/layout.cuh), found %d" % n_asm
    assert "asm" not in re.sub(r"//.*", "", text), "untranslated inline assembly"

    # which functions use collectives, directly or through a device function they name
    device = _functions(text, "__device__")
    flagged = {name for name, body in device if any(c in body for c in COLLECTIVES)}
    grew = True
    while grew:
        grew = False
        for name, body in device:
            if name not in flagged and any(re.search(r"\b%s\b" % f, body) for f in flagged):
                flagged.add(name); grew = True
    kernels = {}
    for name, body in _functions(text, "__global__"):
        kernels[name] = any(c in body for c in COLLECTIVES) or any(re.search(r"\b%s\b" % f, body) for f in flagged)

    out, pos, launches = [], 0, 0
    pattern = re.compile(r"([A-Za-z_][A-Za-z0-9_]*)(<[^<>;()]*>)?<<<")
    while True:
        m = pattern.search(text, pos)
        if m is None:
            break
        cfg_end = text.index(">>>", m.end())
        args_open = cfg_end + 3
        assert text[args_open] == "(", text[m.start():args_open + 20]
        args_end = _matching(text, args_open, "(", ")")
        name, targs = m.group(1), m.group(2) or ""
        assert name in kernels, "launch of an unknown kernel: " + name
        out.append(text[pos:m.start()])
        out.append("emu::launchChecked(emu::Cfg(%s), %s, [&](auto&&... emu_args) { %s%s(emu_args...); }, %s)" % (
            text[m.end():cfg_end], "true" if kernels[name] else "false", name, targs, text[args_open + 1:args_end - 1]))
        pos = args_end
        launches += 1
    out.append(text[pos:])
    assert launches > 0
    return "".join(out), kernels


def build(force=False, verbose=False):
    os.makedirs(OUT, exist_ok=True)
    device_cu = [os.path.join(CSRC, s) for s in DEVICE_SOURCES]
    device = [os.path.join(CSRC, "device", f) for f in sorted(os.listdir(os.path.join(CSRC, "device"))) if f.endswith(".cuh")]
    sources = device_cu + [os.path.join(CSRC, "internal.h"), os.path.join(CSRC, "engine.h"), os.path.join(HERE, "cuda_emu.h"), os.path.abspath(__file__),
               os.path.join(ROOT, "include", "gcsa2_b200.h")] + device + [os.path.join(CSRC, s) for s in HOST_SOURCES]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in sources):
        return LIB
    flags = ["-O1", "-g", "-std=c++17", "-fopenmp", "-fPIC", "-fno-strict-aliasing", "-I", HERE, "-w"]
    run = lambda cmd: subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    objs = []
    for i, source in enumerate(device_cu):
        stem = os.path.basename(source)[:-3]
        obj = os.path.join(OUT, stem + "_emu.o")
        objs.append(obj)
        deps = [source, os.path.join(CSRC, "internal.h"), os.path.join(CSRC, "engine.h"), os.path.join(HERE, "cuda_emu.h"), os.path.abspath(__file__),
                os.path.join(ROOT, "include", "gcsa2_b200.h")] + device
        if not force and os.path.exists(obj) and all(os.path.getmtime(s) <= os.path.getmtime(obj) for s in deps):
            continue
        with open(source) as f:
            translated, kernels = translate(f.read(), main=(i == 0))
        # the translated file sits in _build/, the relative includes of the source are resolved against csrc/
        for relative in ('#include "../../include/gcsa2_b200.h"', '#include "../../../include/gcsa2_b200.h"'):
            translated = translated.replace(relative, '#include "%s"' % os.path.join(ROOT, "include", "gcsa2_b200.h"))
        translated = translated.replace('#include "internal.h"', '#include "%s"' % os.path.join(CSRC, "internal.h"))
        cpp = os.path.join(OUT, stem + "_emu.cpp")
        with open(cpp, "w") as f:
            f.write("// GENERATED by tests/emu/build_emu.py from gcsa2_b200/csrc/%s -- test infrastructure, do not edit\n" % os.path.basename(source))
            f.write(translated)
        run([CXX] + flags + ["-c", cpp, "-o", obj])
    for s in HOST_SOURCES:
        obj = os.path.join(OUT, s[:-4] + ".o")
        run([CXX, "-O2", "-march=x86-64-v2", "-std=c++17", "-fopenmp", "-fPIC", "-c", os.path.join(CSRC, s), "-o", obj])
        objs.append(obj)
    run([CXX, "-shared", "-o", LIB] + objs + ["-fopenmp"])
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
