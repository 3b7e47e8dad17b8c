"""Builds gcsa2_b200/libgcsa2_b200.so in-tree: nvcc for sm_100a (the CUDA translation units; each includes the kernels
it launches from csrc/device/*.cuh) + g++ (the host-only sources).

The shared library is self-contained (static cudart), so it travels to the GPU box with the
snapshot and loads on a CPU-only machine too (symbol checks in the CPU test-suite)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgcsa2_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include", "gcsa2_b200.h")
DEVICE_SOURCES = ["engine.cu", "find.cu", "ops.cu", "locate.cu", "lcp.cu", "kmers.cu", "verify.cu", "linear_builder.cu"]   # CUDA translation units (nvcc, sm_100a)
HOST_SOURCES = ["builder.cpp", "gcsa_file.cpp", "kmer_file.cpp", "pack.cpp"]      # host-side C++ (g++)

NVCC = os.environ.get("GCSA_B200_NVCC", "nvcc")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fopenmp", "-Wno-deprecated-gpu-targets"]
CXX_FLAGS = ["-O2", "-march=x86-64-v2", "-std=c++17", "-fopenmp", "-fPIC", "-Wall", "-Wextra"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    device_cu = [os.path.join(CSRC, name) for name in DEVICE_SOURCES]
    host_cpp = [os.path.join(CSRC, name) for name in HOST_SOURCES]
    device_o = [src[:-3] + ".o" for src in device_cu]
    host_o = [src[:-4] + ".o" for src in host_cpp]
    headers = [os.path.join(CSRC, "device", f) for f in sorted(os.listdir(os.path.join(CSRC, "device"))) if f.endswith(".cuh")]
    headers += [INCLUDE, os.path.join(CSRC, "internal.h"), os.path.join(CSRC, "engine.h")]
    if not force and not _stale(LIB, device_cu + headers + host_cpp):
        return LIB
    run = lambda cmd: subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    # only what changed is recompiled, the translation units side by side
    jobs = [subprocess.Popen([NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj], stdout=None if verbose else subprocess.DEVNULL)
            for src, obj in zip(device_cu, device_o) if force or _stale(obj, [src] + headers)]
    failed = [job.args for job in jobs if job.wait() != 0]
    if failed:
        raise subprocess.CalledProcessError(1, failed[0])
    for src, obj in zip(host_cpp, host_o):
        if force or _stale(obj, [src] + headers):
            run([CXX] + CXX_FLAGS + ["-c", src, "-o", obj])
    run([NVCC, "-shared", "-Wno-deprecated-gpu-targets", "-o", LIB] + device_o + host_o + ["-Xcompiler", "-fopenmp", "-lgomp"])
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
