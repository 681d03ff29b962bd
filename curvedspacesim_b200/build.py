"""In-tree build of libcurvedspacesim_b200.so (nvcc, sm_100a only).  The library is the product; there is
no Python/CPU fallback for any of its entry points."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcurvedspacesim_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
# (source, extra flags).  exact_kernels.cu must not contract a*b+c into FMA (bit parity with the oracle).
UNITS = [("exact_kernels.cu", ["-fmad=false"]), ("geodesic_kernel.cu", []), ("patch_kernel.cu", []), ("stencil_kernel.cu", []), ("window_kernel.cu", []), ("window_half_kernel.cu", []),
         ("microbench.cu", []), ("css_api.cu", [])]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdrs = [os.path.join(CSRC, h) for h in ("common.cuh", "kernels.h", "window_common.cuh")] + [os.path.join(HERE, "..", "include", "css_api.h")]
    objs, cmds = [], []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmds.append([nvcc] + ARCH + COMMON + extra + ["-c", s, "-o", o])
    if cmds:  # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor

        def run(cmd):
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.check_call(cmd)

        with ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 1)) as ex:
            list(ex.map(run, cmds))
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lnccl", "-Xlinker", "--no-undefined"]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))


def build_host_example(force: bool = False) -> str:
    """g++ build of host/example_simulation.cpp (the C++ host layer above the C ABI) against the in-tree library."""
    root = os.path.dirname(HERE)
    src = os.path.join(root, "host", "example_simulation.cpp")
    hdr = os.path.join(root, "host", "css_host.hpp")
    exe = os.path.join(root, "host", "example_simulation.bin")
    build()
    if force or _stale(exe, [src, hdr, LIB, os.path.join(root, "include", "css_api.h")]):
        subprocess.check_call([os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-Wall", "-Wextra", "-o", exe, src, "-L" + HERE,
                               "-lcurvedspacesim_b200", "-Wl,-rpath," + HERE])
    return exe
