"""In-tree build of the native pieces (no pip, no JIT cache): explicit nvcc / g++ commands.

    python -m nbodyhpc_b200._build          # builds everything that is out of date

Artifacts (git-ignored, shipped to the GPU box with the snapshot):
    nbodyhpc_b200/lib/libnbk.so                  CUDA kernels + the C ABI (include/nbk.h)
    nbodyhpc_b200/kdtree/_impl.<abi>.so          pybind11 module mirroring the reference's _impl
    nbodyhpc_b200/lib/kdtree_main                C++ CLI over the wenda::kdtree::KDTree drop-in class
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nbodyhpc_b200")
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
INCLUDE = os.path.join(ROOT, "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # the distance code uses __fmul_rn/__fadd_rn explicitly; this keeps every other float
    # expression in the library uncontracted as well
    "-fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def _csrc_files():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [
        os.path.join(INCLUDE, "nbk.h"),
        os.path.join(INCLUDE, "kdtree", "kdtree.hpp"),
        os.path.join(INCLUDE, "kdtree", "position_array.hpp"),
    ]


def libnbk_path() -> str:
    return os.path.join(LIBDIR, "libnbk.so")


def impl_path() -> str:
    return os.path.join(PKG, "kdtree", "_impl" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_libnbk(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    out = libnbk_path()
    srcs = [f for f in _csrc_files() if f.endswith((".cu", ".cuh", "nbk.h"))]
    if force or _stale(out, srcs):
        _run([_nvcc(), *NVCC_FLAGS, "-shared", "-I", INCLUDE, "-I", CSRC, "-o", out,
              os.path.join(CSRC, "nbk.cu")], verbose)
    return out


def build_pybind(force: bool = False, verbose: bool = False) -> str:
    import pybind11

    out = impl_path()
    srcs = [os.path.join(CSRC, "pybind_module.cpp")] + [f for f in _csrc_files() if f.endswith((".hpp", ".h"))]
    if force or _stale(out, srcs) or _stale(out, [libnbk_path()]):
        _run(["g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-fvisibility=hidden",
              "-I", INCLUDE, "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
              os.path.join(CSRC, "pybind_module.cpp"), "-o", out,
              "-L", LIBDIR, "-lnbk", "-Wl,-rpath,$ORIGIN/../lib"], verbose)
    return out


def build_cli(force: bool = False, verbose: bool = False) -> str:
    out = os.path.join(LIBDIR, "kdtree_main")
    srcs = [os.path.join(CSRC, "kdtree_main.cpp")] + [f for f in _csrc_files() if f.endswith((".hpp", ".h"))]
    if force or _stale(out, srcs) or _stale(out, [libnbk_path()]):
        _run(["g++", "-O2", "-std=c++20", "-I", INCLUDE, os.path.join(CSRC, "kdtree_main.cpp"), "-o", out,
              "-L", LIBDIR, "-lnbk", "-Wl,-rpath,$ORIGIN"], verbose)
    return out


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_libnbk(force, verbose)
    build_pybind(force, verbose)
    if os.path.exists(os.path.join(CSRC, "kdtree_main.cpp")):
        build_cli(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
