"""Builds a variant of libnbk.so with extra nvcc flags for kernel A/B runs:

    python scripts/build_variant.py NAME [-DNBK_QUEUE_CAP=16 ...]   ->  nbodyhpc_b200/lib/variants/libnbk_NAME.so

Select it at run time with NBK_LIBRARY=<path> (nbodyhpc_b200/capi.py).  The .so is git-ignored but travels
to the GPU box with the snapshot."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodyhpc_b200._build import CSRC, INCLUDE, LIBDIR, NVCC_FLAGS, _nvcc  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(LIBDIR, "variants")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, f"libnbk_{name}.so")
cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-shared", "-I", INCLUDE, "-I", CSRC, "-o", out, os.path.join(CSRC, "nbk.cu")]
print(" ".join(cmd), flush=True)
subprocess.run(cmd, check=True)
print(out)
