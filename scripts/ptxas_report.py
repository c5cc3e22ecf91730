"""Register / spill report of the library's kernels: nvcc -Xptxas -v on nbk.cu, demangled.

    python scripts/ptxas_report.py [filter-substring] [extra nvcc flags...]
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodyhpc_b200._build import CSRC, INCLUDE, NVCC_FLAGS, _nvcc  # noqa: E402

filt = sys.argv[1] if len(sys.argv) > 1 else ""
extra = sys.argv[2:]
cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-Xptxas", "-v", "-shared", "-I", INCLUDE, "-I", CSRC, "-o", "/tmp/_ptxas_report.so",
       os.path.join(CSRC, "nbk.cu")]
out = subprocess.run(cmd, capture_output=True, text=True).stderr
names = re.findall(r"Compiling entry function '(\S+)'", out)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
blocks = out.split("Compiling entry function")[1:]
for name, blk in zip(dem, blocks):
    if filt and filt not in name:
        continue
    regs = re.search(r"Used (\d+) registers", blk)
    spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", blk)
    smem = re.search(r"(\d+) bytes smem", blk)
    print(f"{int(regs.group(1)):4d} regs  stack {spill.group(1):>5s}  spill st/ld {spill.group(2)}/{spill.group(3)}  "
          f"smem {smem.group(1) if smem else 0:>6}  {name[:150]}")
