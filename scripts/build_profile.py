#!/usr/bin/env python
"""Builds one side^3 uniform periodic tree from device-resident points (for ncu launch lists).
    python scripts/build_profile.py [side] [repeats]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nbodyhpc_b200 import capi

side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = side ** 3
pts = torch.rand((n, 3), device="cuda")
stream = torch.cuda.current_stream().cuda_stream
capi.Tree.build_device(pts.data_ptr(), 1 << 20, 64, 1.0, stream=stream).close()
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = capi.launch_count()
    e0.record()
    t = capi.Tree.build_device(pts.data_ptr(), n, 64, 1.0, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"side": side, "build_ms": e0.elapsed_time(e1), "launches": capi.launch_count() - l0,
                      "nodes": t.size}), flush=True)
    t.close()
