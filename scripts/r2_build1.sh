V=nbodyhpc_b200/lib/variants
echo "== default"; python scripts/build_profile.py 512 5
echo "== cap4096"; NBK_LIBRARY=$V/libnbk_cap4096.so python scripts/build_profile.py 512 5
echo "== bins 1024"; NBK_BUILD_BINS=1024 python scripts/build_profile.py 512 5
echo "== cap4096 bins 1024"; NBK_BUILD_BINS=1024 NBK_LIBRARY=$V/libnbk_cap4096.so python scripts/build_profile.py 512 5
echo "== cap4096 trace"; NBK_BUILD_TRACE=1 NBK_LIBRARY=$V/libnbk_cap4096.so python scripts/build_profile.py 512 2 2>&1 | grep -v "^{" | tail -8
echo "== cold start trace (first build of the process is the 512^3 one)"
NBK_BUILD_TRACE=1 python - <<'PY' 2>&1 | tail -12
import sys, time
sys.path.insert(0, ".")
import torch
from nbodyhpc_b200 import capi
pts = torch.rand((512**3, 3), device="cuda"); torch.cuda.synchronize()
s = torch.cuda.current_stream().cuda_stream
for i in range(3):
    t0 = time.perf_counter(); t = capi.Tree.build_device(pts.data_ptr(), 512**3, 64, 1.0, stream=s); torch.cuda.synchronize()
    print(f"build {i}: {1e3*(time.perf_counter()-t0):.1f} ms wall", flush=True); t.close()
PY
echo "== cap4096 correctness"; NBK_LIBRARY=$V/libnbk_cap4096.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "build or tree or tiny" 2>&1 | tail -2
