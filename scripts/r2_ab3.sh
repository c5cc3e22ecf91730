V=nbodyhpc_b200/lib/variants
for lib in merged mergedb8; do NBK_LIBRARY=$V/libnbk_$lib.so python scripts/kernel_ab.py; done
echo "== numpy e2e, 16 MB slots, up to 16 threads"; python scripts/e2e_python_probe.py 512 1e8 2>&1 | grep "query 1"
echo "== same, 8 threads"; NBK_HOST_THREADS=8 python scripts/e2e_python_probe.py 512 1e8 2>&1 | grep "query 1"
echo "== cold start"
NBK_BUILD_TRACE=1 python - <<'PY' 2>&1 | grep -v "^$" | head -24
import sys, time
sys.path.insert(0, ".")
import torch
from nbodyhpc_b200 import capi
pts = torch.rand((512**3, 3), device="cuda"); torch.cuda.synchronize()
s = torch.cuda.current_stream().cuda_stream
for i in range(2):
    t0 = time.perf_counter(); t = capi.Tree.build_device(pts.data_ptr(), 512**3, 64, 1.0, stream=s); torch.cuda.synchronize()
    print(f"build {i}: {1e3*(time.perf_counter()-t0):.1f} ms wall", flush=True); t.close()
PY
echo "== config 4"; python bench.py --config 4 --steps 3 --warmup 2 > gpurun_out/r2_config4.json 2> gpurun_out/r2_config4.err; tail -c 400 gpurun_out/r2_config4.err; head -c 600 gpurun_out/r2_config4.json
