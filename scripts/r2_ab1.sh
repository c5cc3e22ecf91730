V=nbodyhpc_b200/lib/variants
python scripts/kernel_ab.py
for lib in q16 q16b8 q12 unstable; do NBK_LIBRARY=$V/libnbk_$lib.so python scripts/kernel_ab.py; done
NBK_MORTON_FIRST_BIT=3 python scripts/kernel_ab.py --queries 12500000
NBK_MORTON_FIRST_BIT=0 python scripts/kernel_ab.py --queries 12500000
for t in 8 16 24; do echo "host threads $t"; NBK_HOST_THREADS=$t python scripts/e2e_python_probe.py 512 1e8 2>&1 | grep "query 1"; done
NBK_LIBRARY=$V/libnbk_q16.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "batched_query or periodic_wrap or clustered_points or odd_k or squared or fixture" 2>&1 | tail -3
