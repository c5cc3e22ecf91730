V=nbodyhpc_b200/lib/variants
NBK_LIBRARY=$V/libnbk_stage.so timeout 300 python scripts/kernel_ab.py
python scripts/kernel_ab.py
NBK_LIBRARY=$V/libnbk_stage.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batched_query or periodic_wrap or clustered_points or fixture or tiny or config1" 2>&1 | tail -2
