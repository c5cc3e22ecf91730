NBK_LIBRARY=nbodyhpc_b200/lib/variants/libnbk_qnh.so python scripts/kernel_ab.py
python scripts/kernel_ab.py
