V=nbodyhpc_b200/lib/variants
for lib in tilecs nodelast both img8; do NBK_LIBRARY=$V/libnbk_$lib.so python scripts/kernel_ab.py; done
python scripts/kernel_ab.py
