V=nbodyhpc_b200/lib/variants
for lib in wide9 wide8 wide7; do NBK_LIBRARY=$V/libnbk_$lib.so python scripts/kernel_ab.py; done
