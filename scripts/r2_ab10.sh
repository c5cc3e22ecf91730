V=nbodyhpc_b200/lib/variants
for lib in sw8 sw12 swmb4 swmb5 sw8mb6; do NBK_LIBRARY=$V/libnbk_$lib.so python scripts/kernel_ab.py --queries 100000000,12500000 --steps 3 | cut -c1-175; done
python scripts/kernel_ab.py --queries 100000000,12500000 --steps 3 | cut -c1-175
