V=nbodyhpc_b200/lib/variants
for lib in b10 b12 t64b18 t64b20 t256b4 pf; do NBK_LIBRARY=$V/libnbk_$lib.so python scripts/kernel_ab.py; done
echo "numpy e2e with recycled result buffers"; python scripts/e2e_python_probe.py 512 1e8 2>&1 | grep "query 1\|build"
echo "cache off"; NBK_RESULT_CACHE_MB=0 python scripts/e2e_python_probe.py 512 1e8 2>&1 | grep "query 1"
