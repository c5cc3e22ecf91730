for k in 64 32 16; do
python scripts/kernel_ab.py --queries 20000000 -k $k --steps 2
NBK_MAX_SHARED_K=8 python scripts/kernel_ab.py --queries 20000000 -k $k --steps 2
done
python scripts/all_kernels_probe.py
