for k in 4 16 32; do
python scripts/e2e_probe.py 1e8 $k
NBK_HOST_FIRST_SLICE=524288 NBK_HOST_SLICE=4194304 python scripts/e2e_probe.py 1e8 $k
NBK_HOST_FIRST_SLICE=2097152 NBK_HOST_SLICE=16777216 python scripts/e2e_probe.py 1e8 $k
done
python scripts/e2e_probe.py 1e8 8
NBK_HOST_FIRST_SLICE=524288 NBK_HOST_SLICE=2097152 python scripts/e2e_probe.py 1e8 8
NBK_HOST_FIRST_SLICE=1048576 NBK_HOST_SLICE=4194304 python scripts/e2e_probe.py 1e8 8
NBK_HOST_FIRST_SLICE=524288 NBK_HOST_SLICE=6291456 python scripts/e2e_probe.py 1e8 8
python scripts/e2e_python_probe.py 512 1e8 2>&1 | grep "query 1"
