python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "recycled or return_squared or device_arrays" 2>&1 | tail -3
python scripts/e2e_python_probe.py 512 1e8 2>&1 | grep "query 1"
echo "== pin off"; NBK_RESULT_PIN=0 python scripts/e2e_python_probe.py 512 1e8 2>&1 | grep "query 1"
