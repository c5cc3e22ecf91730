python scripts/e2e_probe.py
NBK_HOST_FIRST_SLICE=524288 python scripts/e2e_probe.py
NBK_HOST_FIRST_SLICE=262144 python scripts/e2e_probe.py
NBK_HOST_SLICE=8388608 python scripts/e2e_probe.py
NBK_HOST_SLICE=8388608 NBK_HOST_FIRST_SLICE=524288 python scripts/e2e_probe.py
NBK_HOST_SLICE=33554432 python scripts/e2e_probe.py
NBK_HOST_SLICE=4194304 NBK_HOST_FIRST_SLICE=524288 python scripts/e2e_probe.py
