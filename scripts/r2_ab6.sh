python scripts/kernel_ab.py
NBK_SM_QUEUES=0 python scripts/kernel_ab.py
