"""Host<->device copy bandwidth on this box (pinned / pageable), and timing of the host-pointer query."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")

def bw(fn, nbytes, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9

n = 1 << 30
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
pin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
page = torch.empty(n, dtype=torch.uint8)
print("pinned   H2D GB/s", bw(lambda: dev.copy_(pin, non_blocking=True), n))
print("pinned   D2H GB/s", bw(lambda: pin.copy_(dev, non_blocking=True), n))
print("pageable H2D GB/s", bw(lambda: dev.copy_(page), n))
print("pageable D2H GB/s", bw(lambda: page.copy_(dev), n))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
dev2 = torch.empty(n, dtype=torch.uint8, device="cuda"); pin2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
def both():
    with torch.cuda.stream(s1): dev.copy_(pin, non_blocking=True)
    with torch.cuda.stream(s2): pin2.copy_(dev2, non_blocking=True)
print("pinned bidirectional GB/s (sum)", bw(both, 2 * n))
