"""Pinned host -> device copy bandwidth with the pinned buffer allocated (a) wherever the process happens to run and (b) on the
NUMA node local to the GPU (bench.py pins its host batch the second way).  python tools/diag_h2d.py [device]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from biomedkg_b200.hostmem import gpu_local_cpus, pinned_near

dev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
torch.cuda.set_device(dev)
print("allowed cpus:", len(os.sched_getaffinity(0)), "gpu-local cpus:", None if gpu_local_cpus(dev) is None else len(gpu_local_cpus(dev)))
try:
    for n in sorted(os.listdir("/sys/devices/system/node")):
        if n.startswith("node"):
            print(n, open(f"/sys/devices/system/node/{n}/cpulist").read().strip())
except OSError as exc:
    print("no numa info:", exc)
src = torch.empty(300_000_000, dtype=torch.float32)      # 1.2 GB
src.fill_(1.0)
for name, buf in (("default", src.pin_memory()), ("gpu-local", pinned_near(src, dev))):
    dst = torch.empty_like(buf, device="cuda")
    for _ in range(2):
        dst.copy_(buf, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(5):
        dst.copy_(buf, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / 5
    print(f"{name:10s} pinned H2D: {buf.numel() * 4 / dt / 1e9:.1f} GB/s")
    del dst, buf
