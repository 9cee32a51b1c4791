"""SM-clock timeline of the InfoNCE backward's pipeline (CTA 0, first row block) from a tuning build:
    BMKG_NVCC_DEFS="-DBMKG_BWD_TRACE" BMKG_LIB_OUT=$PWD/biomedkg_b200/_lib/libbmkg_trace.so python biomedkg_b200/build.py --force
    BMKG_LIB_PATH=$PWD/biomedkg_b200/_lib/libbmkg_trace.so python tools/trace_bwd.py 28000
Prints, per tile and relative to the tile's own "S ready", when each stage happened, and the steady-state tile period.
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from biomedkg_b200 import _cabi, ops

n = int(sys.argv[1])
torch.manual_seed(0)
h1 = torch.randn(n, 256, device="cuda", requires_grad=True)
h2 = (h1.detach() + torch.randn(n, 256, device="cuda")).requires_grad_(True)
for _ in range(2):
    ops.infonce_loss(h1, h2, 0.2).backward()
torch.cuda.synchronize()
lib = ctypes.CDLL(_cabi.LIB_PATH)
buf = (ctypes.c_longlong * (64 * 24))()
assert lib.bmkg_debug_bwd_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(64, 24)
t0 = t[0, 0]
names = ["m1_wait", "m1_tile_in", "m1_issued", "m2_wait", "p0", "p1", "p2", "p3", "m2_issued",
         "A.wait_s", "A.s_ready", "A.ld_done", "A.calc_done", "A.pub0", "A.pub_last",
         "Z.wait_s", "Z.s_ready", "Z.ld_done", "Z.calc_done", "Z.pub0", "Z.pub_last"]
print("tile " + " ".join(f"{x:>10s}" for x in names))
for i in range(8, 40):
    print(f"{i:4d} " + " ".join(f"{int(v - t0):10d}" if v else f"{'-':>10s}" for v in t[i, :len(names)]))
print("steady-state period per tile (MMA1 issue to MMA1 issue, tiles 16..56):", (t[56, 2] - t[16, 2]) / 40.0)
sl = slice(16, 60)


def d(a, b):   # mean over the tiles this warp served (tile-parity builds leave the other set's slots empty)
    m = (t[sl, a] != 0) & (t[sl, b] != 0)
    return float(np.mean((t[sl, a] - t[sl, b])[m])) if m.any() else float("nan")


print(f"softmax A: s_ready->ld {d(11,10):.0f}, ld->calc {d(12,11):.0f}, calc->published {d(14,12):.0f}; total {d(14,10):.0f}")
print(f"softmax Z: s_ready->ld {d(17,16):.0f}, ld->calc {d(18,17):.0f}, calc->published {d(20,18):.0f}; total {d(20,16):.0f}")
print(f"softmax A idle before S ready (wait_s -> s_ready): {d(10,9):.0f}")
print(f"MMA warp: wait for P0 {d(4,3):.0f}, P0->P3 {d(7,4):.0f}, wait for tile {d(1,0):.0f}")
