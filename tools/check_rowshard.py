"""torchrun --nproc-per-node N tools/check_rowshard.py : row-sharded InfoNCE on N GPUs == single-GPU InfoNCE (loss and grads)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from biomedkg_b200 import ops
from biomedkg_b200.dist import sharded_infonce_loss

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for n in (1000, 28000):
    g = torch.Generator().manual_seed(3)
    h1c = torch.randn(n, 256, generator=g)
    h2c = h1c + torch.randn(n, 256, generator=g)
    h1, h2 = h1c.cuda().requires_grad_(True), h2c.cuda().requires_grad_(True)
    ref = ops.infonce_loss(h1, h2, 0.2)
    ref.backward()
    g1, g2 = h1.grad.clone(), h2.grad.clone()
    h1.grad = h2.grad = None
    loss = sharded_infonce_loss(h1, h2, 0.2)
    loss.backward()
    e = [abs(float(loss) - float(ref)) / abs(float(ref)), float((h1.grad - g1).norm() / g1.norm()), float((h2.grad - g2).norm() / g2.norm())]
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    dist.barrier(); t0.record()
    for _ in range(5):
        h1.grad = h2.grad = None
        sharded_infonce_loss(h1, h2, 0.2).backward()
    t1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"N={n} world={world}: loss rel diff {e[0]:.2e}, grad rel diff {e[1]:.2e} {e[2]:.2e}, sharded fwd+bwd {t0.elapsed_time(t1)/5:.3f} ms")
    assert max(e) < 1e-4, e
dist.destroy_process_group()
