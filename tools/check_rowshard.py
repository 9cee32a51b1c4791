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

# ---- full row-sharded GRACE step (GCN encoder) vs the single-GPU module on the same data / seed ----
import biomedkg_b200 as b
from biomedkg_b200.dist import allreduce_grads, sharded_grace_loss

for (n, e, m, fuse) in ((3000, 40000, 1, "none"), (20000, 400000, 2, "attention")):
    g = torch.Generator().manual_seed(9)
    x = torch.randn(n, m, 768, generator=g) if m > 1 else torch.randn(n, 768, generator=g)
    if m > 1:
        x = x / x.norm(dim=1, keepdim=True)
    ei = torch.randint(0, n, (2, e), generator=g, dtype=torch.int64)
    torch.manual_seed(1)
    mod = b.GRACEModule(768, 256, 256, 2, fuse_method=fuse).cuda().train()
    xd, eid = x.cuda(), ei.cuda()

    class Batch:
        pass

    Batch.x, Batch.edge_index = xd, eid
    params = list(mod.model.parameters())
    res = []
    for sharded in (False, True):
        torch.manual_seed(77)                                   # same device RNG stream for the mask draws
        mod.model.encoder.draws._counter = 0                    # same dropout seeds
        for p in mod.parameters():
            p.grad = None
        if sharded:
            loss = sharded_grace_loss(mod, xd, eid)
            loss.backward()
            allreduce_grads(params)
        else:
            loss = mod.training_step(Batch)
            loss.backward()
        res.append((float(loss.detach()), torch.cat([p.grad.flatten() for p in params]).clone()))
    lerr = abs(res[0][0] - res[1][0]) / abs(res[0][0])
    gerr = float((res[0][1] - res[1][1]).norm() / res[0][1].norm())
    if rank == 0:
        print(f"GRACE step N={n} M={m} fuse={fuse} world={world}: single {res[0][0]:.6f} sharded {res[1][0]:.6f} "
              f"loss rel diff {lerr:.2e}, model-grad rel diff {gerr:.2e}")
    assert lerr < 1e-4 and gerr < 2e-2, (lerr, gerr)
dist.destroy_process_group()
