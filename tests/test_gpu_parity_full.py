"""Full-size training-step parity for the BASELINE.json configurations, through the reference-facing module surface.

Two comparisons per configuration, on the same synthetic inputs, parameters and replayed random draws:

  device vs bf16-emulating oracle (oracle/emu.py, DEVICE_POINTS)  - kernel exactness: the emulation rounds to bf16 exactly
        where the CUDA path stores bf16, so only fp32 accumulation order, ex2.approx and the handful of ReLU decisions that
        sit within fp32 noise of zero differ;
  device vs fp64 oracle                                           - the north star's bound: loss <= 1e-3, whole-model flat
        gradient <= 1e-2 relative (GCN; the GAT extension's bound is stated at its test).

The fp64 oracle at these sizes is the emulation with every rounding point switched off - tests/test_emulation.py pins that
to oracle/models.py (1e-9) - with the blockwise InfoNCE (the dense [N,2N] as-written form needs tens of GB here)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _inputs(n, e, m, seed=42):
    g = torch.Generator().manual_seed(seed)          # bench.py:synth
    if m > 1:
        x = torch.randn(n, m, 768, generator=g)
        x = x / x.norm(dim=1, keepdim=True)
    else:
        x = torch.nn.init.xavier_normal_(torch.empty(n, 768), generator=g)
    return x, torch.randint(0, n, (2, e), generator=g, dtype=torch.int64)


def _flat(named, keys):
    return torch.cat([named[k].detach().double().cpu().flatten() for k in keys])


def _run(n, e, m, fuse, enc):
    import biomedkg_b200 as b
    from biomedkg_b200.draws import ReplayDraws, set_draws
    from oracle import emu
    from oracle import models as om

    torch.set_num_threads(max(1, torch.get_num_threads()))
    x, ei = _inputs(n, e, m)
    torch.manual_seed(7)
    orc = om.GRACEModule(768, 256, 256, 2, fuse_method=fuse, encoder=enc).double().train()
    draws = om.TorchDraws(record=True)
    res = {}
    for name, points in (("fp64", frozenset()), ("emu", emu.DEVICE_POINTS)):
        for p in orc.parameters():
            p.grad = None
        d = draws if name == "fp64" else om.ReplayDraws(draws.log)
        loss = emu.grace_training_step(orc, x.double(), ei, d, points)
        loss.backward()
        res[name] = (float(loss), {k: p.grad.clone() for k, p in orc.named_parameters() if p.grad is not None})

    mod = b.GRACEModule(768, 256, 256, 2, fuse_method=fuse, encoder=enc)
    mod.load_state_dict({k: v.float() for k, v in orc.state_dict().items()})
    mod = mod.to(DEV).train()
    set_draws(mod, ReplayDraws(draws.log, DEV))

    class Batch:
        pass

    Batch.x, Batch.edge_index = x.to(DEV), ei.to(DEV)
    loss = mod.training_step(Batch)
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad for k, p in mod.named_parameters() if p.grad is not None}
    assert set(grads) == set(res["fp64"][1])
    big = max(float(v.norm()) for v in res["fp64"][1].values())
    keys = sorted(k for k, v in res["fp64"][1].items() if float(v.norm()) > 1e-9 * big)   # k_proj.bias: exactly-zero true gradient
    out = {"loss": float(loss)}
    for name in ("fp64", "emu"):
        f, f0 = _flat(grads, keys), _flat(res[name][1], keys)
        out[name] = (abs(float(loss) - res[name][0]) / abs(res[name][0]), float((f - f0).norm() / f0.norm()),
                     float((f * f0).sum() / (f.norm() * f0.norm())))
    fe, f0 = _flat(res["emu"][1], keys), _flat(res["fp64"][1], keys)
    out["format"] = float((fe - f0).norm() / f0.norm())     # emulation vs fp64: the cost of the storage format alone
    return out


def test_cfg1_full_size_step_parity():
    """BASELINE cfg 1: GRACE + 4-conv GCN on the drug subgraph shape (8 000 nodes, 2.67 M directed edges, 768-d features).
    A degree-334 random graph over-smooths completely: loss = ln(2N-1) and the gradient lives in 1e-3-relative deviations."""
    r = _run(8_000, 2_670_000, 1, None, "gcn")
    print("cfg1 parity:", r)
    assert abs(r["loss"] - math.log(2 * 8000 - 1)) < 1e-2
    assert r["fp64"][0] <= 1e-3 and r["emu"][0] <= 1e-5
    assert r["emu"][1] <= 3e-3, r            # kernel exactness
    assert r["fp64"][1] <= 1e-2, r           # north star: gradients within 1e-2 relative


def test_cfg2_shape_step_parity():
    """BASELINE cfg 2: GRACE + GAT + attention fusion of 2 modalities, 28 000 nodes / 650 000 edges.  GAT is this repository's
    extension (no reference symbol).  Measured on B200: device vs fp64 3.3e-2, bf16 emulation vs fp64 3.2e-2, device vs
    emulation 3.4e-2 - three mutually equidistant results, the signature of a discontinuity rather than of rounding: in the
    collapsed regime every attention logit a_src[j] + a_dst[i] is the same number plus 1e-3-relative deviations, and wherever
    that number sits near the LeakyReLU kink any perturbation (fp32 accumulation order included) flips the slope of a
    different set of edges.  GAT kernel exactness is therefore pinned at layer level, on identical bf16 inputs
    (tests/test_gpu_gat.py); this test bounds the full step: loss <= 1e-3, whole-model gradient cosine >= 0.999, error <= 5e-2."""
    r = _run(28_000, 650_000, 2, "attention", "gat")
    print("cfg2 parity:", r)
    assert r["fp64"][0] <= 1e-3 and r["emu"][0] <= 1e-5
    assert r["fp64"][1] <= 5e-2 and r["fp64"][2] >= 0.999, r
    assert r["emu"][1] <= 5e-2 and r["emu"][2] >= 0.999, r


def test_cfg4_like_gcn_attention_step_parity():
    """cfg 4's model (GRACE + GCN + 3-modality attention fusion) at 16 000 nodes / 1 M edges (the full 130k-node fp64 oracle
    does not finish in test time)."""
    r = _run(16_000, 1_000_000, 3, "attention", "gcn")
    print("cfg4-like parity:", r)
    assert r["fp64"][0] <= 1e-3 and r["emu"][1] <= 2e-3 and r["fp64"][1] <= 1e-2, r
