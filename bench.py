#!/usr/bin/env python
"""bench.py - GCL nodes/s of one full-graph GRACE training step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg4] [--impl ours|reference] [--mode rowshard|dp]

A "step" = fusion + 3 encoder passes + projector + InfoNCE + backward + grad all-reduce (N>1) + grad clip + Adam,
exactly BaseGCL.training_step + the Lightning optimiser step of the reference (gcl_module.py:60-64, train_gcl.py:99).

Workload (every N): BASELINE.json configs[3], the PrimeKG++-scale full graph (130k nodes, 8M edges, 3-modality attention
fusion, GRACE + GCN).  N=1 runs it on one GPU; N>1 runs THE SAME graph row-sharded over the N ranks (nodes partitioned by
destination row, layer inputs / InfoNCE operand all-gathered over NVLink, parameter gradients all-reduced - the north star's
partition) -> strong scaling, value = nodes / max-over-ranks step time.  ``--mode dp`` keeps the reference's own DDP regime
(every rank its own graph, weak scaling) as an option.  At N=1 the line also carries, under "also", the device-resident
nodes/s of cfg2 (GAT + attention, CUDA-graph replay) and cfg1, and under roofline.also the CSR aggregation at cfg5 size.

Prints ONE JSON line (see the driver contract in DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIGS = {
    # BASELINE.json configs, SURVEY.md section 8 table
    "cfg1": dict(N=8_000, E=2_670_000, M=1, fuse="none", encoder="gcn", desc="GRACE + 2-layer-hidden GCN, drug subgraph"),
    "cfg2": dict(N=28_000, E=650_000, M=2, fuse="attention", encoder="gat", desc="GRACE + GAT + attention fusion (2 modalities), gene/protein subgraph"),
    "cfg4": dict(N=130_000, E=8_000_000, M=3, fuse="attention", encoder="gcn", desc="PrimeKG++-scale full graph, 3-modality fusion, GRACE"),
    "cfg5": dict(N=1_000_000, E=50_000_000, M=1, fuse="none", encoder="gat", powerlaw=True,
                 desc="synthetic scaling sweep: 1M nodes, 50M power-law edges, hidden 256, GRACE + GAT"),
}
IN_DIM, HID, LAYERS, TAU = 768, 256, 2, 0.2


def synth(cfg, seed, pin=False):
    """Synthetic inputs of the named shape: LM-like features normalised over the modality axis
    (data/node.py:115-117), Erdos-Renyi-like int64 edge_index with duplicates and self-loops left in."""
    g = torch.Generator().manual_seed(seed)
    N, E, M = cfg["N"], cfg["E"], cfg["M"]
    if M > 1:
        x = torch.randn(N, M, IN_DIM, generator=g)
        x = x / x.norm(dim=1, keepdim=True)
    else:
        x = torch.nn.init.xavier_normal_(torch.empty(N, IN_DIM), generator=g)
    if cfg.get("powerlaw"):   # destinations ~ Pareto(alpha = 2.1) degree sequence, sources uniform
        w = (1.0 - torch.rand(N, generator=g, dtype=torch.float64)).pow(-1.0 / 1.1)
        cdf = torch.cumsum(w / w.sum(), 0)
        dst = torch.searchsorted(cdf, torch.rand(E, generator=g, dtype=torch.float64)).clamp_(max=N - 1)
        ei = torch.stack([torch.randint(0, N, (E,), generator=g, dtype=torch.int64), dst])
    else:
        ei = torch.randint(0, N, (2, E), generator=g, dtype=torch.int64)
    if pin:   # pinned on the NUMA node local to this rank's GPU when the topology is known (biomedkg_b200/hostmem.py)
        from biomedkg_b200.hostmem import pinned_near

        d = torch.cuda.current_device()
        x, ei = pinned_near(x, d), pinned_near(ei, d)
    return x, ei


def config_dict(name, world, mode):
    """The ``config`` object of the JSON line - shared by both arms so the driver compares like with like."""
    cfg = CONFIGS[name]
    if world == 1:
        par = "single GPU"
    elif mode == "rowshard":
        par = (f"rowshard{world}: ONE graph, nodes partitioned by destination row over {world} ranks (fusion / GEMMs / aggregation / projector "
               f"on the rank's rows, NCCL all-gather of layer inputs and of the InfoNCE operand, InfoNCE rows of the rank's own nodes, "
               f"parameter gradients all-reduced)" if cfg["encoder"] == "gcn" else
               f"rowshard{world}: ONE graph, replicated GAT encoder, InfoNCE rows split over {world} ranks")
    else:
        par = f"dp{world}: every rank its own graph, NCCL gradient all-reduce (the reference's DDP regime)"
    return {"workload": f"{name}: {cfg['desc']}", "nodes": cfg["N"] * (world if mode == "dp" else 1), "edges": cfg["E"] * (world if mode == "dp" else 1),
            "modalities": cfg["M"], "fuse": cfg["fuse"], "encoder": cfg["encoder"], "objective": "GRACE (InfoNCE L2L, intraview negatives, tau 0.2)",
            "in_dim": IN_DIM, "hidden": HID, "conv_layers": LAYERS + 2, "parallelism": par,
            "unused_view": "computed (faithful to model/gcl.py:44)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, *windows):
        """Summarise the samples that arrived inside the given (t0, t1) wall-clock windows (the timed regions)."""
        if self.proc:
            self.proc.terminate()
        rows = [r for ts, r in self.rows if len(r) >= 9 and (not windows or any(t0 <= ts <= t1 + 0.05 for t0, t1 in windows))]
        sm = sorted(int(float(r[1])) for r in rows if r[1].replace(".", "").isdigit())
        mx = max((int(float(r[2])) for r in rows if r[2].replace(".", "").isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the restated PyG/PyGCL path ("as written"), timed on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_rate(cfg, n_nodes, steps, warmup, seed=42):
    """nodes/s of the oracle's as-written GRACE step (COO gather + scatter_add_, mask-materialising [N,2N] InfoNCE -
    the operations PyG 2.5.3 / PyGCL 0.1.2 execute).  n_nodes == cfg["N"]: the configuration in full; smaller: a bounded
    sample, an n_nodes-node graph of the same shape (same average degree, modalities, encoder, objective).  fp32, all host threads."""
    from oracle import models as om

    torch.set_num_threads(os.cpu_count() or 1)
    sub = dict(cfg)
    sub["N"] = n_nodes
    sub["E"] = max(1, int(cfg["E"] * n_nodes / cfg["N"]))
    x, ei = synth(sub, seed)
    torch.manual_seed(seed)
    mod = om.GRACEModule(IN_DIM, HID, HID, LAYERS, fuse_method=cfg["fuse"], encoder=cfg["encoder"]).train()
    opt = torch.optim.Adam(mod.model.parameters(), lr=1e-3)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = mod.training_step(x, ei)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(mod.model.parameters(), 1.0)
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n_nodes / dt, dt, sub


def _mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1e6
    except Exception:  # noqa: BLE001
        pass
    return 0.0


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (restated: PyG / PyGCL are not installable) on all host cores.
    Headline: args.config with --steps / --warmup honoured, every step a bounded sample of the workload (the as-written
    [N,2N] fp32 InfoNCE needs ~7 x 135 GB at cfg4's 130k nodes, it cannot run in full).  "also": the configurations the host
    CAN run in full - cfg1 always, cfg2 when >= 100 GB of RAM are available - so that same-config pairs exist next to the
    GPU arm's "also" entries."""
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    n_sample = min(args.cpu_sample_nodes, cfg["N"])
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    rate, dt, sub = cpu_reference_step_rate(cfg, n_sample, steps, warmup)
    sample = (f"restated reference path (PyG/PyGCL not installable): oracle as-written GRACE step, fp32, {cores} threads; each step a "
              f"{sub['N']}-node / {sub['E']}-edge sample of {args.config} (same degree, M={cfg['M']}, fuse={cfg['fuse']}, encoder={cfg['encoder']}); "
              f"{steps} steps after {warmup} warm-up, {dt:.2f} s/step.  The full configuration cannot run as written "
              f"([N,2N] fp32 temporaries of {4 * 2 * cfg['N'] ** 2 / 1e9:.0f} GB each); per-node cost grows with N, so the sample favours the CPU")
    also = {}
    if not args.no_also:
        t0 = time.time()
        r1, d1, _ = cpu_reference_step_rate(CONFIGS["cfg1"], CONFIGS["cfg1"]["N"], 3, 1)
        also["cfg1_full"] = {"value": r1, "unit": "nodes/s", "s_per_step": d1, "steps": 3, "warmup": 1, "nodes": CONFIGS["cfg1"]["N"],
                             "edges": CONFIGS["cfg1"]["E"], "same_config_as": "also.cfg1 of the GPU arm"}
        avail = _mem_available_gb()
        if avail >= 200.0 and time.time() - t0 < 120:   # the as-written step peaks near 100 GB at 28k nodes: leave a wide margin, an OOM-killed box helps nobody
            try:
                r2, d2, _ = cpu_reference_step_rate(CONFIGS["cfg2"], CONFIGS["cfg2"]["N"], 1, 0)
                also["cfg2_full"] = {"value": r2, "unit": "nodes/s", "s_per_step": d2, "steps": 1, "warmup": 0, "nodes": CONFIGS["cfg2"]["N"],
                                     "edges": CONFIGS["cfg2"]["E"], "same_config_as": "also.cfg2 of the GPU arm"}
            except (RuntimeError, MemoryError) as exc:
                also["cfg2_full"] = {"unavailable": f"reference OOM ({type(exc).__name__})"}
        else:
            also["cfg2_full"] = {"unavailable": f"reference OOM: the as-written loss needs ~100 GB at 28k nodes, MemAvailable = {avail:.0f} GB (run only with >= 200 GB)"}
    line = {
        "impl": "reference", "metric": "GCL nodes/sec", "value": rate, "unit": "nodes/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong" if args.mode == "rowshard" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args.config, max(1, args.gpus), args.mode),
        "cpu_baseline": {"value": rate, "unit": "nodes/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "also": also,
    }
    _emit(line)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Batch:
    pass


def _make_module(cfg, dev):
    import biomedkg_b200 as b

    torch.manual_seed(42)
    return b.GRACEModule(IN_DIM, HID, HID, LAYERS, scheduler_type="cosine", learning_rate=1e-3, warm_up_ratio=0.2,
                         fuse_method=cfg["fuse"], encoder=cfg["encoder"]).to(dev).train()


def side_workload(name, dev, steps, warmup, use_graph=True, peak_tf=None):
    """Device-resident nodes/s of another BASELINE configuration on this GPU (N=1 'also' entries): same step, same timing rules."""
    from biomedkg_b200.graphed import GraphedStep

    cfg = CONFIGS[name]
    x_host, ei_host = synth(cfg, 42)
    mod = _make_module(cfg, dev)
    params = list(mod.model.parameters())
    opt = torch.optim.Adam(params, lr=1e-3)
    bt = Batch()
    bt.x, bt.edge_index = x_host.to(dev), ei_host.to(dev)

    def eager():
        opt.zero_grad(set_to_none=True)
        loss = mod.training_step(bt)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        return loss

    for _ in range(warmup):
        eager()
    run, graphed = eager, False
    if use_graph:
        try:
            opt.zero_grad(set_to_none=True)
            gs = GraphedStep(mod, bt.x, bt.edge_index, resort=False)

            def run():
                loss = gs()
                torch.nn.utils.clip_grad_norm_(params, 1.0)
                opt.step()
                return loss

            graphed = True
            for _ in range(warmup):
                run()
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] {name}: CUDA-graph capture failed ({type(exc).__name__}: {exc}); eager", file=sys.stderr)
            run = eager
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"value": cfg["N"] / (ms * 1e-3), "unit": "nodes/s", "ms_per_step": ms, "steps": steps, "warmup": warmup, "nodes": cfg["N"],
           "edges": cfg["E"], "cuda_graph": graphed, "final_loss": float(loss.detach()), "workload": f"{name}: {cfg['desc']}"}
    # InfoNCE kernels at this size, timed with CUDA events in three eager steps (events cannot bracket kernels inside a replay)
    from biomedkg_b200 import _cabi

    _cabi.timed_entries.update({"bmkg_infonce_fwd", "bmkg_infonce_bwd"})
    _cabi.timings.clear()
    for _ in range(3):
        eager()
    torch.cuda.synchronize()
    km = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in _cabi.timings.items()}
    _cabi.timed_entries.clear()
    _cabi.timings.clear()
    if "bmkg_infonce_bwd" in km and peak_tf:
        n2d = float(cfg["N"]) ** 2 * HID
        out["infonce"] = {"bwd_ms": km["bmkg_infonce_bwd"], "bwd_frac_of_sustained_bf16_peak (8N^2D credited)": 8.0 * n2d / (km["bmkg_infonce_bwd"] * 1e-3) / 1e12 / peak_tf,
                          "fwd_ms": km.get("bmkg_infonce_fwd"), "fwd_frac (6N^2D credited)": 6.0 * n2d / (km["bmkg_infonce_fwd"] * 1e-3) / 1e12 / peak_tf}
    del mod, opt, bt
    torch.cuda.empty_cache()
    return out


def aggregation_roofline(dev, peaks):
    """bmkg_gcn_aggregate at BASELINE cfg5 size (1M nodes, 50M power-law edges, C=256, one 40%-dropped view): the honest HBM
    test (512 MB of bf16 rows, far beyond L2).  Algorithmic bytes per SURVEY.md 8d: E'(C s + 4) + N C s + 4 (N + 1)."""
    from biomedkg_b200 import ops

    cfg = CONFIGS["cfg5"]
    N, C = cfg["N"], HID
    _, ei = synth(dict(cfg, M=1), 42)
    ei = ei.to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    keep = torch.rand(ei.size(1), device=dev, generator=g) >= 0.4
    sg = ops.sorted_graph(ei, N, cache=False)
    view = sg.view(keep)
    x = torch.randn(N, C, device=dev, generator=g).to(torch.bfloat16)
    bias = torch.zeros(C, device=dev)
    nnz = int(view.nnz.item())
    hub = view.hub_csr if view.hub_possible else None

    def call_fwd():
        return ops.gcn_aggregate(view.rowptr, view.colind, view.dis, x, bias, True, 0.2, 1234, None, False, hub_rows=hub)

    def call_bwd():
        return ops.gcn_aggregate(view.csc_rowptr, view.csc_colind, view.dis, x, hub_rows=view.hub_csc if view.hub_possible else None)

    out = {}
    for name, fn in (("forward (CSR, fused norm+bias+ReLU+dropout)", call_fwd), ("transposed backward (CSC)", call_bwd)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        alg = nnz * (C * 2 + 4) + N * C * 2 + 4 * (N + 1)
        gbs = alg / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "algorithmic_bytes": alg, "achieved_gbs": gbs, "frac_of_hbm": gbs / peaks.get("hbm_gbs", 6538.3)}
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        traffic = tj.get("gcn_aggregate_kernel")
    except Exception:  # noqa: BLE001
        pass
    res = {"kernel": "bmkg_gcn_aggregate", "bound": "hbm", "nodes": N, "edges_in_view_incl_self_loops": nnz, "channels": C,
           "peak_gbs": peaks.get("hbm_gbs", 6538.3), "calls": out,
           "ncu_dram_bytes": traffic or "profiles/r1_ncu_gcn_aggregate_N1M_E50M.txt: 14.2 GB per launch at 1M nodes / 31M kept edges (ER graph)",
           "l2_policy": "10 back-to-back calls over 512 MB of rows + 250 MB of indices (> 126 MB L2)"}
    del sg, view, x, ei, keep
    torch.cuda.empty_cache()
    return res


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist

    from biomedkg_b200 import _cabi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cfg = CONFIGS[args.config]
    N, E, M = cfg["N"], cfg["E"], cfg["M"]
    rowshard = args.mode == "rowshard" and world > 1
    # rowshard: ONE graph on all ranks (strong scaling, full-graph loss - SURVEY.md 8e); dp: every rank its own graph (weak)
    x_host, ei_host = synth(cfg, 42 if (rowshard or world == 1) else 42 + rank, pin=True)
    mod = _make_module(cfg, dev)
    full_shard = rowshard and cfg["encoder"] == "gcn"     # GCN: encoder rows sharded too; GAT: encoder replicated, InfoNCE sharded
    if rowshard and not full_shard:
        from biomedkg_b200.dist import ShardedDualBranchContrast

        mod.contrast_model = ShardedDualBranchContrast(tau=TAU)
    params = [p for p in mod.model.parameters()]
    opt = torch.optim.Adam(params, lr=1e-3)

    # CUDA-graph replay pays off when the step is launch-bound (cfg1 / cfg2: a few ms of 5-50 us kernels).  At cfg4 / cfg5 the step
    # is ~100 ms of long kernels, and a captured graph would pin the InfoNCE E store (up to 135 GB) in a private memory pool per
    # capture - so those run eagerly, where torch's caching allocator hands the same block back every step.
    # Row-sharded steps are short again (cfg4 on 8 GPUs: ~19 ms for ~350 launches per rank, 8 Python processes on the host's
    # cores): issued eagerly the ranks drift apart and wait for each other inside the all-gathers, so the step - NCCL collectives
    # included - is captured and replayed as well.
    use_graph = args.graph == "on" or (args.graph == "auto" and (N <= 50_000 or rowshard))
    shard_loss = None
    if full_shard:
        from biomedkg_b200.dist import sharded_grace_loss as _sgl

        shard_loss = lambda m, bt: _sgl(m, bt.x, bt.edge_index, num_nodes=N)  # noqa: E731

    def all_ranks_ok(ok):
        """a capture that failed on ANY rank must send every rank back to eager launches (mismatched collectives would hang)"""
        if world == 1:
            return ok
        t = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)
    graphed = {}      # "resident" / "e2e" -> GraphedStep (captured lazily, after the eager warm-up)

    def tail():
        if world > 1 and not rowshard:  # DDP-equivalent: average parameter gradients over ranks (NCCL all-reduce, ~2 MB)
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
            flat /= world
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
                off += p.numel()
        torch.nn.utils.clip_grad_norm_(params, 1.0)   # gradient_clip_val=1.0 (train_gcl.py:99)
        opt.step()

    def graph_step(kind, batch):
        """forward + backward replayed from a CUDA graph (biomedkg_b200/graphed.py); all-reduce / clip / Adam eager."""
        gs = graphed[kind]
        loss = gs(batch.x, batch.edge_index) if kind == "e2e" else gs()
        if full_shard:
            from biomedkg_b200.dist import allreduce_grads

            allreduce_grads(params)
        tail()
        return loss

    def step(batch):
        opt.zero_grad(set_to_none=True)
        if full_shard:
            from biomedkg_b200.dist import allreduce_grads, sharded_grace_loss

            loss = sharded_grace_loss(mod, batch.x, batch.edge_index, num_nodes=getattr(batch, "num_nodes", None))
            loss.backward()
            allreduce_grads(params)                 # per-rank partial sums -> full gradient
        else:
            loss = mod.training_step(batch)
            loss.backward()
        tail()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ("value") ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0 and not args.no_clocks:
        sampler.start()                      # started before warm-up so it is already streaming in the timed region
    res = Batch()
    if full_shard:   # every rank holds only its node block of the features (the whole edge_index stays replicated)
        from biomedkg_b200.dist import shard_layout as _sl

        _b0, _b1 = _sl(N, world)[1][rank]
        res.x, res.num_nodes = x_host[_b0:_b1].to(dev), N
    else:
        res.x = x_host.to(dev)
    res.edge_index = ei_host.to(dev)
    for _ in range(args.warmup):
        step(res)
    barrier()
    run = step
    if use_graph:
        from biomedkg_b200.graphed import GraphedStep

        opt.zero_grad(set_to_none=True)
        ok = True
        try:
            graphed["resident"] = GraphedStep(mod, res.x, res.edge_index, resort=False, loss_fn=shard_loss)   # edge list fixed: sorted once, as in the eager loop
        except Exception as exc:  # noqa: BLE001 - a failed capture must not cost the measurement: fall back to eager launches
            print(f"[bench] CUDA-graph capture failed ({type(exc).__name__}: {exc}); running eagerly", file=sys.stderr)
            ok = False
            torch.cuda.synchronize()
        if not all_ranks_ok(ok):
            use_graph = False
            graphed.pop("resident", None)
    timed = {"bmkg_infonce_fwd", "bmkg_infonce_bwd", "bmkg_infonce_fwd_rows", "bmkg_infonce_bwd_rows", "bmkg_gcn_aggregate_rows",
             "bmkg_gat_aggregate", "bmkg_gat_aggregate_bwd"}
    if use_graph:
        run = lambda bt: graph_step("resident", bt)  # noqa: E731
        for _ in range(args.warmup):
            run(res)
        barrier()
        # events cannot bracket kernels inside a replayed graph: per-kernel durations come from eager steps of the same
        # training step, run right here (same clocks, same data), not from the replayed region
        _cabi.timed_entries.update(timed)
        _cabi.timings.clear()
        for _ in range(3):
            step(res)
        barrier()
        kern_ms = {k: sum(a.elapsed_time(bb) for a, bb in v) / len(v) for k, v in _cabi.timings.items()}
        kern_calls = {k: len(v) // 3 for k, v in _cabi.timings.items()}
        _cabi.timed_entries.clear()
        _cabi.timings.clear()
        opt.zero_grad(set_to_none=True)
        for _ in range(2):      # back to the replayed step (GraphedStep re-attaches its captured gradient buffers)
            run(res)
        barrier()
    else:
        _cabi.timed_entries.update(timed)
        _cabi.timings.clear()
    launches0 = _cabi.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        loss = run(res)
    e1.record()
    barrier()
    wall1 = time.time()
    ms = e0.elapsed_time(e1) / args.steps
    if use_graph:
        launches = graphed["resident"].launches_per_replay      # kernels of this package inside one replay of the captured step
    else:
        launches = (_cabi.kernel_launches - launches0) // args.steps
        kern_ms = {k: sum(a.elapsed_time(bb) for a, bb in v) / len(v) for k, v in _cabi.timings.items()}
        kern_calls = {k: len(v) // args.steps for k, v in _cabi.timings.items()}
        _cabi.timed_entries.clear()
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    nodes_total = N if (rowshard or world == 1) else N * world
    value = nodes_total / (ms_max * 1e-3)
    final_loss = float(loss.detach())

    # ---------------- end to end: host batch -> loss on host ----------------
    # What a training loop over host-resident batches does (the reference's Lightning loop moves every batch host -> device):
    # every step copies ITS OWN x and edge_index from pinned host memory (a fresh edge_index tensor, so the radix sort
    # runs every step) and reads the loss back.  The copy of step k+1 is issued on a side stream while step k computes
    # (pinned-memory prefetch, as a DataLoader with pin_memory does); all copies are inside the timed region.
    k2_warm = 2 if (args.e2e_steps is None or args.e2e_steps > 1) else 1
    copy_stream = torch.cuda.Stream(device=dev)
    x_src, ei_src = x_host, ei_host
    if full_shard:   # a sharded loader: every rank reads its node block of the features and 1/world of the edge list from the host
        from biomedkg_b200.dist import edge_chunk, gather_edge_index, shard_layout
        from biomedkg_b200.hostmem import pinned_near

        b0, b1 = shard_layout(N, world)[1][rank]
        x_src = pinned_near(x_host[b0:b1].contiguous(), torch.cuda.current_device())
        per, c0, c1 = edge_chunk(E, world, rank)
        ei_src = torch.zeros(2, per, dtype=torch.int64)
        ei_src[:, : c1 - c0] = ei_host[:, c0:c1]
        ei_src = pinned_near(ei_src, torch.cuda.current_device())

    # Two device staging buffers, as a loader keeps them: the copy of step k+1 lands in the buffer step k-1 used (its
    # "consumed" event is waited for on the copy stream), so the timed loop allocates nothing - with a fresh 1.2 GB tensor per
    # step the caching allocator's cudaMalloc / cudaFree decided, run by run, whether a step cost 3 or 18 ms more.
    stage_x = [torch.empty(x_src.shape, dtype=x_src.dtype, device=dev) for _ in range(2)]
    stage_ei = [torch.empty(ei_src.shape, dtype=ei_src.dtype, device=dev) for _ in range(2)]
    consumed = [None, None]
    fetched = [0]

    def prefetch():
        slot = fetched[0] & 1
        fetched[0] += 1
        with torch.cuda.stream(copy_stream):
            if consumed[slot] is not None:
                copy_stream.wait_event(consumed[slot])
            bt = Batch()
            bt.slot = slot
            bt.x = stage_x[slot]
            bt.x.copy_(x_src, non_blocking=True)
            if full_shard:
                bt.num_nodes = N
            bt.edge_index = stage_ei[slot]
            bt.edge_index.copy_(ei_src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return bt, ev

    e2e_graph = use_graph
    if e2e_graph:
        graphed.pop("resident", None)          # release its memory pool before the second capture
        opt.zero_grad(set_to_none=True)
        ok = True
        try:
            graphed["e2e"] = GraphedStep(mod, res.x, res.edge_index, resort=True, loss_fn=shard_loss)    # every step brings its own edge_index: sort captured too
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] CUDA-graph capture of the end-to-end step failed ({type(exc).__name__}: {exc}); running eagerly", file=sys.stderr)
            ok = False
            torch.cuda.synchronize()
        if not all_ranks_ok(ok):
            e2e_graph = False
            graphed.pop("e2e", None)
    run_e2e = (lambda bt: graph_step("e2e", bt)) if e2e_graph else step

    def e2e_loop(k):
        nxt = prefetch()
        losses_host = torch.empty(k, dtype=torch.float32).pin_memory()   # loss of every step lands here (async D2H per step)
        for i in range(k):
            bt, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            slot = bt.slot
            if full_shard:      # the edge-list chunks of all ranks -> the full int64 [2, E] edge_index, over NVLink
                bt.edge_index = gather_edge_index(bt.edge_index, E)
            if i + 1 < k:
                nxt = prefetch()
            losses_host[i:i + 1].copy_(run_e2e(bt).detach().reshape(1), non_blocking=True)   # device -> host read of the loss, every step
            consumed[slot] = torch.cuda.Event()
            consumed[slot].record(torch.cuda.current_stream())     # the staging buffers of this step may be overwritten
        torch.cuda.current_stream().synchronize()
        return float(losses_host[-1])

    e2e_loop(max(1, min(2, k2_warm)))
    barrier()
    k2 = max(1, args.e2e_steps if args.e2e_steps is not None else args.steps)
    wall2 = time.time()
    e0.record()
    e2e_loop(k2)
    e1.record()
    barrier()
    wall3 = time.time()
    t = torch.tensor([e0.elapsed_time(e1) / k2], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    h2d = x_src.numel() * 4 + ei_src.numel() * 8
    ht = torch.tensor([float(h2d)], device=dev)
    if world > 1:
        dist.all_reduce(ht)                                  # bytes copied by all ranks per step
    e2e = {"value": nodes_total / (e2e_ms * 1e-3), "unit": "nodes/s", "h2d_bytes_per_step": int(ht.item()),
           "d2h_bytes_per_step": 4 * world, "ms_per_step": e2e_ms,
           "note": "pinned-host batch copied every step (prefetched one step ahead on a copy stream; row-sharded ranks copy their own node block of "
                   "x and 1/N of the int64 edge_index, which is then all-gathered over NVLink), edge_index re-sorted every step, loss copied to pinned "
                   "host memory every step (async), one sync at the end"
                   + ("; forward+backward (incl. the sort) replayed from a CUDA graph over static input buffers" if e2e_graph else "")}

    clocks = sampler.stop((wall0, wall1), (wall2, wall3)) if rank == 0 else None    # both timed regions (device-resident and end-to-end)
    if rank != 0:
        return
    # ---------------- roofline of the dominant kernel ----------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)" if peaks else "fallback 1.4 PF sustained (B200_PROFILING.md)"
    D = HID
    flops_bwd, flops_fwd = 8.0 * N * N * D, 6.0 * N * N * D
    roof = None
    for k in ("bmkg_infonce_bwd", "bmkg_infonce_fwd"):   # the row-sharded path launches the *_rows entry points
        if k not in kern_ms and k + "_rows" in kern_ms:
            kern_ms[k] = kern_ms[k + "_rows"] * world       # per-rank time x ranks = single-GPU-equivalent duration
    if "bmkg_infonce_bwd" in kern_ms:
        ach = flops_bwd / (kern_ms["bmkg_infonce_bwd"] * 1e-3) / 1e12
        traffic = None   # DRAM bytes per launch from the committed ncu --set full capture, only if it was taken at this N
        try:
            for fn in ("r2_traffic.json", "r1_traffic.json"):
                path = os.path.join(ROOT, "profiles", fn)
                if os.path.exists(path):
                    allj = json.load(open(path))
                    for tj in (allj.get(f"infonce_bwd_kernel_N{N}"), allj.get("infonce_bwd_kernel")):
                        if tj and tj.get("N") == N and world == 1:     # a capture of the full-range launch at this size
                            traffic = int(tj["dram_bytes_read"] + tj["dram_bytes_write"])
                            break
                    if traffic is not None:
                        break
        except Exception:  # noqa: BLE001
            pass
        roof = {"kernel": "infonce_bwd_kernel (tcgen05)", "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": ach / peak_tf, "traffic": traffic, "algorithmic_flops_per_launch": flops_bwd / (world if rowshard else 1),
                "ms_per_launch": kern_ms["bmkg_infonce_bwd"] / (world if rowshard else 1),
                "peak_source": peak_src,
                # SURVEY 8(d) credits only the four dZ GEMMs (8N^2D); the kernel also recomputes S on the tensor pipe (another
                # 8N^2D with the 2Nx2N Gram form), so the tcgen05 pipe itself runs at twice the credited rate
                "executed_tflops": 2.0 * ach,
                "also": {"infonce_fwd (3 kernels, 6N^2D)": {"ms": kern_ms.get("bmkg_infonce_fwd"),
                                                         "achieved_tflops": flops_fwd / (kern_ms["bmkg_infonce_fwd"] * 1e-3) / 1e12},
                         "infonce fwd+bwd (14N^2D)": {"achieved_tflops": (flops_fwd + flops_bwd) / ((kern_ms["bmkg_infonce_fwd"] + kern_ms["bmkg_infonce_bwd"]) * 1e-3) / 1e12,
                                                      "frac": (flops_fwd + flops_bwd) / ((kern_ms["bmkg_infonce_fwd"] + kern_ms["bmkg_infonce_bwd"]) * 1e-3) / 1e12 / peak_tf}}}
        agg_key = "bmkg_gat_aggregate" if cfg["encoder"] == "gat" else "bmkg_gcn_aggregate_rows"
        if agg_key in kern_ms:
            roof["also"][agg_key + " (in-step, L2-resident rows at this size)"] = {"ms_avg_per_call": kern_ms[agg_key], "calls_per_step": kern_calls[agg_key]}
        if "bmkg_gat_aggregate_bwd" in kern_ms:
            roof["also"]["bmkg_gat_aggregate_bwd"] = {"ms_avg_per_call": kern_ms["bmkg_gat_aggregate_bwd"], "calls_per_step": kern_calls["bmkg_gat_aggregate_bwd"]}

    # ---------------- N=1 extras: other configurations, aggregation roofline, CPU baseline ----------------
    also, cpu = {}, None
    if world == 1 and not args.no_also:
        from biomedkg_b200 import ops as _ops

        res.x = res.edge_index = None
        graphed.clear()
        _ops.drop_e_store_pool()               # the headline workload's E store (up to 126 GiB) goes back to the device
        torch.cuda.empty_cache()
        for name in ("cfg2", "cfg1"):
            if name != args.config:
                try:
                    also[name] = side_workload(name, dev, max(3, min(args.steps, 20)), 3, peak_tf=peak_tf)
                except Exception as exc:  # noqa: BLE001
                    also[name] = {"unavailable": f"{type(exc).__name__}: {exc}"}
        try:
            if roof is not None:
                roof["also"]["aggregation at cfg5 size (HBM)"] = aggregation_roofline(dev, peaks)
        except Exception as exc:  # noqa: BLE001
            roof["also"]["aggregation at cfg5 size (HBM)"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
    if world == 1 and not args.no_cpu_baseline:
        n_s = min(args.cpu_sample_nodes, N)
        rate, dt, sub = cpu_reference_step_rate(cfg, n_s, 2, 1)
        cpu = {"value": rate, "unit": "nodes/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"oracle as-written GRACE step (restated PyG/PyGCL path), {sub['N']}-node / {sub['E']}-edge sample of {args.config}, fp32, 2 steps, {dt:.2f} s/step"}

    config = config_dict(args.config, world, args.mode)
    line = {
        "metric": "GCL nodes/sec", "value": value, "unit": "nodes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak" if (world > 1 and not rowshard) else "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": config,
        "run": {"cuda_graph": bool(use_graph), "tau": TAU,
                "l2_policy": "inputs larger than L2 (x is %.0f MB fp32, Z is %.0f MB bf16); no explicit flush" % (x_host.numel() * 4 / 1e6, 2 * N * D * 2 / 1e6)},
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "final_loss": final_loss,
        "also": also,
    }
    _emit(line)


_REAL_STDOUT = None


def _protect_stdout():
    """The contract is ONE JSON line on stdout: NCCL / c10d print banners to fd 1, so point fd 1 at stderr for the run and
    keep a private duplicate for the final line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg4", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-sample-nodes", type=int, default=8000, help="nodes of the bounded CPU sample (reference arm / cpu_baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the extra N=1 measurements (cfg2 / cfg1 / aggregation roofline; reference arm: cfg1 / cfg2 in full)")
    ap.add_argument("--mode", default="rowshard", choices=["rowshard", "dp"], help="multi-GPU mode (N>1): one row-sharded graph (strong) or per-rank graphs (weak)")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi during the run (A/B of its overhead)")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay forward+backward from a CUDA graph (auto: on unless --mode rowshard with N>1)")
    ap.add_argument("--e2e-steps", type=int, default=None, help="steps of the end-to-end loop (default: --steps)")
    args = ap.parse_args()
    if args.impl == "ours" and args.config != "cfg5":      # W >= 3 (timing rules); cfg5 steps take seconds, 1 warm-up step is enough there
        args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _protect_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
