"""Turn gpurun_out/*.ncu-rep and the launch-list CSV into the small text summaries committed under profiles/.
    python tools/summarize_ncu.py
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

OUT = "profiles"
KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
]


def raw(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    return rows[0], rows[1], rows[2:]


def summarize_rep(rep, out, note=""):
    h, u, data = raw(rep)
    with open(out, "w") as f:
        f.write(f"# {os.path.basename(rep)} - ncu --set full --clock-control none (values per launch){note}\n")
        ik = h.index("Kernel Name")
        for r in data:
            f.write(f"\n## {r[ik][:150]}\n")
            for k in KEYS:
                if k in h:
                    i = h.index(k)
                    f.write(f"{k:75s} {r[i]:>16s} {u[i]}\n")
            try:
                rd, wr = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
                f.write(f"{'traffic = dram read + write (per launch)':75s} {r[rd]} {u[rd]} + {r[wr]} {u[wr]}\n")
            except ValueError:
                pass
    print("wrote", out)


def summarize_launches(csv_path, out, steps_in_file):
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("nsecond", "ns") else v / 1e3 if unit in ("usecond", "us") else v
        n = re.sub(r"\(.*", "", r["Kernel Name"])
        n = re.sub(r"^void ", "", n)[:90]
        tot[n] += ms
        cnt[n] += 1
    S = sum(tot.values())
    nb = sum(c for n, c in cnt.items() if "infonce_bwd_kernel" in n)
    if nb:
        steps_in_file = nb      # one InfoNCE backward per training step: eager warm-up, graph warm-up, replays, end-to-end
    with open(out, "w") as f:
        f.write(f"# launch list of `bench.py --steps 2 --warmup 3 --no-cpu-baseline` under ncu --metrics gpu__time_duration.sum --clock-control none\n")
        f.write(f"# {len(rows)} launches over {steps_in_file} training steps (one InfoNCE backward each).  Under the profiler bench.py's CUDA-graph\n"
                f"# capture is not in play (the list is the eager sequence of the same kernels); per-launch times are cold-cache and\n")
        f.write(f"# serialised, so compare SHARES.  total {S:.2f} ms = {S / steps_in_file:.2f} ms/step, {len(rows) / steps_in_file:.0f} launches/step\n\n")
        f.write(f"{'ms/step':>9s} {'launches/step':>14s} {'share':>7s}  kernel\n")
        for n, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"{v / steps_in_file:9.3f} {cnt[n] / steps_in_file:14.1f} {100 * v / S:6.1f}%  {n}\n")
        ours = sum(v for n, v in tot.items() if n.startswith(("bmkg::", "nce::")) or "bmkg::" in n)
        f.write(f"\nour kernels (bmkg::*): {100 * ours / S:.1f}% of GPU time; cuBLAS GEMMs (nvjet/cublasLt/internal::kernel): "
                f"{100 * sum(v for n, v in tot.items() if 'nvjet' in n or 'cublas' in n.lower() or 'internal::kernel' in n) / S:.1f}%; "
                f"torch elementwise/optimizer: the rest\n")
    print("wrote", out)


if __name__ == "__main__":
    g = "gpurun_out"
    if os.path.exists(f"{g}/launches_r1.csv"):
        summarize_launches(f"{g}/launches_r1.csv", f"{OUT}/r1_launches_bench_cfg2.txt", 7)
    for rep, out, note in [("prof_infonce_r1", "r1_ncu_infonce_N28000.txt", "; python tools/prof_kernels.py infonce 28000"),
                           ("prof_gcn_agg_1M_r1", "r1_ncu_gcn_aggregate_N1M_E50M.txt", "; python tools/prof_kernels.py gcn 1000000 3 50000000 (kept edges 31M, uniform)"),
                           ("prof_gat_r1", "r1_ncu_gat_cfg2.txt", "; GAT kernels inside bench.py cfg2")]:
        if os.path.exists(f"{g}/{rep}.ncu-rep"):
            summarize_rep(f"{g}/{rep}.ncu-rep", f"{OUT}/{out}", note)
