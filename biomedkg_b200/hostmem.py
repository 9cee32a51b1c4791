"""Host-side staging memory for the loaders: pinned buffers placed on the NUMA node the GPU hangs off.

A pinned page lives on the node of the thread that allocates it.  On a two-socket 8-GPU box a process that happens to run on
the other socket pins its batch behind the inter-socket link, and the host->device copies of all ranks then share that link
instead of each using its own PCIe root.  ``pinned_near`` narrows the calling thread's CPU affinity to the CPUs NVML reports as
local to the device for the duration of the allocation only (worker threads, OpenMP pools and the rest of the process keep
their affinity); when NVML, the topology or the container's cpuset give no usable answer it pins wherever the thread runs."""
from __future__ import annotations

import os

import torch


def gpu_local_cpus(device: int):
    """CPUs NVML reports as local to CUDA device ``device`` that this process may run on, or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(device)
        bus = f"{p.pci_domain_id:08X}:{p.pci_bus_id:02X}:{p.pci_device_id:02X}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:      # no NVML, no permission, unknown topology: not an error, just no placement hint
        return None


def pinned_near(t: torch.Tensor, device: int) -> torch.Tensor:
    """``t.pin_memory()`` with the pinned pages allocated on the NUMA node local to CUDA device ``device`` when that is known."""
    cpus = gpu_local_cpus(device)
    if not cpus:
        return t.pin_memory()
    old = os.sched_getaffinity(0)
    try:
        os.sched_setaffinity(0, cpus)
        return t.pin_memory()
    finally:
        os.sched_setaffinity(0, old)
