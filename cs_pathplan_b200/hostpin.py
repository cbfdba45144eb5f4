"""Host-side placement for the host-pointer (end-to-end) path: keep the threads that drive a GPU, and the pinned buffers
they first touch, on the CPU cores / NUMA node the GPU hangs off.  With eight ranks on one box the device -> host copies of
all GPUs otherwise land on whichever node the processes happened to start on and share one memory controller / one PCIe
root's worth of host bandwidth (round-1 finding: 50 GB/s per GPU alone, 11 GB/s per GPU with eight)."""
from __future__ import annotations

import os


def gpu_cpu_affinity(index: int):
    """CPU ids NVML reports as local to GPU `index` (empty list if NVML cannot tell)."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        return [c for c in cpus if c < ncpu]
    except Exception:
        return []


def pin_to_gpu_numa(index: int):
    """Restrict the calling process to the CPUs local to GPU `index` (intersected with what it is allowed to use).  Returns
    the CPU list applied, or None if nothing was changed.  Call BEFORE allocating pinned host buffers: pages are placed on
    the node of the thread that first touches them."""
    try:
        allowed = os.sched_getaffinity(0)
        local = [c for c in gpu_cpu_affinity(index) if c in allowed]
        if not local or len(local) == len(allowed):
            return None
        os.sched_setaffinity(0, local)
        return local
    except Exception:
        return None
