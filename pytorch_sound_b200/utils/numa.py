"""Bind a rank's host threads (and therefore its first-touch / pinned allocations) to the NUMA node of its GPU.

The input path of the reference is DataLoader(pin_memory=True) -> `x.cuda(non_blocking=True)` (data/dataset.py:180,
utils/tensor.py:15): the batch sits in pinned host memory and crosses PCIe every step.  On a two-socket box the
copies of the GPUs behind the OTHER socket cross the inter-socket link when every rank allocates on node 0 — the
8-GPU end-to-end number of round 1 (2.26x one GPU) was bound by exactly that.  Linux exposes the GPU's local CPUs in
sysfs; pinning the process to them before it allocates makes `cudaHostAlloc` place the pages on the GPU's own node.
"""
import os
from typing import Optional


def gpu_local_cpus(device_index: int) -> Optional[list]:
    """CPUs local to the GPU's PCIe root (from /sys/bus/pci/devices/<bdf>/local_cpulist), or None if unknown."""
    try:
        import torch

        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            text = f.read().strip()
    except Exception:
        return None
    cpus = []
    for part in text.split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus or None


def bind_to_gpu(device_index: int) -> dict:
    """Restrict this process to the GPU-local CPUs (intersection with the CPUs it is allowed to use).
    Returns what was done, for the bench record; never raises."""
    info = {"bound": False}
    try:
        allowed = os.sched_getaffinity(0)
        local = gpu_local_cpus(device_index)
        if not local:
            info["reason"] = "no local_cpulist for the GPU"
            return info
        target = sorted(allowed & set(local))
        if not target:
            info["reason"] = "GPU-local CPUs are outside this process's cpuset"
            return info
        if set(target) != allowed:
            os.sched_setaffinity(0, target)
        info.update(bound=True, cpus=len(target), first_cpu=target[0])
        try:
            with open(f"/sys/devices/system/cpu/cpu{target[0]}/topology/physical_package_id") as f:
                info["socket"] = int(f.read())
        except Exception:
            pass
    except Exception as ex:
        info["reason"] = f"{type(ex).__name__}: {ex}"
    return info
