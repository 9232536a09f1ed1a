"""Host placement of a rank: run on (and so allocate page-locked buffers from) the CPUs next to its GPU.

The result matrices of kmer.pos leave the device over PCIe (2.9 GB per 250 Mbp index); with one process per GPU on
a two-socket host, a rank whose pinned buffers sit on the other socket drains through the inter-socket link and the
ranks' aggregate host bandwidth collapses ("e2e").  Linux allocates a process's pages on the node of the
CPU that touches them, so binding the process to the GPU's own CPUs BEFORE the buffers are made is all it takes.
Best effort: without NVML / sysfs information nothing is changed.
"""
from __future__ import annotations

import os


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_cpus(pci_bus_id: str) -> tuple[set[int], int | None]:
    """(CPUs local to the PCI device, its NUMA node) from sysfs; (empty, None) if unknown."""
    base = f"/sys/bus/pci/devices/{pci_bus_id.lower()}"
    try:
        cpus = _parse_cpulist(open(os.path.join(base, "local_cpulist")).read())
    except OSError:
        return set(), None
    node = None
    try:
        node = int(open(os.path.join(base, "numa_node")).read().strip())
    except (OSError, ValueError):
        pass
    return cpus, node


def bind_near_gpu(cuda_index: int) -> dict:
    """Restrict this process (and the threads it starts later) to the CPUs closest to CUDA device `cuda_index`.
    Returns what was found and done, for the bench line."""
    info = {"gpu": int(cuda_index), "node": None, "cpus_local": 0, "cpus_allowed": 0, "bound": False}
    try:
        import torch
        p = torch.cuda.get_device_properties(cuda_index)
        bus = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception as e:  # noqa: BLE001
        info["why"] = f"no PCI address: {e!r}"
        return info
    cpus, node = gpu_cpus(bus)
    if not cpus:
        try:                                                  # sysfs hidden (containers): ask NVML
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            words = ((os.cpu_count() or 1) + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        except Exception as e:  # noqa: BLE001
            info["why"] = f"no locality information: {e!r}"
            return info
    allowed = os.sched_getaffinity(0)
    use = cpus & allowed
    info.update(node=node, cpus_local=len(cpus), cpus_allowed=len(allowed), pci=bus)
    if use and use != allowed:
        try:
            os.sched_setaffinity(0, use)                      # this thread; threads started from now on inherit it
            info["bound"] = True
            for t in os.listdir("/proc/self/task"):           # and the ones that already run (best effort)
                try:
                    os.sched_setaffinity(int(t), use)
                except (OSError, ValueError):
                    pass
        except OSError as e:
            info["why"] = repr(e)
    elif not use:
        info["why"] = "none of the GPU's CPUs is in this process's allowed set"
    else:
        info["why"] = "already confined to the GPU's CPUs"
    return info
