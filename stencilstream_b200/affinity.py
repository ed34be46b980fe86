"""Host-side placement of a rank next to its GPU.

On a multi-socket 8 x B200 box every GPU hangs off one CPU socket's PCIe root complex. A rank whose
host threads (the Python thread that fills and reads the cell images, the runtime's copy workers) and
whose pinned buffers live on the OTHER socket moves every byte across the inter-socket link on its way
to the GPU. `bind_to_gpu_numa_node(device)` restricts the calling process to the cores of the GPU's
NUMA node (memory then follows by first touch), keeping whatever restriction the launcher already
imposed. The reference has no counterpart (single device, SYCL runtime placement).
"""
from __future__ import annotations

import os
from pathlib import Path


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for item in text.strip().split(","):
        if not item:
            continue
        if "-" in item:
            lo, hi = item.split("-")
            cpus.update(range(int(lo), int(hi) + 1))
        else:
            cpus.add(int(item))
    return cpus


def gpu_numa_cpus(device: int) -> tuple[int | None, set[int]]:
    """(NUMA node of CUDA device `device`, cores of that node), or (None, empty set) where the
    platform does not say (single-socket hosts report node -1)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device)
        bus_id = pynvml.nvmlDeviceGetPciInfo(handle).busId
        if isinstance(bus_id, bytes):
            bus_id = bus_id.decode()
    except Exception:
        return None, set()
    # NVML reports an 8-digit PCI domain ("00000000:1B:00.0"), sysfs uses 4 digits
    domain, rest = bus_id.split(":", 1)
    sysfs = Path("/sys/bus/pci/devices") / f"{domain[-4:]}:{rest}".lower()
    try:
        node = int((sysfs / "numa_node").read_text())
        if node < 0:
            return None, set()
        cpus = _parse_cpulist((Path("/sys/devices/system/node") / f"node{node}" / "cpulist").read_text())
        return node, cpus
    except (OSError, ValueError):
        return None, set()


def bind_to_gpu_numa_node(device: int, rank_on_node: int = 0, ranks_on_node: int = 1) -> dict:
    """Restrict this process to the cores of `device`'s NUMA node (intersected with its current
    affinity mask; if several ranks share the node, each takes a disjoint share of those cores).
    Returns a description of what was done, for logging. Does nothing where the information is
    missing or the restriction would leave no core."""
    before = os.sched_getaffinity(0)
    node, cpus = gpu_numa_cpus(device)
    usable = sorted(before & cpus)
    if node is None or not usable:
        return {"numa_node": node, "bound": False, "cpus": sorted(before)}
    if ranks_on_node > 1 and len(usable) >= ranks_on_node:
        share = len(usable) // ranks_on_node
        usable = usable[rank_on_node * share:(rank_on_node + 1) * share]
    os.sched_setaffinity(0, usable)
    return {"numa_node": node, "bound": True, "cpus": usable}
