"""Host-side placement for one-process-per-GPU jobs: bind the process (and therefore the pinned buffers it allocates
afterwards - first touch) to the NUMA node its GPU hangs off.  The end-to-end path is a stream of pinned-memory H2D
copies; with every rank on node 0 (the default affinity of a torchrun job) the ranks whose GPUs sit on the other
socket pull their bytes across the inter-socket link and the aggregate H2D rate collapses."""
from __future__ import annotations

import os
from typing import Dict, List, Optional


def _read(path: str) -> Optional[str]:
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(s: str) -> List[int]:
    out: List[int] = []
    for part in s.split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of a CUDA device (sysfs), None when the platform does not say."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:                                            # noqa: BLE001
        return None
    v = _read(f"/sys/bus/pci/devices/{bus}/numa_node")
    if v is None:
        return None
    n = int(v)
    return n if n >= 0 else None


def topology(device_index: int) -> Dict:
    node = gpu_numa_node(device_index)
    nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()) \
        if os.path.isdir("/sys/devices/system/node") else []
    cpus = _read(f"/sys/devices/system/node/node{node}/cpulist") if node is not None else None
    return {"gpu": device_index, "gpu_numa_node": node, "numa_nodes": nodes, "node_cpulist": cpus,
            "affinity": sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None}


def bind_to_gpu_node(device_index: int) -> Dict:
    """Restrict this process to the CPUs of the GPU's NUMA node that it is allowed to use (never widens the affinity the
    launcher / container gave it).  Call BEFORE allocating pinned host buffers.  -> what was done (for the bench line)."""
    info = topology(device_index)
    info["bound"] = False
    if info["gpu_numa_node"] is None or not info["node_cpulist"] or info["affinity"] is None or len(info["numa_nodes"]) < 2:
        return info
    want = set(_parse_cpulist(info["node_cpulist"])) & set(info["affinity"])
    if want and want != set(info["affinity"]):
        try:
            os.sched_setaffinity(0, want)
            info["bound"] = True
            info["affinity"] = sorted(want)
        except OSError:
            pass
    return info
