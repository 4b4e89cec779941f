"""Host-side placement for the host-pointer entry points: run the calling process on the CPUs next to its GPU before it
allocates pinned buffers, so first-touch puts the pages on the GPU's own NUMA node and the DMA does not cross the socket
interconnect (on an 8-GPU box every rank otherwise shares whichever node the launcher started it on).  Plumbing only."""
import os


def _pci_bdf(device):
    import torch
    p = torch.cuda.get_device_properties(device)
    dom = getattr(p, "pci_domain_id", 0)
    return f"{dom:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa_node(device, spread=None):
    """Restrict this process to the CPUs local to `device` (sysfs local_cpulist).  spread = (index, count): take the
    index-th of `count` equal slices of that list (ranks that share a node do not pile on the same cores).  Returns a dict
    describing what was done; never raises (a VM without NUMA information is left alone)."""
    info = {"device": int(device), "numa_node": None, "cpus_before": None, "cpus_after": None, "bound": False}
    try:
        bdf = _pci_bdf(device)
        base = f"/sys/bus/pci/devices/{bdf}"
        info["pci"] = bdf
        with open(os.path.join(base, "numa_node")) as f:
            info["numa_node"] = int(f.read().strip())
        with open(os.path.join(base, "local_cpulist")) as f:
            local = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        info["cpus_before"] = len(allowed)
        local &= allowed
        if not local or info["numa_node"] < 0 or local == allowed:
            info["cpus_after"] = len(allowed)
            return info                          # single node / no information: nothing to gain
        if spread is not None:
            idx, cnt = spread
            ordered = sorted(local)
            per = max(1, len(ordered) // max(1, cnt))
            mine = ordered[(idx % cnt) * per:(idx % cnt) * per + per]
            local = set(mine) or local
        os.sched_setaffinity(0, local)
        info["cpus_after"] = len(local)
        info["bound"] = True
    except Exception as e:                       # plumbing must never take a run down
        info["error"] = repr(e)
    return info
