"""Keep a rank's host threads -- and therefore its pinned staging buffers, which Linux places on the node of the thread
that first touches them -- on the NUMA node its GPU hangs off.  A host->device copy out of the other socket's memory
crosses the inter-socket link and shares it with every other rank doing the same (measured at 8 ranks: 20 GB/s per
rank without binding).  Launch-side plumbing; a no-op when sysfs / NVML do not expose the topology."""
import os


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


def gpu_numa_node(device_index):
    """NUMA node of a CUDA device (index within CUDA_VISIBLE_DEVICES), or None."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if vis:
            ids = vis.split(",")
            if device_index < len(ids) and ids[device_index].strip().isdigit():
                phys = int(ids[device_index])
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        bus = nv.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_node(device_index):
    """Restrict this process to the CPUs of the GPU's NUMA node.  Returns a small report dict (also when nothing was
    done: {"bound": False, "why": ...})."""
    node = gpu_numa_node(device_index)
    if node is None:
        return {"bound": False, "why": "no NUMA information for the device"}
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return {"bound": False, "node": node, "why": "no allowed CPU on the device's node"}
        if cpus == allowed:
            return {"bound": False, "node": node, "why": "already confined to the device's node"}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "node": node, "cpus": len(cpus)}
    except Exception as e:  # pragma: no cover - topology dependent
        return {"bound": False, "node": node, "why": repr(e)[:120]}
