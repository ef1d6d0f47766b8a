"""Host-side placement for the batch API: run this process on the CPU cores of the GPU's NUMA node, so that the
pinned staging buffers encode_batch() / decode_batch() allocate afterwards (first touch) live in the memory
attached to the socket the GPU's PCIe root port hangs off.  With one process per GPU on a two-socket box this
keeps the D2H streams of the eight GPUs from all landing on one socket's memory controllers.  Linux only; a
no-op (with the reason returned) wherever sysfs does not say."""
import os


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _cpulist(text):
    cpus = set()
    for part in text.split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index):
    """NUMA node of a CUDA device from its PCI address (None if unknown)."""
    import torch
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
    except Exception:
        return None
    node = _read("/sys/bus/pci/devices/%s/numa_node" % bus)
    if node is None or int(node) < 0:
        return None
    return int(node)


def bind_to_gpu(device_index):
    """Restrict this process to the cores of the GPU's NUMA node (within its current affinity mask).
    Returns a short note for logs / the bench line."""
    node = gpu_numa_node(device_index)
    if node is None:
        return "numa node of cuda:%d unknown; affinity unchanged" % device_index
    text = _read("/sys/devices/system/node/node%d/cpulist" % node)
    if not text:
        return "no cpulist for node %d; affinity unchanged" % node
    try:
        allowed = os.sched_getaffinity(0)
        want = _cpulist(text) & allowed
        if not want:
            return "node %d has no allowed cores; affinity unchanged" % node
        os.sched_setaffinity(0, want)
    except (AttributeError, OSError) as e:
        return "sched_setaffinity failed (%s); affinity unchanged" % e
    return "cuda:%d -> numa node %d, %d cores" % (device_index, node, len(want))
