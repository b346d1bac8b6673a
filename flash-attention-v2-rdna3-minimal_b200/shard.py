"""Multi-GPU plan for the forward path: an NCCL-free batch shard.

Every (batch, head, Q-tile) of the forward is independent (reference grid
``(b, h, Tr)``, /root/reference/rocwmma_fattn/kernel_fp16.cu:803-806), so N GPUs simply take
contiguous slices of the batch axis and run the single-GPU kernel; nothing is exchanged on the data
path.  The only cross-rank traffic is the scalar timing / launch-count reduction bench.py does with
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations


def shard_batch(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous slice ``(start, count)`` of ``total`` batch elements owned by ``rank``.
    The first ``total % world`` ranks get one extra element; counts may be 0 if world > total."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request total={total} world={world} rank={rank}")
    base, extra = divmod(total, world)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def reduce_max_time(local_ms: float, dist=None, device=None) -> float:
    """Job time = max over ranks of the device time each rank measured."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_ms)
    import torch

    t = torch.tensor([local_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_tflops(flops_all_ranks: float, local_ms: float, dist=None, device=None) -> float:
    """Whole-job TFLOPS = FLOPs of all ranks / max-over-ranks time."""
    ms = reduce_max_time(local_ms, dist, device)
    return flops_all_ranks / (ms * 1e-3) / 1e12


def device_local_cpus(pci_domain: int, pci_bus: int, pci_device: int, sysfs: str = "/sys/bus/pci/devices") -> list[int]:
    """CPUs on the NUMA node the GPU hangs off (``local_cpulist`` of its PCI device), or [] if unknown."""
    path = f"{sysfs}/{pci_domain:04x}:{pci_bus:02x}:{pci_device:02x}.0/local_cpulist"
    try:
        with open(path) as fh:
            text = fh.read().strip()
    except OSError:
        return []
    cpus: list[int] = []
    for part in text.split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_device_numa(device_index: int) -> list[int]:
    """Pin the calling process to the CPUs local to GPU ``device_index`` so that the pinned host buffers it
    allocates afterwards (first touch) and the threads that feed the copy engines sit on the GPU's NUMA node.
    One process per GPU: without this, ranks on a two-socket host stage half of their PCIe traffic through the
    inter-socket link.  Returns the CPU list it bound to ([] = left unchanged: unknown topology, or the local
    list is not a subset of what the process may use)."""
    import os

    import torch

    prop = torch.cuda.get_device_properties(device_index)
    cpus = device_local_cpus(prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
    try:
        allowed = os.sched_getaffinity(0)
    except AttributeError:
        return []
    cpus = [c for c in cpus if c in allowed]
    if not cpus or len(cpus) == len(allowed):
        return []
    os.sched_setaffinity(0, cpus)
    return cpus
