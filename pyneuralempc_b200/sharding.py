"""Multi-GPU sharding of independent MPC problems: contiguous split of the batch over ranks, weights and index
tables replicated, NO collective on the evaluation path.  ``torch.distributed`` is only used for the
plumbing around it: rendezvous, a barrier, the max-over-ranks of the timed region, and an optional final gather."""
from __future__ import annotations

import os


def shard_range(total, rank, world):
    """problems [lo, hi) owned by ``rank``: sizes differ by at most one, earlier ranks get the remainder."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(total), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device_index, sysfs="/sys/bus/pci/devices"):
    """pin this process to the CPU cores next to GPU ``device_index`` (its PCIe root's NUMA node) BEFORE pinned host buffers are
    allocated, so that first-touch places them in that node's memory: with one process per GPU the host <-> device copies of
    ``nempc_eval_host`` then stay on the GPU's own socket instead of crossing the inter-socket link.  Returns the CPU set used, or None
    when the topology is not exposed (containers without sysfs NUMA information): never fatal."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index)
        pci = getattr(bus, "pci_bus_id", None)
        dom = getattr(bus, "pci_domain_id", 0)
        dev = getattr(bus, "pci_device_id", 0)
        if pci is None:
            return None
        path = os.path.join(sysfs, f"{dom:04x}:{pci:02x}:{dev:02x}.0", "local_cpulist")
        with open(path) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (device timing is reported as the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_solutions(local, dst=0):
    """optional final gather of per-rank result arrays (torch tensors, same trailing shape) onto ``dst``; off the
    timed path.  Returns the concatenated tensor on ``dst`` and None elsewhere."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device))
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[: int(s.item())] for b, s in zip(bufs, sizes)], dim=0)
