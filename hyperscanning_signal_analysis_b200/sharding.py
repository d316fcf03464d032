"""Multi-GPU layer: one process per GPU (torch.distributed), units sharded by dyad/task, no collective inside
the computation (every (dyad, task, window) is independent: run_pipeline's outer loops,
eeg_alpha_ibi_ffdtf.py:665-666).  The only exchange is the optional final all-gather of the result tensor
(NCCL over NVLink on GPUs; the same code runs on gloo/CPU tensors in the tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_units(n_units: int, rank: int, world: int) -> range:
    """Contiguous block partition: ranks 0..(n_units % world)-1 get one extra unit."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def shard_by_cost(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time assignment for uneven units (tasks of different duration): returns, per rank,
    the unit indices it owns (deterministic)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    owned: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        owned[r].append(i)
        load[r] += costs[i]
    return [sorted(o) for o in owned]


def all_gather_windows(local, counts: Sequence[int]):
    """Gather per-rank result tensors ``(n_local, ...)`` (possibly different n_local) into ``(sum(counts), ...)``
    on every rank.  Equal counts use one ``all_gather_into_tensor``; ragged counts pad to the maximum."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    assert len(counts) == world
    tail = tuple(local.shape[1:])
    mx = max(counts)
    if all(c == mx for c in counts):
        out = torch.empty((world * mx,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    padded = torch.zeros((mx,) + tail, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty((world * mx,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded)
    return torch.cat([buf[r * mx: r * mx + counts[r]] for r in range(world)], dim=0)


def unit_table(n_dyads: int, tasks: Sequence[str]) -> List[Tuple[int, str]]:
    """(dyad, task) units in the order run_pipeline visits them."""
    return [(d, t) for d in range(n_dyads) for t in tasks]


def bind_to_gpu_numa(device_index):
    """Pin the calling process to the CPUs local to GPU ``device_index`` (sysfs ``local_cpulist`` of its PCI function), so that
    pinned host buffers allocated afterwards (first touch) sit on the GPU's own NUMA node and the result stream of every rank
    crosses its own root complex instead of the socket interconnect -- what ``numactl --cpunodebind`` does per rank in a
    deployment.  Best effort: returns the CPU set, or None when the topology cannot be read (nothing is changed then)."""
    import os
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        addr = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{addr}/local_cpulist") as fh:
            text = fh.read().strip()
        cpus = set()
        for part in text.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
