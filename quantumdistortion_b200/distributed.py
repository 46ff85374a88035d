"""Multi-GPU sharding of a batch of clips: one process per GPU, clips are independent units.

There is NO collective on the data path: every rank renders its own contiguous
slice of the batch with the same parameters; the only communication is the optional final gather of the
finished clips to one rank (``torch.distributed.gather``: NCCL for CUDA tensors, gloo for CPU tensors).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of ``n_items`` over ``world`` ranks; the first ``n_items % world`` ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, rem = divmod(int(n_items), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def render_sharded(x, render: Callable, *, gather_to: Optional[int] = 0, group=None):
    """Render ``x[lo:hi]`` of the (replicated) batch ``x[B, n]`` on this rank with ``render(shard) -> shard_out``
    and gather the results in rank order on ``gather_to`` (None: no gather, every rank keeps its shard).

    Returns ``(out, (lo, hi))``: ``out`` is the full ``[B, n]`` result on rank ``gather_to``, the local shard's
    result elsewhere (or everywhere when ``gather_to`` is None)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return render(x), (0, int(x.shape[0]))
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(int(x.shape[0]), rank, world)
    y_local = render(x[lo:hi])
    if gather_to is None:
        return y_local, (lo, hi)
    # ranks may hold one clip more or less: pad to the largest shard for the fixed-size gather
    most = shard_bounds(int(x.shape[0]), 0, world)
    most = most[1] - most[0]
    padded = y_local
    if hi - lo < most:
        pad = torch.zeros((most - (hi - lo),) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
        padded = torch.cat([y_local, pad], dim=0)
    padded = padded.contiguous()
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == gather_to else None
    dist.gather(padded, bufs, dst=gather_to, group=group)
    if rank != gather_to:
        return y_local, (lo, hi)
    parts = []
    for r in range(world):
        a, b = shard_bounds(int(x.shape[0]), r, world)
        parts.append(bufs[r][: b - a])
    return torch.cat(parts, dim=0), (lo, hi)


def process_batch_sharded(x, sr: int = 48000, *, gather_to: Optional[int] = 0, **kwargs):
    """``process_batch`` over the ranks of the default process group (one rank per GPU).  ``x`` is the same
    CUDA (or CPU) ``[B, n]`` tensor on every rank; each rank renders only its slice.  ``seeds`` (random spectral
    FX) keeps its single-process meaning: a per-clip list is sliced per rank, and with ``seeds=None`` every rank
    steps its np.random state past the clips of the lower ranks first (``process_batch(shard=...)``), so the result
    equals the unsharded ``process_batch`` call made from the same np.random state."""
    import torch.distributed as dist
    from .pipeline import process_batch
    total = int(x.shape[0])
    lo, hi = 0, total
    if dist.is_available() and dist.is_initialized():
        lo, hi = shard_bounds(total, dist.get_rank(), dist.get_world_size())
    return render_sharded(x, lambda part: process_batch(part, sr, shard=(lo, hi, total), **kwargs)[0], gather_to=gather_to)


def bind_to_gpu_numa(device_index: int) -> Optional[str]:
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (sysfs ``local_cpulist`` of its PCI
    function), so that pinned host buffers allocated afterwards are node-local (first touch) and the
    host<->device copies do not cross the socket interconnect.  Returns the cpulist used, or None when the
    topology is not visible (containers without sysfs PCI nodes) -- in which case nothing is changed."""
    import os
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        dom = getattr(pr, "pci_domain_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/local_cpulist"
        with open(path) as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpulist
    except Exception:  # noqa: BLE001
        return None
