"""Host-side multi-GPU plumbing (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The path shards by Gaussian index: contiguous index ranges are spatial slabs because the
Gaussians are stored in cell order.  The node solve is replicated, so the only per-step exchange is an
all-gather of the deformed SoA (pos 3 + rot 4 + scale 3 + shs 48 floats = 232 B / Gaussian).
"""
from __future__ import annotations

import numpy as np

SOA_WIDTHS = (("pos", 3), ("rot", 4), ("scale", 3), ("shs", 48))
SOA_BYTES_PER_GAUSSIAN = 4 * sum(w for _, w in SOA_WIDTHS)


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced index range of `rank` (first n % world ranks get one extra element)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world: int) -> list[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


class DevArray:
    """Expose a raw device pointer owned by an arap_ctx to torch via __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)


class SoAGather:
    """All-gather of equally sized per-rank SoA shards into full arrays (one collective per attribute)."""

    def __init__(self, parts: dict, world: int):
        import torch
        self.parts = parts
        self.outs = {k: torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for k, t in parts.items()}

    def __call__(self):
        import torch.distributed as dist
        for k, t in self.parts.items():
            dist.all_gather_into_tensor(self.outs[k], t)
        return self.outs


def allgather_variable(local: "np.ndarray", world: int):
    """All-gather of unequal shards (host arrays, any backend): pads to the largest shard."""
    import torch
    import torch.distributed as dist
    n = torch.tensor([local.shape[0]], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = np.zeros((mx,) + local.shape[1:], local.dtype)
    pad[:local.shape[0]] = local
    bufs = [torch.zeros_like(torch.from_numpy(pad)) for _ in range(world)]
    dist.all_gather(bufs, torch.from_numpy(pad))
    return np.concatenate([b.numpy()[:s] for b, s in zip(bufs, sizes)])
