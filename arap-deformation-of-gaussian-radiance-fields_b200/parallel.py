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


class SoAGatherPush:
    """All-gather of equally sized per-rank SoA shards by peer stores: every rank copies its shard straight into the
    gathered arrays of all ranks (symmetric memory, peer-mapped over NVLink) with device-to-device copies, which run on
    the copy engines — no SM is needed, so the gather also overlaps kernels that fill the GPU (the cooperative solve).
    A device-side barrier at the end orders the pushes of all ranks before any consumer.
    Measured (8 B200s, 1.39 GB shard): ~270 GB/s per rank, against ~900 GB/s for NCCL's SM kernels — kept as an option
    (`ARAP_GATHER=push`), NCCL is the default."""

    def __init__(self, parts: dict, world: int, rank: int):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.parts, self.outs, self.dst, self.streams = parts, {}, {}, None
        group = dist.group.WORLD
        for k, t in parts.items():
            n = t.shape[0]
            out = symm.empty((world * n,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            hdl = symm.rendezvous(out, group.group_name)
            self.dst[k] = [hdl.get_buffer(r, out.shape, out.dtype)[rank * n:(rank + 1) * n] for r in range(world)]
            self.outs[k] = out
            self.hdl = hdl

    def __call__(self):
        import torch
        cur = torch.cuda.current_stream()
        if self.streams is None:   # one stream per destination: the copies to different peers use different copy engines
            self.streams = [torch.cuda.Stream(priority=-1) for _ in range(len(next(iter(self.dst.values()))))]
        for r, st in enumerate(self.streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                for k, t in self.parts.items():
                    self.dst[k][r].copy_(t, non_blocking=True)
        for st in self.streams:
            cur.wait_stream(st)
        self.hdl.barrier()
        return self.outs


class SoAGatherPose:
    """All-gather that sends only the pose (pos 3 + rot 4 + scale 3 floats = 40 of the 232 bytes per Gaussian).

    Every rank keeps last step's full SoA of all ranks.  The SH rows of the remote shards are not received: the receiver
    repeats the owner's SH update — R = (q_new q_old^-1).normalized, `sh_rotate` — from the old and the new rotation with
    the same kernel arithmetic (`arapk_replay_shs`), so its copy stays bit-identical to the owner's (checked against a full
    NCCL gather in `bench.py` with ARAP_GATHER_CHECK=1).  8 GPUs x 6M Gaussians: 1.7 instead of 9.7 GB received per rank
    per step.  Needs one call per drag step (a skipped step would skip a rotation) and `refresh_static()` after the
    control blocks change (static Gaussians are not rotated by their owner).  Opt-in: `ARAP_GATHER=pose`."""

    def __init__(self, parts: dict, world: int, rank: int, lib, static_flags=None, replay=None):
        """`replay(rot_old, rot_new, static, shs)` updates `shs` in place for a contiguous index range; default: the CUDA kernel
        `arapk_replay_shs` of `lib`.  (The gloo tests pass a host stand-in with the owner's update rule.)"""
        import torch
        import torch.distributed as dist
        self.parts, self.world, self.rank, self.lib = parts, world, rank, lib
        self.replay = replay or self._replay_cuda
        self.n = parts["pos"].shape[0]
        self.outs = {k: torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for k, t in parts.items()}
        for k, t in parts.items():                       # initial state: everything, once
            dist.all_gather_into_tensor(self.outs[k], t)
        self.rot_prev = torch.empty_like(self.outs["rot"])
        self.static = torch.zeros(world * self.n, dtype=torch.uint8, device=parts["pos"].device)
        if static_flags is not None:
            self.refresh_static(static_flags)

    def refresh_static(self, local_flags):
        import torch.distributed as dist
        dist.all_gather_into_tensor(self.static, local_flags.contiguous())

    def __call__(self):
        import torch.distributed as dist
        n, lo = self.n, self.rank * self.n
        self.rot_prev.copy_(self.outs["rot"])
        for k in ("pos", "rot", "scale"):
            dist.all_gather_into_tensor(self.outs[k], self.parts[k])
        self.outs["shs"][lo:lo + n].copy_(self.parts["shs"])          # own shard: the owner's rows
        for a, b in ((0, lo), (lo + n, self.world * n)):              # remote shards: replay the rotation
            if b > a:
                self.replay(self.rot_prev[a:b], self.outs["rot"][a:b], self.static[a:b], self.outs["shs"][a:b])
        return self.outs

    def _replay_cuda(self, rot_old, rot_new, static, shs):
        import ctypes as C
        import torch
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = self.lib.arapk_replay_shs(C.c_longlong(rot_new.shape[0]), C.c_void_p(rot_old.data_ptr()), C.c_void_p(rot_new.data_ptr()),
                                       C.c_void_p(static.data_ptr()), C.c_void_p(shs.data_ptr()), st)
        if rc != 0:
            raise RuntimeError(f"arapk_replay_shs failed ({rc})")


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
