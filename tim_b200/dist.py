"""Data-parallel plumbing: clips shard over ranks (one process per GPU). The forward needs NO data-path collective
(clips are independent: no BatchNorm, and the mask makes even queries independent) — the only
collectives here are the timing / result-gathering ones used by bench.py and the tests. Mirrors the reference's
DistributedSampler split (datasets/loader.py:48-50) and utils/distributed.py:15-53 helpers."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process (one per GPU) to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates pinned host
    buffers: with 8 ranks streaming inputs / results over PCIe at once, buffers first-touched on the other socket push
    every copy through the inter-socket link. Returns the node, or None when the topology cannot be read (no-op then)."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most one) clip range of `rank`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device: str = "cpu") -> float:
    """Device-timed durations are reported as the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(local: torch.Tensor, rows_per_clip: int, B: int) -> torch.Tensor:
    """Concatenate per-rank [b_r * rows_per_clip, C] results in clip order (ragged shards allowed)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    C_ = local.shape[1]
    maxb = max(shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world))
    pad = torch.zeros((maxb * rows_per_clip, C_), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    parts = []
    for r in range(world):
        lo, hi = shard_range(B, r, world)
        parts.append(bufs[r][:(hi - lo) * rows_per_clip])
    return torch.cat(parts, 0)


class FlatGrads:
    """Groundwork for the training leg: ONE contiguous fp32 buffer holds the gradient of
    every parameter, `param.grad` are views into it, and one all-reduce per step (sum, then divide by the world size) replaces
    DDP's bucketed all-reduce (`recognition/time_interval_machine/models/build.py:58-63`). The backward kernels will write their
    weight gradients straight into this buffer (`tim_encoder_bwd`'s flat dW argument); AdamW keeps seeing ordinary `.grad` tensors.

    Layout: parameters in the order given (state_dict order), each segment padded to 128 bytes so that every view is aligned for
    vector / TMA access. `offsets[name] = (first element, number of elements)`.
    """

    ALIGN_ELEMS = 32            # 128 bytes of fp32

    def __init__(self, named_params, device=None):
        named = [(n, p) for n, p in named_params if p.requires_grad]
        if not named:
            raise ValueError("FlatGrads: no parameter requires a gradient")
        device = device if device is not None else named[0][1].device
        self.offsets = {}
        at = 0
        for n, p in named:
            self.offsets[n] = (at, p.numel())
            at += (p.numel() + self.ALIGN_ELEMS - 1) // self.ALIGN_ELEMS * self.ALIGN_ELEMS
        self.buffer = torch.zeros((at,), dtype=torch.float32, device=device)
        self._params = named
        for n, p in named:
            lo, cnt = self.offsets[n]
            p.grad = self.buffer[lo:lo + cnt].view_as(p)

    def reattach(self) -> int:
        """Both reference train loops call optimizer.zero_grad() (recognition/scripts/train.py:354, detection/scripts/train.py:372),
        which by default sets param.grad to None: the next backward would then allocate fresh gradients OUTSIDE the flat buffer and
        all_reduce() would exchange stale data. Call this before every backward (patch_model does): a parameter whose .grad is no
        longer its view is re-attached - a detached gradient counts as zero, a foreign one is copied in. Returns how many were
        re-attached."""
        n = 0
        for name, p in self._params:
            lo, cnt = self.offsets[name]
            view = self.buffer[lo:lo + cnt].view_as(p)
            g = p.grad
            if g is not None and g.data_ptr() == view.data_ptr():
                continue
            if g is None:
                view.zero_()
            else:
                view.copy_(g)
            p.grad = view
            n += 1
        return n

    def zero_(self) -> None:
        self.buffer.zero_()

    def all_reduce(self, group=None, average: bool = True) -> None:
        """The one data-path collective of a training step: NCCL on GPUs (NVLink / NVSwitch), gloo in the CPU tests."""
        if not (dist.is_available() and dist.is_initialized()):
            return
        self.reattach()
        dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.buffer.div_(dist.get_world_size(group))

    def view(self, name: str) -> torch.Tensor:
        lo, cnt = self.offsets[name]
        return self.buffer[lo:lo + cnt]
