"""Multi-GPU host logic: images are independent, so the batch is sharded over ranks (weights replicated) and the only
exchange on the path is ONE gather of the fixed-capacity detections at the end.

The functions below are backend-agnostic (`nccl` on GPUs over NVLink/NVSwitch, `gloo` in the CPU tests)."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist

OVERFLOW_BIT = 1 << 30


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a batch of n images owned by `rank`; the first n % world ranks get one more."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_detections(boxes: torch.Tensor, scores: torch.Tensor, classes: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """[B,cap,4] f32, [B,cap] f32, [B,cap] i64 -> one [B,cap,6] f32 payload (x1,y1,x2,y2,score,class)."""
    B, cap = scores.shape
    if out is None:
        out = torch.empty((B, cap, 6), dtype=torch.float32, device=boxes.device)
    out[..., :4] = boxes
    out[..., 4] = scores
    out[..., 5] = classes.to(torch.float32)
    return out


def gather_detections(packed: torch.Tensor, counts: torch.Tensor, group=None, out: torch.Tensor = None,
                      out_counts: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather equal-sized shards: packed [b,cap,6] -> [world*b,cap,6], counts [b] i32 -> [world*b] (rank order).
    Two collectives; `gather_packed` below ships the kernel-written payload (counts in its header rows) in ONE."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world * packed.shape[0],) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    if out_counts is None:
        out_counts = torch.empty((world * counts.shape[0],), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(out, packed.contiguous(), group=group)
    dist.all_gather_into_tensor(out_counts, counts.contiguous(), group=group)
    return out, out_counts


def gather_packed(packed: torch.Tensor, group=None, out: torch.Tensor = None, async_op: bool = False):
    """The path's one exchange: ONE all-gather of the [b, cap+1, 6] payload the postprocess kernel
    wrote itself (row 0 of every image = count / overflow flag / K).  -> ([world*b, cap+1, 6] in rank order, work handle)."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world * packed.shape[0],) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    work = dist.all_gather_into_tensor(out, packed, group=group, async_op=async_op)
    return out, work


def unpack_detections(packed: torch.Tensor, counts: torch.Tensor) -> List[Dict[str, torch.Tensor]]:
    """Re-slice the gathered payload by the counts into per-image (boxes, scores, classes) like tools/infer.py produces."""
    res = []
    for b, c in enumerate(counts.cpu().tolist()):
        if c & OVERFLOW_BIT:
            raise RuntimeError(f"image {b}: more detections than the gather capacity {packed.shape[1]}")
        p = packed[b, :c]
        res.append({"boxes": p[:, :4], "scores": p[:, 4], "classes": p[:, 5].to(torch.int64)})
    return res
