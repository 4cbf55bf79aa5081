"""Image-sharded multi-GPU inference plumbing: one process per GPU, contiguous image
ranges per rank (detectron2 InferenceSampler semantics, glass/data/build.py:99), and ONE all-gather of
fixed-size packed detection records replacing the reference's gloo gather of pickled lists
(glass/evaluation/text_evaluator.py:246-249).  No data-path collective exists anywhere else."""
from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of rank's items; earlier ranks take the remainder (InferenceSampler)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_gather_records(rec: torch.Tensor) -> torch.Tensor:
    """rec [n_local, max_det, F] (same shape on every rank) -> [world, n_local, max_det, F]."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rec.unsqueeze(0)
    world = dist.get_world_size()
    out = torch.empty((world * rec.shape[0],) + tuple(rec.shape[1:]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous())  # concatenated along dim 0 (works for nccl and gloo)
    return out.view((world,) + tuple(rec.shape))


def unpack_detections(rec: torch.Tensor, steps: int = 26, num_classes: int = 97) -> List[Dict[str, torch.Tensor]]:
    """Inverse of B200GlassRCNN.pack_detections for records [..., max_det, 10 + steps*classes]."""
    flat = rec.reshape(-1, rec.shape[-2], rec.shape[-1])
    out = []
    for r in flat:
        k = int(r[:, 0].sum().item())
        out.append({"pred_boxes": r[:k, 1:6], "scores": r[:k, 6], "pred_classes": r[:k, 7].long(),
                    "orientations": r[:k, 8:10], "pred_text_prob": r[:k, 10:].reshape(k, steps, num_classes)})
    return out
