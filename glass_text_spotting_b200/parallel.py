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


def gather_sharded(rec: torch.Tensor, n_items: int) -> torch.Tensor:
    """The evaluation loop's gather for a dataset of ``n_items`` images that the ranks split with ``shard_range``
    (shards differ by at most one image): every rank pads its records with empty ones (valid flag 0) to the largest
    shard, ONE all-gather moves them, and the padding is dropped again -> [n_items, max_det, F] in dataset order on every
    rank.  Replaces ``comm.gather`` of pickled per-rank lists + ``itertools.chain`` (glass/evaluation/text_evaluator.py:
    246-252), whose order is the same rank-major order."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    begin, end = shard_range(n_items, rank, world)
    assert rec.shape[0] == end - begin, f"rank {rank} holds {rec.shape[0]} records, its shard is [{begin}, {end})"
    largest = (n_items + world - 1) // world
    if rec.shape[0] < largest:
        rec = torch.cat((rec, rec.new_zeros((largest - rec.shape[0],) + tuple(rec.shape[1:]))), 0)
    g = all_gather_records(rec)                        # [world, largest, max_det, F]
    parts = []
    for r in range(world):
        b, e = shard_range(n_items, r, world)
        parts.append(g[r, : e - b])
    return torch.cat(parts, 0) if parts else g.reshape((0,) + tuple(rec.shape[1:]))


def unpack_detections(rec: torch.Tensor, steps: int = 26, num_classes: int = 97) -> List[Dict[str, torch.Tensor]]:
    """Inverse of B200GlassRCNN.pack_detections for records [..., max_det, 10 + steps*classes]."""
    flat = rec.reshape(-1, rec.shape[-2], rec.shape[-1])
    out = []
    for r in flat:
        k = int(r[:, 0].sum().item())
        out.append({"pred_boxes": r[:k, 1:6], "scores": r[:k, 6], "pred_classes": r[:k, 7].long(),
                    "orientations": r[:k, 8:10], "pred_text_prob": r[:k, 10:].reshape(k, steps, num_classes)})
    return out
