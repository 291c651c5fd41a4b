"""Fixed-size per-image result records and their cross-rank gather.

The inference path shards images across GPUs with no collective inside forward; the only exchange is
one all-gather of compact per-image records after predict (the B200 counterpart of mmengine's
``collect_results`` after ``CocoMetric.process``, mmdet/evaluation/metrics/coco_metric.py:346-400).
A record row is (x1, y1, x2, y2, score, label); each image has ``max_per_img`` rows plus a count."""
from __future__ import annotations

import torch
import torch.distributed as dist


def pack_records(bboxes: torch.Tensor, scores: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """[B,M,4], [B,M], [B,M] -> fp32 [B, M, 6]."""
    return torch.cat([bboxes.float(), scores.float()[..., None], labels.float()[..., None]], dim=2).contiguous()


def gather_records(records: torch.Tensor, counts: torch.Tensor):
    """All-gather [B, M, 6] records and int32 [B] counts over the default process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return records, counts
    world = dist.get_world_size()
    out = torch.empty((world * records.shape[0],) + tuple(records.shape[1:]), dtype=records.dtype, device=records.device)
    cnt = torch.empty(world * counts.shape[0], dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(out, records.contiguous())
    dist.all_gather_into_tensor(cnt, counts.contiguous())
    return out, cnt


def gather_mask_logits(mask_logits: torch.Tensor) -> torch.Tensor:
    """Optional second part of the record: the low-resolution mask logits [B*M, h, w] as fp16, gathered
    over the default process group; the consumer resizes / thresholds them.  100 x 256^2 fp16 = 13 MB per image."""
    x = mask_logits.to(torch.float16).contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    out = torch.empty((dist.get_world_size() * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x)
    return out


def unpack_records(records: torch.Tensor, counts: torch.Tensor) -> list:
    out = []
    for r, n in zip(records, counts.tolist()):
        out.append(dict(bboxes=r[:n, :4], scores=r[:n, 4], labels=r[:n, 5].long()))
    return out
