"""Full-scene cross-tile merge NMS, sharded by class over the ranks of one node.

Replaces the CPU stage `mergebypoly` -> `Pool(16).map(mergesingle)` -> `py_cpu_nms_poly_fast`
(python/jdet/data/devkits/result_merge.py:206-264, 66-127; `Pool(37)` in tools/merge_results.py:38-45).
Suppression is independent per (scene, class) (one file per class, :45-48 of data_merge.py; one NMS
call per scene, result_merge.py:177-193), so:

  * every rank takes a subset of the classes, chosen by longest-processing-time bin packing on the
    per-class pair counts (Vehicle / Ship dominate FAIR1M; a round-robin split would leave most GPUs
    idle);
  * each rank runs ONE batched device NMS over all (scene, class) groups it owns
    (rs_detection_b200.core.nms with the float64 merge predicate and per-group thresholds);
  * ONE exchange step returns the survivors to every rank: an all-gather of the per-rank counts and a
    padded all-gather of the kept row indices (NCCL over NVLink on the GPU box; the same code runs over
    gloo with CPU tensors in the tests).  Payload is KB-MB: latency bound.

The NMS function is injected (`nms_fn`) so that the host logic (planning, padding, gather, ordering)
is testable on CPU with world_size 2 over gloo; the product default is the CUDA engine and raises
without a GPU.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def plan_class_shards(class_counts: Sequence[int], world: int) -> list:
    """LPT bin packing of classes onto ranks.  Cost model: n*(n-1)/2 candidate pairs + n (sort/scan).
    Deterministic (ties by class id), so every rank computes the same plan from the same counts."""
    cost = [(c * (c - 1) // 2 + c, k) for k, c in enumerate(class_counts)]
    cost.sort(key=lambda t: (-t[0], t[1]))
    load = [0] * world
    owner = [0] * len(class_counts)
    for c, k in cost:
        r = min(range(world), key=lambda i: (load[i], i))
        owner[k] = r
        load[r] += c
    return owner


def _device_nms(polys, scores, groups, thr, group_thr):
    from . import core
    from ._lib import NMS_MERGE
    res = core.nms(NMS_MERGE, polys, scores, float(thr), labels=groups, thr_per_label=group_thr, want_mask=False,
                   want_sorted=False, want_score=True, ws_tag="merge")
    return res.score_idx


def merge_sharded(polys: torch.Tensor, scores: torch.Tensor, labels: torch.Tensor, scene_ids: Optional[torch.Tensor] = None,
                  thr: float = 0.1, class_thr: Optional[Sequence[float]] = None, num_classes: Optional[int] = None,
                  group=None, nms_fn: Optional[Callable] = None) -> torch.Tensor:
    """Class-sharded merge NMS of a replicated detection set.

    polys (n,8) float64 scene coordinates, scores (n,), labels (n,) class ids, scene_ids (n,) optional
    (several scenes in one call); all ranks hold the same tensors (on the GPU box: CUDA tensors).
    Returns the kept row indices (int64, same device), identical on every rank, ordered by
    (class, scene, descending score) -- the order of the reference's per-class output files.
    """
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    dev = polys.device
    n = polys.shape[0]
    labels = labels.to(torch.int64)
    C = int(num_classes) if num_classes is not None else (int(labels.max().item()) + 1 if n else 1)
    counts = torch.bincount(labels, minlength=C).cpu().tolist() if n else [0] * C
    owner = plan_class_shards(counts, world)
    mine = torch.tensor([owner[c] == rank for c in range(C)], device=dev)
    sel = torch.nonzero(mine[labels])[:, 0] if n else torch.zeros((0,), dtype=torch.int64, device=dev)

    if scene_ids is None:
        scene = torch.zeros((n,), dtype=torch.int64, device=dev)
        S = 1
    else:
        scene = scene_ids.to(torch.int64)
        S = int(scene.max().item()) + 1 if n else 1
    groups_all = labels * S + scene  # one NMS group per (class, scene)
    group_thr = None
    if class_thr is not None:
        group_thr = torch.as_tensor(np.repeat(np.asarray(class_thr, np.float64), S), device=dev)

    fn = nms_fn or _device_nms
    if sel.numel():
        keep_local = fn(polys[sel], scores[sel], groups_all[sel].to(torch.int32), thr, group_thr)
        kept = sel[keep_local.to(torch.int64)]
    else:
        kept = torch.zeros((0,), dtype=torch.int64, device=dev)

    if world > 1:
        cnt = torch.tensor([kept.numel()], dtype=torch.int64, device=dev)
        cnts = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(cnts, cnt, group=group)
        sizes = [int(c.item()) for c in cnts]
        pad = max(max(sizes), 1)
        buf = torch.full((pad,), -1, dtype=torch.int64, device=dev)
        buf[: kept.numel()] = kept
        bufs = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(bufs, buf, group=group)
        kept = torch.cat([b[:s] for b, s in zip(bufs, sizes)])
    # canonical order: (class, scene) ascending, score descending
    if kept.numel():
        key_g = groups_all[kept]
        order = torch.argsort(-scores[kept].to(torch.float64), stable=True)
        order = order[torch.argsort(key_g[order], stable=True)]
        kept = kept[order]
    return kept
