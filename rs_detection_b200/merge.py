"""Full-scene cross-tile merge NMS, sharded by class over the ranks of one node.

Replaces the CPU stage `mergebypoly` -> `Pool(16).map(mergesingle)` -> `py_cpu_nms_poly_fast`
(python/jdet/data/devkits/result_merge.py:206-264, 66-127; `Pool(37)` in tools/merge_results.py:38-45).
Suppression is independent per (scene, class) (one file per class, :45-48 of data_merge.py; one NMS
call per scene, result_merge.py:177-193), so:

  * every rank takes a subset of the classes, chosen by longest-processing-time bin packing on the
    per-class pair counts (Vehicle / Ship dominate FAIR1M; a round-robin split would leave most GPUs
    idle).  The per-class counts are host knowledge in the pipeline (one `before_nms/<Class>.txt` per class:
    its line count) and are passed in as `class_counts`; only when they are missing does the function pay one
    device->host copy for a histogram;
  * each rank runs ONE batched device NMS over all (scene, class) groups it owns
    (rs_detection_b200.core.nms with the float64 merge predicate and per-group thresholds);
  * ONE collective returns the survivors to every rank: `all_gather_into_tensor` of a fixed-capacity int32
    buffer per rank -- slot 0 holds the rank's survivor count, the rest its kept row indices in descending
    score order, padded with -1.  The capacity (the largest shard) follows from the class counts, so no size
    exchange is needed (NCCL over NVLink on the GPU box; the same code runs over gloo with CPU tensors in
    the tests).  Payload is KB-MB: latency bound;
  * no host synchronisation anywhere between the inputs and the result: shard selection is a prefix sum +
    scatter into a buffer whose size the plan already fixed, the canonical output order is one stable device
    sort of the gathered buffer by group id (each rank's run is already score-descending).  The caller pays
    one device->host copy when it reads the result (`MergeResult.indices()`).

The NMS function is injected (`nms_fn`) so that the host logic (planning, selection, gather, ordering)
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


class MergeResult:
    """`padded` (capacity,) int64 kept row indices in canonical order -- (class, scene) ascending, score
    descending, the order of the reference's per-class output files -- followed by -1 padding; `count` (0-dim
    int64 tensor on the same device).  Identical on every rank.  `indices()` synchronises once and trims."""

    def __init__(self, padded: torch.Tensor, count: torch.Tensor):
        self.padded, self.count = padded, count

    def indices(self) -> torch.Tensor:
        return self.padded[: int(self.count.item())]


def _device_nms(polys, scores, groups, thr, group_thr):
    """-> (kept positions in descending score order, (n,) int64 with a garbage tail; count (1,) int32)"""
    from . import core
    from ._lib import NMS_MERGE
    res = core.nms(NMS_MERGE, polys, scores, float(thr), labels=groups, thr_per_label=group_thr, want_mask=False,
                   want_sorted=False, want_score=True, ws_tag="merge")
    return res.score_idx_padded, res.num_keep


def merge_sharded(polys: torch.Tensor, scores: torch.Tensor, labels: torch.Tensor, scene_ids: Optional[torch.Tensor] = None,
                  thr: float = 0.1, class_thr: Optional[Sequence[float]] = None, num_classes: Optional[int] = None,
                  group=None, nms_fn: Optional[Callable] = None, class_counts: Optional[Sequence[int]] = None,
                  num_scenes: Optional[int] = None) -> MergeResult:
    """Class-sharded merge NMS of a replicated detection set.

    polys (n,8) float64 scene coordinates, scores (n,), labels (n,) class ids, scene_ids (n,) optional
    (several scenes in one call); all ranks hold the same tensors (on the GPU box: CUDA tensors).
    `class_counts` (host ints per class) and `num_scenes` make the call free of host synchronisation.
    """
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    dev = polys.device
    n = polys.shape[0]
    labels = labels.to(torch.int64)
    if class_counts is None:      # one device->host copy; the pipeline knows these from the per-class files
        C = int(num_classes) if num_classes is not None else (int(labels.max().item()) + 1 if n else 1)
        class_counts = torch.bincount(labels, minlength=C).cpu().tolist() if n else [0] * C
    class_counts = [int(c) for c in class_counts]
    C = len(class_counts)
    owner = plan_class_shards(class_counts, world)
    shard = [sum(c for c, o in zip(class_counts, owner) if o == r) for r in range(world)]
    m, cap = shard[rank], max(max(shard), 1)

    if scene_ids is None:
        scene, S = None, 1
    else:
        scene = scene_ids.to(torch.int64)
        S = int(num_scenes) if num_scenes is not None else (int(scene.max().item()) + 1 if n else 1)
    groups_all = labels * S + scene if scene is not None else labels      # one NMS group per (class, scene)
    group_thr = None
    if class_thr is not None:
        group_thr = torch.as_tensor(np.repeat(np.asarray(class_thr, np.float64), S), device=dev)

    # this rank's rows, in original order, without a host round trip: the plan fixes their number (m)
    payload = torch.full((cap + 1,), -1, dtype=torch.int32, device=dev)
    payload[0] = 0
    if m:
        mine = torch.as_tensor([o == rank for o in owner], device=dev)[labels]
        pos = torch.cumsum(mine, 0) - 1
        slot = torch.where(mine, pos, torch.full_like(pos, m))
        sel = torch.empty((m + 1,), dtype=torch.int64, device=dev).scatter_(0, slot, torch.arange(n, device=dev))[:m]
        fn = nms_fn or _device_nms
        keep_local, num = fn(polys[sel], scores[sel], groups_all[sel].to(torch.int32), thr, group_thr)
        num = num.reshape(-1)[:1].to(torch.int32)
        live = torch.arange(m, device=dev) < num
        kept = sel[keep_local.to(torch.int64).clamp_(0, m - 1)]
        payload[1:m + 1] = torch.where(live, kept, torch.full_like(kept, -1)).to(torch.int32)
        payload[:1] = num

    if world > 1:
        gathered = torch.empty((world * (cap + 1),), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(gathered, payload, group=group)      # the ONE collective of the merge stage
        gathered = gathered.view(world, cap + 1)
    else:
        gathered = payload.view(1, cap + 1)
    count = gathered[:, 0].sum(dtype=torch.int64)
    flat = gathered[:, 1:].reshape(-1).to(torch.int64)
    # canonical order: every rank's run is score-descending and ranks own disjoint groups, so one stable sort by
    # group id (padding last) yields (class, scene) ascending, score descending
    key = torch.where(flat >= 0, groups_all[flat.clamp(min=0)] if n else flat, torch.full_like(flat, C * S))
    order = torch.sort(key, stable=True).indices
    return MergeResult(flat[order], count)
