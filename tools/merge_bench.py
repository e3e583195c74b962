#!/usr/bin/env python
"""Full-scene merge NMS (BASELINE config 5), class-sharded over the ranks of one node.
    python tools/merge_bench.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/merge_bench.py
Every rank builds the same synthetic 10k x 10k scene; rank r runs the device NMS for its classes; one
all-gather (NCCL) returns the survivors to everyone.  Rank 0 prints one JSON line (device time, max over
ranks) and checks the result against the single-GPU engine call."""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as W
from rs_detection_b200.jdet.data.devkits.result_merge import merge_detections, nms_threshold_1
from rs_detection_b200.merge import merge_sharded, plan_class_shards

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nobj = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
sc = W.merge_scene(num_objects=nobj, scene=10000, seed=1)
thr = [nms_threshold_1[c] for c in W.FAIR1M_CLASSES]
p, s, l = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (sc["polys"], sc["scores"], sc["labels"])]


def run():
    return merge_sharded(p, s, l, class_thr=thr, num_classes=10)


for _ in range(2):
    kept = run()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
best = 1e30
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); kept = run(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
t = torch.tensor([best], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ref = merge_detections(p, s, l, group_thresh=thr)  # single-GPU engine call, all classes in one launch
    same = sorted(ref.tolist()) == sorted(kept.tolist())
    counts = np.bincount(sc["labels"], minlength=10).tolist()
    print(json.dumps({"workload": "config 5: 10k x 10k scene, 504 tiles (1024/200 @ 0.5/1.0/1.5)", "objects": nobj,
                      "detections": int(p.shape[0]), "n_gpus": world, "ms": float(t.item()),
                      "boxes_per_s": p.shape[0] / float(t.item()) * 1e3, "kept": int(kept.numel()),
                      "matches_single_gpu": bool(same), "class_counts": counts, "owner": plan_class_shards(counts, world)}))
if world > 1:
    dist.destroy_process_group()
