"""N>1 path of the merge stage on CPU: world_size 2 over gloo.  The host logic under test (class-shard
plan, sync-free shard selection, the single fixed-capacity all-gather of survivors, canonical ordering) is the
product's; the per-shard NMS is injected and played by the CPU oracle here (on the GPU box the default is the
CUDA engine, covered by tests/test_gpu_merge.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import workloads as W


def _oracle_nms(polys, scores, groups, thr, group_thr):
    from oracle import oracle as O
    p, s, g = polys.numpy(), scores.numpy(), groups.numpy()
    keep = []
    for gid in np.unique(g):
        idx = np.nonzero(g == gid)[0]
        t = float(group_thr[gid]) if group_thr is not None else thr
        dets = np.concatenate([p[idx], s[idx, None]], 1)
        keep += [int(idx[k]) for k in O.py_cpu_nms_poly_fast(dets, t)]
    keep.sort(key=lambda i: -s[i])
    out = torch.full((len(s),), 12345, dtype=torch.int64)        # garbage tail like the engine's fixed-size output
    out[: len(keep)] = torch.tensor(keep, dtype=torch.int64)
    return out, torch.tensor([len(keep)], dtype=torch.int32)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rs_detection_b200.jdet.data.devkits.result_merge import nms_threshold_1
    from rs_detection_b200.merge import merge_sharded
    sc = W.merge_scene(num_objects=250, scene=2500, seed=6)
    thr = [nms_threshold_1[c] for c in W.FAIR1M_CLASSES]
    scene_ids = torch.from_numpy((np.arange(sc["scores"].size) % 2).astype(np.int64))  # two scenes in one call
    counts = np.bincount(sc["labels"], minlength=10).tolist()    # host knowledge in the pipeline (per-class files)
    calls = []
    real = dist.all_gather_into_tensor
    dist.all_gather_into_tensor = lambda *a, **k: (calls.append(1), real(*a, **k))[1]
    res = merge_sharded(torch.from_numpy(sc["polys"]), torch.from_numpy(sc["scores"]), torch.from_numpy(sc["labels"]),
                        scene_ids, class_thr=thr, nms_fn=_oracle_nms, class_counts=counts, num_scenes=2)
    dist.all_gather_into_tensor = real
    assert len(calls) == 1, "the merge stage must use exactly one collective"
    kept = res.indices()
    assert int((res.padded >= 0).sum()) == int(res.count) == kept.numel()
    q.put((rank, kept.tolist()))
    # without the host-side hints the result is the same (one histogram copy instead)
    res2 = merge_sharded(torch.from_numpy(sc["polys"]), torch.from_numpy(sc["scores"]), torch.from_numpy(sc["labels"]),
                         scene_ids, class_thr=thr, num_classes=10, nms_fn=_oracle_nms)
    assert res2.indices().tolist() == kept.tolist()
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(300)
def test_class_sharded_merge_world2_matches_single_process():
    from oracle import oracle as O
    from rs_detection_b200.jdet.data.devkits.result_merge import nms_threshold_1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0] == res[1], "ranks disagree after the all-gather"
    # single-process truth: one reference-style call per (class, scene)
    sc = W.merge_scene(num_objects=250, scene=2500, seed=6)
    scene = np.arange(sc["scores"].size) % 2
    want = []
    for c in range(10):
        for s_ in range(2):
            idx = np.nonzero((sc["labels"] == c) & (scene == s_))[0]
            dets = np.concatenate([sc["polys"][idx], sc["scores"][idx, None]], 1)
            want += [int(idx[k]) for k in O.py_cpu_nms_poly_fast(dets, nms_threshold_1[W.FAIR1M_CLASSES[c]])]
    assert res[0] == want  # same survivors in the canonical (class, scene, score) order
    assert 0 < len(want) < sc["scores"].size
