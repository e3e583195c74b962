#!/usr/bin/env python
"""The e2e leg's upload pattern in isolation: 8 slabs of 89.7 MB per step, torch-pinned vs cudaHostAlloc'd, one copy
stream vs two alternating ones: python tools/h2d_slabs.py"""
import ctypes
import numpy as np
import torch
rt = ctypes.CDLL("libcudart.so.12")
torch.cuda.init()
n = 22370000  # floats per slab (~89.5 MB)
dev = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(2)]


def host_alloc(n_floats):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n_floats * 4), ctypes.c_uint(0)) == 0
    arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(n_floats,))
    t = torch.from_numpy(arr)
    t.fill_(1.0)
    return t


for kind in ("torch pin_memory", "cudaHostAlloc"):
    slabs = [torch.ones(n, dtype=torch.float32).pin_memory() if kind.startswith("torch") else host_alloc(n) for _ in range(8)]
    print(kind, "is_pinned:", slabs[0].is_pinned())
    for nstreams in (1, 2):
        ss = [torch.cuda.Stream() for _ in range(nstreams)]
        def step():
            for i, h in enumerate(slabs):
                with torch.cuda.stream(ss[i % nstreams]):
                    dev[i % 2].copy_(h, non_blocking=True)
        step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in ss:
            s.wait_event(a)
        for _ in range(5):
            step()
        for s in ss:
            torch.cuda.current_stream().wait_stream(s)
        b.record()
        torch.cuda.synchronize()
        print(f"  {nstreams} copy stream(s): {5 * 8 * n * 4 / 1e9 / (a.elapsed_time(b) * 1e-3):.2f} GB/s")
