#!/usr/bin/env python
"""Pinned host -> device copy bandwidth with default vs write-combined pinned memory (cudaHostAlloc flags 0 / 4):
python tools/h2d_wc.py"""
import ctypes
import torch
rt = ctypes.CDLL("libcudart.so.12")
torch.cuda.init()
d = torch.empty(716 * 1024 * 1024, dtype=torch.uint8, device="cuda")
n = d.numel()
for name, flag in (("default", 0), ("write-combined", 4), ("portable|mapped", 3)):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flag))
    assert rc == 0, rc
    ctypes.memset(p, 1, n)
    st = torch.cuda.current_stream().cuda_stream
    def cp():
        return rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), p, ctypes.c_size_t(n), ctypes.c_int(1), ctypes.c_void_p(st))
    for _ in range(2):
        cp()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(8):
        assert cp() == 0
    b.record()
    torch.cuda.synchronize()
    print(f"{name}: {8 * n / 1e9 / (a.elapsed_time(b) * 1e-3):.2f} GB/s")
    rt.cudaFreeHost(p)
