#!/usr/bin/env python
"""Pinned host -> device copy bandwidth of the box (the ceiling of bench.py's e2e leg): python tools/h2d_peak.py"""
import torch
for mb in (64, 256, 716):
    h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device="cuda")
    for streams in (1, 2):
        ss = [torch.cuda.Stream() for _ in range(streams)]
        chunks = h.chunk(streams), d.chunk(streams)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            for s, hc, dc in zip(ss, *chunks):
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    dc.copy_(hc, non_blocking=True)
            for s in ss:
                torch.cuda.current_stream().wait_stream(s)
        b.record()
        torch.cuda.synchronize()
        print(f"{mb} MB, {streams} stream(s): {5 * mb / 1024 / (a.elapsed_time(b) * 1e-3):.1f} GiB/s = {5 * mb * 1.048576 / (a.elapsed_time(b)):.1f} GB/s")
