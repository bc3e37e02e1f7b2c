#!/usr/bin/env python
"""Experiment: does splitting a large pinned D2H copy over several streams (copy engines) raise the PCIe rate?"""
import time
import torch

n = 490_266_740  # jac_coord! values of config 2
d = torch.empty(n, dtype=torch.float64, device="cuda")
h = torch.empty(n, dtype=torch.float64, pin_memory=True)
d.zero_()
for parts in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k, s in enumerate(streams):
            a, b = n * k // parts, n * (k + 1) // parts
            with torch.cuda.stream(s):
                h[a:b].copy_(d[a:b], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"parts={parts}: {best*1e3:.1f} ms  {8*n/best/1e9:.1f} GB/s")
