#!/usr/bin/env python
"""Development aid: DRAM throughput of scattered small reads (the access pattern of the cell
overlay) against streaming, measured with torch gathers over a footprint far larger than L2."""
import torch

dev = torch.device("cuda:0")
GB = 3.5
for row_bytes in (32, 64, 128, 256):
    cols = row_bytes // 4
    n = int(GB * 1e9) // row_bytes
    x = torch.empty((n, cols), dtype=torch.float32, device=dev).normal_()
    m = n // 4
    idx = torch.randint(0, n, (m,), device=dev)
    out = torch.empty((m, cols), dtype=torch.float32, device=dev)
    for _ in range(2):
        torch.index_select(x, 0, idx, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        torch.index_select(x, 0, idx, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("rows of %3d B: %.3f ms for %d random rows -> %.0f M rows/s, read %.0f GB/s (+ %.0f GB/s sequential write)"
          % (row_bytes, ms, m, m / ms / 1e3, m * row_bytes / ms / 1e6, m * row_bytes / ms / 1e6))
    del x, out, idx
