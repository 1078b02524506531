#!/usr/bin/env python
"""Per-launch cost of back-to-back kernels inside a CUDA graph (development aid): K identical calls captured in one
graph, replay time / K, for a persistent tcgen05 kernel, the tiled aggregation engine and a trivial kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graphcast-lite_b200")]
import torch  # noqa: E402

from gcl_b200 import ops  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)


def graph_time(fn, K=20, reps=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(K):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / K * 1e3


for R in (1376272, 327696, 40960, 4096):
    C = 128
    x = torch.randn(R, C, device=dev)
    W = torch.randn(C, C, device=dev) / C ** 0.5
    b = torch.randn(C, device=dev)
    y = torch.empty_like(x)
    us = graph_time(lambda: ops.linear_fwd_raw(x, W, b, None, False))
    print(f"linear fwd 128->128 R={R}: {us:8.1f} us per launch in a graph of 20 ({4 * R * 2 * C / us / 1e3:6.0f} GB/s)", flush=True)
    us = graph_time(lambda: ops.linear_bwd_dw_raw(x, y, True))
    print(f"dW 128x128 R={R}:          {us:8.1f} us per call (3 kernels)", flush=True)
    us = graph_time(lambda: torch.add(x, 1.0, out=y))
    print(f"torch add R={R}:            {us:8.1f} us per launch ({4 * R * 2 * C / us / 1e3:6.0f} GB/s)", flush=True)
