#!/usr/bin/env python
"""Launches the wide-layer dense kernels once each at the 512x256 workload's shape (for an ncu capture)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graphcast-lite_b200")]
import torch  # noqa: E402

from gcl_b200 import ops  # noqa: E402

R, C = 1376272, 128
dev = "cuda:0"
torch.manual_seed(0)
x = torch.randn(R, C, device=dev)
W = torch.randn(C, C, device=dev) / C ** 0.5
b = torch.randn(C, device=dev)
slope = torch.tensor([0.25], device=dev)
zin = torch.randn(R, C, device=dev)
for _ in range(2):
    ops.linear_fwd_raw(x, W, b, slope, False)
    ops.linear_fwd_raw(x, W, b, slope, True)
    ops.linear_bwd_dx_prelu_raw(x, W, zin, slope)
    ops.linear_bwd_dw_raw(x, zin, True)
torch.cuda.synchronize()
print("ok")
