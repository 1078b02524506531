#!/usr/bin/env python
"""One single-head GATConv forward + backward on the multi-mesh [3,5] at the attention config's shape (for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graphcast-lite_b200")]
import torch  # noqa: E402

from gcl_b200 import ops  # noqa: E402
from gcl_b200.graph import CSR_LOOPS, CSRGraph  # noqa: E402
from gcl_b200.graphs_build import ModelGraphs  # noqa: E402

dev = "cuda:0"
mg = ModelGraphs(32, 64, [3, 5], 0.6, dev)
n, B, C = mg.num_mesh, 64, 64
g = CSRGraph(mg.processing_graph, n, CSR_LOOPS)
torch.manual_seed(0)
z = torch.randn(B, n, C, device=dev, requires_grad=True)
a_s = torch.randn(1, 1, C, device=dev, requires_grad=True)
a_d = torch.randn(1, 1, C, device=dev, requires_grad=True)
bia = torch.randn(C, device=dev, requires_grad=True)
w = torch.randn(B, n, C, device=dev)
for _ in range(3):
    out, _ = ops.gat_attend(z, a_s, a_d, bia, g, 1, False, 0.2)
    (out * w).sum().backward()
torch.cuda.synchronize()
print("ok")
