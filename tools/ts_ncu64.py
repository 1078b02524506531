import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graphcast-lite_b200")]
import torch
from gcl_b200 import ops
R, C = 786560, 64
dev = "cuda:0"
torch.manual_seed(0)
x = torch.randn(R, C, device=dev); W = torch.randn(C, C, device=dev) / 8; zin = torch.randn(R, C, device=dev)
slope = torch.tensor([0.25], device=dev)
for _ in range(3):
    ops.linear_bwd_dx_prelu_raw(x, W, zin, slope)
torch.cuda.synchronize()
print("ok")
