import sys, os
sys.path[:0] = ["/root/repo", "/root/repo/graphcast-lite_b200"]
import torch
from gcl_b200 import ops, _cabi
lib = _cabi.load()
torch.manual_seed(0)
def rel(a, b): return float((a.double()-b.double()).abs().max() / b.double().abs().max())
for (R, cin, cout) in [(2048, 64, 64), (4096, 64, 64), (5000, 128, 128), (3000, 72, 48), (2500, 66, 64), (2500, 64, 33), (3333, 30, 48), (2100, 96, 15), (40000, 128, 64), (2048, 200, 256), (100000, 64, 64)]:
    x = torch.randn(R, cin, device="cuda"); W = torch.randn(cout, cin, device="cuda") / cin**0.5; b = torch.randn(cout, device="cuda"); a = torch.tensor([0.25], device="cuda")
    ref = torch.nn.functional.linear(x.double(), W.double(), b.double())
    refp = torch.where(ref > 0, ref, ref * 0.25)
    out = {}
    for mode in (1, 0):
        lib.gcl_set_dense_mode(mode)
        y, z = ops.linear_fwd_raw(x, W, b, a, True)
        dy = torch.randn(R, cout, device="cuda")
        dx = ops.linear_bwd_dx_raw(dy, W)
        torch.cuda.synchronize()
        out[mode] = (rel(y, refp), rel(z, ref), rel(dx, dy.double() @ W.double()))
    print(f"R={R:6d} {cin:3d}->{cout:3d}  ffma y/z/dx {out[1][0]:.1e} {out[1][1]:.1e} {out[1][2]:.1e} | umma y/z/dx {out[0][0]:.1e} {out[0][1]:.1e} {out[0][2]:.1e}", flush=True)
