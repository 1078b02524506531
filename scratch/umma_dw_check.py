import sys
sys.path[:0] = ["/root/repo", "/root/repo/graphcast-lite_b200"]
import torch
from gcl_b200 import ops, _cabi
lib = _cabi.load()
torch.manual_seed(0)
def rel(a, b): return float((a.double()-b.double()).abs().max() / b.double().abs().max())
for (R, cin, cout) in [(2048, 64, 64), (5000, 128, 128), (3000, 72, 48), (2500, 66, 64), (2500, 64, 33), (3333, 30, 48), (2100, 96, 15), (40000, 128, 64), (100001, 64, 64), (2048, 200, 128)]:
    x = torch.randn(R, cin, device="cuda"); dy = torch.randn(R, cout, device="cuda")
    refW = dy.double().T @ x.double(); refb = dy.double().sum(0)
    res = {}
    for mode in (1, 0):
        lib.gcl_set_dense_mode(mode)
        dW, db = ops.linear_bwd_dw_raw(dy, x, True)
        torch.cuda.synchronize()
        res[mode] = (rel(dW, refW), rel(db, refb))
    print(f"R={R:6d} dy[{cout}] x[{cin}]  ffma dW/db {res[1][0]:.1e} {res[1][1]:.1e} | umma dW/db {res[0][0]:.1e} {res[0][1]:.1e}", flush=True)
