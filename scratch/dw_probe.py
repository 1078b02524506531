import sys
sys.path[:0] = ["/root/repo", "/root/repo/graphcast-lite_b200", "/root/repo/tools"]
import torch
from gcl_b200 import ops, _cabi
from kbench import timeit
lib = _cabi.load()
R = 786560
for (cout, cin) in [(128, 12), (128, 60), (128, 124), (16, 124), (64, 60), (64, 28)]:
    x = torch.randn(R, cin, device="cuda"); dy = torch.randn(R, cout, device="cuda")
    for mode in (0, 1):
        lib.gcl_set_dense_mode(mode)
        us = timeit(lambda: ops.linear_bwd_dw_raw(dy, x, True))
        print(f"dW dy[{cout}] x[{cin}] mode={'umma' if mode==0 else 'ffma'}: {us:8.1f} us   {4*R*(cin+cout)/us/1e3:7.0f} GB/s", flush=True)
