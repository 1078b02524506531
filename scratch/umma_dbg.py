import sys, ctypes
sys.path[:0] = ["/root/repo", "/root/repo/graphcast-lite_b200"]
import torch
from gcl_b200 import ops, _cabi
lib = _cabi.load()
sys.path.insert(0, "/root/repo/tools")
from kbench import timeit
for (R, cin, cout) in [(786560, 64, 64), (786560, 128, 128)]:
    x = torch.randn(R, cin, device="cuda"); W = torch.randn(cout, cin, device="cuda"); b = torch.randn(cout, device="cuda")
    for mask, name in [(0, "full"), (1, "no epilogue stores"), (2, "no A loads"), (4, "no MMA"), (3, "no loads, no stores"), (7, "nothing but sync")]:
        lib.gcl_debug_set_umma_mask(mask)
        us = timeit(lambda: ops.linear_fwd_raw(x, W, b))
        print(f"{cin}->{cout} {name:22s} {us:8.1f} us", flush=True)
lib.gcl_debug_set_umma_mask(0)
