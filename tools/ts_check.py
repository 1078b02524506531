#!/usr/bin/env python
"""Development check of the wide-layer dense kernels (weights-stationary tcgen05 path, umma_linear_ts_kernel):
errors against an fp64 reference and timings at the 512x256 workload's row counts.

  python tools/ts_check.py            # this build
  GCL_UMMA_NO_TS=1 python tools/ts_check.py   # the column-split kernel, for comparison
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graphcast-lite_b200")]

import torch  # noqa: E402

from gcl_b200 import ops  # noqa: E402

PEAK = 6533.5


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def relerr(a, ref):
    return float((a.double() - ref).abs().max() / ref.abs().max())


def main():
    dev = "cuda:0"
    torch.manual_seed(0)
    print("mode:", "column-split (GCL_UMMA_NO_TS=1)" if os.environ.get("GCL_UMMA_NO_TS") == "1" else "weights-stationary")
    worst = 0.0
    for (R, cin, cout) in [(1000, 128, 128), (64, 128, 128), (12345, 96, 96), (777, 64, 128), (5000, 128, 96),
                           (300, 32, 72), (4097, 128, 100), (3000, 64, 64), (5000, 128, 64), (2100, 64, 32), (4000, 64, 20), (4000, 20, 64), (3000, 12, 8), (50000, 20, 128), (200000, 128, 128)]:
        x = torch.randn(R, cin, device=dev)
        W = torch.randn(cout, cin, device=dev) / cin ** 0.5
        b = torch.randn(cout, device=dev)
        slope = torch.tensor([0.25], device=dev)
        ref_z = x.double() @ W.double().t() + b.double()
        ref_y = torch.where(ref_z > 0, ref_z, 0.25 * ref_z)
        y, z = ops.linear_fwd_raw(x, W, b, slope, want_z=True)
        e1, e2 = relerr(z, ref_z), relerr(y, ref_y)
        y2, _ = ops.linear_fwd_raw(x, W, None, None, want_z=False)
        e3 = relerr(y2, x.double() @ W.double().t())
        dy = torch.randn(R, cout, device=dev)
        dx = ops.linear_bwd_dx_raw(dy, W)
        ref_dx = dy.double() @ W.double()
        e4 = relerr(dx, ref_dx)
        zin = torch.randn(R, cin, device=dev)
        dz, dsl, dcs = ops.linear_bwd_dx_prelu_raw(dy, W, zin, slope, True)
        ref_dz = torch.where(zin > 0, ref_dx, 0.25 * ref_dx)
        ref_dsl = (ref_dx * torch.where(zin > 0, torch.zeros_like(ref_dx), zin.double())).sum()
        e5 = relerr(dz, ref_dz)
        e6 = float((dsl.double() - ref_dsl).abs() / ref_dsl.abs().clamp_min(1e-30))
        e9 = relerr(dcs, ref_dz.sum(0))
        print(f"   dx_prelu colsum {e9:.2e}")
        worst = max(worst, e9)
        dW, db = ops.linear_bwd_dw_raw(dy, x, True)
        e7 = relerr(dW, dy.double().t() @ x.double())
        e8 = relerr(db, dy.double().sum(0))
        print(f"   dW {e7:.2e} dbias {e8:.2e}")
        worst = max(worst, e7, e8)
        print(f"R{R} {cin}->{cout}: z {e1:.2e} y {e2:.2e} plain {e3:.2e} dx {e4:.2e} dx_prelu {e5:.2e} dslope {e6:.2e}", flush=True)
        worst = max(worst, e1, e2, e3, e4, e5)
    print("worst rel err", worst)
    assert worst < 5e-6, worst
    for (R, cin, cout) in [(1376272, 64, 20), (1376272, 20, 64)]:
        x = torch.randn(R, cin, device=dev)
        W = torch.randn(cout, cin, device=dev) / cin ** 0.5
        b = torch.randn(cout, device=dev)
        dy = torch.randn(R, cout, device=dev)
        slope = torch.tensor([0.25], device=dev)
        nb = 4 * R * (cin + cout)
        for name, fn in [("fwd", lambda: ops.linear_fwd_raw(x, W, b, None, False)), ("dx", lambda: ops.linear_bwd_dx_raw(dy, W)),
                         ("dW+dbias", lambda: ops.linear_bwd_dw_raw(dy, x, True))]:
            us = timeit(fn)
            print(f"R{R} {cin}->{cout} {name:10s} {us:8.1f} us  {nb / us / 1e3:7.0f} GB/s  {nb / us / 1e3 / PEAK:5.2f}", flush=True)
    for (R, C) in [(1376272, 128), (327696, 128), (2752544, 96), (1376272, 64), (786560, 64)]:
        x = torch.randn(R, C, device=dev)
        W = torch.randn(C, C, device=dev) / C ** 0.5
        b = torch.randn(C, device=dev)
        slope = torch.tensor([0.25], device=dev)
        zin = torch.randn(R, C, device=dev)
        n1 = 4 * R * 2 * C
        for name, fn, nb in [
            ("fwd", lambda: ops.linear_fwd_raw(x, W, b, None, False), n1),
            ("fwd+prelu", lambda: ops.linear_fwd_raw(x, W, b, slope, False), n1),
            ("fwd+prelu+z", lambda: ops.linear_fwd_raw(x, W, b, slope, True), 4 * R * 3 * C),
            ("dx", lambda: ops.linear_bwd_dx_raw(x, W), n1),
            ("dx_prelu", lambda: ops.linear_bwd_dx_prelu_raw(x, W, zin, slope), 4 * R * 3 * C),
            ("dx_prelu+cs", lambda: ops.linear_bwd_dx_prelu_raw(x, W, zin, slope, True), 4 * R * 3 * C),
            ("dW+dbias", lambda: ops.linear_bwd_dw_raw(x, zin, True), n1),
        ]:
            us = timeit(fn)
            print(f"R{R}xC{C} {name:12s} {us:8.1f} us  {nb / us / 1e3:7.0f} GB/s  {nb / us / 1e3 / PEAK:5.2f}", flush=True)


if __name__ == "__main__":
    main()
