#!/usr/bin/env python
"""Kernel micro-benchmarks (CUDA events, inputs >> L2 or rotated) for the gcl_b200 aggregate / dense
kernels at the BASELINE shapes.  Development aid; the judged numbers come from bench.py.

  python tools/kbench.py [spmm] [gat] [linear] [--mesh 35|46] [--grid 64x32|512x256] [-B 64] [-C 64]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graphcast-lite_b200")]

import torch  # noqa: E402

from gcl_b200 import ops  # noqa: E402
from gcl_b200.graph import CSR_LOOPS, NORM_GCN, CSRGraph  # noqa: E402
from gcl_b200.graphs_build import ModelGraphs  # noqa: E402

PEAK = 6533.5


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def report(name, us, nbytes, extra=""):
    gbs = nbytes / us / 1e3
    print(f"{name:46s} {us:9.1f} us  {gbs:8.0f} GB/s  {gbs / PEAK:5.2f} of HBM peak  {extra}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="*", default=["spmm", "gat", "linear"])
    ap.add_argument("--mesh", default="35")
    ap.add_argument("--grid", default="64x32")
    ap.add_argument("-B", type=int, default=64)
    ap.add_argument("-C", type=int, default=64)
    args = ap.parse_args()
    dev = "cuda:0"
    nlon, nlat = map(int, args.grid.split("x"))
    levels = [int(c) for c in args.mesh]
    mg = ModelGraphs(nlat, nlon, levels, 0.6, dev)
    B, C = args.B, args.C
    M, N = mg.num_mesh, mg.num_grid + mg.num_mesh
    graphs = {"mesh": (mg.processing_graph, M), "g2m": (mg.encoding_graph, N), "m2g": (mg.decoding_graph, N)}
    torch.manual_seed(0)
    if "spmm" in args.what:
        for gname, (ei, n) in graphs.items():
            g = CSRGraph(ei, n, CSR_LOOPS)
            w, wt = g.weights(NORM_GCN)
            x = torch.randn(B, n, C, device=dev)
            bias = torch.randn(C, device=dev)
            nbytes = 4 * B * C * 2 * n + 8 * g.nnz + 4 * (n + 1)
            us = timeit(lambda: ops.spmm_raw(g.rowptr, g.col, w, x, n, bias))
            report(f"spmm fwd {gname} N={n} nnz={g.nnz} C={C} B={B}", us, nbytes, f"{B * g.nnz / us / 1e3:.1f} Gedge/s")
            us = timeit(lambda: ops.spmm_raw(g.rowptr_t, g.col_t, wt, x, n))
            report(f"spmm bwd {gname} (sender-grouped)", us, nbytes)
    if "gat" in args.what:
        ei, n = graphs["mesh"]
        g = CSRGraph(ei, n, CSR_LOOPS)
        for H in (1,):
            z = torch.randn(B, n, H * C, device=dev, requires_grad=True)
            a_s = torch.randn(1, H, C, device=dev, requires_grad=True)
            a_d = torch.randn(1, H, C, device=dev, requires_grad=True)
            bias = torch.randn(C, device=dev, requires_grad=True)
            nb_f = 4 * B * (n * (H * C + C + 2 * H) + g.nnz * H) + 4 * g.nnz + 4 * (n + 1)
            with torch.no_grad():
                us = timeit(lambda: ops.gat_attend(z, a_s, a_d, bias, g, H, False, 0.2))
            report(f"gat fwd mesh N={n} nnz={g.nnz} H={H} C={C} B={B}", us, nb_f, f"{B * g.nnz / us / 1e3:.1f} Gedge/s")
            out, _ = ops.gat_attend(z, a_s, a_d, bias, g, H, False, 0.2)
            go = torch.randn_like(out)
            nb_b = 4 * B * (n * (2 * H * C + C + 4 * H) + 3 * g.nnz * H) + 16 * g.nnz
            us = timeit(lambda: torch.autograd.grad(out, (z, a_s, a_d, bias), go, retain_graph=True))
            report("gat bwd (scores+dst+src+datt+colsum)", us, nb_b)
    if "linear" in args.what:
        R = B * N
        for cin, cout in ((C, C), (128, 128)):
            x = torch.randn(R, cin, device=dev)
            W = torch.randn(cout, cin, device=dev) / cin ** 0.5
            b = torch.randn(cout, device=dev)
            dy = torch.randn(R, cout, device=dev)
            nbytes = 4 * R * (cin + cout)
            flops = 2 * R * cin * cout
            us = timeit(lambda: ops.linear_fwd_raw(x, W, b))
            report(f"linear fwd R={R} {cin}->{cout}", us, nbytes, f"{flops / us / 1e6:.1f} TFLOP/s")
            us = timeit(lambda: ops.linear_bwd_dx_raw(dy, W))
            report("linear bwd dx", us, nbytes, f"{flops / us / 1e6:.1f} TFLOP/s")
            us = timeit(lambda: ops.linear_bwd_dw_raw(dy, x, True))
            report("linear bwd dW", us, nbytes, f"{flops / us / 1e6:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
