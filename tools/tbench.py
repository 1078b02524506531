#!/usr/bin/env python
"""A/B micro-benchmark of the tiled aggregation kernels (tile.cu / gat_tile.cu) against the row-gather kernels
(spmm.cu / gat.cu) on the model's real graphs.  CUDA events, inputs rotated over several buffers (> L2).

  python tools/tbench.py [--grid 64x32|512x256] [--mesh 35|46] [-B 64] [-C 64] [--what spmm,gat] [--rows 64 --union 128]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graphcast-lite_b200")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from gcl_b200 import graph as gg, ops  # noqa: E402
from gcl_b200.graph import CSR_LOOPS, NORM_GCN, CSRGraph, TilePlan  # noqa: E402
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
    return e0.elapsed_time(e1) / iters * 1e3


def report(name, us, nbytes, extra=""):
    gbs = nbytes / us / 1e3
    print(f"{name:64s} {us:9.1f} us {gbs:8.0f} GB/s {gbs / PEAK:5.2f} of HBM peak  {extra}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="64x32")
    ap.add_argument("--mesh", default="35")
    ap.add_argument("-B", type=int, default=64)
    ap.add_argument("-C", type=int, default=64)
    ap.add_argument("--what", default="spmm,gat")
    ap.add_argument("--rows", default="64")
    ap.add_argument("--union", default="128")
    ap.add_argument("--renumber", action="store_true", help="physically renumber the mesh nodes along the hint order")
    ap.add_argument("--only", default="", help="comma list of graphs (mesh,g2m,m2g)")
    args = ap.parse_args()
    dev = "cuda:0"
    nlon, nlat = map(int, args.grid.split("x"))
    mg = ModelGraphs(nlat, nlon, [int(c) for c in args.mesh], 0.6, dev)
    B, C = args.B, args.C
    M, N = mg.num_mesh, mg.num_grid + mg.num_mesh
    graphs = {"mesh": (mg.processing_graph, M), "g2m": (mg.encoding_graph, N), "m2g": (mg.decoding_graph, N)}
    hints = dict(gg.ORDER_HINTS)
    if args.renumber:
        rank = torch.empty(M, dtype=torch.int64, device=dev)
        rank[torch.as_tensor(hints[M], device=dev).long()] = torch.arange(M, device=dev)
        graphs["mesh"] = (rank[mg.processing_graph], M)
        hints.pop(M)
    if args.only:
        graphs = {k: v for k, v in graphs.items() if k in args.only.split(",")}
    what = args.what.split(",")
    nrot = 3
    for gname, (ei, n) in graphs.items():
        g = CSRGraph(ei, n, CSR_LOOPS)
        w, wt = g.weights(NORM_GCN)
        xs = [torch.randn(B, n, C, device=dev) for _ in range(nrot)]
        bias = torch.randn(C, device=dev)
        nbytes = 4 * B * C * 2 * n + 8 * g.nnz + 4 * (n + 1)
        it = [0]

        def nx():
            it[0] += 1
            return xs[it[0] % nrot]
        if "spmm" in what:
            for tname, rp, co, ww in (("fwd", g.rowptr, g.col, w), ("bwd", g.rowptr_t, g.col_t, wt)):
                us = timeit(lambda: ops.spmm_raw(rp, co, ww, nx(), n, bias))
                report(f"spmm {tname} {gname} N={n} nnz={g.nnz} C={C} B={B}: row-gather", us, nbytes)
                ref = ops.spmm_raw(rp, co, ww, xs[0], n, bias)[0]
                for hint_name, order in (("natural", None), ("hint", hints.get(n))):
                    if hint_name == "hint" and order is None:
                        continue
                    for mr in map(int, args.rows.split(",")):
                        for mu in map(int, args.union.split(",")):
                            pl = TilePlan(rp, co, g.nnz, n, n, n, order, mr, mu, 1024)
                            us = timeit(lambda: ops.spmm_raw(rp, co, ww, nx(), n, bias, plan=pl, wkey=tname))
                            got = ops.spmm_raw(rp, co, ww, xs[0], n, bias, plan=pl, wkey=tname)[0]
                            err = float((got - ref).abs().max())
                            report(f"   tiled {hint_name} R<={mr} U<={mu}: tiles {pl.n_tiles} heavy {pl.n_heavy} "
                                   f"U/R {pl.union_per_row:.2f}", us, nbytes, f"maxdiff {err:.1e}")
        if "bf16" in what and C % 8 == 0:
            xh = [t.to(torch.bfloat16) for t in xs]
            pl = g.plan(False)
            us = timeit(lambda: ops.spmm_bf16_raw(g.rowptr, g.col, w, xh[it[0] % nrot], n, pl, bias, wkey="bf"))
            report(f"spmm fwd {gname} bf16 rows (tiled)", us, nbytes // 2, f"{nbytes / us / 1e3 / PEAK:.2f} of HBM peak in fp32-equivalent bytes")
        if "gat" in what and gname == "mesh":
            z = torch.randn(B, n, C, device=dev, requires_grad=True)
            a_s = torch.randn(1, 1, C, device=dev, requires_grad=True)
            a_d = torch.randn(1, 1, C, device=dev, requires_grad=True)
            bia = torch.randn(C, device=dev, requires_grad=True)
            nb_f = 4 * B * (n * (C + C + 2) + g.nnz) + 4 * g.nnz + 4 * (n + 1)
            nb_b = 4 * B * (n * (2 * C + C + 4) + 3 * g.nnz) + 16 * g.nnz
            for mode in ("row-gather", "tiled natural", "tiled hint"):
                gg.TILED = mode != "row-gather"
                gg.ORDER_HINTS.clear()
                if mode == "tiled hint":
                    gg.ORDER_HINTS.update(hints)
                g._plans.clear()
                g._gat_ws = None
                prof = ops.KernelProfiler()
                with torch.no_grad():
                    us = timeit(lambda: ops.gat_attend(z, a_s, a_d, bia, g, 1, False, 0.2))
                report(f"gat fwd mesh N={n} nnz={g.nnz} C={C} B={B}: {mode}", us, nb_f)
                out, _ = ops.gat_attend(z, a_s, a_d, bia, g, 1, False, 0.2)
                go = torch.randn_like(out)
                torch.autograd.grad(out, (z, a_s, a_d, bia), go, retain_graph=True)
                ops.PROFILER = prof
                for _ in range(5):
                    torch.autograd.grad(out, (z, a_s, a_d, bia), go, retain_graph=True)
                ops.PROFILER = None
                for (name, tag), a in prof.summary().items():
                    if name.startswith("gcl_gat_bwd") or name.startswith("gcl_gat_fwd"):
                        report(f"   {name} ({mode})", 1e3 * a["ms"] / a["calls"], nb_b)
            gg.TILED = True
            gg.ORDER_HINTS.update(hints)


if __name__ == "__main__":
    main()
