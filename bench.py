#!/usr/bin/env python
"""bench.py -- forecast training throughput of the gcl_b200 hot path on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--batch B] [--ar-steps A] [--no-workloads]
  python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

A "step" = one training step (zero_grad, AR forward(s), latitude-weighted MSE, backward, gradient
all-reduce, Adam) over B synthetic forecast samples per GPU (ERA5-shaped N(0,1) fields, random-init
weights, seed 42).  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every key.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "graphcast-lite_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "forecast samples/sec (training step: fwd + bwd + Adam)"
UNIT = "samples/s"
DEFAULT_WORKLOAD = "wb2_512x256_19f_ar"   # the largest BASELINE.json config that fits one GPU (16 GB at B = 8)
DEFAULT_BATCH = {"baseline": 64, "attention": 64, "sparse_attention": 64,
                 "wb2_64x32_ar_15f_4obs_4pred": 16, "wb2_512x256_19f_ar": 8}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="gcl", choices=["gcl", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU per step (0 = workload default)")
    ap.add_argument("--ar-steps", type=int, default=1)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--no-graph", action="store_true",
                    help="eager launches instead of CUDA-graph replay (for ncu: it fails with LaunchFailed on graph "
                         "nodes that take a CUtensorMap parameter); the numbers of such a run are not bench values")
    ap.add_argument("--kernel-rows", type=int, default=12, help="rows of the per-kernel table kept in the JSON line")
    ap.add_argument("--min-timed-ms", type=float, default=1000.0,
                    help="blocks of --steps steps are repeated until the timed region is at least this long")
    ap.add_argument("--no-workloads", action="store_true",
                    help="only the headline workload (skip the other BASELINE.json configs, the drop-in batch-1 path and "
                         "the per-layer CPU numbers that the N = 1 run adds)")
    return ap.parse_args()


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def workload_config(args):
    from gcl_b200.workloads import get_workload
    cfg = get_workload(args.workload)
    return cfg


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(cfg, workload, ar_steps, warmup, steps, budget_s=None):
    """The reference's own path on the host: restated PyG (oracle/pyg_shim) under the restated model glue
    (oracle/model.py), batch 1 exactly like /root/reference/src/train.py:160-235, all host threads."""
    import torch
    from oracle import model as om
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(42)
    nlat, nlon = cfg["nlat"], cfg["nlon"]
    m = om.WeatherPrediction(cfg, nlat, nlon)
    opt = torch.optim.Adam(m.parameters(), lr=cfg["learning_rate"])
    lw = om.lat_weights(nlat, nlon)
    G = nlat * nlon
    F, T, P = cfg["data"]["num_features_used"], cfg["data"]["obs_window_used"], cfg["data"]["pred_window_used"]
    gen = torch.Generator().manual_seed(42)
    X, y = torch.randn(1, G, T * F, generator=gen), torch.randn(1, G, max(P, ar_steps) * F, generator=gen)
    kw = dict(batch_num=1) if workload == "sparse_attention" else {}

    def one():
        opt.zero_grad()
        loss = om.training_loss(m, X, y, ar_steps, lw, **kw)
        loss.backward()
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        one()
    times = []
    t_end = time.perf_counter() + budget_s if budget_s else None
    n = 0
    while True:
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
        n += 1
        if steps and n >= steps and not t_end:
            break
        if t_end and (time.perf_counter() >= t_end and n >= 3 or n >= 2000):
            break
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload_config(args)
    times = cpu_reference_run(cfg, args.workload, args.ar_steps, args.warmup, args.steps)
    sec = sum(times) / len(times)
    val = 1.0 / sec
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_step": 1, "ar_steps": args.ar_steps,
                   "note": "reference's CPU path: restated torch_geometric 2.5.3 + restated model glue (oracle/), "
                           "batch 1 as in the reference, torch CPU kernels"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{args.steps} batch-1 training steps after {args.warmup} warm-up"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [s.strip() for s in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _barrier(world, dev):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def _max_over_ranks(v, world, dev):
    import torch
    import torch.distributed as dist
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bench_workload(name, args, world, rank, dev, ar_steps=1, batch=0, headline=False, cpu_seconds=0.0,
                   sample_clocks=False):
    """One workload on this process's GPU: device-resident step time (blocks of `steps` steps until >= 1 s, median
    block), end-to-end step from pinned host buffers, per-kernel roofline, inference, CPU port beside it."""
    import torch
    from gcl_b200 import _cabi, ops
    from gcl_b200.model import WeatherPrediction
    from gcl_b200.train import Trainer
    from gcl_b200.workloads import get_workload

    lib = _cabi.load()
    cfg = get_workload(name)
    B = batch or DEFAULT_BATCH[name]
    nlat, nlon = cfg["nlat"], cfg["nlon"]
    G = nlat * nlon
    F, T, P = cfg["data"]["num_features_used"], cfg["data"]["obs_window_used"], cfg["data"]["pred_window_used"]
    A = ar_steps
    tf, yf = T * F, max(P, A) * F
    local = dev.index

    torch.manual_seed(42)
    torch.cuda.reset_peak_memory_stats(dev)
    model = WeatherPrediction(cfg, nlat, nlon, dev)
    tr = Trainer(model, nlat, nlon, lr=cfg["learning_rate"], ar_steps=A)
    tr.capture(B, tf, yf)
    gen = torch.Generator().manual_seed(42 + rank)
    hosts = [(torch.randn(B, G, tf, generator=gen).pin_memory(), torch.randn(B, G, yf, generator=gen).pin_memory())
             for _ in range(2)]                      # e2e alternates between distinct pinned batches
    tr.static_x.copy_(hosts[0][0])
    tr.static_y.copy_(hosts[0][1])

    cpu_base = None
    if cpu_seconds > 0 and rank == 0:      # the oracle port on the host cores, bounded sample (CPU leg only)
        t = cpu_reference_run(cfg, name, A, 1, 0, budget_s=cpu_seconds)
        sec = statistics.median(t)
        cpu_base = {"value": 1.0 / sec, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                    "sample": f"{len(t)} batch-1 training steps (~{sum(t):.0f} s) of the same workload, median"}

    if args.no_graph:                      # profiling aid: same kernels, launched eagerly
        tr.step_captured = lambda: tr.step(tr.static_x, tr.static_y)

    # ---- device-resident throughput: blocks of K steps, repeated until the timed region is >= 1 s; median block
    K = args.steps
    for _ in range(max(args.warmup, 3)):
        tr.step_captured()
    sampler = ClockSampler(local) if sample_clocks else None
    _barrier(world, dev)
    if sampler:
        sampler.start()
    launches0 = lib.gcl_launch_count()
    blocks, total_ms = [], 0.0
    while True:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            tr.step_captured()
        e1.record()
        _barrier(world, dev)
        blk = _max_over_ranks(e0.elapsed_time(e1), world, dev)     # identical on every rank: same loop count
        blocks.append(blk / K)
        total_ms += blk
        if total_ms >= args.min_timed_ms or len(blocks) >= 200:
            break
    eager_launches = (lib.gcl_launch_count() - launches0) // len(blocks)
    ms = statistics.median(blocks)
    last_loss = float(tr.static_loss.item())

    # ---- end to end: every step copies ITS inputs from pinned host memory (issued one step ahead on a copy stream,
    # alternating between two distinct host batches), runs the captured step and reads the loss back to the host
    for i in range(2):
        tr.step_from_host(*hosts[i % 2], next_batch=hosts[(i + 1) % 2])
    _barrier(world, dev)
    n_e2e = max(K, int(len(blocks) * K * 0.5))
    t0 = time.perf_counter()
    for i in range(n_e2e):
        tr.step_from_host(*hosts[i % 2], next_batch=hosts[(i + 1) % 2])
    torch.cuda.synchronize(dev)
    e2e_s = _max_over_ranks((time.perf_counter() - t0) / n_e2e, world, dev)
    _barrier(world, dev)
    clocks = sampler.stop() if sampler else None

    # ---- inference: forecast steps without autograd (eager launches), device-resident inputs
    infer = None
    if rank == 0:
        with torch.no_grad():
            for _ in range(3):
                model(X=tr.static_x)
            n_inf = max(K // 2, 5)
            i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            i0.record()
            for _ in range(n_inf):
                model(X=tr.static_x)
            i1.record()
            torch.cuda.synchronize(dev)
        infer = {"value": B / (i0.elapsed_time(i1) / n_inf * 1e-3), "unit": "forecast samples/s (forward only, one GPU)",
                 "ms_per_step": i0.elapsed_time(i1) / n_inf}
        if headline or name == "attention":
            # device-resident 4-step autoregressive rollout + streaming metrics (gcl_b200.predict)
            from gcl_b200 import predict as gp
            y4 = torch.randn(B, G, 4 * F, device=dev)
            sm = gp.StreamingMetrics(F, device=dev)
            for _ in range(2):
                sm.update(y4, gp.rollout(model, tr.static_x, 4, F, T))
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            for _ in range(n_inf):
                sm.update(y4, gp.rollout(model, tr.static_x, 4, F, T))
            r1.record()
            torch.cuda.synchronize(dev)
            infer["rollout4"] = {"value": B / (r0.elapsed_time(r1) / n_inf * 1e-3),
                                 "unit": "4-step forecasts/s incl. streaming RMSE/ACC (one GPU)",
                                 "ms_per_rollout": r0.elapsed_time(r1) / n_inf, "rmse": sm.result()["rmse"]}

    # ---- per-kernel CUDA-event timing inside real (eager) steps: roofline of the dominant kernel
    peak, peak_src = measured_peak_gbs()
    roof, by_kernel, edges = None, [], None
    if rank == 0 and args.profile_steps > 0:
        world_saved, tr.world = tr.world, 1         # rank 0 profiles alone: no collective in these eager steps
        tr.step(tr.static_x, tr.static_y)           # eager warm-up
        ops.PROFILER = ops.KernelProfiler()
        for _ in range(args.profile_steps):
            tr.step(tr.static_x, tr.static_y)
        agg = ops.PROFILER.summary()
        ops.PROFILER = None
        tr.world = world_saved
        total = sum(a["ms"] for a in agg.values()) or 1.0
        rows = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])
        for (kname, tag), a in rows[:args.kernel_rows]:
            gbs = a["bytes"] / (a["ms"] * 1e-3) / 1e9 if a["ms"] > 0 else 0.0
            by_kernel.append({"kernel": kname, "shape": tag, "calls_per_step": a["calls"] / args.profile_steps,
                              "us_per_call": 1e3 * a["ms"] / a["calls"], "share": a["ms"] / total,
                              "algo_GBps": gbs, "frac_hbm": gbs / peak})
        (kname, tag), a = rows[0]
        gbs = a["bytes"] / (a["ms"] * 1e-3) / 1e9
        traffic = None
        try:    # measured DRAM bytes per launch of this kernel, from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as f:
                traffic = json.load(f).get(f"{kname}[{tag}]")
        except (OSError, ValueError):
            pass
        roof = {"kernel": f"{kname}[{tag}]", "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s",
                "frac": gbs / peak, "traffic": traffic, "peak_source": peak_src,
                "share_of_step_kernel_time": a["ms"] / total,
                "algorithmic_bytes_per_launch": a["bytes"] / a["calls"], "us_per_launch": 1e3 * a["ms"] / a["calls"]}
        # mesh message passing edges/s of the processor's aggregation kernel (forward)
        M = model._num_mesh_nodes
        for (kname, tag), a in rows:
            if kname.startswith(("gcl_gat_fwd", "gcl_spmm")) and tag.startswith(f"N{M}x"):
                nnz = model.processing_graph.shape[1] + (M if not model.using_sparse_gat else 0)
                edges = {"kernel": f"{kname}[{tag}]", "edges_incl_self_loops_per_launch": B * nnz,
                         "us_per_launch": 1e3 * a["ms"] / a["calls"],
                         "edges_per_s": B * nnz / (a["ms"] / a["calls"] * 1e-3),
                         "algo_GBps": a["bytes"] / (a["ms"] * 1e-3) / 1e9,
                         "frac_hbm": a["bytes"] / (a["ms"] * 1e-3) / 1e9 / peak}
                # forward + backward of one processor layer: the GAT backward entry, or (GCN) the same SpMM entry,
                # which is called once forward and once backward per layer (its average covers both directions)
                bwd = next((v for (kn, tg), v in agg.items() if kn.startswith("gcl_gat_bwd") and tg == tag), None)
                t_fb = (a["ms"] / a["calls"] + bwd["ms"] / bwd["calls"]) if bwd else 2 * a["ms"] / a["calls"]
                edges["fwd_bwd_edges_per_s"] = B * nnz / (t_fb * 1e-3)
                break

    act_mb = B * (G + model._num_mesh_nodes) * cfg["pipeline"]["encoder"]["gcn"]["output_dim"] * 4 / 1e6
    res = {
        "workload": name, "value": world * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
        "timed_blocks": len(blocks), "timed_region_ms": total_ms,
        "config": {"workload": name, "batch_per_gpu": B, "global_batch": world * B, "ar_steps": A,
                   "grid": f"{nlon}x{nlat}", "mesh_levels": cfg["graph"]["mesh_levels"],
                   "parallelism": f"dp{world}", "params": tr.num_params,
                   "l2": f"one activation tensor is {act_mb:.0f} MB per step and ~60 are live (> 126 MB L2): "
                         "inputs larger than L2, no flush needed",
                   "cuda_graph": ("off (--no-graph, profiling run)" if args.no_graph
                                  else ("whole step captured: fwd + bwd + gradient all-reduce + Adam = one graph launch"
                                        if getattr(tr, "_graph_has_tail", False)
                                        else "fwd+bwd captured; all-reduce + Adam eager")),
                   "vs_cpu": "the GPU arm batches B samples per step; the reference (and the CPU port beside it) is "
                             "batch-1 only (models.py:822), so both are compared in samples/s"},
        "clocks": clocks,
        "e2e": {"value": world * B / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": world * int(hosts[0][0].numel() + hosts[0][1].numel()) * 4,
                "d2h_bytes_per_step": world * 4,
                "api": "gcl_b200.train.Trainer.step_from_host(X_pinned, y_pinned, next_batch) -> float loss; "
                       "the next batch's H2D copy runs on a copy stream during the step"},
        "hbm_peak_gb": round(torch.cuda.max_memory_allocated(dev) / 1e9, 2),
        "gpu_launches": int((tr.launches_in_graph * K + eager_launches)),
        "gpu_launches_per_step": int(tr.launches_in_graph + eager_launches // max(K, 1)),
        "loss": last_loss,
        "roofline": roof, "kernels": by_kernel, "mesh_message_passing": edges, "inference": infer,
        "cpu_baseline": cpu_base,
    }
    del tr, model
    torch.cuda.empty_cache()
    return res


def bench_dropin_b1(name, args, dev, cpu_value=None):
    """What changing the import at models.py:21 gives: the reference's unfused batch-1 forward on gcl_b200.nn
    (gcl_b200.dropin), training step = forward + MSE + backward + torch.optim.Adam; eager and CUDA-graph replayed."""
    import torch
    from gcl_b200.dropin import reference_forward
    from gcl_b200.model import WeatherPrediction
    from gcl_b200.workloads import get_workload
    cfg = get_workload(name)
    nlat, nlon = cfg["nlat"], cfg["nlon"]
    G = nlat * nlon
    F, T = cfg["data"]["num_features_used"], cfg["data"]["obs_window_used"]
    torch.manual_seed(42)
    model = WeatherPrediction(cfg, nlat, nlon, dev)
    opt = torch.optim.Adam(model.parameters(), lr=cfg["learning_rate"], capturable=True)
    X, y = torch.randn(1, G, T * F, device=dev), torch.randn(G, F, device=dev)

    def step():
        opt.zero_grad(set_to_none=False)
        out = X[0, :, -F:] + reference_forward(model, X)
        loss = (out - y).square().mean()
        loss.backward()
        opt.step()
        return loss

    def timeit(fn, n):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n
    eager_ms = timeit(step, 30)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    graph_ms = timeit(graph.replay, 100)
    res = {"workload": name, "batch": 1, "api": "gcl_b200.dropin.reference_forward (reference glue order, torch.nn.Linear "
           "MLPs, gcl_b200.nn convs) + torch.optim.Adam", "eager": {"value": 1e3 / eager_ms, "unit": UNIT, "ms_per_step": eager_ms},
           "cuda_graph": {"value": 1e3 / graph_ms, "unit": UNIT, "ms_per_step": graph_ms}}
    if cpu_value:
        res["vs_cpu_port_b1"] = {"eager": (1e3 / eager_ms) / cpu_value, "cuda_graph": (1e3 / graph_ms) / cpu_value}
    del model, opt, graph
    torch.cuda.empty_cache()
    return res


def bench_input_pipeline(name, args, dev):
    """f2: a synthetic dataset in the reference's raw on-disk format (float16 memmap (T, lon, lat, F) + scalers.npz),
    read through gcl_b200.data.ChunkedWindowLoader: host staging of raw windows -> H2D -> gcl_window_assemble; alone,
    and feeding the captured training step (every batch comes from disk / page cache, none is re-sent)."""
    import shutil
    import numpy as np
    import torch
    from gcl_b200.data import ChunkedWindowLoader
    from gcl_b200.model import WeatherPrediction
    from gcl_b200.train import Trainer
    from gcl_b200.workloads import get_workload
    cfg = get_workload(name)
    nlat, nlon = cfg["nlat"], cfg["nlon"]
    F, T, P = cfg["data"]["num_features_used"], cfg["data"]["obs_window_used"], cfg["data"]["pred_window_used"]
    B = DEFAULT_BATCH[name]
    n_batches = 12
    frames = B * n_batches + T + P - 1
    tmp = tempfile.mkdtemp(prefix="gcl_bench_data_")
    try:
        rng = np.random.default_rng(0)
        mm = np.memmap(os.path.join(tmp, "data.npy"), dtype=np.float16, mode="w+", shape=(frames, nlon, nlat, F))
        block = rng.standard_normal((8, nlon, nlat, F), dtype=np.float32).astype(np.float16)
        for t0 in range(0, frames, 8):
            mm[t0:t0 + 8] = block[: min(8, frames - t0)]
        mm.flush()
        del mm
        json.dump({"n_time": frames, "n_lon": nlon, "n_lat": nlat, "n_feat": F}, open(os.path.join(tmp, "dataset_info.json"), "w"))
        np.savez(os.path.join(tmp, "scalers.npz"), mean=np.zeros(F, np.float32), std=np.ones(F, np.float32), n=np.int64(frames))
        ld = ChunkedWindowLoader(tmp, T, P, "all", None, device=dev)
        raw_bytes = B * (T + P) * nlon * nlat * F * 2
        for _ in ld.batches(B, drop_last=True):       # warm the page cache and the kernel
            pass
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        n = 0
        for X, Y in ld.batches(B, drop_last=True):
            n += X.shape[0]
        torch.cuda.synchronize(dev)
        loader_s = time.perf_counter() - t0
        torch.manual_seed(42)
        model = WeatherPrediction(cfg, nlat, nlon, dev)
        tr = Trainer(model, nlat, nlon, lr=cfg["learning_rate"], ar_steps=1)
        tr.capture(B, T * F, P * F)
        for X, Y in ld.batches(B, drop_last=True):
            tr.step_from_device(X, Y)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        m = 0
        for X, Y in ld.batches(B, drop_last=True):
            loss = tr.step_from_device(X, Y)
            m += X.shape[0]
        last = float(loss.item())
        train_s = time.perf_counter() - t0
        res = {"workload": name, "batch": B, "format": "raw float16 memmap (T, lon, lat, F), page cache",
               "api": "ChunkedWindowLoader.batches() -> Trainer.step_from_device()",
               "loader": {"value": n / loader_s, "unit": UNIT, "raw_GBps": raw_bytes * (n / B) / loader_s / 1e9},
               "train_from_loader": {"value": m / train_s, "unit": UNIT, "ms_per_step": 1e3 * train_s / (m / B),
                                     "h2d_bytes_per_step": raw_bytes, "loss": last}}
        del tr, model
        torch.cuda.empty_cache()
        return res
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def bench_cpu_layers(seconds=3.0):
    """Per-layer mesh message passing on the host cores: the oracle's restated PyG GCNConv / GATConv on
    the [3,5] multi-mesh, 64 channels, batch 1, edges incl. self loops per second, forward and forward + backward."""
    import torch
    for p in ("pyg_shim", "trimesh_shim"):
        q = os.path.join(ROOT, "oracle", p)
        if q not in sys.path:
            sys.path.insert(0, q)
    import torch_geometric.nn as onn            # the oracle (CPU leg only)
    from oracle import graphs as og
    torch.set_num_threads(os.cpu_count() or 1)
    g = og.build_graphs(32, 64, [3, 5], 0.5)
    ei = torch.as_tensor(g["mesh"])
    M = g["num_mesh"]
    E = ei.shape[1] + M
    out = {"graph": "multi-mesh [3,5], 10 242 nodes, 75 522 edges incl. self loops, 64 channels, batch 1",
           "cores": os.cpu_count(), "kind": "port"}
    torch.manual_seed(0)
    for lname, layer in (("gcn", onn.GCNConv(64, 64)), ("gat", onn.GATConv(64, 64, heads=1, concat=False))):
        x = torch.randn(M, 64, requires_grad=True)

        def fwd():
            with torch.no_grad():
                layer(x, ei)

        def fwd_bwd():
            layer(x, ei).square().sum().backward()
        for kind, fn in (("fwd", fwd), ("fwd_bwd", fwd_bwd)):
            fn()
            ts, t_end = [], time.perf_counter() + seconds / 4
            while time.perf_counter() < t_end or len(ts) < 3:
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            out[f"{lname}_{kind}_edges_per_s"] = E / statistics.median(ts)
    return out


def run_gcl(args):
    import torch
    import torch.distributed as dist
    from gcl_b200 import _cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl gcl) needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _cabi.load()
    cpu_s = 0.0 if (args.no_cpu_baseline or world > 1) else args.cpu_baseline_seconds
    head = bench_workload(args.workload, args, world, rank, dev, ar_steps=args.ar_steps, batch=args.batch,
                          headline=True, cpu_seconds=cpu_s, sample_clocks=True)
    out = None
    if rank == 0:
        out = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        out.update({k: v for k, v in head.items() if k not in ("workload", "value", "unit", "ms_per_step")})

    # ---- the other BASELINE.json configs (one GPU, rank 0's process only when N = 1), the drop-in batch-1 path, the
    # per-layer CPU numbers of BASELINE.md 2
    if world == 1 and not args.no_workloads:
        keep = ("workload", "value", "unit", "ms_per_step", "timed_region_ms", "e2e", "roofline", "cpu_baseline",
                "mesh_message_passing", "hbm_peak_gb", "gpu_launches_per_step", "loss")
        extra_cpu = 0.0 if args.no_cpu_baseline else min(args.cpu_baseline_seconds, 6.0)
        wl = [dict({k: head[k] for k in keep}, batch_per_gpu=head["config"]["batch_per_gpu"], ar_steps=args.ar_steps,
                   inference=head["inference"])]
        for name, ar in (("baseline", 1), ("attention", 1), ("sparse_attention", 1),
                         ("wb2_64x32_ar_15f_4obs_4pred", 1), ("wb2_64x32_ar_15f_4obs_4pred", 4),
                         ("wb2_512x256_19f_ar", 1)):
            if name == args.workload and ar == args.ar_steps:
                continue
            r = bench_workload(name, args, world, rank, dev, ar_steps=ar, cpu_seconds=extra_cpu)
            wl.append(dict({k: r[k] for k in keep}, batch_per_gpu=r["config"]["batch_per_gpu"], ar_steps=ar,
                           inference=r["inference"]))
        out["workloads"] = wl
        cpu_att = next((w["cpu_baseline"]["value"] for w in wl if w["workload"] == "attention" and w["cpu_baseline"]), None)
        out["dropin_b1"] = bench_dropin_b1("attention", args, dev, cpu_att)
        out["input_pipeline"] = bench_input_pipeline("attention", args, dev)
        if not args.no_cpu_baseline:
            out["cpu_layers"] = bench_cpu_layers()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gcl(args)


if __name__ == "__main__":
    main()
