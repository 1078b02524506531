#!/usr/bin/env python
"""bench.py -- forecast training throughput of the gcl_b200 hot path on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--batch B] [--ar-steps A]
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
DEFAULT_WORKLOAD = "attention"          # BASELINE.json configs[1]
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


def run_gcl(args):
    import torch
    import torch.distributed as dist
    from gcl_b200 import _cabi, ops
    from gcl_b200.model import WeatherPrediction
    from gcl_b200.train import Trainer

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
    lib = _cabi.load()
    cfg = workload_config(args)
    B = args.batch or DEFAULT_BATCH[args.workload]
    nlat, nlon = cfg["nlat"], cfg["nlon"]
    G = nlat * nlon
    F, T, P = cfg["data"]["num_features_used"], cfg["data"]["obs_window_used"], cfg["data"]["pred_window_used"]
    A = args.ar_steps
    tf, yf = T * F, max(P, A) * F

    torch.manual_seed(42)
    model = WeatherPrediction(cfg, nlat, nlon, dev)
    tr = Trainer(model, nlat, nlon, lr=cfg["learning_rate"], ar_steps=A)
    tr.capture(B, tf, yf)
    gen = torch.Generator().manual_seed(42 + rank)
    hx = torch.randn(B, G, tf, generator=gen).pin_memory()
    hy = torch.randn(B, G, yf, generator=gen).pin_memory()
    tr.static_x.copy_(hx)
    tr.static_y.copy_(hy)

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores, bounded sample
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        t = cpu_reference_run(cfg, args.workload, A, 1, 0, budget_s=args.cpu_baseline_seconds)
        sec = statistics.median(t)
        cpu_base = {"value": 1.0 / sec, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                    "sample": f"{len(t)} batch-1 training steps (~{sum(t):.0f} s) of the same workload, median"}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.no_graph:                      # profiling aid: same kernels, launched eagerly
        tr.step_captured = lambda: tr.step(tr.static_x, tr.static_y)

    # ---- device-resident throughput
    for _ in range(max(args.warmup, 3)):
        tr.step_captured()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = lib.gcl_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tr.step_captured()
    e1.record()
    barrier()
    eager_launches = lib.gcl_launch_count() - launches0
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    last_loss = float(tr.static_loss.item())

    # ---- end to end: pinned host inputs -> H2D -> step -> loss D2H, every step
    # every step: H2D of ITS inputs from pinned memory (issued one step ahead on a copy stream, so it overlaps the
    # previous step's kernels), the captured step, and the loss read back to the host
    for _ in range(2):
        tr.step_from_host(hx, hy, next_batch=(hx, hy))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tr.step_from_host(hx, hy, next_batch=(hx, hy))
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    barrier()
    clocks = sampler.stop()

    # ---- inference: forecast steps without autograd (eager launches), device-resident inputs
    infer = None
    if rank == 0:
        with torch.no_grad():
            for _ in range(3):
                model(X=tr.static_x)
            n_inf = max(args.steps, 5)
            i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            i0.record()
            for _ in range(n_inf):
                model(X=tr.static_x)
            i1.record()
            torch.cuda.synchronize(dev)
        infer = {"value": B / (i0.elapsed_time(i1) / n_inf * 1e-3), "unit": "forecast samples/s (forward only, one GPU)",
                 "ms_per_step": i0.elapsed_time(i1) / n_inf}
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
        for (name, tag), a in rows[:args.kernel_rows]:
            gbs = a["bytes"] / (a["ms"] * 1e-3) / 1e9 if a["ms"] > 0 else 0.0
            by_kernel.append({"kernel": name, "shape": tag, "calls_per_step": a["calls"] / args.profile_steps,
                              "us_per_call": 1e3 * a["ms"] / a["calls"], "share": a["ms"] / total,
                              "algo_GBps": gbs, "frac_hbm": gbs / peak})
        (name, tag), a = rows[0]
        gbs = a["bytes"] / (a["ms"] * 1e-3) / 1e9
        traffic = None
        try:    # measured DRAM bytes per launch of this kernel, from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as f:
                traffic = json.load(f).get(f"{name}[{tag}]")
        except (OSError, ValueError):
            pass
        roof = {"kernel": f"{name}[{tag}]", "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s",
                "frac": gbs / peak, "traffic": traffic, "peak_source": peak_src,
                "share_of_step_kernel_time": a["ms"] / total,
                "algorithmic_bytes_per_launch": a["bytes"] / a["calls"], "us_per_launch": 1e3 * a["ms"] / a["calls"]}
        # mesh message passing edges/s of the processor's aggregation kernel (forward)
        M = model._num_mesh_nodes
        for (name, tag), a in rows:
            if name in ("gcl_gat_fwd_f32", "gcl_spmm_f32") and tag.startswith(f"N{M}x"):
                nnz = model.processing_graph.shape[1] + (M if not model.using_sparse_gat else 0)
                edges = {"kernel": f"{name}[{tag}]", "edges_incl_self_loops_per_launch": B * nnz,
                         "us_per_launch": 1e3 * a["ms"] / a["calls"],
                         "edges_per_s": B * nnz / (a["ms"] / a["calls"] * 1e-3),
                         "algo_GBps": a["bytes"] / (a["ms"] * 1e-3) / 1e9,
                         "frac_hbm": a["bytes"] / (a["ms"] * 1e-3) / 1e9 / peak}
                # forward + backward of one processor layer: the GAT backward entry, or (GCN) the same SpMM entry,
                # which is called once forward and once backward per layer (its average covers both directions)
                bwd = agg.get(("gcl_gat_bwd_f32", tag))
                t_fb = (a["ms"] / a["calls"] + bwd["ms"] / bwd["calls"]) if bwd else 2 * a["ms"] / a["calls"]
                edges["fwd_bwd_edges_per_s"] = B * nnz / (t_fb * 1e-3)
                break

    if rank == 0:
        act_mb = B * (G + model._num_mesh_nodes) * cfg["pipeline"]["encoder"]["gcn"]["output_dim"] * 4 / 1e6
        out = {
            "metric": METRIC, "value": world * B / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "batch_per_gpu": B, "global_batch": world * B, "ar_steps": A,
                       "grid": f"{nlon}x{nlat}", "mesh_levels": cfg["graph"]["mesh_levels"],
                       "parallelism": f"dp{world}", "params": tr.num_params,
                       "l2": f"one activation tensor is {act_mb:.0f} MB per step and ~60 are live (> 126 MB L2): "
                             "inputs larger than L2, no flush needed",
                       "cuda_graph": ("off (--no-graph, profiling run)" if args.no_graph
                                      else "fwd+bwd captured; all-reduce + Adam eager")},
            "clocks": clocks,
            "e2e": {"value": world * B / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": world * int(hx.numel() + hy.numel()) * 4, "d2h_bytes_per_step": world * 4,
                    "api": "gcl_b200.train.Trainer.step_from_host(X_pinned, y_pinned, next_batch) -> float loss; "
                           "the next batch's H2D copy runs on a copy stream during the step"},
            "hbm_peak_gb": round(torch.cuda.max_memory_allocated(dev) / 1e9, 2),
            "gpu_launches": int(tr.launches_in_graph * args.steps + eager_launches),
            "gpu_launches_per_step": int(tr.launches_in_graph + eager_launches // max(args.steps, 1)),
            "loss": last_loss,
            "roofline": roof, "kernels": by_kernel, "mesh_message_passing": edges, "inference": infer,
            "cpu_baseline": cpu_base,
        }
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
