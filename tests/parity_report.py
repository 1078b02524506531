#!/usr/bin/env python
"""Writes profiles/r02_parity.json (run on the GPU box): for every BASELINE config x dense engine, the worst relative
error (max-norm, tests/helpers.rel_err) of the forecast step, the loss and every gradient of the product against the
fp32 oracle AND against the same oracle in fp64, next to the fp32 oracle's own distance from fp64.  Test
infrastructure (imports the oracle), not product code.

  python tests/parity_report.py [--out profiles/r02_parity.json] [--configs a,b,...]
"""
import argparse
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graphcast-lite_b200"), os.path.join(ROOT, "tests"),
                os.path.join(ROOT, "oracle", "pyg_shim"), os.path.join(ROOT, "oracle", "trimesh_shim")]

import torch  # noqa: E402

from helpers import rel_err  # noqa: E402

ALL = ["baseline", "attention", "sparse_attention", "wb2_64x32_ar_15f_4obs_4pred", "wb2_512x256_19f_ar"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_parity.json"))
    ap.add_argument("--configs", default=",".join(ALL))
    args = ap.parse_args()
    from gcl_b200 import _cabi
    from gcl_b200.model import WeatherPrediction
    from gcl_b200.train import Trainer
    from gcl_b200.workloads import get_workload
    from oracle import graphs as og, model as om
    dev = "cuda:0"
    report = {"tolerance_north_star": 1e-4, "metric": "max |a - b| / max |b|", "batch": 1, "configs": {}}
    for name in args.configs.split(","):
        cfg = get_workload(name)
        nlat, nlon = cfg["nlat"], cfg["nlon"]
        torch.manual_seed(0)
        ref = om.WeatherPrediction(cfg, nlat, nlon, graphs=og.build_graphs(nlat, nlon, cfg["graph"]["mesh_levels"],
                                                                            cfg["graph"]["grid2mesh_radius_query"]))
        with torch.no_grad():
            gen = torch.Generator().manual_seed(1)
            for p in ref.parameters():
                p.add_(0.05 * torch.randn(p.shape, generator=gen))
        G = nlat * nlon
        F, T = cfg["data"]["num_features_used"], cfg["data"]["obs_window_used"]
        gen = torch.Generator().manual_seed(3)
        X, y = torch.randn(1, G, T * F, generator=gen), torch.randn(1, G, F, generator=gen)
        kw = dict(batch_num=1) if name == "sparse_attention" else {}
        lw = om.lat_weights(nlat, nlon)
        out32 = ref(X=X, attention_threshold=0.0, **kw).detach()
        l32 = om.training_loss(ref, X, y, 1, lw, **kw)
        l32.backward()
        ref64 = copy.deepcopy(ref).double()
        ref64.zero_grad()
        ref64.init_grid_features, ref64.init_mesh_features = ref64.init_grid_features.double(), ref64.init_mesh_features.double()
        out64 = ref64(X=X.double(), attention_threshold=0.0, **kw).detach()
        l64 = om.training_loss(ref64, X.double(), y.double(), 1, lw.double(), **kw)
        l64.backward()
        g32 = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
        g64 = {k: p.grad.float() for k, p in ref64.named_parameters() if p.grad is not None}
        nonzero = [k for k in g32 if float(g64[k].abs().max()) >= 1e-12]      # skip analytically-zero gradients
        entry = {"params": sum(p.numel() for p in ref.parameters()),
                 "oracle_fp32_vs_fp64": {"output": rel_err(out32, out64.float()),
                                         "loss": abs(float(l32.detach()) - float(l64.detach())) / abs(float(l64.detach())),
                                         "worst_gradient": max((rel_err(g32[k], g64[k]), k) for k in nonzero),
                                         "gradients_over_1e-4": sorted(k for k in nonzero if rel_err(g32[k], g64[k]) > 1e-4)}}
        for engine in ("ffma", "tcgen05"):
            _cabi.load().gcl_set_dense_mode(1 if engine == "ffma" else 0)
            mine = WeatherPrediction(cfg, nlat, nlon, dev)
            mine.load_state_dict(ref.state_dict())
            out = mine(X=X.to(dev), attention_threshold=0.0, **kw).detach()
            tr = Trainer(mine, nlat, nlon, lr=cfg["learning_rate"], ar_steps=1)
            tr.zero_grad()
            loss = tr.loss(X.to(dev), y.to(dev), 0.0, **kw)
            loss.backward()
            gm = {k: p.grad.detach().cpu() for k, p in mine.named_parameters() if k in g32}
            skip = lambda k: float(g64[k].abs().max()) < 1e-12           # analytically-zero gradients
            e32 = {k: rel_err(gm[k], g32[k]) for k in gm if not skip(k)}
            e64 = {k: rel_err(gm[k], g64[k]) for k in gm if not skip(k)}
            w32, w64 = max(e32, key=e32.get), max(e64, key=e64.get)
            entry[engine] = {
                "output_vs_fp32": rel_err(out, out32), "output_vs_fp64": rel_err(out, out64.float()),
                "loss_vs_fp32": abs(float(loss.detach()) - float(l32.detach())) / abs(float(l32.detach())),
                "loss_vs_fp64": abs(float(loss.detach()) - float(l64.detach())) / abs(float(l64.detach())),
                "worst_gradient_vs_fp32": {"name": w32, "err": e32[w32], "oracle_fp32_vs_fp64": rel_err(g32[w32], g64[w32])},
                "worst_gradient_vs_fp64": {"name": w64, "err": e64[w64], "oracle_fp32_vs_fp64": rel_err(g32[w64], g64[w64])},
                "gradients_over_1e-4_vs_fp32": sorted(k for k, v in e32.items() if v > 1e-4),
                "gradients_over_1e-4_vs_fp64": sorted(k for k, v in e64.items() if v > 1e-4),
                "n_gradients": len(e32)}
            del mine, tr
            torch.cuda.empty_cache()
        _cabi.load().gcl_set_dense_mode(0)
        report["configs"][name] = entry
        print(name, json.dumps(entry)[:600], flush=True)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
