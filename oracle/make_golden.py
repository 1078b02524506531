"""ORACLE / TEST INFRASTRUCTURE ONLY.  Build-container script: runs the UNMODIFIED reference
(/root/reference/src/models.py, create_graphs.py, mesh/*, utils.py, train.py's loss) on top of the
oracle's restated torch_geometric / trimesh shims and writes the fixtures under tests/golden/ that the
GPU box (where /root/reference does not exist) checks against:

  graphs.json          per graph config: node/edge counts and sha256 digests of the three edge lists
                       (grid->mesh as a sorted set), vertices and static features
  model_<name>.npz     small variants of the baseline / attention / sparse_attention workloads:
                       state_dict, input, target, forward output, loss, every gradient

Usage (from the repo root, build container):  python -m oracle.make_golden
"""
import contextlib
import hashlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = "/root/reference"
for p in (os.path.join(HERE, "pyg_shim"), os.path.join(HERE, "trimesh_shim"), REFERENCE, ROOT,
          os.path.join(ROOT, "graphcast-lite_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GRAPH_CONFIGS = {           # name -> (nlat, nlon, mesh_levels, radius factor)
    "notebook_level0": (32, 64, [0], 0.5),          # notebooks/src/main.ipynb cell 5
    "64x32_l35_rq0.5": (32, 64, [3, 5], 0.5),       # baseline / attention / sparse_attention
    "64x32_l46_rq0.65": (32, 64, [4, 6], 0.65),     # wb2_64x32_ar_15f_4obs_4pred
    "512x256_l46_rq0.6": (256, 512, [4, 6], 0.6),   # wb2_512x256_19f_ar
    "small_32x16_l13_rq0.6": (16, 32, [1, 3], 0.6), # golden model fixtures
}
SMALL = dict(nlat=16, nlon=32, mesh_levels=[1, 3], rq=0.6)


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sorted_edges(ei: np.ndarray) -> np.ndarray:
    order = np.lexsort((ei[1], ei[0]))
    return np.ascontiguousarray(ei[:, order])


def reference_model(cfg: dict, nlat: int, nlon: int, seed: int = 42):
    """Instantiate the unmodified reference WeatherPrediction from a workloads-style dict."""
    from src.config import DataConfig, GraphBuildingConfig, PipelineConfig
    from src.models import WeatherPrediction
    graph = GraphBuildingConfig(**cfg["graph"])
    pipe = PipelineConfig(**cfg["pipeline"])
    data = DataConfig(dataset_name="wb2_512x256_19f_ar", want_feats_flattened=True, **cfg["data"])
    lats = np.linspace(-90, 90, nlat)
    lons = np.linspace(0, 360, nlon, endpoint=False)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        return WeatherPrediction(cordinates=(lats, lons), graph_config=graph, pipeline_config=pipe,
                                 data_config=data, device="cpu")


def graph_entry(m) -> dict:
    g2m, mesh, m2g = (t.numpy() for t in (m.encoding_graph, m.processing_graph, m.decoding_graph))
    return {
        "num_grid": int(m._num_grid_nodes), "num_mesh": int(m._num_mesh_nodes),
        "E_g2m": int(g2m.shape[1]), "E_mesh": int(mesh.shape[1]), "E_m2g": int(m2g.shape[1]),
        "g2m_sorted_sha256": digest(sorted_edges(g2m)), "mesh_sha256": digest(mesh), "m2g_sha256": digest(m2g),
        "vertices_sha256": digest(m._finest_mesh.vertices), "faces_sha256": digest(m._finest_mesh.faces),
        "grid_feats_sha256": digest(m.init_grid_features.numpy()),
        "mesh_feats_sha256": digest(m.init_mesh_features.numpy()),
        "max_in_degree_g2m": int(np.bincount(g2m[1]).max()),
    }


def main():
    from gcl_b200.workloads import get_workload
    from src.train import get_lat_weights, weighted_mse_loss
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    graphs = {}
    base = get_workload("sparse_attention")      # cheapest model; only its graphs matter here
    for name, (nlat, nlon, levels, rq) in GRAPH_CONFIGS.items():
        cfg = json.loads(json.dumps(base))
        cfg["graph"]["mesh_levels"], cfg["graph"]["grid2mesh_radius_query"] = levels, rq
        m = reference_model(cfg, nlat, nlon)
        graphs[name] = dict(nlat=nlat, nlon=nlon, mesh_levels=levels, radius_factor=rq, **graph_entry(m))
        print(name, {k: v for k, v in graphs[name].items() if not k.endswith("sha256")})
    with open(os.path.join(out_dir, "graphs.json"), "w") as f:
        json.dump({"generated_by": "oracle/make_golden.py (unmodified reference on the oracle shims)",
                   "numpy": np.__version__, "graphs": graphs}, f, indent=1)

    for name in ("baseline", "attention", "sparse_attention"):
        cfg = get_workload(name)
        cfg["graph"]["mesh_levels"], cfg["graph"]["grid2mesh_radius_query"] = SMALL["mesh_levels"], SMALL["rq"]
        nlat, nlon = SMALL["nlat"], SMALL["nlon"]
        m = reference_model(cfg, nlat, nlon)
        with torch.no_grad():       # move every parameter off its init value so all gradients are informative
            gen = torch.Generator().manual_seed(7)
            for p in m.parameters():
                p.add_(0.05 * torch.randn(p.shape, generator=gen))
        G = nlat * nlon
        F, T = cfg["data"]["num_features_used"], cfg["data"]["obs_window_used"]
        gen = torch.Generator().manual_seed(123)
        X = torch.randn(1, G, T * F, generator=gen)
        y = torch.randn(1, G, F, generator=gen)
        kw = dict(batch_num=1) if name == "sparse_attention" else {}
        with contextlib.redirect_stdout(io.StringIO()):
            delta = m(X=X, attention_threshold=0.0, **kw)
        out = X.view(1, G, T, F)[:, :, -1, :] + delta.unsqueeze(0)
        lw = get_lat_weights(nlat, nlon, "cpu")
        loss = weighted_mse_loss(out, y, lw)
        loss.backward()
        arrays = {"X": X.numpy(), "y": y.numpy(), "delta": delta.detach().numpy(), "loss": loss.detach().numpy()}
        for k, v in m.state_dict().items():
            if k != "_processing_edge_features":
                arrays["param/" + k] = v.numpy()
        for k, p in m.named_parameters():
            if p.grad is not None:      # e.g. the sparse_gat GraphLayer's PReLU is constructed but never called
                arrays["grad/" + k] = p.grad.numpy()
        if name == "sparse_attention":   # one pruning call: which edges survive alpha >= threshold
            with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
                m(X=X, attention_threshold=0.05, batch_num=0)
            arrays["pruned_edge_index"] = m.processing_graph.numpy()
        meta = dict(workload=name, nlat=nlat, nlon=nlon, mesh_levels=SMALL["mesh_levels"], radius_factor=SMALL["rq"])
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        path = os.path.join(out_dir, f"model_{name}.npz")
        np.savez_compressed(path, **arrays)
        print(name, "loss", float(loss), "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
