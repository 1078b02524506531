"""The five BASELINE.json workloads, restated as plain dicts in the reference's config.json schema
(graph / pipeline / data sections only -- everything the hot path reads).

Sources: /root/reference/experiments/{baseline,attention,sparse_attention,
wb2_64x32_ar_15f_4obs_4pred,wb2_512x256_19f_ar}/config.json.  Grid sizes come from the dataset
metadata the reference pairs with each config (src/data/data_configs.py): 64x32 and 512x256.
A user's own config.json (same schema) can be passed anywhere one of these dicts is accepted.
"""
from copy import deepcopy


def _mlp(hidden, out, ln):
    return {"mlp_hidden_dims": list(hidden), "output_dim": out, "use_layer_norm": ln,
            "layer_norm_mode": "node" if ln else None}


def _gcn(layer_type, hidden, out, ln=False, heads=None):
    d = {"layer_type": layer_type, "hidden_dims": list(hidden), "output_dim": out,
         "use_layer_norm": ln, "layer_norm_mode": "node" if ln else None, "activation": "prelu"}
    if heads is not None:
        d["gat_props"] = {"num_heads": heads, "sparsity_thresholds": [0.0, 0.0]}
    return d


def _graph(levels, rq):
    return {"grid2mesh_edge_creation": "radius", "mesh2grid_edge_creation": "contained",
            "grid2mesh_radius_query": rq, "mesh_levels": list(levels)}


def _data(f, obs, pred):
    return {"num_features_used": f, "obs_window_used": obs, "pred_window_used": pred}


WORKLOADS = {
    "baseline": {
        "nlat": 32, "nlon": 64, "max_ar_steps": 1, "learning_rate": 1e-3,
        "graph": _graph([3, 5], 0.5), "data": _data(33, 2, 1),
        "pipeline": {
            "encoder": {"mlp": _mlp([48, 48], 64, True), "gcn": _gcn("conv_gcn", [64, 64], 64)},
            "processor": {"gcn": _gcn("conv_gcn", [64, 64], 64, ln=True, heads=1)},
            "decoder": {"mlp": _mlp([64, 64], 64, False), "gcn": _gcn("conv_gcn", [48, 48], 33)}}},
    "attention": {
        "nlat": 32, "nlon": 64, "max_ar_steps": 1, "learning_rate": 1e-3,
        "graph": _graph([3, 5], 0.5), "data": _data(33, 2, 1),
        "pipeline": {
            "encoder": {"mlp": _mlp([48, 48], 64, True), "gcn": _gcn("conv_gcn", [64, 64], 64)},
            "processor": {"gcn": _gcn("conv_gat", [64, 64], 64, ln=True, heads=1)},
            "decoder": {"mlp": _mlp([64, 64], 64, False), "gcn": _gcn("conv_gcn", [48, 48], 33)}}},
    "sparse_attention": {
        "nlat": 32, "nlon": 64, "max_ar_steps": 1, "learning_rate": 1e-5,
        "graph": _graph([3, 5], 0.5), "data": _data(12, 2, 1),
        "pipeline": {
            "encoder": {"mlp": _mlp([48, 48], 64, True), "gcn": _gcn("simple_conv", [64, 64], 64)},
            "processor": {"gcn": _gcn("sparse_gat", [], 64, ln=True, heads=1)},
            "decoder": {"mlp": _mlp([64, 64], 12, False), "gcn": _gcn("simple_conv", [48, 48], 12)}}},
    "wb2_64x32_ar_15f_4obs_4pred": {
        "nlat": 32, "nlon": 64, "max_ar_steps": 4, "learning_rate": 1e-3,
        "graph": _graph([4, 6], 0.65), "data": _data(15, 4, 4),
        "pipeline": {
            "encoder": {"mlp": _mlp([64, 64], 64, True), "gcn": _gcn("conv_gcn", [96, 96], 96)},
            "processor": {"gcn": _gcn("conv_gcn", [96, 96, 96], 96, ln=True)},
            "decoder": {"mlp": _mlp([64, 64], 64, False), "gcn": _gcn("conv_gcn", [48, 48], 15)}}},
    "wb2_512x256_19f_ar": {
        "nlat": 256, "nlon": 512, "max_ar_steps": 4, "learning_rate": 5e-4,
        "graph": _graph([4, 6], 0.6), "data": _data(19, 2, 1),
        "pipeline": {
            "encoder": {"mlp": _mlp([128, 128], 128, True), "gcn": _gcn("conv_gcn", [128, 128], 128)},
            "processor": {"gcn": _gcn("conv_gcn", [128, 128, 128, 128], 128, ln=True)},
            "decoder": {"mlp": _mlp([128, 64], 64, False), "gcn": _gcn("conv_gcn", [64, 64], 19)}}},
}

# The InteractionNet family: experiments/wb2_512x256_19f_ar_v2/config.json -- 256-wide, swish,
# 12 unshared message-passing steps with 4-d mesh edge features.  Not a BASELINE.json config; kept beside them for the
# parity tests and an optional bench line (`bench.py --workload wb2_512x256_19f_ar_v2`).
WORKLOADS["wb2_512x256_19f_ar_v2"] = {
    "nlat": 256, "nlon": 512, "max_ar_steps": 4, "learning_rate": 3e-4,
    "graph": _graph([4, 6], 0.6), "data": _data(19, 2, 1),
    "pipeline": {
        "encoder": {"mlp": _mlp([256, 256], 256, True), "gcn": dict(_gcn("conv_gcn", [256, 256], 256), activation="swish")},
        "processor": {"gcn": {"layer_type": "interaction_net", "output_dim": 256, "activation": "swish",
                              "use_layer_norm": True, "num_message_passing_steps": 12, "edge_feature_dim": 4}},
        "decoder": {"mlp": _mlp([256, 128], 128, False), "gcn": dict(_gcn("conv_gcn", [128, 128], 19), activation="swish")}}}

# name -> expected trainable parameter count (209 882 pinned by README_RU.MD:141)
PARAM_COUNTS = {"baseline": 53784, "attention": 54168, "sparse_attention": 20625,
                "wb2_64x32_ar_15f_4obs_4pred": 95782, "wb2_512x256_19f_ar": 209882}


def get_workload(name: str) -> dict:
    if name not in WORKLOADS:
        raise KeyError(f"unknown workload {name!r}; choose from {sorted(WORKLOADS)}")
    return deepcopy(WORKLOADS[name])
