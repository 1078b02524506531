"""ORACLE / TEST INFRASTRUCTURE ONLY -- not product code, never imported by gcl_b200.

CPU restatement of the reference's encode-process-decode glue so that the whole forecast step can
be checked (and timed as the CPU baseline) on the GPU box, where /root/reference does not exist:

  MLP                /root/reference/src/models.py:54-109
  SparseGATConv      /root/reference/src/models.py:112-151
  GraphLayer         /root/reference/src/models.py:289-440   (SimpleConv / ConvGCN / GATConv / SparseGATConv)
  Model              /root/reference/src/models.py:443-473
  InteractionNet     /root/reference/src/models.py:166-285
  WeatherPrediction  /root/reference/src/models.py:476-601, 776-874  (no product graph)
  weighted MSE + AR rollout of one training step   /root/reference/src/train.py:85-102, 160-233

The conv / norm arithmetic comes from oracle/pyg_shim (restated torch_geometric 2.5.3).  Module and
parameter names equal the reference's, so state_dicts are interchangeable.  Pinned in the build
container by tests/test_oracle.py::test_model_glue_matches_unmodified_reference (bit-exact against
/root/reference/src/models.py run on the same shims) and on the GPU box through tests/golden/*.npz
written by oracle/make_golden.py.  The conv arithmetic itself stays PARITY UNPINNED (see pyg_shim).

Extension: inputs may be [B, G, T*F]; samples are processed one at a time exactly as batch 1.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

_here = os.path.dirname(os.path.abspath(__file__))
for _p in ("pyg_shim", "trimesh_shim"):
    if os.path.join(_here, _p) not in sys.path:
        sys.path.insert(0, os.path.join(_here, _p))
from torch_geometric.nn import GATConv, GCNConv, LayerNorm, SimpleConv  # noqa: E402

from . import graphs as og  # noqa: E402


def _truthy(v):
    return v is True or (isinstance(v, str) and v.lower() == "true")


class MLP(nn.Module):
    def __init__(self, cfg: dict, input_dim: int):
        super().__init__()
        self.MLP = nn.ModuleList()
        last = input_dim
        for h in (cfg.get("mlp_hidden_dims") or []):
            self.MLP.extend([nn.Linear(last, h), nn.PReLU()])
            last = h
        self.MLP.append(nn.Linear(last, cfg["output_dim"]))
        if _truthy(cfg.get("use_layer_norm")):
            self.MLP.append(LayerNorm(cfg["output_dim"], mode=cfg.get("layer_norm_mode")))

    def forward(self, X):
        for layer in self.MLP:
            X = layer(X)
        return X


class SparseGATConv(GATConv):
    def __init__(self, in_channels, out_channels, heads=1, concat=False, dropout=0.0, bias=True, **kw):
        super().__init__(in_channels, out_channels, heads, concat=concat, dropout=dropout, bias=bias, **kw)

    def forward(self, x, edge_index, attention_threshold=0.0, **kwargs):
        batch_num = kwargs.get("batch_num", 1)
        out, (edge_index, att) = super().forward(x, edge_index, return_attention_weights=True)
        att = att.squeeze()
        if batch_num == 0:
            mask = (att >= attention_threshold).type(torch.bool)
            edge_index, att = edge_index[:, mask], att[mask]
        return out, (edge_index, att)


_ACTS = {"prelu": nn.PReLU, "relu": nn.ReLU, "silu": nn.SiLU, "swish": nn.SiLU}


class InteractionNetLayer(nn.Module):
    """models.py:166-237: edge MLP on [x_s | x_r | e], scatter-mean onto receivers, node MLP on [x | agg], residuals,
    LayerNorm(graph) on the edges and LayerNorm(node) on the nodes."""

    def __init__(self, node_dim, edge_dim, hidden_dim, activation="swish", use_layer_norm=True):
        super().__init__()
        act = _ACTS[activation]()
        self.edge_mlp = nn.Sequential(nn.Linear(node_dim * 2 + edge_dim, hidden_dim), act, nn.Linear(hidden_dim, edge_dim))
        self.node_mlp = nn.Sequential(nn.Linear(node_dim + edge_dim, hidden_dim), act, nn.Linear(hidden_dim, node_dim))
        self.use_layer_norm = use_layer_norm
        if use_layer_norm:
            self.edge_norm = LayerNorm(edge_dim, mode="graph")
            self.node_norm = LayerNorm(node_dim, mode="node")

    def forward(self, x, edge_index, edge_attr):
        from torch_geometric.utils import scatter
        senders, receivers = edge_index[0], edge_index[1]
        edge_update = self.edge_mlp(torch.cat([x[senders], x[receivers], edge_attr], dim=-1))
        aggregated = scatter(edge_update, receivers, dim=0, dim_size=x.size(0), reduce="mean")
        node_update = self.node_mlp(torch.cat([x, aggregated], dim=-1))
        new_edge_attr, new_x = edge_attr + edge_update, x + node_update
        if self.use_layer_norm:
            new_edge_attr, new_x = self.edge_norm(new_edge_attr), self.node_norm(new_x)
        return new_x, new_edge_attr


class InteractionNetProcessor(nn.Module):
    """models.py:239-285."""

    def __init__(self, node_dim, raw_edge_dim, edge_latent_dim, hidden_dim, num_steps, activation="swish",
                 use_layer_norm=True):
        super().__init__()
        self.edge_encoder = nn.Sequential(nn.Linear(raw_edge_dim, edge_latent_dim), _ACTS[activation]())
        self.steps = nn.ModuleList([InteractionNetLayer(node_dim, edge_latent_dim, hidden_dim, activation, use_layer_norm)
                                    for _ in range(num_steps)])

    def forward(self, x, edge_index, edge_attr_raw):
        edge_attr = self.edge_encoder(edge_attr_raw)
        for step in self.steps:
            x, edge_attr = step(x, edge_index, edge_attr)
        return x


class GraphLayer(nn.Module):
    def __init__(self, cfg: dict, input_dim: int):
        super().__init__()
        self.layer_type = cfg["layer_type"]
        if self.layer_type == "simple_conv":
            self.output_dim = input_dim
            self.layers = SimpleConv(aggr="mean")
            return
        if self.layer_type == "interaction_net":                    # models.py:376-398
            self.output_dim = cfg["output_dim"]
            assert self.output_dim == input_dim
            use_ln = cfg.get("use_layer_norm")
            self.layers = InteractionNetProcessor(input_dim, cfg.get("edge_feature_dim") or 4, input_dim, input_dim,
                                                  cfg.get("num_message_passing_steps") or 4,
                                                  cfg.get("activation") or "swish",
                                                  True if use_ln is None else _truthy(use_ln))
            return
        assert self.layer_type in ("conv_gcn", "conv_gat", "sparse_gat"), self.layer_type
        self.activation = _ACTS[cfg.get("activation") or "prelu"]()
        self.output_dim = cfg["output_dim"]
        self.layers = nn.ModuleList()
        hid = cfg.get("hidden_dims") or []
        if self.layer_type == "sparse_gat":
            self.layers.append(SparseGATConv(input_dim, self.output_dim,
                                             heads=cfg["gat_props"]["num_heads"], concat=False))
        else:
            def conv(i, o):
                if self.layer_type == "conv_gcn":
                    return GCNConv(i, o)
                return GATConv(i, o, heads=cfg["gat_props"]["num_heads"], concat=False)
            dims = [input_dim] + list(hid)
            for i in range(len(hid)):
                self.layers.append(conv(dims[i], dims[i + 1]))
                self.layers.append(self.activation)
            self.layers.append(conv(dims[-1], self.output_dim))
        if _truthy(cfg.get("use_layer_norm")):
            self.layers.append(LayerNorm(self.output_dim, mode=cfg.get("layer_norm_mode")))

    def forward(self, X, edge_index, attention_threshold=0.0, **kwargs):
        if self.layer_type == "simple_conv":
            return self.layers(x=X, edge_index=edge_index)
        if self.layer_type == "interaction_net":
            return self.layers(x=X, edge_index=edge_index, edge_attr_raw=kwargs["edge_attr"])
        if self.layer_type == "sparse_gat":
            for layer in self.layers:
                if type(layer) is SparseGATConv:
                    X, (edge_index, _) = layer.forward(X, edge_index, attention_threshold, **kwargs)
                else:
                    X = layer(X)
            return X, edge_index
        for layer in self.layers:
            X = layer(X, edge_index) if type(layer) in (GCNConv, GATConv) else layer(X)
        return X


class Model(nn.Module):
    def __init__(self, cfg: dict, input_dim: int):
        super().__init__()
        self.mlp = MLP(cfg["mlp"], input_dim) if cfg.get("mlp") else None
        gin = cfg["mlp"]["output_dim"] if cfg.get("mlp") else input_dim
        self.graph_layer = GraphLayer(cfg["gcn"], gin)
        self.output_dim = self.graph_layer.output_dim

    def forward(self, X, edge_index, attention_threshold=0.0, **kwargs):
        if self.mlp is not None:
            X = self.mlp(X)
        return self.graph_layer(X=X, edge_index=edge_index, attention_threshold=attention_threshold, **kwargs)


class WeatherPrediction(nn.Module):
    """cfg = {"graph": ..., "pipeline": ..., "data": ...} with the reference's config.json schema."""

    def __init__(self, cfg: dict, nlat: int, nlon: int, graphs: dict = None):
        super().__init__()
        g = graphs or og.build_graphs(nlat, nlon, cfg["graph"]["mesh_levels"],
                                      cfg["graph"]["grid2mesh_radius_query"])
        self.G, self.M = g["num_grid"], g["num_mesh"]
        self.obs_window = cfg["data"]["obs_window_used"]
        self.num_features = cfg["data"]["num_features_used"]
        self.total_feature_size = self.obs_window * self.num_features
        self.encoding_graph = torch.as_tensor(g["g2m"])
        self.processing_graph = torch.as_tensor(g["mesh"])
        self.decoding_graph = torch.as_tensor(g["m2g"])
        self.init_grid_features = torch.as_tensor(g["grid_feats"])
        self.init_mesh_features = torch.as_tensor(g["mesh_feats"])
        pipe = cfg["pipeline"]
        self.using_sparse_gat = pipe["processor"]["gcn"]["layer_type"] == "sparse_gat"
        self.using_interaction_net = pipe["processor"]["gcn"]["layer_type"] == "interaction_net"
        self.register_buffer("_processing_edge_features",
                             torch.as_tensor(g["mesh_edge_feats"]) if self.using_interaction_net else None)
        self.encoder = Model(pipe["encoder"], self.total_feature_size + 6)
        self.processor = Model(pipe["processor"], self.encoder.output_dim)
        self.decoder = Model(pipe["decoder"], self.processor.output_dim)

    def _one(self, X, attention_threshold, **kwargs):
        X = torch.cat((X, self.init_grid_features), dim=-1)
        mesh = torch.cat((torch.zeros(self.M, self.total_feature_size, device=X.device), self.init_mesh_features), dim=-1)
        enc = self.encoder(X=torch.cat((X, mesh), dim=0), edge_index=self.encoding_graph)
        grid_lat, mesh_lat = enc[: self.G], enc[self.G:]
        if self.using_sparse_gat:
            proc, new_ei = self.processor(X=mesh_lat, edge_index=self.processing_graph,
                                          attention_threshold=attention_threshold, **kwargs)
            self.processing_graph = new_ei
        elif self.using_interaction_net:
            proc = self.processor(X=mesh_lat, edge_index=self.processing_graph, attention_threshold=attention_threshold,
                                  edge_attr=self._processing_edge_features)
        else:
            proc = self.processor(X=mesh_lat, edge_index=self.processing_graph,
                                  attention_threshold=attention_threshold)
        dec = self.decoder(X=torch.cat((grid_lat, proc), dim=0), edge_index=self.decoding_graph)
        return dec[: self.G]

    def forward(self, X, attention_threshold=0.0, **kwargs):
        if X.dim() == 3 and X.size(0) > 1:
            return torch.stack([self._one(x, attention_threshold, **kwargs) for x in X])
        return self._one(X.squeeze(0) if X.dim() == 3 else X, attention_threshold, **kwargs)


def lat_weights(nlat: int, nlon: int) -> torch.Tensor:
    """train.py:53-72 (regular grid): cos(lat)/mean, expanded [lon, lat] then flattened -> [1,G,1]."""
    w = torch.cos(torch.deg2rad(torch.linspace(-90, 90, nlat)))
    w = w / w.mean()
    return w.view(1, -1).expand(nlon, nlat).reshape(-1).view(1, -1, 1)


def weighted_mse(pred, target, lat_w=None, channel_mask=None, spatial_mask=None):
    """train.py:85-102: squared error weighted by channel mask [C], spatial mask [1,G,1] and latitude weights
    [1,G,1], normalised by the sum of the weights."""
    diff = (pred - target) ** 2
    w = torch.ones_like(diff)
    if channel_mask is not None:
        w = w * channel_mask.view(1, 1, -1)
    if spatial_mask is not None:
        w = w * spatial_mask
    if lat_w is not None:
        w = w * lat_w
    return (diff * w).sum() / w.sum().clamp_min(1e-12)


def training_loss(model: WeatherPrediction, X, y, ar_steps: int, lat_w=None, threshold=0.0,
                  use_residual=True, static_channels=(), forcing_channels=(), channel_mask=None, spatial_mask=None,
                  **kwargs):
    """Loss of one train_epoch iteration (train.py:173-231): AR rollout with static / forcing carry-forward
    (:218-226), BPTT through it."""
    if X.dim() == 2:
        X, y = X.unsqueeze(0), y.unsqueeze(0)
    Bn, G, _ = X.shape
    obs = model.obs_window
    C = X.shape[-1] // obs
    tsteps = y.shape[-1] // C
    ys = y.view(Bn, G, tsteps, C)
    state = X.view(Bn, G, obs, C)
    steps = min(ar_steps, tsteps)
    loss = 0
    for s in range(steps):
        delta = model(X=state.reshape(Bn, G, -1), attention_threshold=threshold, **kwargs)
        if delta.dim() == 2:
            delta = delta.unsqueeze(0)
        out = state[:, :, -1, :] + delta if use_residual else delta
        loss = loss + weighted_mse(out, ys[:, :, s, :], lat_w, channel_mask, spatial_mask)
        if static_channels:
            static_vals = state[:, :, -1, :]
            for ch in static_channels:
                out[:, :, ch] = static_vals[:, :, ch]
        if forcing_channels and s < tsteps:
            forcing_vals = ys[:, :, s, :]
            for ch in forcing_channels:
                out[:, :, ch] = forcing_vals[:, :, ch]
        state = torch.cat([state[:, :, 1:, :], out.unsqueeze(2)], dim=2)
    return loss / steps
