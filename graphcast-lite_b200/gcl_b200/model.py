"""Batched, fused mirror of the reference's encode-process-decode model.

Same module tree and parameter names as /root/reference/src/models.py (MLP :54-109, GraphLayer
:289-440, Model :443-473, WeatherPrediction :476-874), so a reference ``state_dict`` loads unchanged
(the InteractionNet-only buffer ``_processing_edge_features`` is ignored), but

  * inputs may be [B, G, T*F]: B forecast samples share the three static graphs (the reference is
    batch-1 only, models.py:822);
  * Linear+bias+PReLU and aggregate+bias+PReLU are single kernels, graphs are device CSR built once,
    the encoder input is assembled by one kernel instead of zeros + 3 cats (models.py:776-806).

The InteractionNet processor of the v2 configs (models.py:166-285) is built from the same kernels (see
InteractionNetLayer).  Product graph and regional meshes are out of scope.
"""
import os
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi, ops
from .graph import CSR_LOOPS, CSR_RAW, GLOBAL_CACHE, NORM_GCN, NORM_MEAN, EdgeOps
from .graphs_build import ModelGraphs
from .nn import GATConv, GCNConv, LayerNorm, SimpleConv


def _truthy(v) -> bool:
    return v is True or (isinstance(v, str) and v.lower() == "true")


class MLP(nn.Module):
    """Linear(+PReLU) stack (+ LayerNorm); the attribute is called ``MLP`` as in the reference."""

    def __init__(self, cfg: dict, input_dim: int):
        super().__init__()
        self.MLP = nn.ModuleList()
        last = input_dim
        for h in (cfg.get("mlp_hidden_dims") or []):
            self.MLP.extend([nn.Linear(last, h), nn.PReLU()])
            last = h
        self.MLP.append(nn.Linear(last, cfg["output_dim"]))
        if _truthy(cfg.get("use_layer_norm")):
            self.MLP.append(LayerNorm(cfg["output_dim"], mode=cfg.get("layer_norm_mode") or "graph"))

    def forward(self, X):
        # Linear + bias + PReLU is one kernel that writes the pre-activation z and a = PReLU(z); the chain is
        # differentiated through z, and each PReLU's backward runs inside the NEXT Linear's dX kernel
        # (ops._ActLinear).  (z, slope) is the pending activation whose output X currently is.
        mods = list(self.MLP)
        i, z, slope = 0, None, None
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Linear):
                nxt = mods[i + 1] if i + 1 < len(mods) else None
                so = nxt.weight if isinstance(nxt, nn.PReLU) else None
                res = ops.act_linear(z, X, slope, _fit(m.weight, X), m.bias, so)
                if so is not None:
                    (z, X), slope = res, so
                    i += 2
                else:
                    X, z, slope = res, None, None
                    i += 1
                continue
            if z is not None:                  # someone other than a Linear consumes the activation: make it a
                X, z, slope = ops.prelu(z, slope), None, None      # differentiable tensor again
            if isinstance(m, nn.PReLU):
                X = ops.prelu(X, m.weight)
            else:
                X = m(X)
            i += 1
        if z is not None:
            X = ops.prelu(z, slope)
        return X


class ReLU(nn.Module):
    """_get_activation("relu") (models.py:154-163) on the gcl kernels."""

    def forward(self, x):
        return ops.act(x, ops.ACT_RELU)


class SiLU(nn.Module):
    """_get_activation("swish" | "silu")."""

    def forward(self, x):
        return ops.act(x, ops.ACT_SILU)


def _get_activation(name: str = "prelu") -> nn.Module:
    if name in ("swish", "silu"):
        return SiLU()
    if name == "prelu":
        return nn.PReLU()
    if name == "relu":
        return ReLU()
    raise ValueError(f"Unknown activation: {name}")


def _act_apply(m: nn.Module, x):
    return ops.prelu(x, m.weight) if isinstance(m, nn.PReLU) else m(x)


class InteractionNetLayer(nn.Module):
    """One InteractionNetwork step (models.py:166-237), same parameters (edge_mlp.{0,2}, node_mlp.{0,2}, edge_norm,
    node_norm), for B samples at once.  The first Linear of each MLP acts on a concatenation; it is evaluated as the
    sum of its blocks (W [x_s | x_r | e] = W_s x_s + W_r x_r + W_e e), which moves the two node blocks in front of the
    gather: two node-level GEMMs + two row gathers instead of a GEMM over [E, 3C]."""

    def __init__(self, node_dim: int, edge_dim: int, hidden_dim: int, activation: str = "swish",
                 use_layer_norm: bool = True):
        super().__init__()
        act = _get_activation(activation)
        self.edge_mlp = nn.Sequential(nn.Linear(node_dim * 2 + edge_dim, hidden_dim), act, nn.Linear(hidden_dim, edge_dim))
        self.node_mlp = nn.Sequential(nn.Linear(node_dim + edge_dim, hidden_dim), act, nn.Linear(hidden_dim, node_dim))
        self.use_layer_norm = use_layer_norm
        self.node_dim, self.edge_dim = node_dim, edge_dim
        if use_layer_norm:
            self.edge_norm = LayerNorm(edge_dim, mode="graph")
            self.node_norm = LayerNorm(node_dim, mode="node")

    def forward(self, x, eo: EdgeOps, edge_attr):
        nd = self.node_dim
        W1, b1 = self.edge_mlp[0].weight, self.edge_mlp[0].bias
        ps = ops.spmm_fixed(ops.linear(x, W1[:, :nd]), *eo.gather_src)                 # (W_s x)[senders]
        pr = ops.spmm_fixed(ops.linear(x, W1[:, nd:2 * nd]), *eo.gather_dst)           # (W_r x)[receivers]
        h = ops.add(ops.add(ps, pr), ops.linear(edge_attr, W1[:, 2 * nd:], b1))
        edge_update = ops.linear(_act_apply(self.edge_mlp[1], h), self.edge_mlp[2].weight, self.edge_mlp[2].bias)
        aggregated = ops.spmm_fixed(edge_update, *eo.mean_dst)                           # scatter(..., reduce="mean")
        Wn, bn = self.node_mlp[0].weight, self.node_mlp[0].bias
        hn = ops.add(ops.linear(x, Wn[:, :nd], bn), ops.linear(aggregated, Wn[:, nd:]))
        node_update = ops.linear(_act_apply(self.node_mlp[1], hn), self.node_mlp[2].weight, self.node_mlp[2].bias)
        new_edge, new_x = ops.add(edge_attr, edge_update), ops.add(x, node_update)
        if self.use_layer_norm:
            new_edge, new_x = self.edge_norm(new_edge), self.node_norm(new_x)
        return new_x, new_edge


class InteractionNetProcessor(nn.Module):
    """N unshared InteractionNet steps behind an edge-feature encoder (models.py:239-285)."""

    def __init__(self, node_dim: int, raw_edge_dim: int, edge_latent_dim: int, hidden_dim: int, num_steps: int,
                 activation: str = "swish", use_layer_norm: bool = True):
        super().__init__()
        self.edge_encoder = nn.Sequential(nn.Linear(raw_edge_dim, edge_latent_dim), _get_activation(activation))
        self.steps = nn.ModuleList([InteractionNetLayer(node_dim, edge_latent_dim, hidden_dim, activation, use_layer_norm)
                                    for _ in range(num_steps)])
        self._edge_ops = None

    def forward(self, x, edge_index, edge_attr_raw):
        if self._edge_ops is None or self._edge_ops[0] is not edge_index:
            self._edge_ops = (edge_index, EdgeOps(edge_index, x.size(-2)))
        eo = self._edge_ops[1]
        e = _act_apply(self.edge_encoder[1], ops.linear(edge_attr_raw, self.edge_encoder[0].weight, self.edge_encoder[0].bias))
        if x.dim() == 3:                      # the encoded edge features are the same for every sample
            e = e.unsqueeze(0).expand(x.size(0), -1, -1).contiguous()
        for step in self.steps:
            x, e = step(x, eo, e)
        return x


class SparseGATConv(GATConv):
    """models.py:112-151.  For B > 1 the pruning decision uses the batch-mean attention (over the global batch when
    data-parallel, so every replica prunes to the same graph)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=False, dropout=0.0, bias=True, **kw):
        super().__init__(in_channels, out_channels, heads, concat=concat, dropout=dropout, bias=bias, **kw)
        self.process_group = None        # set by Trainer(process_group=...): the replicas that prune together

    def forward(self, x, edge_index, attention_threshold=0.0, **kwargs):
        batch_num = kwargs.get("batch_num", 1)
        out, (edge_index, att) = super().forward(x, edge_index, return_attention_weights=True)
        if att.dim() == 3:
            att = att.mean(dim=0)
        att = att.squeeze()
        if batch_num == 0:
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1:
                # data-parallel replicas must keep the SAME edge set: prune on the attention averaged over the
                # global batch of the trainer's process group (the reference is single-process, models.py:140-149)
                att = att.contiguous()
                dist.all_reduce(att, op=dist.ReduceOp.SUM, group=self.process_group)
                att = att / dist.get_world_size(self.process_group)
            edge_index = ops.edge_prune(edge_index, att, float(attention_threshold))
        return out, (edge_index, att)


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
            if self.output_dim != input_dim:
                raise ValueError(f"InteractionNet requires output_dim ({self.output_dim}) == input_dim ({input_dim})")
            use_ln = cfg.get("use_layer_norm")
            self.layers = InteractionNetProcessor(
                node_dim=input_dim, raw_edge_dim=cfg.get("edge_feature_dim") or 4, edge_latent_dim=input_dim,
                hidden_dim=input_dim, num_steps=cfg.get("num_message_passing_steps") or 4,
                activation=cfg.get("activation") or "swish", use_layer_norm=True if use_ln is None else _truthy(use_ln))
            return
        if self.layer_type not in ("conv_gcn", "conv_gat", "sparse_gat"):
            raise NotImplementedError(f"gcl_b200: layer type {self.layer_type!r} is out of scope (SURVEY.md 8)")
        self.activation = _get_activation(cfg.get("activation") or "prelu")
        self.output_dim = cfg["output_dim"]
        self.layers = nn.ModuleList()
        hid = list(cfg.get("hidden_dims") or [])
        if self.layer_type == "sparse_gat":
            self.layers.append(SparseGATConv(input_dim, self.output_dim, heads=cfg["gat_props"]["num_heads"],
                                             concat=False))
        else:
            def conv(i, o):
                if self.layer_type == "conv_gcn":
                    return GCNConv(i, o)
                return GATConv(i, o, heads=cfg["gat_props"]["num_heads"], concat=False)
            dims = [input_dim] + hid
            for i in range(len(hid)):
                self.layers.append(conv(dims[i], dims[i + 1]))
                self.layers.append(self.activation)      # ONE shared PReLU, appended several times (models.py:316)
            self.layers.append(conv(dims[-1], self.output_dim))
        if _truthy(cfg.get("use_layer_norm")):
            self.layers.append(LayerNorm(self.output_dim, mode=cfg.get("layer_norm_mode") or "graph"))

    def forward(self, X, edge_index, attention_threshold=0.0, rows_out=None, **kwargs):
        """rows_out = n: the caller only uses nodes 0..n-1 of the result (the decoder: models.py:852 keeps the grid
        rows).  Honoured when the last module is a GCNConv -- its aggregation then produces just those receivers."""
        if self.layer_type == "simple_conv":
            return self.layers(x=X, edge_index=edge_index)
        if self.layer_type == "interaction_net":                     # models.py:435-439
            edge_attr = kwargs.get("edge_attr")
            if edge_attr is None:
                raise ValueError("InteractionNet requires edge_attr (edge features)")
            return self.layers(X, edge_index, edge_attr)
        if self.layer_type == "sparse_gat":
            for layer in self.layers:
                if type(layer) is SparseGATConv:
                    X, (edge_index, _) = layer.forward(X, edge_index, attention_threshold, **kwargs)
                else:
                    X = layer(X)
            return X, edge_index
        mods = list(self.layers)
        n = X.size(-2)
        i = 0
        # (z, slope): X is PReLU(z) coming out of a fused GCNConv + PReLU whose consumer is the next GCNConv's
        # Linear -- the chain is differentiated through z and that PReLU's backward runs in the epilogue of the
        # next layer's dX GEMM (ops.act_linear), as in MLP.forward
        z, slope, sink = None, None, None
        chain = os.environ.get("GCL_NO_GCN_CHAIN") != "1"
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            fuse = isinstance(nxt, nn.PReLU)
            if type(m) is GCNConv:
                g = GLOBAL_CACHE.get(edge_index, n, CSR_LOOPS)
                W, b, C = _fit(m.lin.weight, X), m.bias, m.out_channels
                if C % 4:
                    # rows of 4k+r floats would force every kernel of this layer onto its scalar path: compute a
                    # zero-padded 4-aligned layer (extra output channels are exactly 0) and slice the result
                    W = F.pad(W, (0, 0, 0, 4 - C % 4))
                    b = F.pad(b, (0, 4 - C % 4)) if b is not None else None
                last = i + (2 if fuse else 1) >= len(mods)
                h = ops.act_linear(z, X, slope, W, sink=sink) if z is not None else ops.linear(X, W)
                z, slope, sink = None, None, None
                after = mods[i + 2] if i + 2 < len(mods) else None
                # worth it when the next layer's dX runs on the tensor-core kernels with the fused epilogues
                if chain and fuse and type(after) is GCNConv and C % 4 == 0 and C >= 32 and \
                        after.out_channels >= 32 and after.out_channels % 4 == 0 and X.dtype == torch.float32:
                    sink = ops.ColsumSink() if b is not None else None
                    z, X = ops.aggregate_pre(h, g, NORM_GCN, b, nxt.weight, sink)
                    slope = nxt.weight
                    i += 2
                    continue
                X = ops.aggregate(h, g, NORM_GCN, b, nxt.weight if fuse else None,     # + bias + PReLU fused
                                  rows_out if last else None)
                if C % 4:
                    X = ops.resize_channels(X, C)
                i += 2 if fuse else 1
            elif type(m) is GATConv:
                if fuse and m.heads == 1:                    # GATConv + the shared PReLU in one aggregation kernel
                    g = GLOBAL_CACHE.get(edge_index, n, CSR_LOOPS if m.add_self_loops else CSR_RAW)
                    z, a_s, a_d = ops.linear_scores(X, m.lin.weight, m.att_src, m.att_dst)
                    X, _ = ops.gat_attend(z, m.att_src, m.att_dst, m.bias, g, 1, m.concat, m.negative_slope,
                                          prelu_slope=nxt.weight, scores=(a_s, a_d))
                    i += 2
                else:
                    X = m(X, edge_index)
                    i += 1
            elif isinstance(m, nn.PReLU):
                X = ops.prelu(X, m.weight)
                i += 1
            else:
                X = m(X)                      # ReLU / SiLU (v2 configs), LayerNorm
                i += 1
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


def _fit(W, X):
    """Weight [out, in] widened with zero columns to the (zero-padded) input width."""
    extra = X.shape[-1] - W.shape[1]
    return F.pad(W, (0, extra)) if extra > 0 else W


class _AssembleInput(torch.autograd.Function):
    """[B,G,TF] -> [B,G+M,TF+S]: grid rows get their static features appended, mesh rows are
    zeros + static features (models.py:776-806), written by one kernel."""

    @staticmethod
    def forward(ctx, x, grid_static, mesh_static, width=None):
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("gcl_b200: model input must be a float32 CUDA tensor; no CPU fallback")
        x = x.contiguous()
        B, G, TF = x.shape
        M, S = mesh_static.shape
        width = TF + S if width is None else int(width)
        out = torch.empty((B, G + M, width), dtype=torch.float32, device=x.device)
        lib = _cabi.load()
        with torch.cuda.device(x.device):
            _cabi.check(lib.gcl_assemble_input_f32(x.data_ptr(), grid_static.data_ptr(), mesh_static.data_ptr(),
                                                   out.data_ptr(), B, G, M, TF, S, width,
                                                   torch.cuda.current_stream().cuda_stream),
                        "gcl_assemble_input_f32")
        ctx.G, ctx.TF = G, TF
        return out

    @staticmethod
    def backward(ctx, d):
        return d[:, : ctx.G, : ctx.TF].contiguous(), None, None, None


class WeatherPrediction(nn.Module):
    """cfg = {"graph": ..., "pipeline": ..., "data": ...} in the reference's config.json schema
    (see gcl_b200.workloads).  forward(X [B,G,T*F] | [G,T*F]) -> [B,G,F_out] | [G,F_out]."""

    def __init__(self, cfg: dict, nlat: int, nlon: int, device="cuda:0", graphs: Optional[ModelGraphs] = None):
        super().__init__()
        if cfg["pipeline"].get("product_graph"):
            raise NotImplementedError("gcl_b200: the product graph is out of scope (SURVEY.md 8)")
        self.device = torch.device(device)
        g = graphs or ModelGraphs(nlat, nlon, cfg["graph"]["mesh_levels"], cfg["graph"]["grid2mesh_radius_query"],
                                  self.device)
        self.graphs = g
        self._num_grid_nodes, self._num_mesh_nodes = g.num_grid, g.num_mesh
        self.obs_window = cfg["data"]["obs_window_used"]
        self.num_features = cfg["data"]["num_features_used"]
        self.total_feature_size = self.obs_window * self.num_features
        # plain attributes, not buffers: like the reference they stay out of the state_dict (models.py:597-601)
        self.encoding_graph, self.processing_graph, self.decoding_graph = (
            g.encoding_graph, g.processing_graph, g.decoding_graph)
        self.init_grid_features, self.init_mesh_features = g.init_grid_features, g.init_mesh_features
        pipe = cfg["pipeline"]
        self.using_sparse_gat = pipe["processor"]["gcn"]["layer_type"] == "sparse_gat"
        self.using_interaction_net = pipe["processor"]["gcn"]["layer_type"] == "interaction_net"
        # 4-d mesh edge features (create_graphs.py:37-91); a buffer in the reference (models.py:558), so it is one here
        self.register_buffer("_processing_edge_features",
                             g.processing_edge_features.clone() if self.using_interaction_net else None)
        self.encoder = Model(pipe["encoder"], self.total_feature_size + self.init_grid_features.shape[1])
        self.processor = Model(pipe["processor"], self.encoder.output_dim)
        self.decoder = Model(pipe["decoder"], self.processor.output_dim)
        first = self.encoder.mlp.MLP[0] if self.encoder.mlp is not None else self.encoder.graph_layer.layers[0]
        self._pad_input = type(first) in (nn.Linear, GCNConv)       # both take a zero-padded weight (_fit)
        self.to(self.device)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        # the reference registers the edge features as a buffer for every model (models.py:558); here only the
        # InteractionNet models keep them
        sd = {k: v for k, v in state_dict.items() if k != "_processing_edge_features" or self.using_interaction_net}
        return super().load_state_dict(sd, strict=strict, **kw)

    def forward(self, X: torch.Tensor, attention_threshold=0.0, **kwargs):
        squeeze = False
        if X.dim() == 2:
            X, squeeze = X.unsqueeze(0), True
        elif X.dim() == 3 and X.size(0) == 1:
            squeeze = True                       # the reference returns [G, F] for its batch of one
        G = self._num_grid_nodes
        # rows padded to a multiple of 4 floats (zeros) when the first consumer can take a zero-padded weight
        width = self.total_feature_size + self.init_grid_features.shape[1]
        if self._pad_input:
            width = (width + 3) // 4 * 4
        enc_in = _AssembleInput.apply(X, self.init_grid_features, self.init_mesh_features, width)
        enc = self.encoder(X=enc_in, edge_index=self.encoding_graph)
        # models.py:841-842 / :865 slice the encoder output into grid and mesh rows and concatenate the grid rows with
        # the processed mesh rows.  Here only the mesh rows move: they are copied out for the processor and its
        # result is written back over them, so the decoder reads [grid_lat ; proc] where enc already lies.
        in_place = os.environ.get("GCL_NO_ROW_BRIDGE") != "1"        # A/B switch: the two-copy split / concat
        if in_place:
            bridge = ops.RowBridge()
            mesh_lat = ops.take_rows(enc, G, bridge)
        else:
            grid_lat, mesh_lat = ops.split_rows(enc, G)
        if self.using_sparse_gat:
            proc, new_ei = self.processor(X=mesh_lat, edge_index=self.processing_graph,
                                          attention_threshold=attention_threshold, **kwargs)
            self.processing_graph = new_ei           # models.py:846
        elif self.using_interaction_net:                              # models.py:847-853
            proc = self.processor(X=mesh_lat, edge_index=self.processing_graph,
                                  attention_threshold=attention_threshold, edge_attr=self._processing_edge_features)
        else:
            proc = self.processor(X=mesh_lat, edge_index=self.processing_graph,
                                  attention_threshold=attention_threshold)
        dec_in = ops.put_rows(enc, proc, G, bridge) if in_place else ops.concat_rows(grid_lat, proc)
        dec = self.decoder(X=dec_in, edge_index=self.decoding_graph, rows_out=G)
        out = dec[:, :G]
        return out.squeeze(0) if squeeze else out
