"""ORACLE / TEST INFRASTRUCTURE ONLY -- not product code, never imported by gcl_b200.

CPU restatement (plain PyTorch ops) of the ``torch_geometric.nn`` symbols the reference imports
at /root/reference/src/models.py:21,25,182 -- ``GCNConv``, ``GATConv``, ``SimpleConv``,
``LayerNorm``, ``summary`` -- following torch_geometric==2.5.3 (requirements.txt:6; not vendored,
not installable here).  Same algorithm as upstream: linear -> index_select gather -> multiply ->
scatter_add_.  With this package on sys.path, /root/reference/src/models.py imports and runs
UNMODIFIED on CPU, which is how tests/golden fixtures are produced (oracle/make_golden.py).

PARITY UNPINNED: the reference has no tests or golden vectors for this arithmetic (SURVEY.md 4,
8c).  Pins that do exist and are checked in tests/test_oracle.py: parameter counts 4 288 / 1 488 /
128 / 4 417 and edge counts 1512 / 60->72 / 6144 from notebooks/src/main.ipynb cell 5, 75 522 from
README.md:126, 209 882 params from README_RU.MD:141; plus dense-matrix cross-checks.

Extension (not in PyG): GATConv here also accepts ``x`` of shape [B, N, C] and applies the layer
to every sample independently (GCNConv / SimpleConv take [*, N, C] upstream already, node_dim=-2).
"""
import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor
from torch.nn import Parameter

from torch_geometric.utils import (add_remaining_self_loops, add_self_loops, remove_self_loops,
                                   scatter, softmax)


def glorot_(t: Tensor) -> None:
    stdv = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-stdv, stdv)


class _GlorotLinear(torch.nn.Module):
    """torch_geometric.nn.dense.linear.Linear(in, out, bias=False, weight_initializer='glorot')."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = Parameter(torch.empty(out_channels, in_channels))
        glorot_(self.weight)

    def forward(self, x: Tensor) -> Tensor:
        return F.linear(x, self.weight)


class MessagePassing(torch.nn.Module):
    pass


def gcn_norm(edge_index: Tensor, edge_weight: Optional[Tensor], num_nodes: int,
             improved: bool = False, add_loops: bool = True, dtype=torch.float32):
    """Symmetric normalisation, flow=source_to_target (upstream gcn_conv.gcn_norm)."""
    fill = 2.0 if improved else 1.0
    if add_loops:
        edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill, num_nodes)
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    row, col = edge_index[0], edge_index[1]
    deg = scatter(edge_weight, col, 0, dim_size=num_nodes, reduce="sum")
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float("inf"), 0)
    return edge_index, dis[row] * edge_weight * dis[col]


class GCNConv(MessagePassing):
    def __init__(self, in_channels: int, out_channels: int, improved: bool = False,
                 cached: bool = False, add_self_loops: Optional[bool] = None,
                 normalize: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        if add_self_loops is None:
            add_self_loops = normalize
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached = improved, cached
        self.add_self_loops, self.normalize = add_self_loops, normalize
        self.lin = _GlorotLinear(in_channels, out_channels)
        if bias:
            self.bias = Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def forward(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None) -> Tensor:
        n = x.size(-2)
        if self.normalize:
            edge_index, edge_weight = gcn_norm(edge_index, edge_weight, n, self.improved,
                                               self.add_self_loops, x.dtype)
        x = self.lin(x)
        x_j = x.index_select(-2, edge_index[0])
        if edge_weight is not None:
            x_j = edge_weight.view(-1, 1) * x_j
        out = scatter(x_j, edge_index[1], dim=-2, dim_size=n, reduce="sum")
        if self.bias is not None:
            out = out + self.bias
        return out


class GATConv(MessagePassing):
    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2, dropout: float = 0.0, add_self_loops: bool = True,
                 edge_dim: Optional[int] = None, fill_value="mean", bias: bool = True, **kwargs):
        super().__init__()
        assert edge_dim is None, "reference never passes edge_dim (models.py:336-364)"
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.add_self_loops, self.fill_value = add_self_loops, fill_value
        self.lin = _GlorotLinear(in_channels, heads * out_channels)
        self.att_src = Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = Parameter(torch.empty(1, heads, out_channels))
        glorot_(self.att_src)
        glorot_(self.att_dst)
        if bias:
            self.bias = Parameter(torch.zeros(heads * out_channels if concat else out_channels))
        else:
            self.register_parameter("bias", None)

    def _forward_one(self, x: Tensor, edge_index: Tensor):
        H, C = self.heads, self.out_channels
        n = x.size(0)
        z = self.lin(x).view(-1, H, C)
        a_src = (z * self.att_src).sum(dim=-1)
        a_dst = (z * self.att_dst).sum(dim=-1)
        if self.add_self_loops:
            edge_index, _ = remove_self_loops(edge_index)
            edge_index, _ = add_self_loops(edge_index, num_nodes=n)
        src, dst = edge_index[0], edge_index[1]
        alpha = a_src.index_select(0, src) + a_dst.index_select(0, dst)
        alpha = F.leaky_relu(alpha, self.negative_slope)
        alpha = softmax(alpha, dst, num_nodes=n)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        msg = alpha.unsqueeze(-1) * z.index_select(0, src)
        out = scatter(msg, dst, dim=0, dim_size=n, reduce="sum")
        out = out.view(-1, H * C) if self.concat else out.mean(dim=1)
        if self.bias is not None:
            out = out + self.bias
        return out, edge_index, alpha

    def forward(self, x: Tensor, edge_index: Tensor, edge_attr=None, size=None,
                return_attention_weights=None):
        if x.dim() == 3:  # extension, see module docstring
            outs = [self._forward_one(xb, edge_index) for xb in x]
            out = torch.stack([o[0] for o in outs])
            ei, alpha = outs[0][1], torch.stack([o[2] for o in outs])
        else:
            out, ei, alpha = self._forward_one(x, edge_index)
        if isinstance(return_attention_weights, bool):
            return out, (ei, alpha)
        return out


class SimpleConv(MessagePassing):
    def __init__(self, aggr: str = "sum", combine_root: Optional[str] = None, **kwargs):
        super().__init__()
        assert combine_root is None
        self.aggr = aggr

    def forward(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None,
                size=None) -> Tensor:
        x_j = x.index_select(-2, edge_index[0])
        if edge_weight is not None:
            x_j = edge_weight.view(-1, 1) * x_j
        return scatter(x_j, edge_index[1], dim=-2, dim_size=x.size(-2), reduce=self.aggr)


class LayerNorm(torch.nn.Module):
    def __init__(self, in_channels: int, eps: float = 1e-5, affine: bool = True,
                 mode: str = "graph"):
        super().__init__()
        self.in_channels, self.eps, self.affine, self.mode = in_channels, eps, affine, mode
        if affine:
            self.weight = Parameter(torch.ones(in_channels))
            self.bias = Parameter(torch.zeros(in_channels))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def forward(self, x: Tensor, batch: Optional[Tensor] = None, batch_size=None) -> Tensor:
        if self.mode == "graph":
            assert batch is None, "reference never passes batch"
            x = x - x.mean()
            out = x / (x.std(unbiased=False) + self.eps)
            if self.weight is not None and self.bias is not None:
                out = out * self.weight + self.bias
            return out
        if self.mode == "node":
            return F.layer_norm(x, (self.in_channels,), self.weight, self.bias, self.eps)
        raise ValueError(f"Unknown normalization mode: {self.mode}")


def summary(model: torch.nn.Module, *args, max_depth: int = 3, leaf_module="MessagePassing",
            **kwargs) -> str:
    """Runs one eval/no_grad forward with hooks and tabulates module, shapes and #params."""
    rows, hooks = [], []

    def shape_of(o):
        if isinstance(o, Tensor):
            return str(list(o.shape))
        if isinstance(o, (tuple, list)):
            return ", ".join(s for s in (shape_of(v) for v in o) if s)
        return ""

    def register(name, mod, depth):
        def hook(m, inp, out, name=name, depth=depth):
            n_par = sum(p.numel() for p in m.parameters())
            rows.append(("  " * depth + name, shape_of(inp), shape_of(out), n_par))
        hooks.append(mod.register_forward_hook(hook))
        if depth < max_depth and not isinstance(mod, MessagePassing):
            for cname, child in mod.named_children():
                register(f"({cname}){type(child).__name__}", child, depth + 1)

    register(type(model).__name__, model, 0)
    training = model.training
    model.eval()
    try:
        with torch.no_grad():
            model(*args, **kwargs)
    finally:
        model.train(training)
        for h in hooks:
            h.remove()
    lines = ["| Layer | Input Shape | Output Shape | #Param |", "|---|---|---|---|"]
    for name, i, o, n in rows:
        lines.append(f"| {name} | {i} | {o} | {n:,} |")
    return "\n".join(lines)
