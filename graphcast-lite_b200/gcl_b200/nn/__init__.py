"""Drop-in replacements for the ``torch_geometric.nn`` symbols graphcast-lite imports
(/root/reference/src/models.py:21,25):

    from gcl_b200.nn import GCNConv, SimpleConv, GATConv, LayerNorm, summary

Same class names (models.py:418,424,430 test ``type(layer) == GCNConv``), constructor arguments, call
signatures ``conv(x, edge_index[, edge_weight]) -> [N, C]`` and state_dict keys as PyG 2.5.3, so
reference checkpoints load; ``GATConv`` is subclassable the way ``SparseGATConv`` does it
(models.py:112-151).  Arithmetic runs in the hand-written sm_100a kernels behind libgcl_b200.so;
CPU tensors raise (no fallback).  ``x`` may also be [B, N, C]: B samples sharing the graph.
"""
import math
from typing import Optional

import torch
from torch import Tensor
from torch.nn import Parameter

from .. import ops
from ..graph import CSR_LOOPS, CSR_RAW, GLOBAL_CACHE, NORM_GCN, NORM_MEAN, NORM_NONE

__all__ = ["GCNConv", "GATConv", "SimpleConv", "LayerNorm", "Linear", "MessagePassing", "summary"]


def _glorot_(t: Tensor) -> None:
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)


class Linear(torch.nn.Module):
    """Bias-free glorot linear, parameter name ``weight`` (PyG ``nn.dense.linear.Linear``)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = Parameter(torch.empty(out_channels, in_channels))
        self.reset_parameters()

    def reset_parameters(self):
        _glorot_(self.weight)

    def forward(self, x: Tensor) -> Tensor:
        return ops.linear(x, self.weight)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, bias=False"


class MessagePassing(torch.nn.Module):
    """Marker base class (summary() treats subclasses as leaves, like PyG)."""


def _num_nodes(x: Tensor) -> int:
    return x.size(-2)


class GCNConv(MessagePassing):
    """out = D^-1/2 (A + I) D^-1/2 (x W^T) + b   (PyG GCNConv, improved=False, cached=False).

    Replaces the layers built at /root/reference/src/models.py:323-330 and called at :419."""

    def __init__(self, in_channels: int, out_channels: int, improved: bool = False, cached: bool = False,
                 add_self_loops: Optional[bool] = None, normalize: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        if improved:
            raise NotImplementedError("gcl_b200.GCNConv: improved=True is not used by graphcast-lite")
        if add_self_loops is None:
            add_self_loops = normalize
        if add_self_loops and not normalize:
            raise ValueError("GCNConv does not support adding self-loops without normalization (as in PyG)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached = improved, cached
        self.add_self_loops, self.normalize = add_self_loops, normalize
        self.lin = Linear(in_channels, out_channels)
        if bias:
            self.bias = Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def reset_parameters(self):
        self.lin.reset_parameters()
        if self.bias is not None:
            torch.nn.init.zeros_(self.bias)

    def forward(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None) -> Tensor:
        mode = CSR_LOOPS if self.add_self_loops else CSR_RAW
        g = GLOBAL_CACHE.get(edge_index, _num_nodes(x), mode, edge_weight)
        if x.dtype == torch.bfloat16:
            # optional bf16 feature rows (tolerance rel 2e-2): the dense transform runs in fp32 (3xTF32), the
            # aggregation -- the HBM-bound part -- reads and writes bf16 rows
            h = ops.linear(x.float(), self.lin.weight).to(torch.bfloat16)
            return ops.aggregate(h, g, NORM_GCN if self.normalize else NORM_NONE, self.bias)
        h = ops.linear(x, self.lin.weight)
        return ops.aggregate(h, g, NORM_GCN if self.normalize else NORM_NONE, self.bias)

    def __repr__(self):
        return f"{type(self).__name__}({self.in_channels}, {self.out_channels})"


class GATConv(MessagePassing):
    """PyG GATConv (v1 attention) with int in_channels, no edge features, dropout 0.

    Replaces models.py:336-357 (called at :425) and is the base class of the reference's
    SparseGATConv (models.py:112-151), which calls
    ``super().forward(x, edge_index, return_attention_weights=True)``."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2, dropout: float = 0.0, add_self_loops: bool = True,
                 edge_dim: Optional[int] = None, fill_value="mean", bias: bool = True, **kwargs):
        super().__init__()
        if edge_dim is not None:
            raise NotImplementedError("gcl_b200.GATConv: edge_dim is not used by graphcast-lite")
        if dropout != 0.0:
            raise NotImplementedError("gcl_b200.GATConv: attention dropout is not used by graphcast-lite")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.add_self_loops, self.edge_dim, self.fill_value = add_self_loops, edge_dim, fill_value
        self.lin = Linear(in_channels, heads * out_channels)
        self.att_src = Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = Parameter(torch.empty(1, heads, out_channels))
        if bias:
            self.bias = Parameter(torch.zeros(heads * out_channels if concat else out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        self.lin.reset_parameters()
        _glorot_(self.att_src)
        _glorot_(self.att_dst)
        if self.bias is not None:
            torch.nn.init.zeros_(self.bias)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # PyG <= 2.4 checkpoints name the shared projection lin_src / lin_dst.
        for old in ("lin_src.weight", "lin_dst.weight"):
            if prefix + old in state_dict:
                w = state_dict.pop(prefix + old)
                state_dict.setdefault(prefix + "lin.weight", w)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def forward(self, x: Tensor, edge_index: Tensor, edge_attr=None, size=None, return_attention_weights=None):
        if edge_attr is not None or size is not None:
            raise NotImplementedError("gcl_b200.GATConv: edge_attr / size are not used by graphcast-lite")
        mode = CSR_LOOPS if self.add_self_loops else CSR_RAW
        g = GLOBAL_CACHE.get(edge_index, _num_nodes(x), mode)
        want = isinstance(return_attention_weights, bool)
        scores = None
        if self.heads == 1:              # logits' node terms from the epilogue of the `lin` GEMM
            z, a_s, a_d = ops.linear_scores(x, self.lin.weight, self.att_src, self.att_dst)
            scores = (a_s, a_d)
        else:
            z = ops.linear(x, self.lin.weight)
        out, alpha = ops.gat_attend(z, self.att_src, self.att_dst, self.bias, g, self.heads, self.concat,
                                    self.negative_slope, want_alpha=want, scores=scores)
        if want:
            return out, (g.edge_index_with_loops, alpha)
        return out

    def __repr__(self):
        return f"{type(self).__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})"


class SimpleConv(MessagePassing):
    """Parameter-free neighbourhood aggregation (models.py:309, called at :414).  aggr: mean | sum."""

    def __init__(self, aggr: str = "sum", combine_root: Optional[str] = None, **kwargs):
        super().__init__()
        if combine_root is not None:
            raise NotImplementedError("gcl_b200.SimpleConv: combine_root is not used by graphcast-lite")
        if aggr not in ("mean", "sum", "add"):
            raise NotImplementedError(f"gcl_b200.SimpleConv: aggr={aggr!r} unsupported (mean | sum)")
        self.aggr = aggr

    def forward(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None, size=None) -> Tensor:
        g = GLOBAL_CACHE.get(edge_index, _num_nodes(x), CSR_RAW, edge_weight)
        return ops.aggregate(x, g, NORM_MEAN if self.aggr == "mean" else NORM_NONE)

    def __repr__(self):
        return f"{type(self).__name__}(aggr={self.aggr})"


class LayerNorm(torch.nn.Module):
    """torch_geometric.nn.LayerNorm (models.py:103,370; InteractionNet's edge_norm :201).  mode='node' -- the mode of
    every BASELINE config -- normalises each row; mode='graph' normalises by the statistics of the whole sample
    ([N, C], or every [N, C] slice of a [B, N, C] batch).  Both are kernels (gcl_layernorm_*), CUDA only."""

    def __init__(self, in_channels: int, eps: float = 1e-5, affine: bool = True, mode: str = "graph"):
        super().__init__()
        if mode not in ("graph", "node"):
            raise ValueError(f"Unknown normalization mode: {mode}")
        self.in_channels, self.eps, self.affine, self.mode = in_channels, eps, affine, mode
        if affine:
            self.weight = Parameter(torch.ones(in_channels))
            self.bias = Parameter(torch.zeros(in_channels))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def reset_parameters(self):
        if self.affine:
            torch.nn.init.ones_(self.weight)
            torch.nn.init.zeros_(self.bias)

    def forward(self, x: Tensor, batch: Optional[Tensor] = None, batch_size=None) -> Tensor:
        if self.mode == "node":
            return ops.layer_norm(x, self.weight, self.bias, self.eps)
        if batch is not None:
            raise NotImplementedError("gcl_b200.LayerNorm(mode='graph'): `batch` vectors are not used by graphcast-lite")
        return ops.layer_norm_graph(x, self.weight, self.bias, self.eps)

    def __repr__(self):
        return f"{type(self).__name__}({self.in_channels}, affine={self.affine}, mode={self.mode})"


def summary(model: torch.nn.Module, *args, max_depth: int = 3, leaf_module="MessagePassing", **kwargs) -> str:
    """torch_geometric.nn.summary stand-in (models.py:607-655): one eval/no_grad forward with hooks,
    returns a table of module, input/output shapes and parameter counts."""
    rows, hooks = [], []

    def shp(o):
        if isinstance(o, Tensor):
            return str(list(o.shape))
        if isinstance(o, (tuple, list)):
            return ", ".join(s for s in map(shp, o) if s)
        return ""

    def visit(name, mod, depth):
        def hook(m, inp, out, name=name, depth=depth):
            rows.append(("  " * depth + name, shp(inp), shp(out), sum(p.numel() for p in m.parameters())))
        hooks.append(mod.register_forward_hook(hook))
        if depth < max_depth and not isinstance(mod, MessagePassing):
            for cname, child in mod.named_children():
                visit(f"({cname}){type(child).__name__}", child, depth + 1)

    visit(type(model).__name__, model, 0)
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            model(*args, **kwargs)
    finally:
        model.train(was_training)
        for h in hooks:
            h.remove()
    w = [max(len(str(r[i])) for r in rows + [("Layer", "Input Shape", "Output Shape", "#Param")]) for i in range(4)]
    line = "+" + "+".join("-" * (k + 2) for k in w) + "+"
    fmt = "| " + " | ".join("{:<%d}" % k for k in w) + " |"
    out = [line, fmt.format("Layer", "Input Shape", "Output Shape", "#Param"), line]
    out += [fmt.format(r[0], r[1], r[2], f"{r[3]:,}") for r in rows]
    out.append(line)
    return "\n".join(out)
