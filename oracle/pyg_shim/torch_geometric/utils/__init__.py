"""ORACLE / TEST INFRASTRUCTURE ONLY -- not product code, never imported by gcl_b200.

CPU restatement (plain PyTorch ops) of the ``torch_geometric.utils`` symbols the
reference imports:

  * ``dense_to_sparse``, ``softmax``  -- /root/reference/src/models.py:24
  * ``scatter``                        -- /root/reference/src/models.py:220, src/dual_mesh.py:23
  * ``remove_self_loops`` / ``add_self_loops`` / ``add_remaining_self_loops`` -- used inside
    the restated GCNConv / GATConv (torch_geometric/nn/conv/{gcn_conv,gat_conv}.py upstream).

torch_geometric==2.5.3 (requirements.txt:6) is NOT vendored in /root/reference and is not
installable here (no network), so this file restates its published algorithm.
PARITY UNPINNED: the reference ships no golden vectors for this arithmetic; the restatement is
cross-checked against dense-matrix formulations in tests/test_oracle.py.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor


def maybe_num_nodes(edge_index: Tensor, num_nodes: Optional[int] = None) -> int:
    if num_nodes is not None:
        return int(num_nodes)
    return int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0


def _expand_index(index: Tensor, src: Tensor, dim: int) -> Tensor:
    dim = src.dim() + dim if dim < 0 else dim
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src)


def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None,
            reduce: str = "sum") -> Tensor:
    """torch_geometric.utils.scatter: reduce rows of ``src`` that share an ``index``."""
    dim = src.dim() + dim if dim < 0 else dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    size = list(src.shape)
    size[dim] = dim_size
    if reduce in ("sum", "add"):
        return src.new_zeros(size).scatter_add_(dim, _expand_index(index, src, dim), src)
    if reduce == "mean":
        count = src.new_zeros(dim_size)
        count.scatter_add_(0, index, src.new_ones(src.size(dim)))
        count = count.clamp(min=1)
        out = src.new_zeros(size).scatter_add_(dim, _expand_index(index, src, dim), src)
        shape = [1] * src.dim()
        shape[dim] = -1
        return out / count.view(shape)
    if reduce in ("max", "min", "amax", "amin"):
        red = "amax" if reduce in ("max", "amax") else "amin"
        return src.new_zeros(size).scatter_reduce_(dim, _expand_index(index, src, dim), src,
                                                   reduce=red, include_self=False)
    raise ValueError(f"unsupported reduce {reduce!r}")


def softmax(src: Tensor, index: Optional[Tensor] = None, ptr: Optional[Tensor] = None,
            num_nodes: Optional[int] = None, dim: int = 0) -> Tensor:
    """Segment softmax: exp(src - max_seg) / (sum_seg + 1e-16), max detached."""
    assert index is not None, "oracle restates the index form only"
    n = maybe_num_nodes(index, num_nodes)
    src_max = scatter(src.detach(), index, dim, dim_size=n, reduce="max")
    out = src - src_max.index_select(dim, index)
    out = out.exp()
    out_sum = scatter(out, index, dim, dim_size=n, reduce="sum") + 1e-16
    return out / out_sum.index_select(dim, index)


def remove_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None
                      ) -> Tuple[Tensor, Optional[Tensor]]:
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    return edge_index, (None if edge_attr is None else edge_attr[mask])


def add_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None, fill_value=None,
                   num_nodes: Optional[int] = None) -> Tuple[Tensor, Optional[Tensor]]:
    n = maybe_num_nodes(edge_index, num_nodes)
    loop = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    loop = loop.unsqueeze(0).repeat(2, 1)
    if edge_attr is not None:
        fill = 1.0 if fill_value is None or isinstance(fill_value, str) else fill_value
        loop_attr = edge_attr.new_full((n,) + tuple(edge_attr.shape[1:]), fill)
        edge_attr = torch.cat([edge_attr, loop_attr], dim=0)
    return torch.cat([edge_index, loop], dim=1), edge_attr


def add_remaining_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None,
                             fill_value=1.0, num_nodes: Optional[int] = None
                             ) -> Tuple[Tensor, Optional[Tensor]]:
    """Drop existing self loops, append one loop per node (existing loop weights are kept)."""
    n = maybe_num_nodes(edge_index, num_nodes)
    mask = edge_index[0] != edge_index[1]
    loop = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    loop = loop.unsqueeze(0).repeat(2, 1)
    if edge_attr is not None:
        loop_attr = edge_attr.new_full((n,) + tuple(edge_attr.shape[1:]), fill_value)
        inv = ~mask
        loop_attr[edge_index[0][inv]] = edge_attr[inv]
        edge_attr = torch.cat([edge_attr[mask], loop_attr], dim=0)
    return torch.cat([edge_index[:, mask], loop], dim=1), edge_attr


def dense_to_sparse(adj: Tensor, mask: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    assert adj.dim() == 2, "oracle restates the 2-D form only (reference: models.py:772)"
    edge_index = adj.nonzero().t().contiguous()
    return edge_index, adj[edge_index[0], edge_index[1]]
