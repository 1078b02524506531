"""Stand-ins for the torch_geometric.utils symbols graphcast-lite imports
(/root/reference/src/models.py:24 ``dense_to_sparse, softmax``; models.py:220 and
src/dual_mesh.py:23 ``scatter``).  None of them is on the measured hot path of the BASELINE configs
(``softmax`` is imported but never called; ``scatter`` is used by InteractionNet / dual mesh only;
``dense_to_sparse`` once at product-graph construction), so they are thin device-side torch
compositions with PyG's semantics."""
from typing import Optional, Tuple

import torch
from torch import Tensor


def _expand(index: Tensor, src: Tensor, dim: int) -> Tensor:
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src)


def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None, reduce: str = "sum") -> Tensor:
    dim = src.dim() + dim if dim < 0 else dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    size = list(src.shape)
    size[dim] = dim_size
    if reduce in ("sum", "add"):
        return src.new_zeros(size).scatter_add_(dim, _expand(index, src, dim), src)
    if reduce == "mean":
        cnt = src.new_zeros(dim_size).scatter_add_(0, index, src.new_ones(src.size(dim))).clamp_(min=1)
        shape = [1] * src.dim()
        shape[dim] = -1
        return src.new_zeros(size).scatter_add_(dim, _expand(index, src, dim), src) / cnt.view(shape)
    if reduce in ("max", "min"):
        return src.new_zeros(size).scatter_reduce_(dim, _expand(index, src, dim), src,
                                                   reduce="amax" if reduce == "max" else "amin", include_self=False)
    raise ValueError(f"unsupported reduce {reduce!r}")


def softmax(src: Tensor, index: Optional[Tensor] = None, ptr: Optional[Tensor] = None,
            num_nodes: Optional[int] = None, dim: int = 0) -> Tensor:
    if index is None:
        raise NotImplementedError("gcl_b200.utils.softmax: index form only")
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    m = scatter(src.detach(), index, dim, n, "max")
    out = (src - m.index_select(dim, index)).exp()
    return out / (scatter(out, index, dim, n, "sum") + 1e-16).index_select(dim, index)


def dense_to_sparse(adj: Tensor, mask: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    if adj.dim() != 2:
        raise NotImplementedError("gcl_b200.utils.dense_to_sparse: 2-D adjacency only (models.py:772)")
    ei = adj.nonzero().t().contiguous()
    return ei, adj[ei[0], ei[1]]
