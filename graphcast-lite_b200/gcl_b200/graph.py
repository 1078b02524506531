"""Device-resident CSR of a PyG ``edge_index`` (built once by gcl_csr_build, cached).

PyG recomputes self-loop handling and gcn_norm inside every GCNConv/GATConv.forward (reference call
sites /root/reference/src/models.py:414,419,425,431).  Here that work happens once per
(edge_index, num_nodes, flavour) and is cached; SparseGATConv's pruned edge lists
(models.py:148, 846) are new tensors and therefore get their own entry.
"""
import weakref
from collections import OrderedDict
from typing import Optional

import torch

from . import _cabi

CSR_RAW, CSR_LOOPS = 0, 1
NORM_NONE, NORM_GCN, NORM_MEAN = 0, 1, 2


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"gcl_b200: {what} must be a CUDA tensor (got {t.device}); there is no CPU fallback")


class CSRGraph:
    """Receiver-grouped and sender-grouped CSR of one edge list, plus per-kind aggregation weights."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int, mode: int,
                 edge_weight: Optional[torch.Tensor] = None):
        _require_cuda(edge_index, "edge_index")
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise ValueError("edge_index must be an int64 tensor of shape [2, E]")
        lib = _cabi.load()
        ei = edge_index.contiguous()
        dev = ei.device
        E, N = int(ei.size(1)), int(num_nodes)
        if E > 0:
            lo, hi = int(ei.min()), int(ei.max())
            if lo < 0 or hi >= N:
                raise ValueError(f"edge_index values must lie in [0, {N}); got [{lo}, {hi}]")
        ew = None
        if edge_weight is not None:
            _require_cuda(edge_weight, "edge_weight")
            ew = edge_weight.detach().to(torch.float32).contiguous()
            if ew.numel() != E:
                raise ValueError("edge_weight must have one entry per edge")
        cap = E + N if mode == CSR_LOOPS else E
        self.num_nodes, self.num_input_edges, self.mode, self.cap = N, E, mode, cap
        i32 = dict(dtype=torch.int32, device=dev)
        self.ei_pyg = torch.empty((2, max(cap, 1)), dtype=torch.int64, device=dev)
        self.w_pyg = torch.empty(max(cap, 1), dtype=torch.float32, device=dev) if ew is not None else None
        self.rowptr = torch.empty(N + 1, **i32)
        self.rowptr_t = torch.empty(N + 1, **i32)
        self.col, self.perm = torch.empty(max(cap, 1), **i32), torch.empty(max(cap, 1), **i32)
        self.col_t, self.perm_t = torch.empty(max(cap, 1), **i32), torch.empty(max(cap, 1), **i32)
        self.t2r = torch.empty(max(cap, 1), **i32)
        self.nnz_dev = torch.zeros(1, **i32)
        ws_bytes = lib.gcl_csr_workspace_bytes(E, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.gcl_csr_build(ei.data_ptr(), ew.data_ptr() if ew is not None else None, E, N, mode,
                                   self.ei_pyg.data_ptr(),
                                   self.w_pyg.data_ptr() if self.w_pyg is not None else None,
                                   self.rowptr.data_ptr(), self.col.data_ptr(), self.perm.data_ptr(),
                                   self.rowptr_t.data_ptr(), self.col_t.data_ptr(), self.perm_t.data_ptr(),
                                   self.t2r.data_ptr(), self.nnz_dev.data_ptr(), ws.data_ptr(), ws_bytes,
                                   _stream())
        _cabi.check(rc, "gcl_csr_build")
        self.nnz = int(self.nnz_dev.item())  # one-time sync at graph build
        self.max_in_degree = int((self.rowptr[1:] - self.rowptr[:-1]).max()) if N > 0 else 0
        self._weights = {}
        self._ei_view = None

    @property
    def edge_index_with_loops(self) -> torch.Tensor:
        """PyG-order [2, nnz] int64 edge list (kept edges, then one loop per node).  Always the same
        tensor object, so feeding it back (SparseGAT, models.py:846) hits the graph cache."""
        if self._ei_view is None:
            self._ei_view = self.ei_pyg[:, : self.nnz]
        return self._ei_view

    def weights(self, kind: int):
        """(w_csr, w_csr_t) for NORM_GCN / NORM_MEAN / NORM_NONE; None for unit weights."""
        if kind == NORM_NONE and self.w_pyg is None:
            return None, None
        if kind not in self._weights:
            lib = _cabi.load()
            dev = self.rowptr.device
            n = max(self.cap, 1)
            w = torch.empty(n, dtype=torch.float32, device=dev)
            wt = torch.empty(n, dtype=torch.float32, device=dev)
            dis = torch.empty(self.num_nodes, dtype=torch.float32, device=dev) if kind == NORM_GCN else None
            with torch.cuda.device(dev):
                rc = lib.gcl_csr_weights(self.rowptr.data_ptr(), self.col.data_ptr(), self.perm.data_ptr(),
                                         self.t2r.data_ptr(),
                                         self.w_pyg.data_ptr() if self.w_pyg is not None else None,
                                         self.num_nodes, self.cap, kind,
                                         dis.data_ptr() if dis is not None else None, w.data_ptr(),
                                         wt.data_ptr(), _stream())
            _cabi.check(rc, "gcl_csr_weights")
            self._weights[kind] = (w, wt, dis)
        w, wt, _ = self._weights[kind]
        return w, wt


class GraphCache:
    """LRU of CSRGraph keyed on the identity + version of the edge_index tensor."""

    def __init__(self, capacity: int = 32):
        self.capacity = capacity
        self._d = OrderedDict()

    def get(self, edge_index: torch.Tensor, num_nodes: int, mode: int,
            edge_weight: Optional[torch.Tensor] = None) -> CSRGraph:
        ew_key = None if edge_weight is None else (edge_weight.data_ptr(), edge_weight._version)
        key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(num_nodes), mode,
               ew_key, edge_index.device.index)
        hit = self._d.get(key)
        if hit is not None:
            ref, g = hit
            if ref() is edge_index:
                self._d.move_to_end(key)
                return g
            del self._d[key]  # address reused by another tensor
        g = CSRGraph(edge_index, num_nodes, mode, edge_weight)
        self._d[key] = (weakref.ref(edge_index), g)
        if mode == CSR_LOOPS and edge_weight is None:
            # dropping and re-adding the loops of g's own PyG-order list reproduces it: alias it to g
            v = g.edge_index_with_loops
            vkey = (v.data_ptr(), tuple(v.shape), v._version, int(num_nodes), mode, None, v.device.index)
            self._d[vkey] = (weakref.ref(v), g)
        while len(self._d) > self.capacity:
            self._d.popitem(last=False)
        return g

    def clear(self):
        self._d.clear()


GLOBAL_CACHE = GraphCache()
