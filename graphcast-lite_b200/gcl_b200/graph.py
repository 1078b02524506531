"""Device-resident CSR of a PyG ``edge_index`` (built once by gcl_csr_build, cached).

PyG recomputes self-loop handling and gcn_norm inside every GCNConv/GATConv.forward (reference call
sites /root/reference/src/models.py:414,419,425,431).  Here that work happens once per
(edge_index, num_nodes, flavour) and is cached; SparseGATConv's pruned edge lists
(models.py:148, 846) are new tensors and therefore get their own entry.
"""
import ctypes
import os
import weakref
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch

from . import _cabi

CSR_RAW, CSR_LOOPS = 0, 1
NORM_NONE, NORM_GCN, NORM_MEAN = 0, 1, 2


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"gcl_b200: {what} must be a CUDA tensor (got {t.device}); there is no CPU fallback")


# Row orders with spatial locality, keyed by node count: tiles of a plan take rows in this order, which keeps the
# union of their source rows small.  The model registers the orders it knows from the mesh geometry (graphs_build);
# graphs without a hint are tiled in natural row order.  Purely a scheduling hint -- results do not depend on it.
ORDER_HINTS = {}

TILE_ROWS = int(os.environ.get("GCL_TILE_ROWS", "64"))
TILE_UNION = int(os.environ.get("GCL_TILE_UNION", "128"))
TILE_ENTRIES = int(os.environ.get("GCL_TILE_ENTRIES", "1024"))
TILED = os.environ.get("GCL_TILED", "1") != "0"          # A/B switch: 0 = first-generation row-gather kernels


class TilePlan:
    """Tile plan of one CSR orientation (gcl_tile_plan_host): device index arrays + the C struct that names them.
    pad = 2: rows padded to entry pairs, as the pipelined SpMM kernel wants; pad = 1 for the attention kernels."""

    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, nnz: int, n_rows: int, n_rows_out: int,
                 n_rows_in: int, order: Optional[np.ndarray] = None, max_rows: int = None, max_union: int = None,
                 max_entries: int = None, pad: int = 2):
        lib = _cabi.load()
        dev = rowptr.device
        rp = np.ascontiguousarray(rowptr.cpu().numpy(), dtype=np.int32)
        co = np.ascontiguousarray(col[: max(nnz, 1)].cpu().numpy(), dtype=np.int32)
        od = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
        if od is not None and od.shape != (n_rows,):
            raise ValueError(f"gcl_b200: row order hint has shape {od.shape}, graph has {n_rows} rows")
        n1, ne = n_rows + 1, max(nnz, 1) + (pad - 1) * n_rows
        out = dict(tile_rowptr=np.zeros(n1, np.int32), tile_uptr=np.zeros(n1, np.int32), rows=np.zeros(n1, np.int32),
                   eptr=np.zeros(n1, np.int32), lidx=np.zeros(ne, np.uint16), ek=np.zeros(ne, np.int32),
                   usrc=np.zeros(max(nnz, 1), np.int32), heavy_rows=np.zeros(n1, np.int32),
                   tile_desc=np.zeros(8 * n1, np.int32))
        counts = np.zeros(8, np.int64)
        ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        rc = lib.gcl_tile_plan_host(ptr(rp), ptr(co), None if od is None else ptr(od), n_rows, n_rows_out, n_rows_in,
                                    max_rows or TILE_ROWS, max_union or TILE_UNION, max_entries or TILE_ENTRIES, pad,
                                    ptr(out["tile_rowptr"]), ptr(out["tile_uptr"]), ptr(out["rows"]), ptr(out["eptr"]),
                                    ptr(out["lidx"]), ptr(out["ek"]), ptr(out["usrc"]), ptr(out["heavy_rows"]),
                                    ptr(out["tile_desc"]), ptr(counts))
        _cabi.check(rc, "gcl_tile_plan_host")
        T, nplan, nu, nent, nheavy, mr, mu, me = (int(v) for v in counts)
        self.n_tiles, self.n_plan_rows, self.n_union, self.n_entries, self.n_heavy = T, nplan, nu, nent, nheavy
        self.max_rows, self.max_union, self.max_entries, self.pad = mr, mu, me, pad
        used = dict(tile_rowptr=T + 1, tile_uptr=T + 1, rows=max(nplan, 1), eptr=nplan + 1, lidx=max(nent, 1),
                    ek=max(nent, 1), usrc=max(nu, 1), heavy_rows=max(nheavy, 1), tile_desc=8 * max(T, 1))
        self.t = {}
        for k, n in used.items():
            a = out[k][:n]
            if a.dtype == np.uint16:                  # torch has no uint16 arithmetic; the kernels only need the bytes
                a = a.view(np.int16)
            self.t[k] = torch.from_numpy(a.copy()).to(dev)
        self.struct = _cabi.TilePlanStruct(*(self.t[k].data_ptr() for k in
                                             ("tile_rowptr", "tile_uptr", "rows", "eptr", "lidx", "ek", "usrc",
                                              "heavy_rows", "tile_desc")), T, nheavy, mr, mu, me, pad)
        self.ref = ctypes.byref(self.struct)
        self._ent = {}

    @property
    def union_per_row(self) -> float:
        return self.n_union / max(self.n_plan_rows, 1)

    def entries(self, w: Optional[torch.Tensor], key=None) -> torch.Tensor:
        """int32 [n_entries, 2] = {lidx, weight bits} in plan order for per-CSR-entry weights w (None = 1); pad and
        masked entries carry lidx = max_union (the kernel's all-zero row), pads weight 0.  Cached per `key` (the weight kind); the gather w[ek] runs once."""
        hit = self._ent.get(key) if key is not None else None
        if hit is not None:
            return hit
        ek = self.t["ek"].long()
        real = ek >= 0
        if w is None:
            wv = real.to(torch.float32)
        else:
            wv = torch.where(real, w.to(torch.float32)[ek.clamp_min(0)], torch.zeros((), device=ek.device))
        li = self.t["lidx"].to(torch.int32) & 0xFFFF
        li = torch.where(li == 0xFFFF, torch.full_like(li, self.max_union), li)   # pads / masked -> the zero row
        ent = torch.stack([li, wv.view(torch.int32)], dim=1).contiguous()
        if key is not None:
            self._ent[key] = ent
        return ent


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
        self._plans = {}
        self._gat_ws = None

    def gat_ws(self):
        """What the warp-specialised single-head GAT kernels need besides the two plans (pad = 2): the packed entries
        of both plans and f2t, the sender-grouped plan entry of every receiver-grouped plan entry (-1 for pads).
        None if a plan has heavy rows."""
        if self._gat_ws is None:
            pf, pt = self.plan(False, pad=2), self.plan(True, pad=2)
            if pf.n_heavy or pt.n_heavy or pf.n_tiles == 0:
                self._gat_ws = False
            else:
                dev, nnz = self.rowptr.device, self.nnz
                ek_f, ek_t = pf.t["ek"].long(), pt.t["ek"].long()
                r2t = torch.empty(nnz, dtype=torch.long, device=dev)          # CSR position -> sender-grouped position
                r2t[self.t2r[:nnz].long()] = torch.arange(nnz, device=dev)
                real_t = ek_t >= 0
                inv_t = torch.empty(nnz, dtype=torch.long, device=dev)        # sender-grouped position -> plan_t entry
                inv_t[ek_t[real_t]] = torch.nonzero(real_t).squeeze(1)
                f2t = torch.where(ek_f >= 0, inv_t[r2t[ek_f.clamp_min(0)]], torch.full_like(ek_f, -1)).to(torch.int32)
                pcol = torch.where(ek_f >= 0, self.col.long()[ek_f.clamp_min(0)], torch.full_like(ek_f, -1))
                self._gat_ws = (pf, pt, pf.entries(None, "unit"), pt.entries(None, "unit"), f2t.contiguous(),
                                pcol.to(torch.int32).contiguous())
        return self._gat_ws or None

    def plan(self, transposed: bool = False, n_rows_out: Optional[int] = None,
             n_rows_in: Optional[int] = None, pad: int = 2) -> TilePlan:
        """Tile plan of the receiver-grouped (forward) or sender-grouped (backward) CSR, built once and cached.
        n_rows_out: only rows below it are produced; n_rows_in: columns at or past it are masked (zero rows);
        pad: 2 for the SpMM kernel (entry pairs), 1 for the attention kernels."""
        n = self.num_nodes
        key = (bool(transposed), n if n_rows_out is None else int(n_rows_out),
               n if n_rows_in is None else int(n_rows_in), int(pad))
        pl = self._plans.get(key)
        if pl is None:
            rp, co = (self.rowptr_t, self.col_t) if transposed else (self.rowptr, self.col)
            pl = TilePlan(rp, co, self.nnz, n, key[1], key[2], ORDER_HINTS.get(n), pad=pad)
            self._plans[key] = pl
        return pl

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


class EdgeOps:
    """Gathers node rows onto edges and reduces edge rows onto nodes for a fixed edge list -- the index side of the
    InteractionNet processor (models.py:213-219: x[senders], x[receivers], scatter(..., receivers, reduce='mean')).
    Each direction is a CSR (one entry per edge row for the gathers, edges grouped by node for the reductions) with its
    tile plan, so the work runs on the same deterministic SpMM kernels as the convolutions: a gather's backward is
    the reduction over the same index and vice versa, no atomics."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int):
        _require_cuda(edge_index, "edge_index")
        dev = edge_index.device
        E, N = int(edge_index.shape[1]), int(num_nodes)
        self.num_edges, self.num_nodes = E, N
        i32 = torch.int32
        ar = torch.arange(E + 1, dtype=i32, device=dev)
        ones = None

        def grouped(idx):                       # edges grouped by node idx[e], ascending edge id inside a node
            order = torch.argsort(idx, stable=True)
            cnt = torch.bincount(idx, minlength=N)
            rp = torch.zeros(N + 1, dtype=torch.int64, device=dev)
            rp[1:] = torch.cumsum(cnt, 0)
            return rp.to(i32).contiguous(), order.to(i32).contiguous(), cnt

        src, dst = edge_index[0].contiguous(), edge_index[1].contiguous()
        rp_s, col_s, _ = grouped(src)
        rp_d, col_d, cnt_d = grouped(dst)
        inv = 1.0 / cnt_d.clamp(min=1).to(torch.float32)
        w_mean_csr = inv[dst[col_d.long()]].contiguous()          # weight per (node-grouped) entry
        w_mean_edge = inv[dst].contiguous()                        # the same weights, one per edge row

        def pack(rowptr, col, w, n_rows, n_cols):
            plan = TilePlan(rowptr, col, int(col.numel()), n_rows, n_rows, n_cols, None, pad=2) if col.numel() else None
            return (rowptr, col, w, n_rows, plan)
        # gathers: edge row e <- node row idx[e];  reductions: node row i <- sum / mean of its edge rows
        self.gather_src = (pack(ar, src.to(i32).contiguous(), ones, E, N), pack(rp_s, col_s, ones, N, E))
        self.gather_dst = (pack(ar, dst.to(i32).contiguous(), ones, E, N), pack(rp_d, col_d, ones, N, E))
        self.mean_dst = (pack(rp_d, col_d, w_mean_csr, N, E), pack(ar, dst.to(i32).contiguous(), w_mean_edge, E, N))
