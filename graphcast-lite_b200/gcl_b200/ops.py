"""torch.autograd.Function custom ops over the gcl_b200 C ABI (fp32, deterministic, CUDA only).

Every op checks its inputs and raises instead of falling back to PyTorch/CPU.  Tensors are
[B, N, C] (B samples on one shared graph) or the reference's [N, C].
"""
from typing import Optional

import os

import torch

from . import _cabi
from . import graph as _graph
from .graph import CSRGraph, NORM_GCN, NORM_MEAN, NORM_NONE  # noqa: F401


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"gcl_b200: {name} must be a tensor")
    if not t.is_cuda:
        raise RuntimeError(f"gcl_b200: {name} is on {t.device}; the kernels are CUDA (sm_100a) only, no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"gcl_b200: {name} has dtype {t.dtype}; this build computes in float32")
    return t.contiguous()


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# A/B switch: autograd's default zero-fills a full-size gradient for every unused / non-differentiable output
_MATERIALIZE = os.environ.get("GCL_MATERIALIZE_GRADS") == "1"

class KernelProfiler:
    """CUDA-event timing of every C-ABI call on the launching stream (bench.py's roofline numbers).
    `nbytes` is the ALGORITHMIC traffic of the call (DESIGN.md / SURVEY.md 8d), not a counter."""

    def __init__(self):
        self.records = []   # (name, tag, nbytes, start, end)

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, tag, nbytes, s, e in self.records:
            a = agg.setdefault((name, tag), dict(calls=0, ms=0.0, bytes=0))
            a["calls"] += 1
            a["ms"] += s.elapsed_time(e)
            a["bytes"] += nbytes
        return agg


PROFILER: Optional[KernelProfiler] = None


def _call(name: str, *args, nbytes: int = 0, tag: str = "") -> None:
    lib = _cabi.load()
    fn = getattr(lib, name)
    if PROFILER is None:
        _cabi.check(fn(*args), name)
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    rc = fn(*args)
    e.record()
    PROFILER.records.append((name, tag, int(nbytes), s, e))
    _cabi.check(rc, name)


def _as3(x: torch.Tensor):
    """[N, C] -> ([1, N, C], squeeze?)"""
    if x.dim() == 2:
        return x.unsqueeze(0), True
    if x.dim() == 3:
        return x, False
    raise ValueError(f"gcl_b200: node features must be [N, C] or [B, N, C], got {tuple(x.shape)}")


# ------------------------------------------------------------------------------------------- raw calls
def _tileable(C: int, *tensors) -> bool:
    """The tiled kernels take 16-byte aligned rows of 4..128 channels (a multiple of 4)."""
    return (_graph.TILED and C % 4 == 0 and 0 < C <= 128
            and all(t is None or t.data_ptr() % 16 == 0 for t in tensors))


def spmm_raw(rowptr, col, w, x3, n_out, bias=None, slope=None, want_z=False, plan=None, wkey=None):
    """plan: a graph.TilePlan (pad = 2) of (rowptr, col) built for n_rows_out = n_out and n_rows_in = x3.shape[1]: the
    tiled kernel (persistent CTAs, rows staged in a shared-memory ring by asynchronous copies); otherwise the
    row-gather kernel.  wkey: cache key of the plan-ordered copy of w (the weight kind)."""
    B, n_in, C = x3.shape
    out = torch.empty((B, n_out, C), dtype=torch.float32, device=x3.device)
    z = torch.empty_like(out) if want_z else None
    with torch.cuda.device(x3.device):
        nnz = int(col.numel())
        nbytes = 4 * B * C * (n_in + n_out * (2 if want_z else 1)) + nnz * (8 if w is not None else 4) + 4 * (n_out + 1)
        if plan is not None and _tileable(C, x3, out, bias, z):
            ent = plan.entries(w, wkey)
            _call("gcl_spmm_tiled_f32", plan.ref, _p(ent), _p(rowptr), _p(col), _p(w), _p(x3), _p(out), B, n_in, C,
                  n_in * C, n_out * C, _p(bias), _p(slope), _p(z), _stream(), nbytes=nbytes, tag=f"N{n_out}xC{C}xB{B}")
        else:
            _call("gcl_spmm_f32", _p(rowptr), _p(col), _p(w), _p(x3), _p(out), B, n_out, n_in, C, n_in * C, n_out * C,
                  _p(bias), _p(slope), _p(z), nnz, _stream(), nbytes=nbytes, tag=f"N{n_out}xC{C}xB{B}")
    return out, z


def colsum_raw(x2):
    R, C = x2.shape
    out = torch.empty(C, dtype=torch.float32, device=x2.device)
    lib = _cabi.load()
    nb = lib.gcl_colsum_workspace_bytes(R, C)
    ws = _ws(nb, x2.device)
    with torch.cuda.device(x2.device):
        _call("gcl_colsum_f32", _p(x2), _p(out), R, C, _p(ws), nb, _stream(), nbytes=4 * R * C, tag=f"R{R}xC{C}")
    return out


def prelu_bwd_raw(dy, z, slope):
    dx = torch.empty_like(dy)
    dslope = torch.empty(1, dtype=torch.float32, device=dy.device)
    lib = _cabi.load()
    n = dy.numel()
    nb = lib.gcl_prelu_bwd_workspace_bytes(n)
    ws = _ws(nb, dy.device)
    with torch.cuda.device(dy.device):
        _call("gcl_prelu_bwd_f32", _p(dy), _p(z), _p(slope), _p(dx), _p(dslope), n, _p(ws), nb, _stream(),
              nbytes=12 * n, tag=f"n{n}")
    return dx, dslope


def prelu_bwd_colsum_raw(dy3, z, slope):
    """(dx, dslope[1], dbias[C]) in one pass over dy / z."""
    C = dy3.shape[-1]
    R = dy3.numel() // C
    dx = torch.empty_like(dy3)
    dslope = torch.empty(1, dtype=torch.float32, device=dy3.device)
    dbias = torch.empty(C, dtype=torch.float32, device=dy3.device)
    lib = _cabi.load()
    nb = lib.gcl_prelu_bwd_colsum_workspace_bytes(R, C)
    ws = _ws(nb, dy3.device)
    with torch.cuda.device(dy3.device):
        _call("gcl_prelu_bwd_colsum_f32", _p(dy3), _p(z), _p(slope), _p(dx), _p(dslope), _p(dbias), R, C, _p(ws), nb,
              _stream(), nbytes=12 * R * C, tag=f"R{R}xC{C}")
    return dx, dslope, dbias


def linear_fwd_raw(x2, W, bias=None, slope=None, want_z=False):
    R, cin = x2.shape
    cout = W.shape[0]
    y = torch.empty((R, cout), dtype=torch.float32, device=x2.device)
    z = torch.empty_like(y) if want_z else None
    wt = torch.empty(cin * cout, dtype=torch.float32, device=x2.device)
    with torch.cuda.device(x2.device):
        _call("gcl_linear_fwd_f32", _p(x2), _p(W), _p(bias), _p(y), R, cin, cout, _p(slope), _p(z), _p(wt), _stream(),
              nbytes=4 * R * (cin + cout * (2 if want_z else 1)) + 4 * cin * cout, tag=f"R{R}x{cin}->{cout}")
    return y, z


def linear_bwd_dx_raw(dy2, W):
    R, cout = dy2.shape
    cin = W.shape[1]
    dx = torch.empty((R, cin), dtype=torch.float32, device=dy2.device)
    wt = torch.empty(cin * cout, dtype=torch.float32, device=dy2.device)
    with torch.cuda.device(dy2.device):
        _call("gcl_linear_bwd_dx_f32", _p(dy2), _p(W), _p(dx), R, cin, cout, _p(wt), _stream(),
              nbytes=4 * R * (cin + cout) + 4 * cin * cout, tag=f"R{R}x{cout}->{cin}")
    return dx


def linear_bwd_dx_prelu_raw(dy2, W, z_in2, slope, want_colsum=False):
    """dz_in = (dy W) * PReLU'(z_in), dslope (and optionally the column sums of dz_in): backward of PReLU -> Linear
    w.r.t. the PReLU input, one kernel."""
    R, cout = dy2.shape
    cin = W.shape[1]
    dz = torch.empty((R, cin), dtype=torch.float32, device=dy2.device)
    dslope = torch.empty(1, dtype=torch.float32, device=dy2.device)
    dcs = torch.empty(cin, dtype=torch.float32, device=dy2.device) if want_colsum else None
    scratch = torch.empty(cin * cout, dtype=torch.float32, device=dy2.device)
    lib = _cabi.load()
    nb = lib.gcl_linear_bwd_dx_prelu_workspace_bytes(R, cin)
    ws = _ws(nb, dy2.device)
    with torch.cuda.device(dy2.device):
        _call("gcl_linear_bwd_dx_prelu_f32", _p(dy2), _p(W), _p(z_in2), _p(slope), _p(dz), _p(dslope), _p(dcs), R, cin,
              cout, _p(scratch), _p(ws), nb, _stream(), nbytes=4 * R * (2 * cin + cout) + 4 * cin * cout,
              tag=f"R{R}x{cout}->{cin}")
    return (dz, dslope, dcs) if want_colsum else (dz, dslope)


def linear_bwd_dw_raw(dy2, x2, want_bias):
    R, cout = dy2.shape
    cin = x2.shape[1]
    dW = torch.empty((cout, cin), dtype=torch.float32, device=dy2.device)
    db = torch.empty(cout, dtype=torch.float32, device=dy2.device) if want_bias else None
    lib = _cabi.load()
    nb = lib.gcl_linear_bwd_dw_workspace_bytes(R, cin, cout)
    ws = _ws(nb, dy2.device)
    with torch.cuda.device(dy2.device):
        _call("gcl_linear_bwd_dw_f32", _p(dy2), _p(x2), _p(dW), _p(db), R, cin, cout, _p(ws), nb, _stream(),
              nbytes=4 * R * (cin + cout) + 4 * cin * cout, tag=f"R{R}x{cout}x{cin}")
    return dW, db


# ------------------------------------------------------------------------------------------- autograd
class _Aggregate(torch.autograd.Function):
    """out = prelu?( A_w x + bias ), A_w = CSR with per-entry weights of `kind`.  Backward uses the
    sender-grouped CSR (A_w^T), so it is a gather too -- no atomics."""

    @staticmethod
    def forward(ctx, x, bias, slope, graph: CSRGraph, kind: int, rows_out=None):
        x3, squeeze = _as3(_chk(x, "x"))
        if x3.shape[1] != graph.num_nodes:
            raise ValueError(f"gcl_b200: x has {x3.shape[1]} nodes, graph has {graph.num_nodes}")
        n_out = graph.num_nodes if rows_out is None else int(rows_out)   # receivers 0 .. n_out-1 only
        if not 0 < n_out <= graph.num_nodes:
            raise ValueError(f"gcl_b200: rows_out={rows_out} outside 1..{graph.num_nodes}")
        bias_c = _chk(bias, "bias") if bias is not None else None
        slope_c = _chk(slope, "slope") if slope is not None else None
        w, _ = graph.weights(kind)
        need_z = slope_c is not None and any(ctx.needs_input_grad[:3])
        plan = graph.plan(False, n_out, graph.num_nodes) if _tileable(x3.shape[-1]) else None
        out, z = spmm_raw(graph.rowptr, graph.col, w, x3, n_out, bias_c, slope_c, need_z, plan, ("fwd", kind))
        ctx.graph, ctx.kind, ctx.squeeze = graph, kind, squeeze
        ctx.has_bias, ctx.has_slope = bias is not None, slope is not None
        ctx.save_for_backward(z, slope_c)
        return out.squeeze(0) if squeeze else out

    @staticmethod
    def backward(ctx, dout):
        z, slope = ctx.saved_tensors
        g = ctx.graph
        d3, _ = _as3(_chk(dout, "grad_out"))
        dslope = dbias = None
        want_bias = ctx.has_bias and ctx.needs_input_grad[1]
        if ctx.has_slope and want_bias:
            d3, dslope, dbias = prelu_bwd_colsum_raw(d3, z, slope)      # one pass instead of two
            dslope = dslope.view_as(slope)
        elif ctx.has_slope:
            d3, dslope = prelu_bwd_raw(d3, z, slope)
            dslope = dslope.view_as(slope)
        elif want_bias:
            dbias = colsum_raw(d3.view(-1, d3.shape[-1]))
        dx = None
        if ctx.needs_input_grad[0]:
            _, wt = g.weights(ctx.kind)
            plan = g.plan(True, g.num_nodes, d3.shape[1]) if _tileable(d3.shape[-1]) else None
            dx, _ = spmm_raw(g.rowptr_t, g.col_t, wt, d3, g.num_nodes, plan=plan, wkey=("bwd", ctx.kind))
            if ctx.squeeze:
                dx = dx.squeeze(0)
        return dx, dbias, (dslope if ctx.has_slope and ctx.needs_input_grad[2] else None), None, None, None


def spmm_bf16_raw(rowptr, col, w, x3, n_out, plan, bias=None, wkey=None):
    """bf16 feature rows through the tiled engine (fp32 weights / bias / accumulation, bf16 result)."""
    B, n_in, C = x3.shape
    out = torch.empty((B, n_out, C), dtype=torch.bfloat16, device=x3.device)
    ent = plan.entries(w, wkey)
    with torch.cuda.device(x3.device):
        nnz = int(col.numel())
        nbytes = 2 * B * C * (n_in + n_out) + nnz * (8 if w is not None else 4) + 4 * (n_out + 1)
        _call("gcl_spmm_tiled_bf16", plan.ref, _p(ent), _p(rowptr), _p(col), _p(w), _p(x3), _p(out), B, n_in, C,
              n_in * C, n_out * C, _p(bias), None, None, _stream(), nbytes=nbytes, tag=f"N{n_out}xC{C}xB{B}xbf16")
    return out


class _AggregateBF16(torch.autograd.Function):
    """_Aggregate for bf16 feature rows (north_star's optional storage format, tolerance rel 2e-2): x, out and the
    gradients are bf16, weights / bias / accumulation fp32.  C must be a multiple of 8, <= 256 (no fallback)."""

    @staticmethod
    def forward(ctx, x, bias, graph: CSRGraph, kind: int, rows_out=None):
        if not x.is_cuda:
            raise RuntimeError(f"gcl_b200: x is on {x.device}; the kernels are CUDA (sm_100a) only, no CPU fallback")
        x3, squeeze = _as3(x.contiguous())
        B, N, C = x3.shape
        if N != graph.num_nodes:
            raise ValueError(f"gcl_b200: x has {N} nodes, graph has {graph.num_nodes}")
        if C % 8 or C > 256 or x3.data_ptr() % 16:
            raise RuntimeError(f"gcl_b200: bf16 rows need 16-byte aligned rows of 8..256 channels (multiple of 8), got {C}")
        n_out = N if rows_out is None else int(rows_out)
        bias_c = _chk(bias, "bias") if bias is not None else None
        w, _ = graph.weights(kind)
        out = spmm_bf16_raw(graph.rowptr, graph.col, w, x3, n_out, graph.plan(False, n_out, N), bias_c, ("fwd", kind))
        ctx.graph, ctx.kind, ctx.squeeze, ctx.has_bias = graph, kind, squeeze, bias is not None
        return out.squeeze(0) if squeeze else out

    @staticmethod
    def backward(ctx, dout):
        g = ctx.graph
        d3, _ = _as3(dout.to(torch.bfloat16).contiguous())
        dbias = dx = None
        if ctx.has_bias and ctx.needs_input_grad[1]:
            dbias = colsum_raw(d3.float().view(-1, d3.shape[-1]))
        if ctx.needs_input_grad[0]:
            _, wt = g.weights(ctx.kind)
            dx = spmm_bf16_raw(g.rowptr_t, g.col_t, wt, d3, g.num_nodes, g.plan(True, g.num_nodes, d3.shape[1]),
                               wkey=("bwd", ctx.kind))
            if ctx.squeeze:
                dx = dx.squeeze(0)
        return dx, dbias, None, None, None


class _AggregatePre(torch.autograd.Function):
    """(z, a) with z = A_w x + bias and a = PReLU(z), for a consumer that differentiates through the
    PRE-activation (ops.act_linear: the PReLU's backward then runs in the epilogue of the next layer's dX GEMM
    instead of a separate pass over dout, z and dz).  z carries the gradient, a is a buffer."""

    @staticmethod
    def forward(ctx, x, bias, slope, graph: CSRGraph, kind: int, sink=None):
        ctx.sink = sink
        ctx.set_materialize_grads(_MATERIALIZE)     # no [B, N, C] zero fill for the gradient of the non-differentiable `a`
        x3, squeeze = _as3(_chk(x, "x"))
        if x3.shape[1] != graph.num_nodes:
            raise ValueError(f"gcl_b200: x has {x3.shape[1]} nodes, graph has {graph.num_nodes}")
        bias_c = _chk(bias, "bias") if bias is not None else None
        slope_c = _chk(slope, "slope")
        w, _ = graph.weights(kind)
        n = graph.num_nodes
        plan = graph.plan(False, n, n) if _tileable(x3.shape[-1]) else None
        out, z = spmm_raw(graph.rowptr, graph.col, w, x3, n, bias_c, slope_c, True, plan, ("fwd", kind))
        ctx.graph, ctx.kind, ctx.squeeze, ctx.has_bias = graph, kind, squeeze, bias is not None
        if squeeze:
            out, z = out.squeeze(0), z.squeeze(0)
        ctx.mark_non_differentiable(out)
        return z, out

    @staticmethod
    def backward(ctx, dz, _da=None):
        if dz is None:
            return None, None, None, None, None, None
        g = ctx.graph
        d3, _ = _as3(_chk(dz, "grad_out"))
        dbias = dx = None
        if ctx.has_bias and ctx.needs_input_grad[1]:
            dbias = ctx.sink.take(d3) if ctx.sink is not None else None     # summed in the consumer's dX epilogue
            if dbias is None:
                dbias = colsum_raw(d3.view(-1, d3.shape[-1]))
        if ctx.needs_input_grad[0]:
            _, wt = g.weights(ctx.kind)
            plan = g.plan(True, g.num_nodes, d3.shape[1]) if _tileable(d3.shape[-1]) else None
            dx, _ = spmm_raw(g.rowptr_t, g.col_t, wt, d3, g.num_nodes, plan=plan, wkey=("bwd", ctx.kind))
            if ctx.squeeze:
                dx = dx.squeeze(0)
        return dx, dbias, None, None, None, None


def aggregate_pre(x, graph: CSRGraph, kind: int, bias, prelu_slope, sink=None):
    """(z, a): see _AggregatePre.  Feed them to act_linear(z, a, prelu_slope, W_next[, sink=sink])."""
    return _AggregatePre.apply(x, bias, prelu_slope, graph, kind, sink)


def aggregate(x, graph: CSRGraph, kind: int, bias=None, prelu_slope=None, rows_out=None):
    """rows_out = n: only receivers 0..n-1 are produced ([.., n, C]); their gradient flows back to all senders.
    bf16 x: the bf16 feature-row path (no fused PReLU)."""
    if isinstance(x, torch.Tensor) and x.dtype == torch.bfloat16:
        if prelu_slope is not None:
            raise NotImplementedError("gcl_b200: the bf16 feature-row path has no fused PReLU epilogue")
        return _AggregateBF16.apply(x, bias, graph, kind, rows_out)
    return _Aggregate.apply(x, bias, prelu_slope, graph, kind, rows_out)


class _Linear(torch.autograd.Function):
    """y = prelu?( x W^T + b ) over the last dim."""

    @staticmethod
    def forward(ctx, x, W, bias, slope):
        xc, Wc = _chk(x, "x"), _chk(W, "weight")
        if xc.shape[-1] != Wc.shape[1]:
            raise ValueError(f"gcl_b200: linear got x[..., {xc.shape[-1]}] and weight {tuple(Wc.shape)}")
        bias_c = _chk(bias, "bias") if bias is not None else None
        slope_c = _chk(slope, "slope") if slope is not None else None
        x2 = xc.view(-1, xc.shape[-1])
        need_z = slope_c is not None and any(ctx.needs_input_grad)
        y, z = linear_fwd_raw(x2, Wc, bias_c, slope_c, need_z)
        ctx.has_bias, ctx.has_slope = bias is not None, slope is not None
        ctx.save_for_backward(x2, Wc, z, slope_c)
        return y.view(*xc.shape[:-1], Wc.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, W, z, slope = ctx.saved_tensors
        d2 = _chk(dy, "grad_out").view(-1, W.shape[0])
        dslope = None
        if ctx.has_slope:
            d2, dslope = prelu_bwd_raw(d2, z, slope)
            dslope = dslope.view_as(slope)
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = linear_bwd_dx_raw(d2, W).view(*dy.shape[:-1], W.shape[1])
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dW, db = linear_bwd_dw_raw(d2, x2, ctx.has_bias)
        return dx, dW, (db if ctx.has_bias else None), (dslope if ctx.has_slope else None)


class _ActLinear(torch.autograd.Function):
    """y = Linear(a_in) where a_in = PReLU(z_in) was produced (together with z_in) by the previous layer's fused
    epilogue.  Autograd sees the chain through the PRE-activations: the input that carries gradient is z_in, and
    the PReLU's backward runs in the epilogue of this layer's dX GEMM.  With slope_out the layer returns
    (z_out [differentiable], a_out = PReLU(z_out) [buffer for the next _ActLinear]); without, just y.
    z_in = None: a_in is an ordinary differentiable input."""

    @staticmethod
    def forward(ctx, z_in, a_in, slope_in, W, bias, slope_out, sink=None):
        ctx.sink = sink
        ctx.set_materialize_grads(_MATERIALIZE)     # autograd would zero-fill a full-size gradient for the buffer output
        ac, Wc = _chk(a_in, "x"), _chk(W, "weight")
        if ac.shape[-1] != Wc.shape[1]:
            raise ValueError(f"gcl_b200: linear got x[..., {ac.shape[-1]}] and weight {tuple(Wc.shape)}")
        bias_c = _chk(bias, "bias") if bias is not None else None
        so = _chk(slope_out, "slope") if slope_out is not None else None
        x2 = ac.view(-1, ac.shape[-1])
        y, z = linear_fwd_raw(x2, Wc, bias_c, so, so is not None)
        ctx.has_bias, ctx.has_act_in, ctx.lead = bias is not None, z_in is not None, ac.shape[:-1]
        zi = _chk(z_in, "z_in").view(-1, ac.shape[-1]) if z_in is not None else None
        si = _chk(slope_in, "slope_in") if z_in is not None else None
        ctx.save_for_backward(x2, Wc, zi, si)
        yv = y.view(*ac.shape[:-1], Wc.shape[0])
        if so is None:
            return yv
        ctx.mark_non_differentiable(yv)
        return z.view_as(yv), yv

    @staticmethod
    def backward(ctx, d, _da=None):
        if d is None:
            return None, None, None, None, None, None, None
        x2, W, zi, si = ctx.saved_tensors
        d2 = _chk(d, "grad_out").view(-1, W.shape[0])
        dz = dx = dsl = None
        if ctx.has_act_in:
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[2]:
                if ctx.sink is not None:       # the producer of z_in wants colsum(dz) (its bias gradient)
                    dz, dsl, dcs = linear_bwd_dx_prelu_raw(d2, W, zi, si, True)
                    ctx.sink.put(dz, dcs)
                else:
                    dz, dsl = linear_bwd_dx_prelu_raw(d2, W, zi, si)
                dz, dsl = dz.view(*ctx.lead, W.shape[1]), dsl.view_as(si)
        elif ctx.needs_input_grad[1]:
            dx = linear_bwd_dx_raw(d2, W).view(*ctx.lead, W.shape[1])
        dW = db = None
        if ctx.needs_input_grad[3] or (ctx.has_bias and ctx.needs_input_grad[4]):
            dW, db = linear_bwd_dw_raw(d2, x2, ctx.has_bias)
        return dz, dx, dsl, dW, (db if ctx.has_bias else None), None, None


class ColsumSink:
    """Hands colsum(dz) from the backward of the consumer of a pre-activation z (act_linear) to the backward of its
    producer (aggregate_pre), where it is the bias gradient.  Valid only for the very tensor it was computed from."""

    def __init__(self):
        self._ptr, self._cs = None, None

    def put(self, dz, colsum):
        self._ptr, self._cs = dz.data_ptr(), colsum

    def take(self, dz):
        cs = self._cs if (self._cs is not None and self._ptr == dz.data_ptr()) else None
        self._ptr, self._cs = None, None
        return cs


def act_linear(z_in, a_in, slope_in, weight, bias=None, slope_out=None, sink=None):
    """See _ActLinear.  Returns y, or (z_out, a_out) when slope_out is given."""
    return _ActLinear.apply(z_in, a_in, slope_in, weight, bias, slope_out, sink)


class _LinearScores(torch.autograd.Function):
    """z = x W^T plus the attention logits' node terms (z * att).sum(-1) of a single-head GATConv, one kernel.
    The scores are returned as non-differentiable buffers: their gradient path (d att, and the part of dz that
    comes through them) is produced by the GAT backward, which owns the scores."""

    @staticmethod
    def forward(ctx, x, W, att_src, att_dst):
        ctx.set_materialize_grads(_MATERIALIZE)
        xc, Wc = _chk(x, "x"), _chk(W, "weight")
        if xc.shape[-1] != Wc.shape[1]:
            raise ValueError(f"gcl_b200: linear got x[..., {xc.shape[-1]}] and weight {tuple(Wc.shape)}")
        a_s, a_d = _chk(att_src, "att_src").view(-1), _chk(att_dst, "att_dst").view(-1)
        cout, cin = Wc.shape
        if a_s.numel() != cout or a_d.numel() != cout:
            raise ValueError("gcl_b200: linear_scores is for heads == 1 (att vectors of out_channels elements)")
        x2 = xc.view(-1, cin)
        R = x2.shape[0]
        dev = x2.device
        z = torch.empty((R, cout), dtype=torch.float32, device=dev)
        asrc = torch.empty(R, dtype=torch.float32, device=dev)
        adst = torch.empty(R, dtype=torch.float32, device=dev)
        wt = torch.empty(cin * cout + 2 * cout, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _call("gcl_linear_fwd_scores_f32", _p(x2), _p(Wc), _p(z), _p(a_s), _p(a_d), _p(asrc), _p(adst), R, cin, cout,
                  _p(wt), _stream(), nbytes=4 * R * (cin + cout + 2) + 4 * cin * cout, tag=f"R{R}x{cin}->{cout}")
        ctx.save_for_backward(x2, Wc)
        lead = xc.shape[:-1]
        asrc, adst = asrc.view(*lead, 1), adst.view(*lead, 1)
        ctx.mark_non_differentiable(asrc, adst)
        return z.view(*lead, cout), asrc, adst

    @staticmethod
    def backward(ctx, dz, _ds=None, _dd=None):
        if dz is None:
            return None, None, None, None
        x2, W = ctx.saved_tensors
        d2 = _chk(dz, "grad_out").view(-1, W.shape[0])
        dx = dW = None
        if ctx.needs_input_grad[0]:
            dx = linear_bwd_dx_raw(d2, W).view(*dz.shape[:-1], W.shape[1])
        if ctx.needs_input_grad[1]:
            dW, _ = linear_bwd_dw_raw(d2, x2, False)
        return dx, dW, None, None


def linear_scores(x, weight, att_src, att_dst):
    """(z, a_src, a_dst) for a single-head GATConv: pass the scores on to gat_attend(scores=...)."""
    return _LinearScores.apply(x, weight, att_src, att_dst)


def linear(x, weight, bias=None, prelu_slope=None):
    return _Linear.apply(x, weight, bias, prelu_slope)


class _PReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, slope):
        xc, sc = _chk(x, "x"), _chk(slope, "slope")
        if sc.numel() != 1:
            raise RuntimeError("gcl_b200: PReLU with one shared slope only (reference: nn.PReLU())")
        y = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _call("gcl_prelu_fwd_f32", _p(xc), _p(sc), _p(y), xc.numel(), _stream(), nbytes=8 * xc.numel())
        ctx.save_for_backward(xc, sc)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, slope = ctx.saved_tensors
        dx, ds = prelu_bwd_raw(_chk(dy, "grad_out"), x, slope)
        return dx, ds.view_as(slope)


def prelu(x, slope):
    return _PReLU.apply(x, slope)


class _LayerNorm(torch.autograd.Function):
    """F.layer_norm over the last dim (torch_geometric LayerNorm(mode='node'))."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        xc = _chk(x, "x")
        C = xc.shape[-1]
        x2 = xc.view(-1, C)
        g = _chk(gamma, "weight") if gamma is not None else None
        b = _chk(beta, "bias") if beta is not None else None
        R = x2.shape[0]
        y = torch.empty_like(x2)
        mean = torch.empty(R, dtype=torch.float32, device=xc.device)
        rstd = torch.empty(R, dtype=torch.float32, device=xc.device)
        with torch.cuda.device(xc.device):
            _call("gcl_layernorm_fwd_f32", _p(x2), _p(g), _p(b), _p(y), _p(mean), _p(rstd), R, C, float(eps),
                  _stream(), nbytes=8 * R * C + 8 * R, tag=f"R{R}xC{C}")
        ctx.affine = gamma is not None
        ctx.save_for_backward(x2, g, mean, rstd)
        return y.view_as(xc)

    @staticmethod
    def backward(ctx, dy):
        x2, g, mean, rstd = ctx.saved_tensors
        R, C = x2.shape
        d2 = _chk(dy, "grad_out").view(R, C)
        dx = torch.empty_like(x2)
        dg = torch.empty(C, dtype=torch.float32, device=x2.device) if ctx.affine else None
        db = torch.empty(C, dtype=torch.float32, device=x2.device) if ctx.affine else None
        lib = _cabi.load()
        nb = lib.gcl_layernorm_bwd_workspace_bytes(R, C)
        ws = _ws(nb, x2.device)
        with torch.cuda.device(x2.device):
            _call("gcl_layernorm_bwd_f32", _p(d2), _p(x2), _p(g), _p(mean), _p(rstd), _p(dx), _p(dg), _p(db), R, C,
                  _p(ws), nb, _stream(), nbytes=12 * R * C + 8 * R, tag=f"R{R}xC{C}")
        return dx.view_as(dy), dg, db, None


def layer_norm(x, weight, bias, eps=1e-5):
    return _LayerNorm.apply(x, weight, bias, eps)


class _GAT(torch.autograd.Function):
    """Attention logits + LeakyReLU + segment softmax + weighted aggregation + head mean/concat + bias.
    z is the already-transformed [B, N, H*C] (or [N, H*C]) feature matrix."""

    @staticmethod
    def forward(ctx, z, att_src, att_dst, bias, graph: CSRGraph, heads, concat, slope, want_alpha, prelu_slope=None,
                scores=None):
        ctx.set_materialize_grads(_MATERIALIZE)      # alpha (SparseGAT's pruning input) is a non-differentiable output
        z3, squeeze = _as3(_chk(z, "z"))
        if prelu_slope is not None and int(heads) != 1:
            raise NotImplementedError("gcl_b200: PReLU fused into GATConv needs heads == 1")
        ps = _chk(prelu_slope, "prelu_slope") if prelu_slope is not None else None
        B, N, HC = z3.shape
        H = int(heads)
        C = HC // H
        a_s, a_d = _chk(att_src, "att_src").view(-1), _chk(att_dst, "att_dst").view(-1)
        bias_c = _chk(bias, "bias") if bias is not None else None
        dev = z3.device
        if scores is not None:           # produced by the epilogue of the `lin` GEMM (ops.linear_scores)
            asrc, adst = (_chk(t, "scores").view(B, N, H) for t in scores)
        else:
            asrc = torch.empty((B, N, H), dtype=torch.float32, device=dev)
            adst = torch.empty((B, N, H), dtype=torch.float32, device=dev)
        nnz = graph.nnz
        cout = HC if concat else C
        out = torch.empty((B, N, cout), dtype=torch.float32, device=dev)
        alpha = None
        alpha_pyg = torch.empty((B, max(nnz, 1), H), dtype=torch.float32, device=dev) if want_alpha else None
        zpre = torch.empty_like(out) if ps is not None else None
        with torch.cuda.device(dev):
            if scores is None:
                _call("gcl_gat_scores_f32", _p(z3), _p(a_s), _p(a_d), _p(asrc), _p(adst), B * N, H, C, _stream(),
                      nbytes=4 * B * N * (H * C + 2 * H), tag=f"R{B * N}xH{H}xC{C}")
            nbytes = (4 * B * (N * (H * C + cout + 2 * H) + nnz * H * (2 if want_alpha else 1)) + 4 * nnz
                      + 4 * (N + 1))
            ws = None
            if H == 1 and nnz > 0 and _tileable(C, z3, out, bias_c, zpre):
                ws = graph.gat_ws()
                if ws is not None and not _cabi.load().gcl_gat_ws_supported(ws[0].ref, ws[1].ref, C, B):
                    ws = None
            if ws is not None:          # coefficients in plan order + aggregation on the persistent tiled engine
                pf, pt, ent_f, ent_t, f2t, pcol = ws
                alpha = torch.empty((B, pf.n_entries), dtype=torch.float32, device=dev)       # alpha_f
                alr = torch.empty((B, pf.n_entries), dtype=torch.float32, device=dev)
                alpha_t = torch.zeros((B, pt.n_entries), dtype=torch.float32, device=dev)
                _call("gcl_gat_fwd_ws_f32", pf.ref, _p(ent_f), _p(pcol), _p(graph.perm), _p(f2t), _p(z3), _p(asrc),
                      _p(adst), _p(bias_c), _p(out), _p(alpha), _p(alr), _p(alpha_t), _p(alpha_pyg), _p(ps), _p(zpre),
                      B, N, nnz, C, pf.n_entries, pt.n_entries, float(slope), _stream(), nbytes=nbytes,
                      tag=f"N{N}xH{H}xC{C}xB{B}")
                asrc, adst = alr, alpha_t            # what the backward needs in place of the node scores
                mode = "ws"
            else:
                alpha = torch.empty((B, max(nnz, 1), H), dtype=torch.float32, device=dev)
                plan = None
                if H == 1 and nnz > 0 and _tileable(C, z3, out, bias_c, zpre):
                    plan = graph.plan(False, pad=1)        # ring too small (wide rows): one-tile-per-CTA kernels
                    if plan.n_heavy or graph.plan(True, pad=1).n_heavy:
                        plan = None
                if plan is not None:
                    _call("gcl_gat_fwd_tiled_f32", plan.ref, _p(graph.perm), _p(z3), _p(asrc), _p(adst), _p(bias_c),
                          _p(out), _p(alpha), _p(alpha_pyg), _p(ps), _p(zpre), B, N, nnz, C, float(slope), _stream(),
                          nbytes=nbytes, tag=f"N{N}xH{H}xC{C}xB{B}")
                    mode = "tiled"
                else:
                    _call("gcl_gat_fwd_f32", _p(graph.rowptr), _p(graph.col), _p(graph.perm), _p(z3), _p(asrc),
                          _p(adst), _p(bias_c), _p(out), _p(alpha), _p(alpha_pyg), _p(ps), _p(zpre), B, N, nnz, H, C,
                          int(bool(concat)), float(slope), _stream(), nbytes=nbytes, tag=f"N{N}xH{H}xC{C}xB{B}")
                    mode = "rows"
        ctx.ws, ctx.mode = mode == "ws", mode
        ctx.graph, ctx.H, ctx.C, ctx.concat, ctx.slope = graph, H, C, bool(concat), float(slope)
        ctx.squeeze, ctx.has_bias, ctx.has_prelu = squeeze, bias is not None, ps is not None
        ctx.save_for_backward(z3, asrc, adst, alpha, a_s, a_d, zpre, ps)
        if squeeze:
            out = out.squeeze(0)
        if want_alpha:
            ap = alpha_pyg[:, :nnz]
            ctx.mark_non_differentiable(ap)
            return out, (ap.squeeze(0) if squeeze else ap)
        return out, None

    @staticmethod
    def backward(ctx, dout, _dalpha=None):
        if dout is None:
            return (None,) * 11
        z3, asrc, adst, alpha, a_s, a_d, zpre, ps = ctx.saved_tensors
        g = ctx.graph
        B, N, HC = z3.shape
        H, C = ctx.H, ctx.C
        d3, _ = _as3(_chk(dout, "grad_out"))
        dslope = dbias = None
        if ctx.has_prelu and ctx.has_bias:
            d3, dslope, dbias = prelu_bwd_colsum_raw(d3, zpre, ps)        # PReLU backward + bias gradient, one pass
            dslope = dslope.view_as(ps)
        elif ctx.has_prelu:
            d3, dslope = prelu_bwd_raw(d3, zpre, ps)
            dslope = dslope.view_as(ps)
        dev = z3.device
        gbuf = torch.empty_like(alpha) if not ctx.ws else None
        da_s = torch.empty_like(asrc) if not ctx.ws else None
        da_d = torch.empty_like(adst) if not ctx.ws else None
        dz = torch.empty_like(z3)
        datt_s = torch.empty(HC, dtype=torch.float32, device=dev)
        datt_d = torch.empty(HC, dtype=torch.float32, device=dev)
        lib = _cabi.load()
        nb = lib.gcl_gat_datt_workspace_bytes(B * N, H, C)
        ws = _ws(nb, dev)
        with torch.cuda.device(dev):
            nbytes = 4 * B * (N * (2 * H * C + d3.shape[-1] + 4 * H) + 3 * g.nnz * H) + 16 * g.nnz
            if ctx.ws:
                pf, pt, ent_f, ent_t, f2t, _ = g.gat_ws()
                alr, alpha_t = asrc, adst            # (saved in their place by the forward)
                g_t = torch.zeros_like(alpha_t)
                da_s = torch.empty((B, N, 1), dtype=torch.float32, device=dev)
                da_d = torch.empty((B, N, 1), dtype=torch.float32, device=dev)
                _call("gcl_gat_bwd_ws_f32", pf.ref, _p(ent_f), pt.ref, _p(ent_t), _p(f2t), _p(z3), _p(alpha), _p(alr),
                      _p(alpha_t), _p(a_s), _p(a_d), _p(d3), _p(g_t), _p(da_s), _p(da_d), _p(dz), B, N, C,
                      pf.n_entries, pt.n_entries, _stream(), nbytes=nbytes, tag=f"N{N}xH{H}xC{C}xB{B}")
            elif ctx.mode == "tiled":
                _call("gcl_gat_bwd_tiled_f32", g.plan(False, pad=1).ref, g.plan(True, pad=1).ref, _p(g.t2r), _p(z3),
                      _p(asrc), _p(adst), _p(alpha), _p(a_s), _p(a_d), _p(d3), _p(gbuf), _p(da_s), _p(da_d), _p(dz),
                      B, N, g.nnz, C, ctx.slope, _stream(), nbytes=nbytes, tag=f"N{N}xH{H}xC{C}xB{B}")
            else:
                _call("gcl_gat_bwd_f32", _p(g.rowptr), _p(g.col), _p(g.rowptr_t), _p(g.col_t), _p(g.t2r), _p(z3),
                      _p(asrc), _p(adst), _p(alpha), _p(a_s), _p(a_d), _p(d3), _p(gbuf), _p(da_s), _p(da_d), _p(dz),
                      B, N, g.nnz, H, C, int(ctx.concat), ctx.slope, _stream(), nbytes=nbytes,
                      tag=f"N{N}xH{H}xC{C}xB{B}")
            _call("gcl_gat_datt_f32", _p(z3), _p(da_s), _p(da_d), _p(datt_s), _p(datt_d), B * N, H, C, _p(ws), nb,
                  _stream(), nbytes=4 * B * N * (H * C + 2 * H), tag=f"R{B * N}xH{H}xC{C}")
        if ctx.has_bias and dbias is None:
            dbias = colsum_raw(d3.view(-1, d3.shape[-1]))
        if ctx.squeeze:
            dz = dz.squeeze(0)
        return dz, datt_s.view(1, H, C), datt_d.view(1, H, C), dbias, None, None, None, None, None, dslope, None


def gat_attend(z, att_src, att_dst, bias, graph, heads, concat, negative_slope, want_alpha=False, prelu_slope=None,
               scores=None):
    """prelu_slope (heads == 1): PReLU applied behind the bias inside the aggregation kernel.
    scores = (a_src, a_dst) from linear_scores(): skips the separate score pass over z."""
    return _GAT.apply(z, att_src, att_dst, bias, graph, heads, concat, negative_slope, want_alpha, prelu_slope, scores)


def _rows_concat_raw(a, b, B, na, nb, C, dev):
    out = torch.empty((B, na + nb, C), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _call("gcl_rows_concat_f32", _p(a), _p(b), _p(out), B, na, nb, C, _stream(), nbytes=8 * B * (na + nb) * C,
              tag=f"B{B}x({na}+{nb})xC{C}")
    return out


def _rows_split_raw(x, na, want_a=True, want_b=True):
    B, n, C = x.shape
    a = torch.empty((B, na, C), dtype=torch.float32, device=x.device) if want_a else None
    b = torch.empty((B, n - na, C), dtype=torch.float32, device=x.device) if want_b else None
    with torch.cuda.device(x.device):
        _call("gcl_rows_split_f32", _p(x), _p(a), _p(b), B, na, n - na, C, _stream(), nbytes=8 * B * n * C,
              tag=f"B{B}x({na}+{n - na})xC{C}")
    return a, b


class _ConcatRows(torch.autograd.Function):
    """[B, Na, C], [B, Nb, C] -> [B, Na + Nb, C] (torch.cat(dim=1) of models.py:865) in one pass."""

    @staticmethod
    def forward(ctx, a, b):
        ac, bc = _chk(a, "a"), _chk(b, "b")
        if ac.dim() != 3 or bc.dim() != 3 or ac.shape[0] != bc.shape[0] or ac.shape[2] != bc.shape[2]:
            raise ValueError(f"gcl_b200: concat_rows got {tuple(ac.shape)} and {tuple(bc.shape)}")
        ctx.na = ac.shape[1]
        return _rows_concat_raw(ac, bc, ac.shape[0], ac.shape[1], bc.shape[1], ac.shape[2], ac.device)

    @staticmethod
    def backward(ctx, d):
        return _rows_split_raw(_chk(d, "grad_out"), ctx.na, ctx.needs_input_grad[0], ctx.needs_input_grad[1])


class _SplitRows(torch.autograd.Function):
    """[B, N, C] -> contiguous [B, Na, C], [B, N - Na, C] (the slices of models.py:841-842) in one pass."""

    @staticmethod
    def forward(ctx, x, na):
        xc = _chk(x, "x")
        if xc.dim() != 3 or not 0 < int(na) < xc.shape[1]:
            raise ValueError(f"gcl_b200: split_rows got {tuple(xc.shape)}, na={na}")
        ctx.shape = xc.shape
        ctx.na = int(na)
        return _rows_split_raw(xc, int(na))

    @staticmethod
    def backward(ctx, da, db):
        B, n, C = ctx.shape
        da = _chk(da, "grad_a") if da is not None else None
        db = _chk(db, "grad_b") if db is not None else None
        return _rows_concat_raw(da, db, B, ctx.na, n - ctx.na, C, (da if da is not None else db).device), None


def _rows_block_copy_raw(src, dst, rows, src_row0, dst_row0):
    B, ns, C = src.shape
    nd = dst.shape[1]
    with torch.cuda.device(src.device):
        _call("gcl_rows_block_copy_f32", _p(src), _p(dst), B, rows, C, src_row0, ns, dst_row0, nd, _stream(),
              nbytes=8 * B * rows * C, tag=f"B{B}x{rows}of{max(ns, nd)}xC{C}")
    return dst


class RowBridge:
    """Shared state of one take_rows / put_rows pair (see put_rows)."""

    def __init__(self):
        self.d_full = None


class _TakeRows(torch.autograd.Function):
    """x [B, N, C] -> contiguous copy of x[:, n0:] (the mesh rows, models.py:842).  Backward: see put_rows."""

    @staticmethod
    def forward(ctx, x, n0, bridge):
        xc = _chk(x, "x")
        if xc.dim() != 3 or not 0 <= int(n0) < xc.shape[1]:
            raise ValueError(f"gcl_b200: take_rows got {tuple(xc.shape)}, n0={n0}")
        ctx.shape, ctx.n0, ctx.bridge = xc.shape, int(n0), bridge
        out = torch.empty((xc.shape[0], xc.shape[1] - int(n0), xc.shape[2]), dtype=torch.float32, device=xc.device)
        return _rows_block_copy_raw(xc, out, out.shape[1], int(n0), 0)

    @staticmethod
    def backward(ctx, d):
        d = _chk(d, "grad_out")
        full = ctx.bridge.d_full if ctx.bridge is not None else None
        if full is None:               # no put_rows downstream: an ordinary gradient (zeros outside the block)
            g = torch.zeros(ctx.shape, dtype=torch.float32, device=d.device)
            _rows_block_copy_raw(d, g, d.shape[1], 0, ctx.n0)
            return g, None, None
        # put_rows' backward already handed the [B, N, C] gradient of x to autograd with this block left open:
        # complete it in place (it is not consumed before this node has run) instead of materialising a second
        # [B, N, C] tensor for autograd to add
        ctx.bridge.d_full = None
        _rows_block_copy_raw(d, full, d.shape[1], 0, ctx.n0)
        return None, None, None


class _PutRows(torch.autograd.Function):
    """x[:, n0:] = rows IN PLACE, returns x: torch.cat((x[:, :n0], rows), dim=1) of models.py:865 without moving the
    first n0 rows."""

    @staticmethod
    def forward(ctx, x, rows, n0, bridge):
        xc, rc = _chk(x, "x"), _chk(rows, "rows")
        if xc is not x:
            raise ValueError("gcl_b200: put_rows needs a contiguous fp32 CUDA tensor to write into")
        if xc.dim() != 3 or rc.dim() != 3 or rc.shape[0] != xc.shape[0] or rc.shape[2] != xc.shape[2] or \
                rc.shape[1] + int(n0) != xc.shape[1]:
            raise ValueError(f"gcl_b200: put_rows got {tuple(xc.shape)}, {tuple(rc.shape)}, n0={n0}")
        ctx.n0, ctx.bridge, ctx.rows_shape = int(n0), bridge, rc.shape
        _rows_block_copy_raw(rc, xc, rc.shape[1], 0, int(n0))
        ctx.mark_dirty(x)
        return x

    @staticmethod
    def backward(ctx, d):
        d = _chk(d, "grad_out")
        d_rows = None
        if ctx.needs_input_grad[1]:
            d_rows = torch.empty(ctx.rows_shape, dtype=torch.float32, device=d.device)
            _rows_block_copy_raw(d, d_rows, d_rows.shape[1], ctx.n0, 0)
        if not ctx.needs_input_grad[0]:
            return None, d_rows, None, None
        # gradient of the overwritten x: d with rows n0.. zero.  When the rows were taken from x by take_rows and
        # went through a differentiable path, its backward fills exactly those rows of this same tensor.
        if ctx.bridge is not None and ctx.needs_input_grad[1]:
            ctx.bridge.d_full = d
        else:
            d[:, ctx.n0:].zero_()
        return d, d_rows, None, None


def take_rows(x, n0: int, bridge=None):
    return _TakeRows.apply(x, n0, bridge)


def put_rows(x, rows, n0: int, bridge=None):
    """x[:, n0:] = rows in place (x must be a non-leaf that no other op still needs in its old state)."""
    return _PutRows.apply(x, rows, n0, bridge)


def concat_rows(a, b):
    return _ConcatRows.apply(a, b)


def split_rows(x, na: int):
    return _SplitRows.apply(x, na)


class _ResizeChannels(torch.autograd.Function):
    """x[..., :c] (c smaller: drop alignment padding) or zero-pad to c channels, contiguous result, one kernel each
    way (torch's slice backward is a zero fill plus a strided copy)."""

    @staticmethod
    def forward(ctx, x, c):
        xc = _chk(x, "x")
        ctx.c_in = xc.shape[-1]
        rows = xc.numel() // max(ctx.c_in, 1)
        y = torch.empty(xc.shape[:-1] + (int(c),), dtype=torch.float32, device=xc.device)
        with torch.cuda.device(xc.device):
            _call("gcl_resize_channels_f32", _p(xc), _p(y), rows, ctx.c_in, int(c), _stream(),
                  nbytes=4 * rows * (ctx.c_in + int(c)), tag=f"R{rows}x{ctx.c_in}->{int(c)}")
        return y

    @staticmethod
    def backward(ctx, dy):
        d = _chk(dy, "grad_out")
        rows = d.numel() // d.shape[-1]
        dx = torch.empty(d.shape[:-1] + (ctx.c_in,), dtype=torch.float32, device=d.device)
        with torch.cuda.device(d.device):
            _call("gcl_resize_channels_f32", _p(d), _p(dx), rows, d.shape[-1], ctx.c_in, _stream(),
                  nbytes=4 * rows * (ctx.c_in + d.shape[-1]), tag=f"R{rows}x{d.shape[-1]}->{ctx.c_in}")
        return dx, None


def resize_channels(x, c: int):
    return _ResizeChannels.apply(x, c)


ACT_RELU, ACT_SILU = 1, 2


class _Act(torch.autograd.Function):
    """ReLU / SiLU of _get_activation (models.py:154-163)."""

    @staticmethod
    def forward(ctx, x, kind):
        xc = _chk(x, "x")
        y = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _call("gcl_act_fwd_f32", _p(xc), _p(y), xc.numel(), int(kind), _stream(), nbytes=8 * xc.numel())
        ctx.save_for_backward(xc)
        ctx.kind = int(kind)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        d = _chk(dy, "grad_out")
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _call("gcl_act_bwd_f32", _p(d), _p(x), _p(dx), x.numel(), ctx.kind, _stream(), nbytes=12 * x.numel())
        return dx, None


def act(x, kind: int):
    return _Act.apply(x, kind)


class _Add(torch.autograd.Function):
    """a + b (residual connections, models.py:226-227); the gradient passes through unchanged."""

    @staticmethod
    def forward(ctx, a, b):
        ac, bc = _chk(a, "a"), _chk(b, "b")
        if ac.shape != bc.shape:
            raise ValueError(f"gcl_b200: add got {tuple(ac.shape)} and {tuple(bc.shape)}")
        y = torch.empty_like(ac)
        with torch.cuda.device(ac.device):
            _call("gcl_add_f32", _p(ac), _p(bc), _p(y), ac.numel(), _stream(), nbytes=12 * ac.numel())
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


def add(a, b):
    return _Add.apply(a, b)


class _LayerNormGraph(torch.autograd.Function):
    """torch_geometric LayerNorm(mode='graph') without a batch vector: statistics over all elements of a sample
    ([N, C], or each [N, C] slice of [B, N, C]); y = (x - mean) / (std + eps) * weight + bias."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        xc = _chk(x, "x")
        x3, squeeze = _as3(xc)
        B, N, C = x3.shape
        g = _chk(gamma, "weight") if gamma is not None else None
        b = _chk(beta, "bias") if beta is not None else None
        y = torch.empty_like(x3)
        stats = torch.empty((B, 3), dtype=torch.float32, device=xc.device)
        lib = _cabi.load()
        nb = lib.gcl_layernorm_graph_workspace_bytes(B)
        ws = _ws(nb, xc.device)
        with torch.cuda.device(xc.device):
            _call("gcl_layernorm_graph_fwd_f32", _p(x3), _p(g), _p(b), _p(y), _p(stats), B, N * C, C, float(eps), _p(ws), nb,
                  _stream(), nbytes=12 * x3.numel(), tag=f"B{B}xN{N}xC{C}")
        ctx.save_for_backward(x3, g, stats)
        ctx.affine, ctx.squeeze = gamma is not None, squeeze
        return y.squeeze(0) if squeeze else y

    @staticmethod
    def backward(ctx, dy):
        x3, g, stats = ctx.saved_tensors
        B, N, C = x3.shape
        d3, _ = _as3(_chk(dy, "grad_out"))
        dx = torch.empty_like(x3)
        t = torch.empty_like(x3) if ctx.affine else None
        coef = torch.empty((B, 2), dtype=torch.float32, device=x3.device)
        lib = _cabi.load()
        nb = lib.gcl_layernorm_graph_workspace_bytes(B)
        ws = _ws(nb, x3.device)
        with torch.cuda.device(x3.device):
            _call("gcl_layernorm_graph_bwd_f32", _p(d3), _p(x3), _p(g), _p(stats), _p(dx), _p(t), _p(coef), B, N * C, C,
                  _p(ws), nb, _stream(), nbytes=16 * x3.numel(), tag=f"B{B}xN{N}xC{C}")
        dg = db = None
        if ctx.affine:
            dg = colsum_raw(t.view(-1, C))
            db = colsum_raw(d3.reshape(-1, C))
        return (dx.squeeze(0) if ctx.squeeze else dx), dg, db, None


def layer_norm_graph(x, weight, bias, eps=1e-5):
    return _LayerNormGraph.apply(x, weight, bias, eps)


class _SpmmFixed(torch.autograd.Function):
    """out = A x with a fixed sparse A given as CSR (+ optional tile plan) and A^T for the backward: the building block of
    the edge gathers / scatters of the InteractionNet processor.  x [B, n_in, C] -> [B, n_out, C]."""

    @staticmethod
    def forward(ctx, x, fwd, bwd):
        x3, squeeze = _as3(_chk(x, "x"))
        rowptr, col, w, n_out, plan = fwd
        out, _ = spmm_raw(rowptr, col, w, x3, n_out, plan=plan if _tileable(x3.shape[-1]) else None, wkey=("fixed", id(w)))
        ctx.bwd, ctx.squeeze = bwd, squeeze
        return out.squeeze(0) if squeeze else out

    @staticmethod
    def backward(ctx, dout):
        d3, _ = _as3(_chk(dout, "grad_out"))
        rowptr, col, w, n_out, plan = ctx.bwd
        dx, _ = spmm_raw(rowptr, col, w, d3, n_out, plan=plan if _tileable(d3.shape[-1]) else None, wkey=("fixed", id(w)))
        return (dx.squeeze(0) if ctx.squeeze else dx), None, None


def spmm_fixed(x, fwd, bwd):
    return _SpmmFixed.apply(x, fwd, bwd)


def edge_prune(ei_pyg: torch.Tensor, alpha_pyg: torch.Tensor, threshold: float) -> torch.Tensor:
    """SparseGATConv pruning (models.py:140-149): edges with alpha >= threshold, order preserved."""
    if not ei_pyg.is_cuda or not alpha_pyg.is_cuda:
        raise RuntimeError("gcl_b200: edge_prune needs CUDA tensors; no CPU fallback")
    a = alpha_pyg.detach().reshape(-1).to(torch.float32).contiguous()
    nnz = a.numel()
    if ei_pyg.shape[1] != nnz:
        raise ValueError("edge_prune: one attention value per edge expected (heads=1)")
    ei = ei_pyg if ei_pyg.stride(1) == 1 else ei_pyg.contiguous()
    out = torch.empty((2, max(nnz, 1)), dtype=torch.int64, device=ei.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=ei.device)
    lib = _cabi.load()
    nb = lib.gcl_edge_prune_workspace_bytes(nnz)
    ws = _ws(nb, ei.device)
    with torch.cuda.device(ei.device):
        _call("gcl_edge_prune", _p(ei), _p(a), nnz, ei.stride(0), float(threshold), _p(out), out.stride(0),
              _p(cnt), _p(ws), nb, _stream())
    return out[:, : int(cnt.item())].contiguous()
