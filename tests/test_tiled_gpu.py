"""Tiled aggregation kernels (tile.cu, gat_tile.cu) through the C ABI: against the row-gather kernels (same per-row
summation order => identical bits outside heavy rows), against dense fp64 references, with small tile limits that
force many tiles, masked columns, row prefixes, heavy rows (687 entries, the 512x256 polar-row case) and row-order
hints."""
import numpy as np
import pytest
import torch

from helpers import RTOL_F32, assert_close, random_graph

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _graph(n, e, seed, **kw):
    from gcl_b200.graph import CSR_LOOPS, CSRGraph
    return CSRGraph(random_graph(n, e, seed=seed, **kw).to(DEV), n, CSR_LOOPS)


@pytest.mark.parametrize("C,B", [(64, 5), (128, 3), (96, 2), (48, 4), (20, 9), (4, 1)])
@pytest.mark.parametrize("limits", [(64, 128, 1024), (8, 24, 64), (3, 16, 32)])
def test_tiled_spmm_equals_row_gather_bits(C, B, limits):
    from gcl_b200 import ops
    from gcl_b200.graph import NORM_GCN, TilePlan
    n = 1500
    g = _graph(n, 9000, seed=C + B, isolated=30, dups=20)
    w, wt = g.weights(NORM_GCN)
    x = torch.randn(B, n, C, device=DEV)
    bias = torch.randn(C, device=DEV)
    slope = torch.tensor([0.25], device=DEV)
    for rp, co, ww in ((g.rowptr, g.col, w), (g.rowptr_t, g.col_t, wt)):
        order = np.random.default_rng(C).permutation(n).astype(np.int32)
        for od in (None, order):
            pl = TilePlan(rp, co, g.nnz, n, n, n, od, *limits)
            assert pl.n_tiles > 0
            ref, zref = ops.spmm_raw(rp, co, ww, x, n, bias, slope, True)
            got, zgot = ops.spmm_raw(rp, co, ww, x, n, bias, slope, True, plan=pl)
            light = torch.ones(n, dtype=torch.bool, device=DEV)
            light[pl.t["heavy_rows"][: pl.n_heavy].long()] = False        # heavy rows: another (fixed) summation order
            assert torch.equal(got[:, light], ref[:, light]) and torch.equal(zgot[:, light], zref[:, light])
            assert_close(got, ref, 1e-6, "heavy rows")


def test_tiled_spmm_heavy_rows_prefix_and_masks():
    """Rows with 687 (> max_union) entries go to the CTA-per-row kernel; rows_out prefix and masked columns."""
    from gcl_b200 import ops
    from gcl_b200.graph import NORM_GCN, TilePlan
    n, C, B = 4000, 64, 3
    gen = torch.Generator().manual_seed(0)
    ei = random_graph(n, 12000, seed=11)
    heavy_src = torch.randperm(n, generator=gen)[:687]
    ei = torch.cat([ei, torch.stack([heavy_src, torch.full_like(heavy_src, 5)]),
                    torch.stack([heavy_src[:300], torch.full_like(heavy_src[:300], 3000)])], dim=1)
    from gcl_b200.graph import CSR_LOOPS, CSRGraph
    g = CSRGraph(ei.to(DEV), n, CSR_LOOPS)
    w, _ = g.weights(NORM_GCN)
    x = torch.randn(B, n, C, device=DEV)
    bias = torch.randn(C, device=DEV)
    pl = g.plan(False)
    assert pl.n_heavy == 2
    ref, _ = ops.spmm_raw(g.rowptr, g.col, w, x, n, bias)
    got, _ = ops.spmm_raw(g.rowptr, g.col, w, x, n, bias, plan=pl)
    heavy = pl.t["heavy_rows"][: pl.n_heavy].long()
    mask = torch.ones(n, dtype=torch.bool, device=DEV)
    mask[heavy] = False
    assert torch.equal(got[:, mask], ref[:, mask])
    assert_close(got[:, heavy], ref[:, heavy], 1e-6, "heavy rows (different fixed summation order)")
    again, _ = ops.spmm_raw(g.rowptr, g.col, w, x, n, bias, plan=pl)
    assert torch.equal(again, got)                                   # deterministic
    # output restricted to a row prefix, inputs restricted to a row prefix (the decoder's forward / backward)
    n_out, n_in = 2500, 3100
    ref_p, _ = ops.spmm_raw(g.rowptr, g.col, w, x[:, :n_in].contiguous(), n_out, bias)
    pl_p = TilePlan(g.rowptr, g.col, g.nnz, n, n_out, n_in, None, 16, 48, 256)
    got_p, _ = ops.spmm_raw(g.rowptr, g.col, w, x[:, :n_in].contiguous(), n_out, bias, plan=pl_p)
    hp = pl_p.t["heavy_rows"][: pl_p.n_heavy].long()
    mp = torch.ones(n_out, dtype=torch.bool, device=DEV)
    mp[hp] = False
    assert torch.equal(got_p[:, mp], ref_p[:, mp])
    assert_close(got_p, ref_p, 1e-6, "prefix + masks")


@pytest.mark.parametrize("C,B", [(64, 5), (128, 2), (48, 3), (32, 1)])
def test_tiled_gat_forward_backward_matches_row_gather_and_dense(C, B):
    """Single-head GATConv message passing: tiled fused forward / two-pass backward vs the first-generation kernels
    (1e-5: exp and summation orders differ slightly) and vs a dense fp64 softmax reference (1e-4)."""
    from gcl_b200 import graph as gg, ops
    n = 1200
    g = _graph(n, 7000, seed=C * 3 + B, isolated=10)
    gen = torch.Generator().manual_seed(C)
    z0 = torch.randn(B, n, C, generator=gen)
    a_s0, a_d0 = torch.randn(1, 1, C, generator=gen) * 0.3, torch.randn(1, 1, C, generator=gen) * 0.3
    b0 = torch.randn(C, generator=gen)
    go = torch.randn(B, n, C, generator=gen).to(DEV)
    res = {}
    for tiled in (False, True):
        gg.TILED = tiled
        try:
            z, a_s, a_d, bias = (t.clone().to(DEV).requires_grad_(True) for t in (z0, a_s0, a_d0, b0))
            out, alpha = ops.gat_attend(z, a_s, a_d, bias, g, 1, False, 0.2, want_alpha=True)
            grads = torch.autograd.grad(out, (z, a_s, a_d, bias), go)
            res[tiled] = [out.detach(), alpha.detach()] + [t.detach() for t in grads]
        finally:
            gg.TILED = True
    for a, b, nm in zip(res[True], res[False], ("out", "alpha", "dz", "datt_src", "datt_dst", "dbias")):
        assert_close(a, b, 1e-5, f"tiled vs row-gather {nm}", atol=1e-6)
    # dense fp64 reference
    ei = g.edge_index_with_loops.cpu()
    zd = z0.double().requires_grad_(True)
    asd, add_, bd = a_s0.double().requires_grad_(True), a_d0.double().requires_grad_(True), b0.double().requires_grad_(True)
    s_src, s_dst = (zd * asd.view(1, 1, C)).sum(-1), (zd * add_.view(1, 1, C)).sum(-1)
    e = torch.nn.functional.leaky_relu(s_src[:, ei[0]] + s_dst[:, ei[1]], 0.2)              # [B, E]
    dense = torch.full((B, n, n), float("-inf"), dtype=torch.float64)
    out_d = torch.zeros(B, n, C, dtype=torch.float64)
    m = torch.full((B, n), float("-inf"), dtype=torch.float64).scatter_reduce(1, ei[1].expand(B, -1), e.detach(), "amax")
    p = (e - m[:, ei[1]]).exp()
    den = torch.zeros(B, n, dtype=torch.float64).scatter_add(1, ei[1].expand(B, -1), p)
    al = p / (den[:, ei[1]] + 1e-16)
    out_d = out_d.index_add(1, ei[1], al.unsqueeze(-1) * zd[:, ei[0]]) + bd
    gd = torch.autograd.grad(out_d, (zd, asd, add_, bd), go.cpu().double())
    assert_close(res[True][0], out_d.detach().float(), RTOL_F32, "tiled out vs dense fp64")
    assert_close(res[True][1].squeeze(-1), al.detach().float(), RTOL_F32, "tiled alpha (PyG order) vs dense fp64")
    for a, b, nm in zip(res[True][2:], gd, ("dz", "datt_src", "datt_dst", "dbias")):
        assert_close(a, b.float(), RTOL_F32, f"tiled {nm} vs dense fp64", atol=1e-6)


def test_tiled_gat_fused_prelu_and_determinism():
    from gcl_b200 import ops
    n, C, B = 2000, 64, 4
    g = _graph(n, 12000, seed=77)
    z = torch.randn(B, n, C, device=DEV, requires_grad=True)
    a_s = (torch.randn(1, 1, C, device=DEV) * 0.3).requires_grad_(True)
    a_d = (torch.randn(1, 1, C, device=DEV) * 0.3).requires_grad_(True)
    bias = torch.randn(C, device=DEV, requires_grad=True)
    slope = torch.tensor([0.25], device=DEV, requires_grad=True)
    go = torch.randn(B, n, C, device=DEV)
    plain, _ = ops.gat_attend(z, a_s, a_d, bias, g, 1, False, 0.2)
    fused, _ = ops.gat_attend(z, a_s, a_d, bias, g, 1, False, 0.2, prelu_slope=slope)
    assert torch.equal(fused, torch.nn.functional.prelu(plain, slope.detach()))
    runs = []
    for _ in range(2):
        o, _ = ops.gat_attend(z, a_s, a_d, bias, g, 1, False, 0.2, prelu_slope=slope)
        runs.append([o.detach()] + list(torch.autograd.grad(o, (z, a_s, a_d, bias, slope), go)))
    for a, b in zip(*runs):
        assert torch.equal(a, b)


def test_model_registers_order_hints_and_results_do_not_depend_on_them():
    from gcl_b200 import graph as gg, ops
    from gcl_b200.graph import CSR_LOOPS, NORM_GCN, CSRGraph
    from gcl_b200.graphs_build import ModelGraphs
    mg = ModelGraphs(16, 32, [1, 3], 0.6, DEV)
    M, N = mg.num_mesh, mg.num_grid + mg.num_mesh
    assert sorted(gg.ORDER_HINTS[M].tolist()) == list(range(M))
    assert sorted(gg.ORDER_HINTS[N].tolist()) == list(range(N))
    x = torch.randn(2, M, 64, device=DEV)
    outs = []
    for use_hint in (True, False):
        saved = dict(gg.ORDER_HINTS)
        if not use_hint:
            gg.ORDER_HINTS.clear()
        try:
            g = CSRGraph(mg.processing_graph, M, CSR_LOOPS)
            outs.append(ops.aggregate(x, g, NORM_GCN))
        finally:
            gg.ORDER_HINTS.update(saved)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("C,B", [(64, 5), (128, 2), (96, 3), (256, 1), (8, 4)])
def test_bf16_feature_rows(C, B):
    """bf16 storage of the feature rows (north_star: rel 2e-2): the tiled engine with bf16 x / out and fp32
    accumulation equals the fp32 kernel on the same (bf16-representable) inputs up to the final rounding to bf16;
    forward, backward, heavy rows included; GCNConv on bf16 input vs the fp32 oracle within 2e-2."""
    from gcl_b200 import ops
    from gcl_b200.graph import CSR_LOOPS, NORM_GCN, CSRGraph
    n = 3000
    gen = torch.Generator().manual_seed(C)
    ei = random_graph(n, 15000, seed=C + 1, isolated=20)
    heavy_src = torch.randperm(n, generator=gen)[:400]
    ei = torch.cat([ei, torch.stack([heavy_src, torch.full_like(heavy_src, 7)])], dim=1)
    g = CSRGraph(ei.to(DEV), n, CSR_LOOPS)
    assert g.plan(False).n_heavy == 1
    x = torch.randn(B, n, C, generator=gen).to(torch.bfloat16)
    bias = torch.randn(C, generator=gen)
    go = torch.randn(B, n, C, generator=gen).to(torch.bfloat16)
    xb = x.to(DEV).requires_grad_(True)
    bb = bias.to(DEV).requires_grad_(True)
    out = ops.aggregate(xb, g, NORM_GCN, bb)
    assert out.dtype == torch.bfloat16
    out.backward(go.to(DEV))
    xf = x.float().to(DEV).requires_grad_(True)
    bf = bias.to(DEV).requires_grad_(True)
    ref = ops.aggregate(xf, g, NORM_GCN, bf)
    ref.backward(go.float().to(DEV))
    # only the final rounding to bf16 (2^-9 relative per element) separates the two
    assert float((out.float() - ref).abs().max()) <= 2.0 ** -8 * float(ref.abs().max())
    assert float((xb.grad.float() - xf.grad).abs().max()) <= 2.0 ** -8 * float(xf.grad.abs().max())
    assert_close(bb.grad, bf.grad, 1e-5, "dbias")
    assert torch.equal(ops.aggregate(xb.detach(), g, NORM_GCN, bb.detach()), out.detach())     # deterministic
    if C <= 128:
        import torch_geometric.nn as onn
        import gcl_b200.nn as gnn
        torch.manual_seed(C)
        lo, lg = onn.GCNConv(C, C), gnn.GCNConv(C, C).to(DEV)
        lg.load_state_dict(lo.state_dict())
        y_ref = lo(x[0].float(), ei)
        y = lg(x[0].to(DEV), ei.to(DEV))
        assert y.dtype == torch.bfloat16
        assert_close(y.float(), y_ref, 2e-2, "GCNConv on bf16 rows vs fp32 oracle")
