"""GPU parity tests proper: every kernel, through the C ABI (ctypes host layer), against the CPU oracle
(oracle/pyg_shim = restated torch_geometric 2.5.3) on the same seeded inputs.

Bars (BASELINE.json north_star): bit-exact for CSR construction / indexing; rel 1e-4 (max-norm
relative) for fp32 layer outputs and gradients."""
import numpy as np
import pytest
import torch

from helpers import RTOL_F32, assert_close, np_csr_from_pyg, random_graph

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _oracle_nn():
    import torch_geometric.nn as onn  # the oracle shim (tests/conftest.py puts it on sys.path)
    assert "oracle-restatement" in __import__("torch_geometric").__version__
    return onn


GRAPH_CASES = [
    dict(n=1, e=0, seed=0),
    dict(n=7, e=0, seed=1),
    dict(n=50, e=200, seed=2, self_loops=5, dups=10),
    dict(n=300, e=1500, seed=3, isolated=40),
    dict(n=2000, e=9000, seed=4, heavy=700, self_loops=3),
    dict(n=10242, e=65280, seed=5),
]


@pytest.mark.parametrize("case", GRAPH_CASES)
@pytest.mark.parametrize("loops", [True, False])
def test_csr_build_bit_exact(case, loops):
    from gcl_b200.graph import CSR_LOOPS, CSR_RAW, CSRGraph
    kw = dict(case)
    n = kw.pop("n")
    ei = random_graph(n, **kw)
    g = CSRGraph(ei.to(DEV), n, CSR_LOOPS if loops else CSR_RAW)
    ref = np_csr_from_pyg(ei.numpy(), n, loops)
    nnz = len(ref["src"])
    assert g.nnz == nnz
    got_ei = g.edge_index_with_loops.cpu().numpy()
    assert np.array_equal(got_ei, np.stack([ref["src"], ref["dst"]])), "PyG-order edge list differs"
    for name in ("rowptr", "rowptr_t"):
        assert np.array_equal(getattr(g, name).cpu().numpy(), ref[name]), name
    for name in ("col", "perm", "col_t", "perm_t"):
        assert np.array_equal(getattr(g, name).cpu().numpy()[:nnz], ref[name]), name
    inv = np.empty(nnz, dtype=np.int64)
    inv[ref["perm"]] = np.arange(nnz)
    assert np.array_equal(g.t2r.cpu().numpy()[:nnz], inv[ref["perm_t"]]), "t2r"
    # building twice gives identical arrays (atomics only order the scratch, not the result)
    g2 = CSRGraph(ei.to(DEV), n, CSR_LOOPS if loops else CSR_RAW)
    assert torch.equal(g.col[:nnz], g2.col[:nnz]) and torch.equal(g.perm_t[:nnz], g2.perm_t[:nnz])


@pytest.mark.parametrize("weighted", [False, True])
def test_gcn_norm_matches_oracle(weighted):
    from gcl_b200.graph import CSR_LOOPS, NORM_GCN, CSRGraph
    onn = _oracle_nn()
    n = 500
    ei = random_graph(n, 3000, seed=11, self_loops=20, dups=30, isolated=25)
    ew = torch.rand(ei.size(1), generator=torch.Generator().manual_seed(3)) + 0.1 if weighted else None
    ei_o, w_o = onn.gcn_norm(ei, ew, n)
    g = CSRGraph(ei.to(DEV), n, CSR_LOOPS, ew.to(DEV) if weighted else None)
    w, wt = g.weights(NORM_GCN)
    perm = g.perm.cpu().long()[: g.nnz]
    assert torch.equal(g.edge_index_with_loops.cpu(), ei_o)
    got = w.cpu()[: g.nnz]
    assert torch.allclose(got, w_o[perm], rtol=2e-7, atol=0), float((got - w_o[perm]).abs().max())
    assert torch.equal(wt.cpu()[: g.nnz], got[g.t2r.cpu().long()[: g.nnz]])


def _dense_adj(ei, w, n):
    a = torch.zeros(n, n, dtype=torch.float64)
    a.index_put_((ei[1], ei[0]), w.double(), accumulate=True)
    return a


@pytest.mark.parametrize("C", [4, 12, 15, 19, 33, 48, 64, 96, 128, 200, 256])
@pytest.mark.parametrize("B", [1, 3])
def test_spmm_forward_all_widths(C, B):
    """Aggregation kernel vs a dense fp64 matmul, incl. widths with C % 4 != 0 and C > 128."""
    from gcl_b200 import ops
    from gcl_b200.graph import CSR_LOOPS, NORM_GCN, CSRGraph
    onn = _oracle_nn()
    n = 257
    ei = random_graph(n, 1500, seed=C, heavy=100, isolated=10)
    ei_o, w_o = onn.gcn_norm(ei, None, n)
    x = torch.randn(B, n, C, generator=torch.Generator().manual_seed(C + B))
    bias = torch.randn(C, generator=torch.Generator().manual_seed(1))
    want = torch.einsum("ij,bjc->bic", _dense_adj(ei_o, w_o, n), x.double()) + bias.double()
    g = CSRGraph(ei.to(DEV), n, CSR_LOOPS)
    got = ops.aggregate(x.to(DEV), g, NORM_GCN, bias.to(DEV))
    assert_close(got, want.float(), what=f"spmm C={C} B={B}")
    slope = torch.tensor([0.25])
    got2 = ops.aggregate(x.to(DEV), g, NORM_GCN, bias.to(DEV), slope.to(DEV))
    assert_close(got2, torch.nn.functional.prelu(want.float(), slope), what="spmm+prelu")


def _copy_params(dst, src):
    dst.load_state_dict({k: v.clone() for k, v in src.state_dict().items()})


def _grad_check(layer_gpu, layer_cpu, x, run_gpu, run_cpu, tol=RTOL_F32, what=""):
    xg = x.clone().to(DEV).requires_grad_(True)
    xc = x.clone().requires_grad_(True)
    yg, yc = run_gpu(layer_gpu, xg), run_cpu(layer_cpu, xc)
    assert_close(yg, yc, tol, f"{what} forward")
    go = torch.randn(yc.shape, generator=torch.Generator().manual_seed(99))
    yg.backward(go.to(DEV))
    yc.backward(go)
    assert_close(xg.grad, xc.grad, tol, f"{what} dx")
    pc = dict(layer_cpu.named_parameters())
    for name, p in layer_gpu.named_parameters():
        assert p.grad is not None, f"{what}: no grad for {name}"
        assert_close(p.grad, pc[name].grad, tol, f"{what} d{name}")


@pytest.mark.parametrize("cin,cout", [(64, 64), (72, 48), (96, 96), (128, 128), (48, 33), (64, 19), (30, 12)])
@pytest.mark.parametrize("B", [None, 2])
def test_gcnconv_forward_backward(cin, cout, B):
    import gcl_b200.nn as gnn
    onn = _oracle_nn()
    n = 700
    ei = random_graph(n, 4000, seed=cin + cout, heavy=90, isolated=30, dups=7)
    torch.manual_seed(0)
    ref = onn.GCNConv(cin, cout)
    with torch.no_grad():
        ref.bias.uniform_(-0.5, 0.5)
    mine = gnn.GCNConv(cin, cout).to(DEV)
    _copy_params(mine, ref)
    shape = (n, cin) if B is None else (B, n, cin)
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(5))
    eig = ei.to(DEV)
    _grad_check(mine, ref, x, lambda m, t: m(t, eig), lambda m, t: m(t, ei), what=f"GCNConv {cin}->{cout} B={B}")


@pytest.mark.parametrize("C", [64, 12])
def test_simpleconv_mean(C):
    import gcl_b200.nn as gnn
    onn = _oracle_nn()
    n = 400
    ei = random_graph(n, 1200, seed=C, isolated=100, self_loops=4, dups=5)
    x = torch.randn(n, C, generator=torch.Generator().manual_seed(2))
    eig = ei.to(DEV)
    _grad_check(gnn.SimpleConv(aggr="mean"), onn.SimpleConv(aggr="mean"), x,
                lambda m, t: m(x=t, edge_index=eig), lambda m, t: m(x=t, edge_index=ei), what="SimpleConv")
    # rows without incoming edges are exactly zero (encoder grid rows in the sparse_attention config)
    y = gnn.SimpleConv(aggr="mean")(x.to(DEV), eig).cpu()
    indeg = torch.bincount(ei[1], minlength=n)
    assert torch.all(y[indeg == 0] == 0)


@pytest.mark.parametrize("heads,concat,C", [(1, False, 64), (4, False, 64), (2, True, 32), (3, False, 20), (1, False, 128)])
@pytest.mark.parametrize("B", [None, 2])
def test_gatconv_forward_backward_alpha(heads, concat, C, B):
    import gcl_b200.nn as gnn
    onn = _oracle_nn()
    n, cin = 600, 64
    ei = random_graph(n, 3600, seed=heads * 10 + C, heavy=80, self_loops=6, dups=9, isolated=20)
    torch.manual_seed(1)
    ref = onn.GATConv(cin, C, heads=heads, concat=concat)
    with torch.no_grad():
        ref.bias.uniform_(-0.5, 0.5)
    mine = gnn.GATConv(cin, C, heads=heads, concat=concat).to(DEV)
    _copy_params(mine, ref)
    shape = (n, cin) if B is None else (B, n, cin)
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(6))
    eig = ei.to(DEV)
    _grad_check(mine, ref, x, lambda m, t: m(t, eig), lambda m, t: m(t, ei), what=f"GATConv h={heads} concat={concat}")
    out_g, (ei_g, a_g) = mine(x.to(DEV), eig, return_attention_weights=True)
    out_c, (ei_c, a_c) = ref(x, ei, return_attention_weights=True)
    assert torch.equal(ei_g.cpu(), ei_c), "edge_index with self loops must be bit-exact, PyG order"
    assert_close(a_g, a_c, RTOL_F32, "alpha (PyG edge order)")
    assert_close(out_g, out_c, RTOL_F32, "GAT out")
    # softmax rows sum to one
    a1 = a_c if B is None else a_c[0]
    s = torch.zeros(n, heads).index_add_(0, ei_c[1], a1)
    assert torch.allclose(s, torch.ones_like(s), atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("B", [None, 3])
def test_gat_with_fused_prelu_matches_unfused(B):
    """heads == 1: GATConv + PReLU in one aggregation kernel (the model's fused path) against GATConv then
    PReLU of the oracle, forward and all gradients including the shared PReLU slope."""
    import gcl_b200.nn as gnn
    from gcl_b200 import ops
    from gcl_b200.graph import CSR_LOOPS, GLOBAL_CACHE
    onn = _oracle_nn()
    n, cin, C = 500, 64, 64
    ei = random_graph(n, 3000, seed=77, heavy=40, self_loops=5, dups=7, isolated=10)
    torch.manual_seed(3)
    ref = onn.GATConv(cin, C, heads=1, concat=False)
    with torch.no_grad():
        ref.bias.uniform_(-0.5, 0.5)
    ref_act = torch.nn.PReLU()
    mine = gnn.GATConv(cin, C, heads=1, concat=False).to(DEV)
    _copy_params(mine, ref)
    act = torch.nn.PReLU().to(DEV)
    shape = (n, cin) if B is None else (B, n, cin)
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(8))
    xg = x.to(DEV).requires_grad_(True)
    xc = x.clone().requires_grad_(True)
    g = GLOBAL_CACHE.get(ei.to(DEV), n, CSR_LOOPS)
    z = ops.linear(xg, mine.lin.weight)
    out_g, _ = ops.gat_attend(z, mine.att_src, mine.att_dst, mine.bias, g, 1, False, mine.negative_slope,
                              prelu_slope=act.weight)
    out_c = ref_act(ref(xc, ei))
    assert_close(out_g, out_c, RTOL_F32, "GAT+PReLU out")
    w = torch.randn(out_c.shape, generator=torch.Generator().manual_seed(9))
    (out_g * w.to(DEV)).sum().backward()
    (out_c * w).sum().backward()
    assert_close(xg.grad, xc.grad, RTOL_F32, "dx")
    assert_close(act.weight.grad, ref_act.weight.grad, RTOL_F32, "dslope")
    for (k, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
        assert_close(p.grad, q.grad, RTOL_F32, f"d{k}", atol=1e-6 if k == "att_dst" else 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("B,na,nb,C", [(3, 50, 70, 64), (2, 7, 5, 33), (1, 2048, 10242, 48)])
def test_rows_concat_split_are_exact_and_inverse(B, na, nb, C):
    """Node-axis concat / split (models.py:841-842, 865): bit-exact copies, each the other's backward."""
    from gcl_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + C)
    a = torch.randn(B, na, C, generator=g).to(DEV).requires_grad_(True)
    b = torch.randn(B, nb, C, generator=g).to(DEV).requires_grad_(True)
    out = ops.concat_rows(a, b)
    assert torch.equal(out, torch.cat((a, b), dim=1))
    w = torch.randn(out.shape, generator=g).to(DEV)
    (out * w).sum().backward()
    assert torch.equal(a.grad, w[:, :na]) and torch.equal(b.grad, w[:, na:])
    x = torch.randn(B, na + nb, C, generator=g).to(DEV).requires_grad_(True)
    p, q = ops.split_rows(x, na)
    assert p.is_contiguous() and q.is_contiguous()
    assert torch.equal(p, x[:, :na]) and torch.equal(q, x[:, na:])
    (q * w[:, na:]).sum().backward()                       # only one half used: the other half's gradient is zero
    assert torch.equal(x.grad[:, na:], w[:, na:]) and not x.grad[:, :na].any()


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cin,cout", [(1000, 20, 19), (1000, 19, 20), (7, 33, 36), (5, 4, 3), (1, 1, 4),
                                           (4099, 16, 15), (3, 15, 16)])
def test_resize_channels_matches_slice_and_pad(rows, cin, cout):
    """x[..., :c] / zero padding to c channels (the alignment padding of layers with odd widths), forward and backward,
    bit-exact against torch's slice / F.pad."""
    import torch.nn.functional as F
    from gcl_b200 import ops
    g = torch.Generator().manual_seed(rows + 31 * cin + cout)
    x = torch.randn(2, rows, cin, generator=g).to(DEV).requires_grad_(True)
    w = torch.randn(2, rows, cout, generator=g).to(DEV)
    y = ops.resize_channels(x, cout)
    ref = x.detach()[..., :cout] if cout <= cin else F.pad(x.detach(), (0, cout - cin))
    assert y.is_contiguous() and torch.equal(y, ref)
    (y * w).sum().backward()
    gref = F.pad(w, (0, cin - cout)) if cout <= cin else w[..., :cin]
    assert torch.equal(x.grad, gref)


@pytest.mark.gpu
@pytest.mark.parametrize("B,na,nb,C", [(2, 2048, 162, 64), (3, 100, 37, 20), (1, 7, 5, 3)])
def test_take_put_rows_bridge_matches_slice_and_cat(B, na, nb, C):
    """The encoder -> processor -> decoder hand-over (models.py:841-842, :865): mesh rows copied out, the processed
    rows written back over them in place.  Values and ALL gradients equal torch's slice + cat formulation, including
    the gradient of the tensor that was overwritten (grid rows from the consumer of the merged tensor, mesh rows
    through the processor path)."""
    from gcl_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + na + C)
    x0 = torch.randn(B, na + nb, C, generator=g).to(DEV)
    Wp = torch.randn(C, C, generator=g).to(DEV)
    wout = torch.randn(B, na + nb, C, generator=g).to(DEV)

    def run(bridge_path):
        x = x0.clone().requires_grad_(True)
        W = Wp.clone().requires_grad_(True)
        enc = x * 1.5                                   # a non-leaf, like the encoder output
        if bridge_path:
            br = ops.RowBridge()
            mesh = ops.take_rows(enc, na, br)
            proc = torch.tanh(mesh @ W)
            merged = ops.put_rows(enc, proc, na, br)
        else:
            mesh = enc[:, na:]
            proc = torch.tanh(mesh @ W)
            merged = torch.cat((enc[:, :na], proc), dim=1)
        out = (merged * wout).sum() + (merged[:, :na] ** 2).sum()
        out.backward()
        return merged.detach().clone(), x.grad.clone(), W.grad.clone()

    m1, gx1, gw1 = run(True)
    m0, gx0, gw0 = run(False)
    assert torch.equal(m1, m0)
    torch.testing.assert_close(gx1, gx0, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(gw1, gw0, rtol=1e-5, atol=1e-5)
    # without a differentiable path through the taken rows the overwritten block gets a zero gradient
    x = x0.clone().requires_grad_(True)
    enc = x * 1.5
    merged = ops.put_rows(enc, torch.zeros(B, nb, C, device=DEV), na)
    (merged * wout).sum().backward()
    assert torch.equal(x.grad[:, :na], 1.5 * wout[:, :na]) and not x.grad[:, na:].any()
    # take_rows on its own is an ordinary differentiable slice
    x = x0.clone().requires_grad_(True)
    (ops.take_rows(x * 1.0, na) * wout[:, na:]).sum().backward()
    assert torch.equal(x.grad[:, na:], wout[:, na:]) and not x.grad[:, :na].any()


@pytest.mark.gpu
@pytest.mark.parametrize("R,dims", [(3000, (72, 48, 48, 64)), (2500, (64, 128, 128, 64)), (300, (30, 48, 48, 33)),
                                    (4100, (66, 96, 96, 96))])
def test_mlp_chain_with_prelu_backward_fused_into_dx(R, dims):
    """Linear+PReLU -> Linear+PReLU -> Linear differentiated through the pre-activations (ops.act_linear): the
    PReLU backward runs in the epilogue of the next layer's dX kernel (tcgen05 path for R >= 2048, two-kernel
    fallback below).  Checked against torch in fp64."""
    from gcl_b200 import ops
    g = torch.Generator().manual_seed(R + dims[1])
    x = torch.randn(R, dims[0], generator=g)
    Ws = [torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5 for i in range(3)]
    bs = [torch.randn(dims[i + 1], generator=g) * 0.1 for i in range(3)]
    sl = [torch.tensor([0.25]), torch.tensor([0.1])]
    w = torch.randn(R, dims[3], generator=g)

    def leaves(dt, dev):
        return [t.detach().clone().to(dev, dt).requires_grad_(True) for t in [x] + Ws + bs + sl]

    L = leaves(torch.float64, "cpu")
    h = torch.nn.functional.prelu(L[0] @ L[1].t() + L[4], L[7])
    h = torch.nn.functional.prelu(h @ L[2].t() + L[5], L[8])
    (((h @ L[3].t() + L[6]) * w.double()).sum()).backward()
    P = leaves(torch.float32, DEV)
    z1, a1 = ops.act_linear(None, P[0], None, P[1], P[4], P[7])
    z2, a2 = ops.act_linear(z1, a1, P[7], P[2], P[5], P[8])
    y = ops.act_linear(z2, a2, P[8], P[3], P[6])
    assert not a1.requires_grad and z1.requires_grad
    ((y * w.to(DEV)).sum()).backward()
    names = ["x", "W1", "W2", "W3", "b1", "b2", "b3", "slope1", "slope2"]
    for n, a, b in zip(names, P, L):
        assert_close(a.grad, b.grad.float(), RTOL_F32, f"d{n}")


@pytest.mark.gpu
@pytest.mark.parametrize("C,B", [(64, 3), (36, 2), (33, 1)])
def test_aggregate_restricted_to_a_row_prefix(C, B):
    """rows_out = n (the decoder keeps only its grid rows, models.py:852): forward equals the full aggregation
    sliced, and the backward -- a [B, n, C] gradient read through the transposed CSR with n_rows_in masking --
    equals the backward of the sliced full result."""
    from gcl_b200 import ops
    from gcl_b200.graph import CSR_LOOPS, NORM_GCN, CSRGraph
    n, keep = 700, 180
    ei = random_graph(n, 5000, seed=C, heavy=60, self_loops=4, dups=5, isolated=15).to(DEV)
    g = CSRGraph(ei, n, CSR_LOOPS)
    gen = torch.Generator().manual_seed(C + B)
    x1 = torch.randn(B, n, C, generator=gen).to(DEV).requires_grad_(True)
    x2 = x1.detach().clone().requires_grad_(True)
    bias = torch.randn(C, generator=gen).to(DEV)
    slope = torch.tensor([0.2], device=DEV)
    w = torch.randn(B, keep, C, generator=gen).to(DEV)
    part = ops.aggregate(x1, g, NORM_GCN, bias, slope, rows_out=keep)
    full = ops.aggregate(x2, g, NORM_GCN, bias, slope)
    assert part.shape == (B, keep, C)
    assert torch.equal(part, full[:, :keep])
    (part * w).sum().backward()
    (full[:, :keep] * w).sum().backward()
    assert_close(x1.grad, x2.grad, 1e-6, "dx through the row-prefix aggregation")


def test_sparse_gat_subclass_and_prune():
    """The reference's SparseGATConv (models.py:112-151) restated on top of OUR GATConv: subclassing,
    super().forward(..., return_attention_weights=True), threshold mask; plus the fused prune kernel."""
    import gcl_b200.nn as gnn
    from gcl_b200 import ops
    from oracle.model import SparseGATConv as OracleSparse

    class SparseGATConv(gnn.GATConv):
        def __init__(self, i, o, heads=1, concat=False, dropout=0.0, bias=True, **kw):
            super().__init__(i, o, heads, concat=concat, dropout=dropout, bias=bias, **kw)

        def forward(self, x, edge_index, attention_threshold=0.0, **kwargs):
            out, (edge_index, att) = super().forward(x, edge_index, return_attention_weights=True)
            att = att.squeeze()
            if kwargs.get("batch_num", 1) == 0:
                mask = att >= attention_threshold
                edge_index, att = edge_index[:, mask], att[mask]
            return out, (edge_index, att)

    n = 500
    ei = random_graph(n, 3000, seed=77)
    torch.manual_seed(3)
    ref = OracleSparse(64, 64, heads=1, concat=False)
    mine = SparseGATConv(64, 64, heads=1, concat=False).to(DEV)
    _copy_params(mine, ref)
    x = torch.randn(n, 64, generator=torch.Generator().manual_seed(8))
    thr = 0.12
    o_c, (e_c, a_c) = ref(x, ei, thr, batch_num=0)
    o_g, (e_g, a_g) = mine(x.to(DEV), ei.to(DEV), thr, batch_num=0)
    assert_close(o_g, o_c, RTOL_F32, "SparseGAT out")
    # alpha values within 1e-4 of the threshold may legitimately flip; compare away from it
    _, (e_full, a_full) = ref(x, ei, thr, batch_num=1)
    safe = (a_full - thr).abs() > 1e-5
    keep_c = (a_full >= thr)
    _, (ei_g_full, a_g_full) = mine(x.to(DEV), ei.to(DEV), thr, batch_num=1)
    keep_g = (a_g_full.cpu() >= thr)
    assert torch.equal(keep_c[safe], keep_g[safe])
    if bool(safe.all()):
        assert torch.equal(e_g.cpu(), e_c)
    pruned = ops.edge_prune(ei_g_full, a_g_full, thr)
    assert torch.equal(pruned, ei_g_full[:, a_g_full >= thr]), "prune kernel vs boolean mask"
    assert pruned.shape[1] < ei_g_full.shape[1]
    # the pruned list (which still contains the self loops) feeds the next call, as in models.py:846
    o2_c, _ = ref(x, e_c, thr)
    o2_g, _ = mine(x.to(DEV), e_g, thr)
    assert_close(o2_g, o2_c, RTOL_F32, "SparseGAT on the pruned graph")


@pytest.mark.parametrize("C", [64, 96, 128, 33, 200])
@pytest.mark.parametrize("affine", [True, False])
def test_layernorm_node(C, affine):
    import gcl_b200.nn as gnn
    onn = _oracle_nn()
    ref = onn.LayerNorm(C, mode="node", affine=affine)
    mine = gnn.LayerNorm(C, mode="node", affine=affine).to(DEV)
    if affine:
        with torch.no_grad():
            ref.weight.uniform_(0.5, 1.5)
            ref.bias.uniform_(-0.5, 0.5)
        _copy_params(mine, ref)
    x = torch.randn(2, 333, C, generator=torch.Generator().manual_seed(C)) * 3 + 1
    _grad_check(mine, ref, x, lambda m, t: m(t), lambda m, t: m(t), what=f"LayerNorm C={C}")


def test_layernorm_graph_mode():
    import gcl_b200.nn as gnn
    onn = _oracle_nn()
    ref, mine = onn.LayerNorm(32, mode="graph"), gnn.LayerNorm(32, mode="graph").to(DEV)
    x = torch.randn(100, 32, generator=torch.Generator().manual_seed(4))
    _grad_check(mine, ref, x, lambda m, t: m(t), lambda m, t: m(t), what="LayerNorm graph")


@pytest.mark.parametrize("R,cin,cout", [(1, 8, 8), (130, 72, 48), (1000, 66, 64), (777, 128, 128), (513, 64, 33),
                                         (300, 30, 48), (257, 96, 15), (64, 200, 260)])
@pytest.mark.parametrize("act", [False, True])
def test_linear_forward_backward(R, cin, cout, act):
    from gcl_b200 import ops
    gen = torch.Generator().manual_seed(R + cin)
    x = torch.randn(R, cin, generator=gen)
    W = torch.randn(cout, cin, generator=gen) / cin ** 0.5
    b = torch.randn(cout, generator=gen)
    a = torch.tensor([0.25])
    go = torch.randn(R, cout, generator=gen)
    tc = [t.clone().requires_grad_(True) for t in (x, W, b, a)]
    yc = torch.nn.functional.linear(tc[0].double(), tc[1].double(), tc[2].double())
    if act:
        yc = torch.nn.functional.prelu(yc, tc[3].double())
    yc.backward(go.double())
    tg = [t.clone().to(DEV).requires_grad_(True) for t in (x, W, b, a)]
    yg = ops.linear(tg[0], tg[1], tg[2], tg[3] if act else None)
    yg.backward(go.to(DEV))
    assert_close(yg, yc.float(), what="linear fwd")
    for i, nme in enumerate(["dx", "dW", "db"] + (["dslope"] if act else [])):
        assert_close(tg[i].grad, tc[i].grad, what=f"linear {nme} ({R},{cin},{cout})")


def test_results_are_deterministic():
    """No floating-point atomics: two runs give identical bits (forward and every gradient)."""
    import gcl_b200.nn as gnn
    n = 3000
    ei = random_graph(n, 20000, seed=5, heavy=500).to(DEV)
    torch.manual_seed(0)
    layers = [gnn.GCNConv(64, 64).to(DEV), gnn.GATConv(64, 64, heads=2, concat=False).to(DEV)]
    x = torch.randn(2, n, 64, device=DEV)
    for layer in layers:
        outs = []
        for _ in range(2):
            layer.zero_grad()
            xx = x.clone().requires_grad_(True)
            y = layer(xx, ei)
            y.square().sum().backward()
            outs.append([y.detach().clone(), xx.grad.clone()] + [p.grad.clone() for p in layer.parameters()])
        for a, b in zip(*outs):
            assert torch.equal(a, b)


def test_error_paths():
    from gcl_b200 import ops
    import gcl_b200.nn as gnn
    x = torch.randn(10, 8, device=DEV)
    with pytest.raises(ValueError):
        gnn.GCNConv(8, 8).to(DEV)(x, torch.tensor([[0, 11], [1, 2]], device=DEV))  # node id out of range
    with pytest.raises(RuntimeError, match="float32"):
        ops.linear(x.double(), torch.randn(8, 8, device=DEV).double())
    with pytest.raises(ValueError):
        ops.linear(x, torch.randn(8, 9, device=DEV))


@pytest.mark.parametrize("R,cin,cout", [(2048, 64, 64), (5000, 128, 128), (3000, 72, 48), (2500, 66, 64), (2500, 64, 33),
                                         (3333, 30, 48), (2100, 96, 15), (40000, 128, 64), (100001, 64, 64), (4099, 96, 96)])
def test_linear_tcgen05_3xtf32_path(R, cin, cout):
    """rows >= 2048 run on the tcgen05 tensor cores in 3xTF32 (hi/lo split, fp32 accumulate in TMEM): forward
    with bias + PReLU + pre-activation copy, dX and dW/db against fp64, and against the FFMA engine."""
    from gcl_b200 import _cabi, ops
    lib = _cabi.load()
    gen = torch.Generator().manual_seed(R + cout)
    x = torch.randn(R, cin, generator=gen).to(DEV)
    W = (torch.randn(cout, cin, generator=gen) / cin ** 0.5).to(DEV)
    b = torch.randn(cout, generator=gen).to(DEV)
    a = torch.tensor([0.25], device=DEV)
    dy = torch.randn(R, cout, generator=gen).to(DEV)
    pre = torch.nn.functional.linear(x.double(), W.double(), b.double())
    want = dict(y=torch.where(pre > 0, pre, 0.25 * pre), z=pre, dx=dy.double() @ W.double(),
                dW=dy.double().T @ x.double(), db=dy.double().sum(0))
    got = {}
    try:
        for mode, tag in ((0, "tc"), (1, "ffma")):
            assert lib.gcl_set_dense_mode(mode) == 0
            y, z = ops.linear_fwd_raw(x, W, b, a, True)
            dW, db = ops.linear_bwd_dw_raw(dy, x, True)
            got[tag] = dict(y=y, z=z, dx=ops.linear_bwd_dx_raw(dy, W), dW=dW, db=db)
    finally:
        lib.gcl_set_dense_mode(0)
    for k, ref in want.items():
        assert_close(got["tc"][k], ref.float(), 2e-5, f"tcgen05 {k} ({R},{cin},{cout})")
        assert_close(got["ffma"][k], ref.float(), 2e-5, f"ffma {k} ({R},{cin},{cout})")
    # deterministic: same bits on a second call
    y2, _ = ops.linear_fwd_raw(x, W, b, a, True)
    assert torch.equal(y2, got["tc"]["y"])
