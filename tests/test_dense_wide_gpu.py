"""GPU parity of the wide-layer dense kernels (weights / dY^T in tensor memory: umma_linear_ts_kernel,
umma_dw_ts_kernel in csrc/umma_gemm.cu) through the C ABI, against torch in fp64, and of the chains that are
differentiated through pre-activations (ops.aggregate_pre + ops.act_linear with the bias gradient summed in the
dX epilogue) against the unfused formulation.

Bars: the 3xTF32 products keep ~22 mantissa bits, so forward / dX results are held to 2e-6 of the result's scale
(max-norm relative; BASELINE.json's bar is 1e-4), row reductions over up to 2e5 rows (dW, bias gradients) to 5e-6."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (rows, c_in, c_out): tiles with row tails, partial K blocks (K % 32 != 0), channel tails (c_out % 32 != 0),
# fewer rows than one tile, the 32..64-channel variant that stores the pre-activation, narrow K (FFMA / row-major path)
SHAPES = [(2048, 128, 128), (2111, 128, 128), (12345, 96, 96), (2777, 64, 128), (5000, 128, 96), (2300, 32, 72),
          (4097, 128, 100), (3000, 64, 64), (5000, 128, 64), (2100, 64, 32), (2500, 44, 128), (50000, 20, 128),
          (200000, 128, 128), (2049, 36, 68), (2100, 100, 124), (2048, 128, 68), (3333, 40, 128), (9473, 128, 72)]


def _relerr(a, ref):
    return float((a.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("R,cin,cout", SHAPES)
def test_wide_linear_kernels_against_fp64(R, cin, cout):
    from gcl_b200 import ops
    g = torch.Generator().manual_seed(R + cin + cout)
    x = torch.randn(R, cin, generator=g).to(DEV)
    W = (torch.randn(cout, cin, generator=g) / cin ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    slope = torch.tensor([0.25], device=DEV)
    dy = torch.randn(R, cout, generator=g).to(DEV)
    zin = torch.randn(R, cin, generator=g).to(DEV)
    xd, Wd, bd, dyd, zd = x.double(), W.double(), b.double(), dy.double(), zin.double()
    ref_z = xd @ Wd.t() + bd
    ref_y = torch.where(ref_z > 0, ref_z, 0.25 * ref_z)
    y, z = ops.linear_fwd_raw(x, W, b, slope, want_z=True)                 # + bias, PReLU, pre-activation copy
    assert _relerr(z, ref_z) < 2e-6 and _relerr(y, ref_y) < 2e-6
    y1, _ = ops.linear_fwd_raw(x, W, b, slope, want_z=False)
    assert _relerr(y1, y.double()) < 1e-6, "the variant without the pre-activation copy"
    y2, _ = ops.linear_fwd_raw(x, W, None, None, want_z=False)
    assert _relerr(y2, xd @ Wd.t()) < 2e-6
    ref_dx = dyd @ Wd
    dx = ops.linear_bwd_dx_raw(dy, W)
    assert _relerr(dx, ref_dx) < 2e-6
    dz, dsl, dcs = ops.linear_bwd_dx_prelu_raw(dy, W, zin, slope, True)    # dX * PReLU'(z_in), dslope, column sums
    ref_dz = torch.where(zd > 0, ref_dx, 0.25 * ref_dx)
    ref_dsl = (ref_dx * torch.where(zd > 0, torch.zeros_like(zd), zd)).sum()
    assert _relerr(dz, ref_dz) < 2e-6
    assert abs(float(dsl) - float(ref_dsl)) <= 1e-5 * float((ref_dx * zd.clamp(max=0)).abs().sum())
    assert _relerr(dcs, ref_dz.sum(0)) < 5e-6
    dz2, dsl2 = ops.linear_bwd_dx_prelu_raw(dy, W, zin, slope)
    # (for <= 64 channels the variant with column sums runs on the weights-stationary kernel, the one without on the
    # row-major kernel: same three products, different accumulation order)
    assert _relerr(dz2, dz.double()) < 1e-6 and abs(float(dsl2) - float(dsl)) <= 1e-5 * float((ref_dx * zd.clamp(max=0)).abs().sum())
    dW, db = ops.linear_bwd_dw_raw(dy, x, True)
    assert _relerr(dW, dyd.t() @ xd) < 5e-6 and _relerr(db, dyd.sum(0)) < 5e-6
    # deterministic: same bits on a second call
    dW2, db2 = ops.linear_bwd_dw_raw(dy, x, True)
    assert torch.equal(dW, dW2) and torch.equal(db, db2)
    assert torch.equal(ops.linear_fwd_raw(x, W, b, slope, want_z=True)[0], y)


@pytest.mark.parametrize("R,cin,cout", [(2111, 128, 128), (12345, 96, 96), (3000, 64, 64)])
def test_wide_kernels_agree_with_ffma_engine(R, cin, cout):
    """tcgen05 3xTF32 (default) against the CUDA-core fp32 FFMA kernels on the same inputs."""
    from gcl_b200 import _cabi, ops
    lib = _cabi.load()
    g = torch.Generator().manual_seed(7 * R + cout)
    x = torch.randn(R, cin, generator=g).to(DEV)
    W = (torch.randn(cout, cin, generator=g) / cin ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    dy = torch.randn(R, cout, generator=g).to(DEV)
    slope = torch.tensor([0.1], device=DEV)

    def run():
        y, z = ops.linear_fwd_raw(x, W, b, slope, want_z=True)
        return y, z, ops.linear_bwd_dx_raw(dy, W), *ops.linear_bwd_dw_raw(dy, x, True)

    assert lib.gcl_get_dense_mode() == 0
    a = run()
    try:
        _cabi.check(lib.gcl_set_dense_mode(1), "gcl_set_dense_mode")
        f = run()
    finally:
        _cabi.check(lib.gcl_set_dense_mode(0), "gcl_set_dense_mode")
    for name, p, q in zip(("y", "z", "dx", "dW", "dbias"), a, f):
        assert _relerr(p, q.double()) < 1e-5, name


@pytest.mark.parametrize("C", [128, 96, 64])
def test_gcn_chain_through_pre_activation_matches_unfused(C):
    """GCNConv -> PReLU -> GCNConv's Linear: (z, a) from aggregate_pre + act_linear (PReLU', slope and bias gradients
    in the dX epilogue) against aggregate(+PReLU) + linear, values bit-identical, gradients to 1e-5."""
    from gcl_b200 import ops
    from gcl_b200.graph import CSR_LOOPS, NORM_GCN, CSRGraph
    from helpers import random_graph
    n, B = 1500, 2                                     # B * n rows >= 2048: the tensor-core kernels
    ei = random_graph(n, e=6000, seed=11, self_loops=4).to(DEV)
    gph = CSRGraph(ei, n, CSR_LOOPS)
    g = torch.Generator().manual_seed(C)
    h0 = torch.randn(B, n, C, generator=g).to(DEV)
    b0 = torch.randn(C, generator=g).to(DEV)
    s0 = torch.tensor([0.25], device=DEV)
    W0 = (torch.randn(C, C, generator=g) / C ** 0.5).to(DEV)
    wout = torch.randn(B, n, C, generator=g).to(DEV)

    def run(chain):
        h, b, s, W = (t.clone().requires_grad_(True) for t in (h0, b0, s0, W0))
        if chain:
            sink = ops.ColsumSink()
            z, a = ops.aggregate_pre(h, gph, NORM_GCN, b, s, sink)
            y = ops.act_linear(z, a, s, W, sink=sink)
        else:
            y = ops.linear(ops.aggregate(h, gph, NORM_GCN, b, s), W)
        (y * wout).sum().backward()
        return y.detach(), h.grad, b.grad, s.grad, W.grad

    got, ref = run(True), run(False)
    assert torch.equal(got[0], ref[0])
    for name, p, q in zip(("dh", "dbias", "dslope", "dW"), got[1:], ref[1:]):
        assert _relerr(p, q.double()) < 1e-5, name
