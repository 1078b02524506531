"""f3 on the GPU: the InteractionNet processor (models.py:166-285) and what the v2 configs add -- ReLU / SiLU
activations, LayerNorm(mode="graph"), 4-d mesh edge features -- through the C ABI against the oracle (pinned bit-exact
to the unmodified reference in tests/test_oracle.py::test_interaction_net_glue_matches_unmodified_reference)."""
import numpy as np
import pytest
import torch

from helpers import RTOL_F32, assert_close
from test_oracle import _small_inet_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("kind,ref", [(1, torch.nn.functional.relu), (2, torch.nn.functional.silu)])
def test_activations(kind, ref):
    from gcl_b200 import ops
    x = torch.randn(3, 777, 20, generator=torch.Generator().manual_seed(kind))
    xg, xc = x.clone().to(DEV).requires_grad_(True), x.clone().requires_grad_(True)
    yg, yc = ops.act(xg, kind), ref(xc)
    assert_close(yg, yc, 1e-6, "act fwd")
    go = torch.randn(x.shape, generator=torch.Generator().manual_seed(9))
    yg.backward(go.to(DEV))
    yc.backward(go)
    assert_close(xg.grad, xc.grad, 1e-5, "act bwd")


@pytest.mark.parametrize("shape,affine", [((900, 64), True), ((3, 500, 32), True), ((2, 1200, 20), False)])
def test_layernorm_graph_mode(shape, affine):
    """PyG LayerNorm(mode='graph'): (x - mean) / (std + eps) * w + b over a whole sample; batched = per sample."""
    import torch_geometric.nn as onn
    import gcl_b200.nn as gnn
    C = shape[-1]
    gen = torch.Generator().manual_seed(C)
    x = torch.randn(shape, generator=gen) * 3 + 1.5
    lo, lg = onn.LayerNorm(C, affine=affine, mode="graph"), gnn.LayerNorm(C, affine=affine, mode="graph").to(DEV)
    if affine:
        with torch.no_grad():
            lo.weight.copy_(torch.randn(C, generator=gen))
            lo.bias.copy_(torch.randn(C, generator=gen))
        lg.load_state_dict(lo.state_dict())
    xg, xc = x.clone().to(DEV).requires_grad_(True), x.clone().requires_grad_(True)
    yg = lg(xg)
    yc = lo(xc) if x.dim() == 2 else torch.stack([lo(s) for s in xc])
    assert_close(yg, yc, RTOL_F32, "LN graph fwd")
    go = torch.randn(shape, generator=gen)
    yg.backward(go.to(DEV))
    yc.backward(go)
    assert_close(xg.grad, xc.grad, RTOL_F32, "LN graph dx")
    if affine:
        assert_close(lg.weight.grad, lo.weight.grad, RTOL_F32, "LN graph dweight")
        assert_close(lg.bias.grad, lo.bias.grad, RTOL_F32, "LN graph dbias")


@pytest.mark.parametrize("C", [64, 256, 12])
def test_edge_gather_and_mean_reduce(C):
    """EdgeOps: x[senders], x[receivers], scatter(..., receivers, reduce='mean') and their backwards on the SpMM
    kernels (tiled for C <= 128, row-gather beyond), bit-exact against index ops (gathers) / 1e-6 (mean)."""
    from gcl_b200 import ops
    from gcl_b200.graph import EdgeOps
    from helpers import random_graph
    n = 700
    ei = random_graph(n, 5000, seed=C, isolated=30, dups=10)
    E = ei.shape[1]
    eo = EdgeOps(ei.to(DEV), n)
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(2, n, C, generator=gen)
    m = torch.randn(2, E, C, generator=gen)
    xg, mg = x.to(DEV).requires_grad_(True), m.to(DEV).requires_grad_(True)
    xs, xr = ops.spmm_fixed(xg, *eo.gather_src), ops.spmm_fixed(xg, *eo.gather_dst)
    assert torch.equal(xs.cpu(), x[:, ei[0]]) and torch.equal(xr.cpu(), x[:, ei[1]])
    agg = ops.spmm_fixed(mg, *eo.mean_dst)
    cnt = torch.bincount(ei[1], minlength=n).clamp(min=1).float()
    want = torch.zeros(2, n, C).index_add_(1, ei[1], m) / cnt.view(1, -1, 1)
    assert_close(agg, want, 1e-6, "scatter mean")
    go_e, go_n = torch.randn(2, E, C, generator=gen), torch.randn(2, n, C, generator=gen)
    (xs * go_e.to(DEV)).sum().backward()
    assert_close(xg.grad, torch.zeros(2, n, C).index_add_(1, ei[0], go_e), 1e-6, "gather backward = scatter add")
    (agg * go_n.to(DEV)).sum().backward()
    assert_close(mg.grad, (go_n / cnt.view(1, -1, 1))[:, ei[1]], 1e-6, "mean backward = weighted gather")


def test_mesh_edge_features_equal_oracle():
    from gcl_b200.graphs_build import ModelGraphs
    from oracle import graphs as og
    for args in ((16, 32, [1, 3], 0.6), (32, 64, [3, 5], 0.5)):
        g = ModelGraphs(*args, DEV)
        o = og.build_graphs(*args)
        assert np.array_equal(g.processing_edge_features.cpu().numpy(), o["mesh_edge_feats"])


@pytest.mark.parametrize("act,width,B", [("swish", 32, 1), ("swish", 64, 3), ("relu", 32, 2), ("prelu", 32, 2)])
def test_interaction_net_model_forward_and_gradients(act, width, B):
    """Whole v2-style model (GCN encoder / decoder with the config's activation, InteractionNet processor with edge
    features, residuals, graph- and node-LayerNorm): forecast step, loss and every gradient vs the oracle, B samples."""
    from gcl_b200.model import WeatherPrediction
    from gcl_b200.train import Trainer
    from oracle import graphs as og, model as om
    cfg = _small_inet_cfg(width=width, steps=2, act=act)
    nlat, nlon = 16, 32
    torch.manual_seed(3)
    ref = om.WeatherPrediction(cfg, nlat, nlon, graphs=og.build_graphs(nlat, nlon, [1, 3], 0.6))
    mine = WeatherPrediction(cfg, nlat, nlon, DEV)
    assert sorted(mine.state_dict()) == sorted(ref.state_dict())
    mine.load_state_dict(ref.state_dict())
    G, F, T = nlat * nlon, 19, 2
    gen = torch.Generator().manual_seed(4)
    X, y = torch.randn(B, G, T * F, generator=gen), torch.randn(B, G, F, generator=gen)
    out_g, out_c = mine(X=X.to(DEV)), ref(X=X)
    assert_close(out_g, out_c, RTOL_F32, f"InteractionNet forecast step ({act})")
    tr = Trainer(mine, nlat, nlon, ar_steps=1)
    tr.zero_grad()
    lg = tr.loss(X.to(DEV), y.to(DEV))
    lg.backward()
    lc = om.training_loss(ref, X, y, 1, om.lat_weights(nlat, nlon))
    lc.backward()
    assert abs(float(lg.detach()) - float(lc.detach())) <= RTOL_F32 * abs(float(lc.detach()))
    pr = dict(ref.named_parameters())
    for k, p in mine.named_parameters():
        if pr[k].grad is None:             # the last step's edge update feeds nothing (models.py:281-284)
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, k
        assert_close(p.grad, pr[k].grad, 2e-4 if act == "prelu" else RTOL_F32, f"d{k}", atol=1e-9)
