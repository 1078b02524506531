"""GPU parity of the whole path: graph construction (bit-exact edge sets), one forecast step and its
gradients for the five BASELINE workloads, the training step (AR rollout, loss, Adam), and the
import-swap of the reference's model glue onto the gcl_b200 layers.  Checker = the CPU oracle and the
fixtures written by the unmodified reference (tests/golden)."""
import hashlib
import importlib.util
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

from helpers import RTOL_F32, assert_close, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALL = ["baseline", "attention", "sparse_attention", "wb2_64x32_ar_15f_4obs_4pred", "wb2_512x256_19f_ar"]


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _sorted_edges(ei):
    return np.ascontiguousarray(ei[:, np.lexsort((ei[1], ei[0]))])


_GRAPHS = {}


def product_graphs(nlat, nlon, levels, rq):
    from gcl_b200.graphs_build import ModelGraphs
    key = (nlat, nlon, tuple(levels), rq)
    if key not in _GRAPHS:
        _GRAPHS[key] = ModelGraphs(nlat, nlon, levels, rq, DEV)
    return _GRAPHS[key]


_ORACLE_GRAPHS = {}


def oracle_graphs(nlat, nlon, levels, rq):
    from oracle import graphs as og
    key = (nlat, nlon, tuple(levels), rq)
    if key not in _ORACLE_GRAPHS:
        _ORACLE_GRAPHS[key] = og.build_graphs(nlat, nlon, levels, rq)
    return _ORACLE_GRAPHS[key]


@pytest.mark.parametrize("name", ["notebook_level0", "small_32x16_l13_rq0.6", "64x32_l35_rq0.5", "64x32_l46_rq0.65",
                                  "512x256_l46_rq0.6"])
def test_graph_construction_bit_exact(name, golden_dir):
    """Device radius query / closest face + vectorised mesh == oracle == unmodified reference (digests)."""
    gd = json.load(open(os.path.join(golden_dir, "graphs.json")))["graphs"][name]
    args = (gd["nlat"], gd["nlon"], gd["mesh_levels"], gd["radius_factor"])
    g = product_graphs(*args)
    o = oracle_graphs(*args)
    assert np.array_equal(g.mesh_vertices, o["mesh_vertices"]), "mesh vertices (float32) must be bit-identical"
    assert np.array_equal(g.finest_faces, o["finest_faces"])
    g2m, mesh, m2g = (t.cpu().numpy() for t in (g.encoding_graph, g.processing_graph, g.decoding_graph))
    assert np.array_equal(_sorted_edges(g2m), _sorted_edges(o["g2m"])), "grid->mesh edge SET differs"
    assert np.array_equal(g2m, _sorted_edges(g2m)), "grid->mesh must come out (sender, receiver)-sorted"
    assert np.array_equal(mesh, o["mesh"]), "mesh edge list (order included)"
    assert np.array_equal(m2g, o["m2g"]), "mesh->grid edge list (order included)"
    assert np.array_equal(g.init_grid_features.cpu().numpy(), o["grid_feats"])
    assert np.array_equal(g.init_mesh_features.cpu().numpy(), o["mesh_feats"])
    # and the reference-made fixture
    assert (g2m.shape[1], mesh.shape[1], m2g.shape[1]) == (gd["E_g2m"], gd["E_mesh"], gd["E_m2g"])
    assert _digest(mesh) == gd["mesh_sha256"]
    if _digest(g.mesh_vertices) == gd["vertices_sha256"]:      # same BLAS/libm bits as the fixture's machine
        assert _digest(_sorted_edges(g2m)) == gd["g2m_sorted_sha256"]
        assert _digest(m2g) == gd["m2g_sha256"]


def _models(name, nlat=None, nlon=None, levels=None, rq=None, seed=0):
    """(product model on GPU, oracle model on CPU) with identical, randomised weights."""
    from gcl_b200.model import WeatherPrediction
    from gcl_b200.workloads import get_workload
    from oracle import model as om
    cfg = get_workload(name)
    nlat, nlon = nlat or cfg["nlat"], nlon or cfg["nlon"]
    if levels is not None:
        cfg["graph"]["mesh_levels"], cfg["graph"]["grid2mesh_radius_query"] = levels, rq
    gargs = (nlat, nlon, cfg["graph"]["mesh_levels"], cfg["graph"]["grid2mesh_radius_query"])
    torch.manual_seed(seed)
    ref = om.WeatherPrediction(cfg, nlat, nlon, graphs=oracle_graphs(*gargs))
    with torch.no_grad():
        gen = torch.Generator().manual_seed(seed + 1)
        for p in ref.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
    mine = WeatherPrediction(cfg, nlat, nlon, DEV, graphs=product_graphs(*gargs))
    mine.load_state_dict(ref.state_dict())
    return mine, ref, cfg, nlat, nlon


def _check_grads(mine, ref, tol, what):
    pr = dict(ref.named_parameters())
    n = 0
    for k, p in mine.named_parameters():
        if pr[k].grad is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, f"{what}: unexpected grad for {k}"
            continue
        assert p.grad is not None, f"{what}: missing grad for {k}"
        # d att_dst is analytically ~0 (softmax is invariant to a per-receiver shift, only LeakyReLU breaks
        # it), i.e. a cancelling sum: judge it on an absolute floor relative to d att_src
        extra = 1e-4 * float(pr[k.replace("att_dst", "att_src")].grad.abs().max()) if k.endswith("att_dst") else 0.0
        assert_close(p.grad, pr[k].grad, tol, f"{what} d{k}", atol=max(1e-9, extra))
        n += 1
    assert n > 0


def _set_engine(engine):
    from gcl_b200 import _cabi
    _cabi.load().gcl_set_dense_mode(1 if engine == "ffma" else 0)


@pytest.fixture(autouse=True)
def _restore_dense_engine():
    yield
    _set_engine("tcgen05")


@pytest.mark.parametrize("engine", ["ffma", "tcgen05"])
@pytest.mark.parametrize("name", ALL)
def test_forecast_step_and_gradients_full_size(name, engine):
    """One forecast step (fwd), lat-weighted loss and all gradients at the BASELINE sizes, B = 1.

    engine = ffma   : CUDA-core dense kernels; same fp32 operation mix as the reference, so outputs AND
                      gradients must match the fp32 oracle within rel 1e-4.
    engine = tcgen05: 3xTF32 tensor-core dense kernels (the default).  The forecast step must match the fp32
                      oracle within rel 1e-4.  For the gradients the fp32 oracle is itself 1e-4..2e-3 away
                      from the fp64 result on these models (PReLU kinks, long reductions -- measured, see
                      DESIGN.md "Parity"), and a different-but-equally-accurate fp32 evaluation lands
                      elsewhere inside that band; so each gradient is compared with the fp64 oracle and must
                      be within max(1e-4, 3 x the fp32 oracle's own distance from fp64)."""
    import copy
    from gcl_b200.train import Trainer
    from oracle import model as om
    _set_engine(engine)
    mine, ref, cfg, nlat, nlon = _models(name)
    G = nlat * nlon
    F, T = cfg["data"]["num_features_used"], cfg["data"]["obs_window_used"]
    gen = torch.Generator().manual_seed(3)
    X, y = torch.randn(1, G, T * F, generator=gen), torch.randn(1, G, F, generator=gen)
    kw = dict(batch_num=1) if name == "sparse_attention" else {}
    out_c = ref(X=X, attention_threshold=0.0, **kw)
    out_g = mine(X=X.to(DEV), attention_threshold=0.0, **kw)
    assert out_g.shape == out_c.shape == (G, F)
    assert_close(out_g, out_c, RTOL_F32, f"{name} forecast step")
    tr = Trainer(mine, nlat, nlon, lr=cfg["learning_rate"], ar_steps=1)
    tr.zero_grad()
    lg = tr.loss(X.to(DEV), y.to(DEV), 0.0, **kw)
    lg.backward()
    lw = om.lat_weights(nlat, nlon)
    lc = om.training_loss(ref, X, y, 1, lw, **kw)
    lc.backward()
    assert abs(float(lg.detach()) - float(lc.detach())) <= RTOL_F32 * abs(float(lc.detach()))
    if engine == "ffma":
        _check_grads(mine, ref, RTOL_F32, name)
        return
    ref64 = copy.deepcopy(ref).double()
    ref64.zero_grad()
    ref64.init_grid_features, ref64.init_mesh_features = ref64.init_grid_features.double(), ref64.init_mesh_features.double()
    om.training_loss(ref64, X.double(), y.double(), 1, lw.double(), **kw).backward()
    p32, p64 = dict(ref.named_parameters()), dict(ref64.named_parameters())
    checked = 0
    for k, p in mine.named_parameters():
        if p64[k].grad is None:
            continue
        truth = p64[k].grad.float()
        ref_err = rel_err(p32[k].grad, truth)
        got_err = rel_err(p.grad, truth)
        if float((p.grad.detach().cpu() - truth).abs().max()) <= 1e-9:
            continue
        assert got_err <= max(RTOL_F32, 3 * ref_err), f"{name} d{k}: {got_err:.2e} vs fp64 (fp32 oracle: {ref_err:.2e})"
        checked += 1
    assert checked > 0


def _kink_free_inputs(ref, make, run, floor, seeds):
    """Inputs whose smallest |PReLU input| anywhere in the oracle's forward is > floor: a pre-activation within fp32
    rounding of 0 may take different PReLU branches in two equally accurate evaluations and then moves a weight
    gradient by ~1/sqrt(rows) -- not a kernel error.  Returns the first such seed's inputs."""
    seen = []
    hooks = [m.register_forward_hook(lambda mod, inp, out: seen.append(float(inp[0].detach().abs().min())))
             for m in ref.modules() if isinstance(m, torch.nn.PReLU)]
    try:
        best = None
        for seed in seeds:
            data = make(seed)
            seen.clear()
            with torch.no_grad():
                run(*data)
            if best is None or min(seen) > best[0]:
                best = (min(seen), data)
            if min(seen) > floor:
                break
    finally:
        for h in hooks:
            h.remove()
    return best


@pytest.mark.parametrize("name", ALL)
def test_tcgen05_engine_gradients_within_1e4_of_fp32_oracle_on_kink_free_inputs(name):
    """The DEFAULT dense engine (tcgen05 3xTF32) held to the plain north-star bar -- output, loss and EVERY gradient
    within rel 1e-4 of the fp32 oracle -- for the architectures of all five BASELINE configs (their widths, layer
    types and depths; 32x16 grid, mesh [1,3], B = 2 samples so that every dense layer has >= 2048 rows and takes the
    tensor-core kernels), on inputs chosen to keep every PReLU input away from its kink."""
    from gcl_b200 import _cabi
    from gcl_b200.train import Trainer
    from oracle import model as om
    _set_engine("tcgen05")
    assert _cabi.load().gcl_get_dense_mode() == 0
    kw = dict(batch_num=1) if name == "sparse_attention" else {}
    B = 2
    # the mesh rows of the encoder input are constants (static features), so some PReLU inputs do not depend on the
    # data seed: the weight seed is part of the search
    for wseed in range(11, 19):
        mine, ref, cfg, nlat, nlon = _models(name, 16, 32, [1, 3], 0.6, seed=wseed)
        G = nlat * nlon
        F, T = cfg["data"]["num_features_used"], cfg["data"]["obs_window_used"]
        assert B * (G + mine._num_mesh_nodes) >= 2048
        lw = om.lat_weights(nlat, nlon)

        def make(seed):
            gen = torch.Generator().manual_seed(seed)
            return torch.randn(B, G, T * F, generator=gen), torch.randn(B, G, F, generator=gen)
        floor, (X, y) = _kink_free_inputs(ref, make, lambda X, y: om.training_loss(ref, X, y, 1, lw, **kw), 5e-7,
                                          range(100, 140))
        if floor > 1e-7:
            break
    assert floor > 1e-7, f"no kink-free input found (best floor {floor:.1e})"
    out_g = mine(X=X.to(DEV), attention_threshold=0.0, **kw)
    out_c = ref(X=X, attention_threshold=0.0, **kw)
    assert_close(out_g, out_c, RTOL_F32, f"{name} forecast step (tcgen05)")
    tr = Trainer(mine, nlat, nlon, lr=cfg["learning_rate"], ar_steps=1)
    tr.zero_grad()
    lg = tr.loss(X.to(DEV), y.to(DEV), 0.0, **kw)
    lg.backward()
    lc = om.training_loss(ref, X, y, 1, lw, **kw)
    lc.backward()
    assert abs(float(lg.detach()) - float(lc.detach())) <= RTOL_F32 * abs(float(lc.detach()))
    _check_grads(mine, ref, RTOL_F32, f"{name} (tcgen05, kink-free)")


@pytest.mark.parametrize("name", ["baseline", "attention", "sparse_attention"])
def test_against_unmodified_reference_fixture(name, golden_dir):
    """Product vs what the UNMODIFIED reference computed (tests/golden/model_*.npz)."""
    from gcl_b200.model import WeatherPrediction
    from gcl_b200.train import Trainer
    from gcl_b200.workloads import get_workload
    z = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    cfg = get_workload(name)
    cfg["graph"]["mesh_levels"], cfg["graph"]["grid2mesh_radius_query"] = meta["mesh_levels"], meta["radius_factor"]
    m = WeatherPrediction(cfg, meta["nlat"], meta["nlon"], DEV)
    m.load_state_dict({k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")})
    X, y = torch.from_numpy(z["X"]).to(DEV), torch.from_numpy(z["y"]).to(DEV)
    kw = dict(batch_num=1) if name == "sparse_attention" else {}
    assert_close(m(X=X, attention_threshold=0.0, **kw), torch.from_numpy(z["delta"]), RTOL_F32, "delta")
    tr = Trainer(m, meta["nlat"], meta["nlon"], ar_steps=1)
    tr.zero_grad()
    loss = tr.loss(X, y, 0.0, **kw)
    loss.backward()
    assert abs(float(loss) - float(z["loss"])) <= RTOL_F32 * abs(float(z["loss"]))
    for k, p in m.named_parameters():
        if "grad/" + k in z.files:
            extra = 1e-4 * float(np.abs(z["grad/" + k.replace("att_dst", "att_src")]).max()) if k.endswith("att_dst") else 0.0
            assert_close(p.grad, torch.from_numpy(z["grad/" + k]), RTOL_F32, f"d{k}", atol=extra)
    if name == "sparse_attention":
        with torch.no_grad():
            m(X=X, attention_threshold=0.05, batch_num=0)
        got, want = m.processing_graph.cpu().numpy(), z["pruned_edge_index"]
        # attention values within float noise of the threshold may flip; everything else must agree
        a = set(map(tuple, got.T.tolist()))
        b = set(map(tuple, want.T.tolist()))
        assert len(a ^ b) <= 2, f"pruned edge sets differ in {len(a ^ b)} edges"
        assert want.shape[1] < 4080 + 642


def test_batched_equals_per_sample_and_ar_rollout():
    """[B,G,C] inputs == B independent batch-1 runs; AR = 2 rollout with BPTT matches the oracle."""
    from gcl_b200.train import Trainer
    from oracle import model as om
    mine, ref, cfg, nlat, nlon = _models("wb2_64x32_ar_15f_4obs_4pred", 16, 32, [1, 3], 0.6)
    G, F, T = nlat * nlon, 15, 4
    # PReLU has a kink at 0: a pre-activation within fp32 rounding of 0 can take different branches on
    # the two sides and moves the (tiny) BPTT gradients by a few %, which is not a kernel error (seen
    # with seed 9: every kernel call matched its reference, one |z| ~ 1e-8 flipped).  Pick inputs whose
    # smallest |PReLU input| over both rollout steps is well away from 0.
    prelu_min = []
    hooks = [m.register_forward_hook(lambda mod, inp, out: prelu_min.append(float(inp[0].detach().abs().min())))
             for m in ref.modules() if isinstance(m, torch.nn.PReLU)]
    for seed in range(10, 40):
        gen = torch.Generator().manual_seed(seed)
        X, y = torch.randn(3, G, T * F, generator=gen), torch.randn(3, G, 4 * F, generator=gen)
        prelu_min.clear()
        with torch.no_grad():
            om.training_loss(ref, X, y, 2, om.lat_weights(nlat, nlon))
        if min(prelu_min) > 2e-6:
            break
    for h in hooks:
        h.remove()
    out_b = mine(X=X.to(DEV))
    assert out_b.shape == (3, G, F)
    for b in range(3):
        assert_close(out_b[b], mine(X=X[b:b + 1].to(DEV)), 2e-5, "batched vs single")   # FFMA (< 2048 rows) vs tcgen05 engine
    assert_close(out_b, ref(X=X), RTOL_F32, "batched vs oracle")
    tr = Trainer(mine, nlat, nlon, ar_steps=2)
    tr.zero_grad()
    lg = tr.loss(X.to(DEV), y.to(DEV))
    lg.backward()
    lc = om.training_loss(ref, X, y, 2, om.lat_weights(nlat, nlon))
    lc.backward()
    assert abs(float(lg) - float(lc)) <= RTOL_F32 * abs(float(lc))
    _check_grads(mine, ref, RTOL_F32, "AR=2")


def test_training_steps_match_torch_adam():
    """Three optimiser steps (flat buffers + gcl_adam_f32) track torch.optim.Adam on the oracle."""
    from gcl_b200.train import Trainer
    from oracle import model as om
    mine, ref, cfg, nlat, nlon = _models("baseline", 16, 32, [1, 3], 0.6)
    G, F, T = nlat * nlon, 33, 2
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    tr = Trainer(mine, nlat, nlon, lr=1e-3, ar_steps=1)
    lw = om.lat_weights(nlat, nlon)
    gen = torch.Generator().manual_seed(4)
    for it in range(3):
        X, y = torch.randn(2, G, T * F, generator=gen), torch.randn(2, G, F, generator=gen)
        opt.zero_grad()
        lc = om.training_loss(ref, X, y, 1, lw)
        lc.backward()
        opt.step()
        lg = tr.step(X.to(DEV), y.to(DEV))
        assert abs(float(lg) - float(lc)) <= 2e-4 * abs(float(lc)), (it, float(lg), float(lc))
    pr = dict(ref.named_parameters())
    for k, p in mine.named_parameters():
        assert_close(p, pr[k], 2e-3, f"param {k} after 3 Adam steps")   # Adam's m/sqrt(v) amplifies fp noise
    assert int(tr.step_count.item()) == 3


def test_sparse_gat_prunes_and_reuses_graph():
    from oracle import model as om  # noqa: F401
    mine, ref, cfg, nlat, nlon = _models("sparse_attention", 16, 32, [1, 3], 0.6)
    G = nlat * nlon
    X = torch.randn(1, G, 24, generator=torch.Generator().manual_seed(2))
    e0 = mine.processing_graph.shape[1]
    with torch.no_grad():
        o1 = mine(X=X.to(DEV), attention_threshold=0.1, batch_num=0)
        e1 = mine.processing_graph.shape[1]
        c1 = ref(X=X, attention_threshold=0.1, batch_num=0)
        assert_close(o1, c1, RTOL_F32, "output of the pruning call (computed on the unpruned graph)")
        assert e1 < e0 + mine._num_mesh_nodes and abs(e1 - ref.processing_graph.shape[1]) <= 2
        g_before = mine.processing_graph
        o2 = mine(X=X.to(DEV), attention_threshold=0.1, batch_num=5)
        c2 = ref(X=X, attention_threshold=0.1, batch_num=5)
        assert_close(o2, c2, 5e-4, "forward on the pruned graph")
        g_after2 = mine.processing_graph
        assert abs(g_after2.shape[1] - ref.processing_graph.shape[1]) <= 2    # pruned self loops are re-added (PyG)
        o3 = mine(X=X.to(DEV), attention_threshold=0.1, batch_num=6)
        assert torch.equal(o2, o3)
        # non-pruning calls leave the edge set alone, and the edge list they hand back is the cached CSR's own
        # PyG-order view, so the next call reuses that CSR instead of rebuilding it
        from gcl_b200.graph import CSR_LOOPS, GLOBAL_CACHE
        assert mine.processing_graph is g_after2 and g_before is not g_after2
        csr = GLOBAL_CACHE.get(mine.processing_graph, mine._num_mesh_nodes, CSR_LOOPS)
        assert mine.processing_graph is csr.edge_index_with_loops


def test_cuda_graph_capture_of_forward_backward():
    """The whole step is enqueue-only and graph-capturable: a replay reproduces the eager result."""
    from gcl_b200.train import Trainer
    mine, ref, cfg, nlat, nlon = _models("attention", 16, 32, [1, 3], 0.6)
    G, F, T = nlat * nlon, 33, 2
    gen = torch.Generator().manual_seed(6)
    X, y = torch.randn(2, G, T * F, generator=gen).to(DEV), torch.randn(2, G, F, generator=gen).to(DEV)
    tr = Trainer(mine, nlat, nlon, ar_steps=1)
    tr.capture(2, T * F, F, whole_step=False)      # forward + backward only: comparable with plain autograd
    assert tr.launches_in_graph > 50 and not tr._graph_has_tail
    tr.static_x.copy_(X)
    tr.static_y.copy_(y)
    for v in tr._grad_views:
        v.fill_(123.0)                 # every parameter's gradient must be rewritten by the captured step
    tr.graph.replay()
    torch.cuda.synchronize()
    graph_loss, graph_grad = tr.static_loss.detach().clone(), tr.flat_grad.clone()
    p0 = tr.flat_param.clone()
    tr.zero_grad()
    l0 = tr.loss(X, y)
    l0.backward()                      # plain autograd accumulation into the zeroed flat buffer
    assert torch.equal(l0.detach(), graph_loss)
    assert torch.equal(tr.flat_grad, graph_grad)
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(tr.params, tr._grad_views))
    # the full captured step also moves the weights (Adam) and leaves the loss readable
    tr.step_captured()
    assert not torch.equal(tr.flat_param, p0) and int(tr.step_count.item()) == 1
    # whole-step graph (the default): ONE replay = forward + backward + (all-reduce) + Adam, bit-identical to the
    # eager step from the same optimiser state
    state = [t.clone() for t in (tr.flat_param, tr.exp_avg, tr.exp_avg_sq, tr.step_count)]
    n_fb = tr.launches_in_graph
    tr.capture(2, T * F, F)
    assert tr._graph_has_tail and tr.launches_in_graph == n_fb + 2
    assert all(torch.equal(a, b) for a, b in zip(state, (tr.flat_param, tr.exp_avg, tr.exp_avg_sq, tr.step_count)))
    tr.static_x.copy_(X)
    tr.static_y.copy_(y)
    tr.step_captured()
    torch.cuda.synchronize()
    after_graph = tr.flat_param.clone()
    for dst, src in zip((tr.flat_param, tr.exp_avg, tr.exp_avg_sq, tr.step_count), state):
        dst.copy_(src)
    tr.step(X, y)
    assert torch.equal(tr.flat_param, after_graph) and int(tr.step_count.item()) == 2
    # end-to-end entry: pinned host batch in, loss out; a prefetched batch is consumed from the staging buffers
    hx, hy = X.cpu().pin_memory(), y.cpu().pin_memory()
    l1 = tr.step_from_host(hx, hy, next_batch=(hx, hy))
    assert tr._has_staged
    tr.static_x.zero_()
    l2 = tr.step_from_host(hx, hy)
    assert not tr._has_staged and torch.equal(tr.static_x, X)
    assert l1 > 0 and l2 > 0 and l2 < l1 and int(tr.step_count.item()) == 4


def test_device_resident_rollout_and_metrics_match_oracle():
    """f4: a 3-step autoregressive rollout of the batched model on the GPU (gcl_b200.predict.rollout) against the
    oracle model driven by the restated reference loop (scripts/predict.py:534-580), and the streaming metrics
    accumulated on the device against the restated StreamingMetrics."""
    from gcl_b200 import predict as gp
    from oracle import predict as op
    mine, ref, cfg, nlat, nlon = _models("attention", 16, 32, [1, 3], 0.6)
    G, C, OBS, AR, B = nlat * nlon, 33, 2, 3, 2
    gen = torch.Generator().manual_seed(11)
    X, y = torch.randn(B, G, OBS * C, generator=gen), torch.randn(B, G, AR * C, generator=gen)
    out = gp.rollout(mine, X.to(DEV), AR, C, OBS, y=y.to(DEV), static_ch=(0,), forcing_ch=(5,))
    assert out.shape == (B, G, AR * C) and out.is_cuda
    with torch.no_grad():
        want = torch.stack([op.ar_rollout(lambda inp: ref(X=inp), X[b:b + 1].clone(), AR, C, OBS, y=y[b],
                                          static_ch=(0,), forcing_ch=(5,)) for b in range(B)])
    assert_close(out, want, RTOL_F32, "3-step rollout")
    sm, so = gp.StreamingMetrics(C, (0, 5), device=DEV), op.StreamingMetrics(C, [0, 5])
    sm.update(y.to(DEV), out)
    for b in range(B):
        so.update(y[b], want[b])
    r = sm.result()
    assert abs(r["rmse"] - so.rmse) <= 1e-5 * so.rmse and abs(r["acc"] - so.acc) <= 1e-4
    assert np.allclose(r["rmse_per_channel"], so.rmse_per_channel, rtol=1e-5)


def test_import_swap_reference_glue_on_gcl_layers():
    """The reference's model glue (oracle/model.py restates models.py line for line and is pinned
    bit-exact against it) executed with `torch_geometric.nn` resolving to gcl_b200.nn: the drop-in claim
    (type(layer) == GCNConv checks, SparseGATConv subclassing, shared PReLU, summary-free forward)."""
    import gcl_b200.nn as gnn
    from gcl_b200.workloads import get_workload
    from oracle import model as om
    saved = {k: sys.modules.get(k) for k in ("torch_geometric", "torch_geometric.nn")}
    fake = types.ModuleType("torch_geometric")
    fake.nn = gnn
    sys.modules["torch_geometric"], sys.modules["torch_geometric.nn"] = fake, gnn
    try:
        spec = importlib.util.spec_from_file_location("oracle._glue_on_gcl", os.path.join(ROOT, "oracle", "model.py"))
        swapped = importlib.util.module_from_spec(spec)
        swapped.__package__ = "oracle"
        spec.loader.exec_module(swapped)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert swapped.GCNConv is gnn.GCNConv and issubclass(swapped.SparseGATConv, gnn.GATConv)
    for name in ("baseline", "attention", "sparse_attention"):
        cfg = get_workload(name)
        cfg["graph"]["mesh_levels"], cfg["graph"]["grid2mesh_radius_query"] = [1, 3], 0.6
        og_ = oracle_graphs(16, 32, [1, 3], 0.6)
        torch.manual_seed(5)
        ref = om.WeatherPrediction(cfg, 16, 32, graphs=og_)
        sw = swapped.WeatherPrediction(cfg, 16, 32, graphs=og_)
        sw.load_state_dict(ref.state_dict())
        sw.to(DEV)
        for attr in ("encoding_graph", "processing_graph", "decoding_graph", "init_grid_features", "init_mesh_features"):
            setattr(sw, attr, getattr(sw, attr).to(DEV))
        X = torch.randn(1, 512, cfg["data"]["num_features_used"] * 2, generator=torch.Generator().manual_seed(1))
        kw = dict(batch_num=0) if name == "sparse_attention" else {}
        a = ref(X=X, attention_threshold=0.02, **kw)
        b = sw(X=X.to(DEV), attention_threshold=0.02, **kw)
        assert_close(b, a, RTOL_F32, f"import-swap {name}")
        a.square().mean().backward()
        b.square().mean().backward()
        _check_grads(sw, ref, RTOL_F32, f"import-swap {name}")


@pytest.mark.parametrize("residual", [True, False])
def test_ar_rollout_with_static_forcing_channels_and_loss_masks(residual):
    """train_epoch semantics beyond the BASELINE configs (train.py:85-102, 203-207, 218-226): static channels carried
    forward, forcing channels taken from the target, channel / spatial loss masks, use_residual on and off; AR = 3
    with BPTT, against the oracle's restated loop."""
    from gcl_b200.train import Trainer
    from oracle import model as om
    mine, ref, cfg, nlat, nlon = _models("wb2_64x32_ar_15f_4obs_4pred", 16, 32, [1, 3], 0.6, seed=21)
    G, F, T = nlat * nlon, 15, 4
    static, forcing = (2, 7), (0, 11)
    cmask = torch.ones(F)
    cmask[list(static)] = 0.0
    smask = torch.zeros(nlon, nlat)
    smask[2:nlon - 2, 1:nlat - 1] = 1.0                  # build_boundary_mask layout (train.py:74-83)
    smask = smask.reshape(1, -1, 1)
    lw = om.lat_weights(nlat, nlon)

    def make(seed):
        gen = torch.Generator().manual_seed(seed)
        return torch.randn(2, G, T * F, generator=gen), torch.randn(2, G, 4 * F, generator=gen)
    kwargs = dict(use_residual=residual, static_channels=static, forcing_channels=forcing, channel_mask=cmask,
                  spatial_mask=smask)
    floor, (X, y) = _kink_free_inputs(ref, make, lambda X, y: om.training_loss(ref, X, y, 3, lw, **kwargs), 1e-6,
                                      range(40, 90))
    tr = Trainer(mine, nlat, nlon, ar_steps=3, use_residual=residual, static_channels=static, forcing_channels=forcing,
                 channel_mask=cmask, spatial_mask=smask)
    tr.zero_grad()
    lg = tr.loss(X.to(DEV), y.to(DEV))
    lg.backward()
    lc = om.training_loss(ref, X, y, 3, lw, **kwargs)
    lc.backward()
    assert abs(float(lg) - float(lc)) <= RTOL_F32 * abs(float(lc)), (float(lg), float(lc))
    _check_grads(mine, ref, RTOL_F32, f"AR=3 masks residual={residual}")
    # the carried state itself: static channels keep the input's last step, forcing channels follow the target
    from gcl_b200.train import _ARStep
    delta = torch.randn(2, G, F, device=DEV)
    st = X.to(DEV).view(2, G, T, F)
    _, new = _ARStep.apply(delta, st, y.to(DEV).view(2, G, 4, F)[:, :, 1, :], tr.node_w, tr.chan_w, tr._carry,
                           residual, 1.0, 1.0, True)
    assert torch.equal(new[:, :, :-1], st[:, :, 1:])
    assert torch.equal(new[:, :, -1, list(static)], st[:, :, -1, list(static)])
    assert torch.equal(new[:, :, -1, list(forcing)], y.to(DEV).view(2, G, 4, F)[:, :, 1, list(forcing)])
    free = [c for c in range(F) if c not in static + forcing]
    want = delta[:, :, free] + (st[:, :, -1, free] if residual else 0)
    assert torch.equal(new[:, :, -1, free], want)
