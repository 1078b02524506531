"""CPU tests of the product's host logic: the vectorised mesh / edge builders (bit-exact against the
oracle), the workload table, and the data-parallel host logic of the Trainer with gloo, world size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_vectorised_mesh_hierarchy_is_bit_identical_to_the_oracle():
    from gcl_b200 import graphs_build as gb
    from oracle import graphs as og
    ho, hp = og.mesh_hierarchy(5), gb.mesh_hierarchy(5)
    for (vo, fo), (vp, fp) in zip(ho, hp):
        assert vo.dtype == vp.dtype == np.float32
        assert np.array_equal(vo, vp), "vertex coordinates must match to the last bit"
        assert np.array_equal(fo, fp)
    for levels in ([0], [1, 3], [3, 5], [2, 4, 5]):
        eo = og.mesh_edges(og.merged_faces(ho, levels)[1]).astype(np.int64)
        ep = gb.edges_from_faces(gb.merged_faces(hp, levels), len(hp[max(levels)][0]))
        assert np.array_equal(eo, ep), levels
    assert [len(v) for v, _ in hp] == [12, 42, 162, 642, 2562, 10242]
    assert [len(f) for _, f in hp] == [20 * 4 ** k for k in range(6)]


def test_static_features_and_mesh_latlon_match_oracle():
    from gcl_b200 import graphs_build as gb
    from oracle import graphs as og
    v = gb.mesh_hierarchy(4)[-1][0]
    assert all(np.array_equal(a, b) for a, b in zip(gb._mesh_lat_lon(v), og.mesh_lat_lon(v)))
    lat32 = np.linspace(-90, 90, 32).astype(np.float32)
    lon32 = np.linspace(0, 360, 64, endpoint=False).astype(np.float32)
    glon, glat = np.meshgrid(lon32, lat32)
    a = gb._static_features(glat.reshape(-1), glon.reshape(-1))
    b = og.node_features(glat.reshape(-1).astype(np.float32), glon.reshape(-1).astype(np.float32))
    assert a.shape == (2048, 6) and np.array_equal(a, b)
    assert np.array_equal(gb._grid_xyz(lat32, lon32), og.grid_xyz(lat32, lon32))


def test_workload_table_matches_reference_configs():
    """The restated config dicts equal the reference's config.json files where those are mounted."""
    import json
    from gcl_b200.workloads import WORKLOADS
    ref_dir = "/root/reference/experiments"
    if not os.path.isdir(ref_dir):
        pytest.skip("/root/reference not present")

    def norm(v):
        if isinstance(v, str) and v.lower() in ("true", "false"):
            return v.lower() == "true"
        return v

    for name, w in WORKLOADS.items():
        cfg = json.load(open(os.path.join(ref_dir, name, "config.json")))
        assert cfg["graph"]["mesh_levels"] == w["graph"]["mesh_levels"]
        assert cfg["graph"]["grid2mesh_radius_query"] == w["graph"]["grid2mesh_radius_query"]
        assert cfg["data"]["num_features_used"] == w["data"]["num_features_used"]
        assert cfg["data"]["obs_window_used"] == w["data"]["obs_window_used"]
        assert cfg.get("max_ar_steps", 1) == w["max_ar_steps"]
        assert cfg["learning_rate"] == w["learning_rate"]
        for part in ("encoder", "processor", "decoder"):
            rg, wg = cfg["pipeline"][part]["gcn"], w["pipeline"][part]["gcn"]
            assert rg["layer_type"] == wg["layer_type"]
            assert (rg.get("hidden_dims") or []) == (wg.get("hidden_dims") or [])
            for k in ("activation", "num_message_passing_steps", "edge_feature_dim"):      # v2 / InteractionNet keys
                assert (rg.get(k) or {"activation": "prelu"}.get(k)) == (wg.get(k) or {"activation": "prelu"}.get(k)), (name, part, k)
            assert rg.get("output_dim") == wg["output_dim"]
            assert bool(norm(rg.get("use_layer_norm", False))) == bool(wg["use_layer_norm"])
            rm, wm = cfg["pipeline"][part].get("mlp"), w["pipeline"][part].get("mlp")
            assert (rm is None) == (wm is None)
            if rm:
                assert rm["mlp_hidden_dims"] == wm["mlp_hidden_dims"] and rm["output_dim"] == wm["output_dim"]
                assert bool(norm(rm["use_layer_norm"])) == bool(wm["use_layer_norm"])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gcl_b200.train import Trainer
        torch.manual_seed(100 + rank)                      # ranks start from DIFFERENT weights
        model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.PReLU(), torch.nn.Linear(7, 3))
        model.obs_window = 1
        tr = Trainer(model, nlat=4, nlon=8, lr=1e-2, ar_steps=1)
        # construction broadcasts rank 0's weights: replicas are identical afterwards
        w0 = tr.flat_param.clone()
        gathered = [torch.empty_like(w0) for _ in range(world)]
        dist.all_gather(gathered, w0)
        same_weights = all(torch.equal(g, gathered[0]) for g in gathered)
        # parameters are views of the flat buffer and gradients accumulate into the flat gradient buffer
        x = torch.randn(6, 5, generator=torch.Generator().manual_seed(rank))
        tr.zero_grad()
        model(x).square().mean().backward()
        local = tr.flat_grad.clone()
        tr.reduce_gradients()
        gl = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gl, local)
        ok_sum = torch.allclose(tr.flat_grad, sum(gl), rtol=0, atol=1e-7)
        views = all(p.data_ptr() >= tr.flat_param.data_ptr() and
                    p.data_ptr() < tr.flat_param.data_ptr() + 4 * tr.flat_size and p.data_ptr() % 128 == tr.flat_param.data_ptr() % 128
                    for p in model.parameters())
        out[rank] = (same_weights, ok_sum, views, tr.world, tr.num_params)
    finally:
        dist.destroy_process_group()


def test_data_parallel_host_logic_gloo_world2():
    """Rank-0 broadcast of the initial weights, flat parameter/gradient views, and the one all-reduce."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        same_weights, ok_sum, views, w, n = out[rank]
        assert same_weights and ok_sum and views and w == world and n == 5 * 7 + 7 + 1 + 7 * 3 + 3
