"""f2 (input pipeline), host side: the product's ChunkedWindowLoader stages RAW windows (the convert / normalise /
transpose kernel is CUDA only: tests/test_data_gpu.py), so here its sample index, splits, staging and rank sharding
are checked against the oracle restatement, the oracle's arithmetic on those raw windows against the product's
on-disk reading, and -- where /root/reference is mounted -- the oracle against the unmodified reference
TimeseriesChunkDataset, bit for bit, on synthetic datasets in the reference's on-disk formats."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from gcl_b200.data import ChunkedWindowLoader
from oracle import data as od

REF = "/root/reference"


def _make(tmp, kind, T=14, lon=6, lat=4, F=5, seed=0):
    rng = np.random.default_rng(seed)
    np.savez(os.path.join(tmp, "scalers.npz"), mean=rng.normal(size=F).astype(np.float32),
             std=(0.5 + rng.random(F)).astype(np.float32), n=np.int64(T))
    if kind == "raw":
        a = rng.normal(size=(T, lon, lat, F)).astype(np.float16)
        mm = np.memmap(os.path.join(tmp, "data.npy"), dtype=np.float16, mode="w+", shape=a.shape)
        mm[:] = a
        mm.flush()
        json.dump({"n_time": T, "n_lon": lon, "n_lat": lat, "n_feat": F}, open(os.path.join(tmp, "dataset_info.json"), "w"))
        return [a]
    if kind == "flat":
        a = rng.normal(size=(T, lon * lat, F)).astype(np.float16)
        mm = np.memmap(os.path.join(tmp, "data.npy"), dtype=np.float16, mode="w+", shape=a.shape)
        mm[:] = a
        mm.flush()
        json.dump({"n_time": T, "n_nodes": lon * lat, "n_feat": F, "flat": True}, open(os.path.join(tmp, "dataset_info.json"), "w"))
        return [a]
    parts = [rng.normal(size=(t, lon, lat, F)).astype(np.float16) for t in (T, 3, T - 5)]   # one chunk too short
    for i, a in enumerate(parts):
        np.save(os.path.join(tmp, f"chunk_{i}.npy"), a)
    return parts


@pytest.mark.parametrize("kind", ["raw", "flat", "chunks"])
@pytest.mark.parametrize("split,nf,obs,pred", [("train", None, 2, 1), ("test_only", 3, 2, 4), ("val", None, 4, 4), ("all", 4, 1, 2)])
def test_loader_matches_oracle_bit_for_bit(tmp_path, kind, split, nf, obs, pred):
    """Sample index / splits equal the oracle's; the staged raw windows are the stored windows, bit for bit (the
    device kernel turns exactly these into the model tensors)."""
    parts = _make(str(tmp_path), kind)
    ld = ChunkedWindowLoader(str(tmp_path), obs, pred, split, nf, device="cpu")
    want_idx = od.sample_indices([p.shape[0] for p in parts], obs, pred, split)
    assert ld.sample_indices == want_idx and len(ld) == len(want_idx)
    if not want_idx:
        return
    pick = list(range(len(want_idx)))[::2][:5] or [0]
    raw = ld.batch(pick, raw=True)
    assert raw.shape[:2] == (len(pick), obs + pred) and raw.dtype == torch.float16
    for b, i in enumerate(pick):
        ci, t = want_idx[i]
        assert np.array_equal(raw[b].numpy(), parts[ci][t: t + obs + pred]), (kind, split, i)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ld.batch(pick)


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not present")
@pytest.mark.parametrize("kind", ["raw", "flat", "chunks"])
def test_loader_and_oracle_equal_the_reference_dataset(tmp_path, kind):
    sys.path.insert(0, REF)
    try:
        from src.data.dataloader_chunked import TimeseriesChunkDataset
    finally:
        sys.path.remove(REF)
    parts = _make(str(tmp_path), kind, seed=3)
    for split, nf, obs, pred in (("train", None, 2, 1), ("test", 3, 2, 3), ("val", None, 3, 2), ("test_only", 4, 2, 4), ("all", None, 2, 2)):
        ref = TimeseriesChunkDataset(str(tmp_path), obs_window=obs, pred_steps=pred, split=split, n_features=nf)
        ld = ChunkedWindowLoader(str(tmp_path), obs, pred, split, nf, device="cpu")
        assert ld.sample_indices == ref._sample_indices
        assert od.sample_indices([p.shape[0] for p in parts], obs, pred, split) == ref._sample_indices
        if len(ref) == 0:
            continue
        sc = np.load(os.path.join(str(tmp_path), "scalers.npz"))
        F = nf or parts[0].shape[-1]
        raw = ld.batch(range(len(ref)), raw=True).numpy()        # what the product hands to its device kernel
        for i in range(len(ref)):
            xr, yr = ref[i]
            x, y = od.window_sample(raw[i], 0, obs, pred, F, sc["mean"].astype(np.float32), sc["std"].astype(np.float32),
                                    kind == "flat")              # the oracle's arithmetic on the product's staging
            assert np.array_equal(x, xr.numpy()) and np.array_equal(y, yr.numpy()), (kind, split, i)


def test_prefetching_iterator_covers_the_split_deterministically(tmp_path):
    _make(str(tmp_path), "raw", T=23)
    ld = ChunkedWindowLoader(str(tmp_path), 2, 2, "all", None, device="cpu")
    n = len(ld)
    plain = [ld.batch(range(i, min(i + 4, n)), raw=True) for i in range(0, n, 4)]
    got = list(ld.batches(4, raw=True))
    assert len(got) == len(plain) and all(torch.equal(a, b) for a, b in zip(got, plain))
    assert len(list(ld.batches(4, drop_last=True, raw=True))) == n // 4
    # shuffled: same permutation for the same seed, every sample exactly once, ranks take disjoint strides
    a = torch.cat(list(ld.batches(3, shuffle=True, seed=7, raw=True)))
    b = torch.cat(list(ld.batches(3, shuffle=True, seed=7, raw=True)))
    assert torch.equal(a, b) and a.shape[0] == n
    full = torch.cat(list(ld.batches(n, raw=True)))
    key = lambda t: sorted(map(tuple, t.reshape(t.shape[0], -1)[:, :6].tolist()))
    assert key(a) == key(full)
    r0 = torch.cat(list(ld.batches(3, rank=0, world=2, raw=True)))
    r1 = torch.cat(list(ld.batches(3, rank=1, world=2, raw=True)))
    assert r0.shape[0] + r1.shape[0] == n and key(torch.cat([r0, r1])) == key(full)


def test_rank_shards_are_equal_when_the_split_does_not_divide(tmp_path):
    """n % world != 0: every rank must see the same number of batches of the same shapes (each step ends in a
    collective), the order being wrap-padded like DistributedSampler's; every sample is still covered."""
    _make(str(tmp_path), "raw", T=23)
    ld = ChunkedWindowLoader(str(tmp_path), 2, 2, "all", None, device="cpu")
    n = len(ld)
    for world, bs in ((3, 2), (3, 1), (6, 3), (7, 2)):
        assert n % world != 0
        per_rank = [list(ld.batches(bs, shuffle=True, seed=3, rank=r, world=world, raw=True)) for r in range(world)]
        shapes = [[tuple(x.shape) for x in b] for b in per_rank]
        assert all(s == shapes[0] for s in shapes), (world, bs, shapes)
        key = lambda t: set(map(tuple, t.reshape(t.shape[0], -1)[:, :6].tolist()))
        full = torch.cat(list(ld.batches(n, raw=True)))
        seen = set().union(*[key(torch.cat(b)) for b in per_rank])
        assert seen == key(full)
        per_rank_d = [list(ld.batches(bs, rank=r, world=world, drop_last=True, raw=True)) for r in range(world)]
        assert len({len(b) for b in per_rank_d}) == 1


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_legacy_chunks_of_other_float_dtypes_are_not_quantised(tmp_path, dtype):
    """chunk_*.npy keep their stored dtype; the reference does .astype(np.float32) (dataloader_chunked.py:190)."""
    rng = np.random.default_rng(5)
    T, lon, lat, F = 9, 5, 3, 4
    np.savez(os.path.join(tmp_path, "scalers.npz"), mean=rng.normal(size=F).astype(np.float32),
             std=(0.5 + rng.random(F)).astype(np.float32), n=np.int64(T))
    a = (rng.normal(size=(T, lon, lat, F)) * 1e5).astype(dtype)          # beyond float16's range
    np.save(os.path.join(tmp_path, "chunk_0.npy"), a)
    ld = ChunkedWindowLoader(str(tmp_path), 2, 1, "all", None, device="cpu")
    raw = ld.batch([0, 3], raw=True)
    assert raw.dtype == torch.from_numpy(np.empty(0, dtype)).dtype       # staged in the stored dtype, not float16
    for b, t in enumerate((0, 3)):
        assert np.array_equal(raw[b].numpy(), a[t:t + 3])
