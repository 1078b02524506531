"""f2 on the GPU: the device-side convert / normalise / transpose of ChunkedWindowLoader (pinned float16 staging ->
H2D -> float32 math on the device) is bit-identical to the oracle restatement of the reference's CPU dataset."""
import os

import numpy as np
import pytest
import torch

from test_data_cpu import _make

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["raw", "flat", "chunks"])
def test_loader_on_device_is_bit_exact(tmp_path, kind):
    from gcl_b200.data import ChunkedWindowLoader
    from oracle import data as od
    parts = _make(str(tmp_path), kind, T=20, lon=16, lat=8, F=7, seed=5)
    obs, pred, nf = 2, 4, 6
    ld = ChunkedWindowLoader(str(tmp_path), obs, pred, "all", nf, device="cuda:0")
    idx = od.sample_indices([p.shape[0] for p in parts], obs, pred, "all")
    assert ld.sample_indices == idx
    pick = list(range(0, len(idx), 3))
    X, Y = ld.batch(pick)
    assert X.is_cuda and X.dtype == torch.float32 and X.is_contiguous()
    sc = np.load(os.path.join(str(tmp_path), "scalers.npz"))
    for b, i in enumerate(pick):
        ci, t = idx[i]
        x, y = od.window_sample(parts[ci], t, obs, pred, nf, sc["mean"].astype(np.float32), sc["std"].astype(np.float32),
                                kind == "flat")
        assert np.array_equal(X[b].cpu().numpy(), x) and np.array_equal(Y[b].cpu().numpy(), y), (kind, i)
    X2, _ = ld.batch(pick[:2])                      # staging buffer reuse with a smaller batch
    assert torch.equal(X2, X[:2])


def test_prefetching_iterator_on_device(tmp_path):
    """batches(): worker-thread staging + event-guarded buffer reuse gives the same tensors as batch()."""
    from gcl_b200.data import ChunkedWindowLoader
    _make(str(tmp_path), "raw", T=40, lon=16, lat=8, F=7, seed=9)
    ld = ChunkedWindowLoader(str(tmp_path), 2, 1, "all", None, device="cuda:0")
    n = len(ld)
    got = [(x.clone(), y.clone()) for x, y in ld.batches(5)]
    assert sum(x.shape[0] for x, _ in got) == n
    for k, (x, y) in enumerate(got):
        xr, yr = ld.batch(range(5 * k, min(5 * k + 5, n)))
        assert torch.equal(x, xr) and torch.equal(y, yr), k


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_legacy_chunks_of_other_float_dtypes_on_device(tmp_path, dtype):
    """chunk_*.npy stored as float32 / float64 (values beyond float16's range): the kernel converts like the
    reference's .astype(np.float32) (dataloader_chunked.py:190) -- bit-exact, nothing quantised."""
    from gcl_b200.data import ChunkedWindowLoader
    rng = np.random.default_rng(5)
    T, lon, lat, F = 9, 5, 3, 4
    np.savez(os.path.join(tmp_path, "scalers.npz"), mean=rng.normal(size=F).astype(np.float32),
             std=(0.5 + rng.random(F)).astype(np.float32), n=np.int64(T))
    a = (rng.normal(size=(T, lon, lat, F)) * 1e5).astype(dtype)
    np.save(os.path.join(tmp_path, "chunk_0.npy"), a)
    ld = ChunkedWindowLoader(str(tmp_path), 2, 1, "all", None, device="cuda:0")
    X, Y = ld.batch([0, 3])
    sc = np.load(os.path.join(tmp_path, "scalers.npz"))
    for b, t in enumerate((0, 3)):
        w = (a[t:t + 3].astype(np.float32) - sc["mean"]) / sc["std"]
        w = w.transpose(2, 1, 0, 3).reshape(lon * lat, 3, F)
        assert np.array_equal(X[b].cpu().numpy(), w[:, :2].reshape(lon * lat, 2 * F))
        assert np.array_equal(Y[b].cpu().numpy(), w[:, 2:].reshape(lon * lat, F))
        assert np.isfinite(X[b].cpu().numpy()).all()
