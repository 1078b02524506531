"""ORACLE / TEST INFRASTRUCTURE ONLY -- numpy restatement of TimeseriesChunkDataset.__getitem__ and its sample index
(/root/reference/src/data/dataloader_chunked.py:137-223), for the GPU box where /root/reference is absent.  Pinned:
tests/test_data_cpu.py compares it (and the product) with the unmodified reference class on synthetic datasets."""
import numpy as np


def sample_indices(chunk_lengths, obs_window, pred_steps, split, test_fraction=0.2):
    window = obs_window + pred_steps
    idx = [(ci, t) for ci, T in enumerate(chunk_lengths) for t in range(max(T - window + 1, 0))]
    cut = int(len(idx) * (1 - test_fraction))
    if split == "train":
        return idx[:cut]
    if split == "test":
        return idx[cut:]
    if split in ("val", "test_only"):
        test = idx[cut:]
        half = len(test) // 2
        return test[:half] if split == "val" else test[half:]
    if split == "all":
        return idx
    raise ValueError(split)


def window_sample(chunk, local_t, obs_window, pred_steps, n_feat, mean, std, flat):
    """One sample -> (X [G, obs*F], Y [G, pred*F]) float32."""
    w = np.asarray(chunk[local_t: local_t + obs_window + pred_steps])
    w = (w[..., :n_feat].astype(np.float32) - mean[:n_feat]) / std[:n_feat]
    xf, yf = w[:obs_window], w[obs_window:]
    if flat:
        n = w.shape[1]
        return (xf.transpose(1, 0, 2).reshape(n, obs_window * n_feat), yf.transpose(1, 0, 2).reshape(n, pred_steps * n_feat))
    g = w.shape[1] * w.shape[2]
    return (xf.transpose(2, 1, 0, 3).reshape(g, obs_window * n_feat), yf.transpose(2, 1, 0, 3).reshape(g, pred_steps * n_feat))
