"""Input pipeline for the chunked raw-timeseries datasets.

Mirrors TimeseriesChunkDataset of /root/reference/src/data/dataloader_chunked.py:33-223 -- same on-disk format (raw
float16 memmap `data.npy` + `dataset_info.json`, or legacy `chunk_*.npy`; `scalers.npz`), same sliding-window sample
index that never crosses a chunk boundary (:137-149), same time-ordered splits (:151-174) -- but moves the arithmetic
off the host: the reference converts, normalises and transposes every window on CPU workers (:189-223); here the host
only copies the raw float16 window into a pinned staging buffer (half the bytes, no math), and the float32 convert,
(x - mean) / std, the (lat, lon)-major flatten and the obs / target split run on the device for the whole batch.
Results are bit-identical to the reference's (same IEEE float32 operations in the same order).
"""
import glob
import json
import os
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


class ChunkedWindowLoader:
    def __init__(self, data_dir: str, obs_window: int = 2, pred_steps: int = 1, split: str = "train",
                 n_features: Optional[int] = None, test_fraction: float = 0.2, device="cuda:0"):
        self.obs_window, self.pred_steps = int(obs_window), int(pred_steps)
        self.device = torch.device(device)
        sc = np.load(os.path.join(data_dir, "scalers.npz"))
        single, info_file = os.path.join(data_dir, "data.npy"), os.path.join(data_dir, "dataset_info.json")
        if os.path.exists(single) and os.path.exists(info_file):        # raw memmap, no .npy header (:81-97)
            info = json.load(open(info_file))
            self.flat_grid = bool(info.get("flat", False))
            shape = ((info["n_time"], info["n_nodes"], info["n_feat"]) if self.flat_grid
                     else (info["n_time"], info["n_lon"], info["n_lat"], info["n_feat"]))
            self.chunks = [np.memmap(single, dtype=np.float16, mode="r", shape=shape)]
        else:                                                            # legacy chunk_*.npy (:98-113)
            self.flat_grid = False
            files = sorted(glob.glob(os.path.join(data_dir, "chunk_*.npy")))
            if not files:
                raise FileNotFoundError(f"No data.npy or chunk_*.npy found in {data_dir}")
            self.chunks = [np.load(f, mmap_mode="r") for f in files]
        c0 = self.chunks[0]
        # legacy chunk files keep whatever dtype they were saved with (the reference converts with
        # .astype(np.float32), :190); the staging buffers take the stored dtype and the device converts
        self.raw_dtype = torch.from_numpy(np.empty(0, dtype=c0.dtype)).dtype
        if any(ch.dtype != c0.dtype for ch in self.chunks):
            raise ValueError("gcl_b200.data: all chunk files must share one dtype")
        if not self.raw_dtype.is_floating_point:
            raise ValueError(f"gcl_b200.data: unsupported chunk dtype {c0.dtype}")
        self.n_feat_total = c0.shape[-1]
        self.n_feat = int(n_features) if n_features else self.n_feat_total
        self.frame_shape = tuple(c0.shape[1:])                          # (lon, lat, F_total) or (N, F_total)
        self.grid_nodes = c0.shape[1] if self.flat_grid else c0.shape[1] * c0.shape[2]
        self.mean = torch.from_numpy(sc["mean"].astype(np.float32)[: self.n_feat]).to(self.device)
        self.std = torch.from_numpy(sc["std"].astype(np.float32)[: self.n_feat]).to(self.device)
        window = self.obs_window + self.pred_steps
        idx: List[Tuple[int, int]] = []
        for ci, ch in enumerate(self.chunks):                            # windows never cross a chunk (:137-149)
            idx.extend((ci, t) for t in range(max(ch.shape[0] - window + 1, 0)))
        cut = int(len(idx) * (1 - test_fraction))                        # time-ordered splits (:151-174)
        if split == "train":
            idx = idx[:cut]
        elif split == "test":
            idx = idx[cut:]
        elif split in ("val", "test_only"):
            test = idx[cut:]
            half = len(test) // 2
            idx = test[:half] if split == "val" else test[half:]
        elif split != "all":
            raise ValueError(f"Unknown split: {split}")
        self.sample_indices = idx
        self._stage = None
        self._stage_free = None          # event: the H2D copy that last read self._stage has completed

    def __len__(self):
        return len(self.sample_indices)

    def _staging(self, batch: int) -> torch.Tensor:
        window = self.obs_window + self.pred_steps
        if self._stage is None or self._stage.shape[0] < batch:
            t = torch.empty((batch, window) + self.frame_shape, dtype=self.raw_dtype)
            self._stage = t.pin_memory() if self.device.type == "cuda" else t
        return self._stage[:batch]

    def _fill(self, stage: torch.Tensor, indices: Sequence[int]):
        """Host side of a batch: raw float16 window copies memmap -> (pinned) staging, nothing else."""
        window = self.obs_window + self.pred_steps
        host = stage.numpy()
        for b, i in enumerate(indices):
            ci, t = self.sample_indices[int(i)]
            host[b] = self.chunks[ci][t: t + window]

    def batch(self, indices: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
        """X [B, G, obs*F], Y [B, G, pred*F] float32 on self.device for the given sample indices."""
        if self._stage_free is not None:
            self._stage_free.synchronize()       # the previous batch()'s H2D copy may still be reading the buffer
        stage = self._staging(len(indices))
        self._fill(stage, indices)
        out = self._to_device(stage)
        if self.device.type == "cuda":
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream(self.device))
        return out

    def batches(self, batch_size: int, shuffle: bool = False, seed: int = 0, drop_last: bool = False,
                rank: int = 0, world: int = 1):
        """Iterate over the split in batches (rank r of a data-parallel job takes samples r::world).  Every rank gets
        the same number of samples -- the order is wrap-padded to a multiple of `world` first, as
        torch.utils.data.DistributedSampler does -- so all ranks run the same number of (collective) steps.  The host
        copy of batch k+1 runs on a worker thread into a second staging buffer while batch k is being consumed; a
        staging buffer is rewritten only after the H2D copy that read it has completed."""
        order = np.arange(len(self.sample_indices))
        if shuffle:
            order = np.random.default_rng(seed).permutation(order)
        if world > 1 and len(order) % world:
            order = np.concatenate([order, order[: world - len(order) % world]])
        order = order[rank::world]
        groups = [order[i: i + batch_size] for i in range(0, len(order), batch_size)]
        if drop_last and groups and len(groups[-1]) < batch_size:
            groups.pop()
        if not groups:
            return
        window = self.obs_window + self.pred_steps
        cuda = self.device.type == "cuda"
        stages, events = [], []
        for _ in range(2):
            t = torch.empty((batch_size, window) + self.frame_shape, dtype=self.raw_dtype)
            stages.append(t.pin_memory() if cuda else t)
            events.append(torch.cuda.Event() if cuda else None)
        worker = threading.Thread(target=self._fill, args=(stages[0][: len(groups[0])], groups[0]))
        worker.start()
        for k, g in enumerate(groups):
            worker.join()                                            # staging k % 2 holds batch k
            cur = k % 2
            if k + 1 < len(groups):
                nxt = (k + 1) % 2
                if cuda and k >= 1:
                    events[nxt].synchronize()                        # the H2D copy of batch k-1 has left that buffer
                worker = threading.Thread(target=self._fill, args=(stages[nxt][: len(groups[k + 1])], groups[k + 1]))
                worker.start()
            out = self._to_device(stages[cur][: len(g)])
            if cuda:
                events[cur].record(torch.cuda.current_stream(self.device))
            yield out

    def _to_device(self, stage: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        B, window = stage.shape[0], self.obs_window + self.pred_steps
        raw = stage.to(self.device, non_blocking=True)
        w = (raw[..., : self.n_feat].float() - self.mean) / self.std     # dataloader_chunked.py:190-191 / 204-207
        if self.flat_grid:                                               # [B, W, N, F] -> [B, N, W, F]        (:196-199)
            w = w.permute(0, 2, 1, 3)
        else:                                                            # [B, W, lon, lat, F] -> [B, lat, lon, W, F] (:218-221)
            w = w.permute(0, 3, 2, 1, 4)
        w = w.reshape(B, self.grid_nodes, window, self.n_feat)
        X = w[:, :, : self.obs_window].reshape(B, self.grid_nodes, self.obs_window * self.n_feat)
        Y = w[:, :, self.obs_window:].reshape(B, self.grid_nodes, self.pred_steps * self.n_feat)
        return X.contiguous(), Y.contiguous()
