"""Input pipeline for the chunked raw-timeseries datasets.

Mirrors TimeseriesChunkDataset of /root/reference/src/data/dataloader_chunked.py:33-223 -- same on-disk format (raw
float16 memmap `data.npy` + `dataset_info.json`, or legacy `chunk_*.npy`; `scalers.npz`), same sliding-window sample
index that never crosses a chunk boundary (:137-149), same time-ordered splits (:151-174) -- but moves the arithmetic
off the host: the reference converts, normalises and transposes every window on CPU workers (:189-223); here the host
only copies the raw float16 window into a pinned staging buffer (half the bytes, no math), and ONE kernel
(gcl_window_assemble) does the float32 convert, (x - mean) / std, the (lat, lon)-major flatten and the obs / target
split on the device for the whole batch.  Results are bit-identical to the reference's (same IEEE float32 operations).
The device side is CUDA only (no CPU fallback); `raw=True` returns the staged raw windows for host-side use.
"""
import glob
import json
import os
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi


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

    def batch(self, indices: Sequence[int], raw: bool = False):
        """X [B, G, obs*F], Y [B, G, pred*F] float32 on self.device for the given sample indices.
        raw=True: the staged raw windows [B, W, ...] (host tensor, stored dtype) instead -- no device involved."""
        if raw:
            t = torch.empty((len(indices), self.obs_window + self.pred_steps) + self.frame_shape, dtype=self.raw_dtype)
            self._fill(t, indices)
            return t
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
                rank: int = 0, world: int = 1, raw: bool = False):
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
        if raw:                                    # host-side consumers (and the host-logic tests): staged raw windows
            for g in groups:
                yield self.batch(g, raw=True)
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
        if self.device.type != "cuda":
            raise RuntimeError("gcl_b200.data: the convert / normalise / transpose kernel is CUDA only; there is no CPU "
                               "fallback (use raw=True for the staged raw windows)")
        B, window = stage.shape[0], self.obs_window + self.pred_steps
        raw = stage.to(self.device, non_blocking=True)
        code = {torch.float16: 0, torch.float32: 1, torch.float64: 2}.get(self.raw_dtype)
        if code is None:
            raise RuntimeError(f"gcl_b200.data: stored dtype {self.raw_dtype} is not float16 / float32 / float64")
        nlon, nlat = (self.grid_nodes, 1) if self.flat_grid else self.frame_shape[:2]
        X = torch.empty((B, self.grid_nodes, self.obs_window * self.n_feat), dtype=torch.float32, device=self.device)
        Y = torch.empty((B, self.grid_nodes, self.pred_steps * self.n_feat), dtype=torch.float32, device=self.device)
        lib = _cabi.load()
        with torch.cuda.device(self.device):
            _cabi.check(lib.gcl_window_assemble(raw.data_ptr(), code, self.mean.data_ptr(), self.std.data_ptr(),
                                                X.data_ptr(), Y.data_ptr(), B, window, self.obs_window, nlon, nlat,
                                                self.n_feat_total, self.n_feat, int(self.flat_grid),
                                                torch.cuda.current_stream(self.device).cuda_stream),
                        "gcl_window_assemble")
        return X, Y
