"""Device-resident inference: autoregressive rollout + streaming forecast metrics.

Mirrors the inference loop of /root/reference/scripts/predict.py:451-614 -- residual add, static-channel
carry-forward, forcing channels from the ground truth, window slide (:534-580); StreamingMetrics (:53-124) -- for B
samples at once and without the reference's per-step `.cpu()` round trips (:499-559): everything stays on the GPU
until result() is read.  The per-step glue is the training step's kernel (gcl_ar_step_f32: residual add, carry-forward
and window slide in one pass), the metrics one reduction kernel (gcl_forecast_metrics_f32); CUDA only, no CPU
fallback.  The model is any callable [B, G, OBS*C] -> [B, G, C] (gcl_b200.model.WeatherPrediction).
"""
from typing import Callable, Optional, Sequence

import torch

from . import _cabi


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"gcl_b200.predict: {what} is on {t.device}; the kernels are CUDA only, no CPU fallback")


@torch.no_grad()
def rollout(model: Callable, X: torch.Tensor, ar_steps: int, num_channels: int, obs_window: int,
            y: Optional[torch.Tensor] = None, static_ch: Sequence[int] = (), forcing_ch: Sequence[int] = (),
            residual: bool = True, **model_kwargs) -> torch.Tensor:
    """X [B, G, OBS*C] -> forecasts [B, G, ar_steps*C]; y [B, G, steps*C] supplies the forcing channels."""
    if X.dim() != 3 or X.shape[-1] != obs_window * num_channels:
        raise ValueError(f"gcl_b200.predict.rollout: X {tuple(X.shape)} is not [B, G, {obs_window}*{num_channels}]")
    _require_cuda(X, "X")
    B, G, C = X.shape[0], X.shape[1], num_channels
    dev = X.device
    lib = _cabi.load()
    state = X.to(torch.float32).contiguous().view(B, G, obs_window, C)
    y_f = None
    if len(forcing_ch) and y is not None:
        _require_cuda(y, "y")
        y_f = y.to(torch.float32).contiguous().view(B, G, -1, C)
    base = torch.zeros(C, dtype=torch.int32)
    for ch in static_ch:
        base[int(ch)] = 1
    forced = base.clone()
    for ch in forcing_ch:
        forced[int(ch)] = 2
    carry_plain, carry_forced = base.to(dev), forced.to(dev)
    nb = lib.gcl_wmse_workspace_bytes(B, G, C)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    dummy = torch.empty(1, dtype=torch.float32, device=dev)
    out = torch.empty((B, G, ar_steps, C), dtype=torch.float32, device=dev)
    for step in range(ar_steps):
        delta = model(state.view(B, G, obs_window * C), **model_kwargs)
        if delta.dim() == 2:
            delta = delta.unsqueeze(0)
        delta = delta.to(torch.float32).contiguous()
        have_y = y_f is not None and step < y_f.shape[2]               # forcing values exist for this step (:564)
        yv = y_f[:, :, step, :] if have_y else None
        new_state = torch.empty_like(state)
        with torch.cuda.device(dev):
            _cabi.check(lib.gcl_ar_step_f32(delta.data_ptr(), state.data_ptr(), yv.data_ptr() if have_y else None,
                                            yv.stride(1) if have_y else C, None, None,
                                            (carry_forced if have_y else carry_plain).data_ptr(), int(bool(residual)),
                                            1.0, 1.0, new_state.data_ptr(), None, dummy.data_ptr(), 0, B, G,
                                            obs_window, C, ws.data_ptr(), nb,
                                            torch.cuda.current_stream(dev).cuda_stream), "gcl_ar_step_f32")
        state = new_state
        out[:, :, step, :] = state[:, :, -1, :]
    return out.view(B, G, ar_steps * C)


def persistence(X: torch.Tensor, num_channels: int, horizons: int) -> torch.Tensor:
    """Baseline forecast: the last observed step repeated.  X [B, G, OBS*C] -> [B, G, horizons*C]."""
    return X[..., -num_channels:].repeat(1, 1, horizons)


class StreamingMetrics:
    """MSE / RMSE / MAE over the non-excluded channels, per-channel RMSE and spatial anomaly correlation (ACC),
    accumulated in float64 on the device (one reduction kernel per update); update() takes a whole batch and never
    synchronises."""

    def __init__(self, num_channels: int, exclude_channels: Sequence[int] = (), device="cpu"):
        self.C = int(num_channels)
        keep = torch.ones(self.C, dtype=torch.bool)
        for c in exclude_channels:
            keep[int(c)] = False
        f64 = dict(dtype=torch.float64, device=device)
        self.keep = keep.to(device)
        self.n_keep = int(keep.sum())               # host-side count: update() must not synchronise
        self.n = 0
        self.sum_se, self.sum_ae = torch.zeros((), **f64), torch.zeros((), **f64)
        self.total_elem = 0
        self.sum_se_per_ch, self.sum_acc = torch.zeros(self.C, **f64), torch.zeros(self.C, **f64)
        self.elem_per_ch = torch.zeros(self.C, dtype=torch.int64, device=device)
        self.acc_count = torch.zeros(self.C, dtype=torch.int64, device=device)

    @torch.no_grad()
    def update(self, y_true: torch.Tensor, y_pred: torch.Tensor):
        """y_true, y_pred: [B, G, C*P] (or [G, C*P] for one sample)."""
        if y_true.dim() == 2:
            y_true, y_pred = y_true.unsqueeze(0), y_pred.unsqueeze(0)
        _require_cuda(y_true, "y_true")
        _require_cuda(y_pred, "y_pred")
        yt, yp = y_true.to(torch.float32).contiguous(), y_pred.to(torch.float32).contiguous()
        B, G, CP = yt.shape
        P = CP // self.C
        lib = _cabi.load()
        stats = torch.empty((B, CP, 3), dtype=torch.float64, device=yt.device)    # sum err^2, sum |err|, ACC
        nb = lib.gcl_forecast_metrics_workspace_bytes(B, CP)
        ws = torch.empty(nb, dtype=torch.uint8, device=yt.device)
        with torch.cuda.device(yt.device):
            _cabi.check(lib.gcl_forecast_metrics_f32(yt.data_ptr(), yp.data_ptr(), stats.data_ptr(), B, G, CP,
                                                     ws.data_ptr(), nb,
                                                     torch.cuda.current_stream(yt.device).cuda_stream),
                        "gcl_forecast_metrics_f32")
        per_ch = lambda t: t.view(B, P, self.C).sum(dim=(0, 1))                    # columns c -> channel c % C
        se, ae, corr = stats[..., 0], stats[..., 1], stats[..., 2]
        self.sum_se_per_ch += per_ch(se)
        self.sum_acc += per_ch(corr)
        self.elem_per_ch += B * P * G
        self.acc_count += B * P
        keep_cols = self.keep.repeat(P)
        self.sum_se += se[:, keep_cols].sum()
        self.sum_ae += ae[:, keep_cols].sum()
        self.total_elem += B * G * P * self.n_keep
        self.n += B

    def result(self) -> dict:
        """Reads the accumulators back (the only synchronisation)."""
        mse = float(self.sum_se) / max(self.total_elem, 1)
        acc_pc = (self.sum_acc / self.acc_count.clamp(min=1)).cpu()
        rmse_pc = (self.sum_se_per_ch / self.elem_per_ch.clamp(min=1)).sqrt().cpu()
        keep = self.keep.cpu()
        return {"n": self.n, "mse": mse, "rmse": mse ** 0.5, "mae": float(self.sum_ae) / max(self.total_elem, 1),
                "acc": float(acc_pc[keep].mean()) if bool(keep.any()) else 0.0,
                "acc_per_channel": acc_pc.numpy(), "rmse_per_channel": rmse_pc.numpy()}
