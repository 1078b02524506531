"""Device-resident inference: autoregressive rollout + streaming forecast metrics.

Mirrors the inference loop of /root/reference/scripts/predict.py:451-614 -- residual add, static-channel
carry-forward, forcing channels from the ground truth, window slide (:534-580); StreamingMetrics (:53-124) -- for B
samples at once and without the reference's per-step `.cpu()` round trips (:499-559): everything stays on the
tensors' device until result() is read.  Pure tensor code on top of the model's forward; the model is any callable
[B, G, OBS*C] -> [B, G, C] (gcl_b200.model.WeatherPrediction on the GPU).
"""
from typing import Callable, Optional, Sequence

import torch


@torch.no_grad()
def rollout(model: Callable, X: torch.Tensor, ar_steps: int, num_channels: int, obs_window: int,
            y: Optional[torch.Tensor] = None, static_ch: Sequence[int] = (), forcing_ch: Sequence[int] = (),
            residual: bool = True, **model_kwargs) -> torch.Tensor:
    """X [B, G, OBS*C] -> forecasts [B, G, ar_steps*C]; y [B, G, steps*C] supplies the forcing channels."""
    if X.dim() != 3 or X.shape[-1] != obs_window * num_channels:
        raise ValueError(f"gcl_b200.predict.rollout: X {tuple(X.shape)} is not [B, G, {obs_window}*{num_channels}]")
    B, G, C = X.shape[0], X.shape[1], num_channels
    curr = X.view(B, G, obs_window, C)
    y_f = y.view(B, G, -1, C) if (len(forcing_ch) and y is not None) else None
    static_idx = torch.as_tensor(list(static_ch), dtype=torch.long, device=X.device)
    forcing_idx = torch.as_tensor(list(forcing_ch), dtype=torch.long, device=X.device)
    outs = []
    for step in range(ar_steps):
        delta = model(curr.reshape(B, G, obs_window * C), **model_kwargs)
        if delta.dim() == 2:
            delta = delta.unsqueeze(0)
        last = curr[:, :, -1, :]
        step_out = last + delta if residual else delta.clone()
        if static_idx.numel():
            step_out[:, :, static_idx] = last[:, :, static_idx]
        if y_f is not None and step < y_f.shape[2] and forcing_idx.numel():
            step_out[:, :, forcing_idx] = y_f[:, :, step, :][:, :, forcing_idx]
        outs.append(step_out)
        curr = torch.cat([curr[:, :, 1:, :], step_out.unsqueeze(2)], dim=2)
    return torch.cat(outs, dim=-1)


def persistence(X: torch.Tensor, num_channels: int, horizons: int) -> torch.Tensor:
    """Baseline forecast: the last observed step repeated.  X [B, G, OBS*C] -> [B, G, horizons*C]."""
    return X[..., -num_channels:].repeat(1, 1, horizons)


class StreamingMetrics:
    """MSE / RMSE / MAE over the non-excluded channels, per-channel RMSE and spatial anomaly correlation (ACC),
    accumulated in float64 on the device; update() takes a whole batch and never synchronises."""

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
        yt, yp = y_true.float(), y_pred.float()
        B, G, CP = yt.shape
        P = CP // self.C
        err = yp - yt
        se = err.pow(2).sum(dim=1)                                   # [B, CP], fp32 sums like the reference's
        ae = err.abs().sum(dim=1)
        yt_a, yp_a = yt - yt.mean(dim=1, keepdim=True), yp - yp.mean(dim=1, keepdim=True)
        corr = (yt_a * yp_a).sum(dim=1) / (yt_a.norm(dim=1) * yp_a.norm(dim=1) + 1e-8)
        per_ch = lambda t: t.double().view(B, P, self.C).sum(dim=(0, 1))          # columns c -> channel c % C
        self.sum_se_per_ch += per_ch(se)
        self.sum_acc += per_ch(corr)
        self.elem_per_ch += B * P * G
        self.acc_count += B * P
        keep_cols = self.keep.repeat(P)
        self.sum_se += se.double()[:, keep_cols].sum()
        self.sum_ae += ae.double()[:, keep_cols].sum()
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
