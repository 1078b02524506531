"""ORACLE / TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's inference rollout and streaming
metrics.  Only tests/ may import this; the product is gcl_b200/predict.py.

Follows /root/reference/scripts/predict.py:
  * StreamingMetrics            :53-124  (per-column squared error, spatial anomaly correlation, aggregate MSE/MAE
                                          over the non-excluded channels, float64 host accumulators)
  * the AR rollout of main()    :534-580 (residual add, static-channel carry-forward, forcing channels taken from
                                          the ground truth, window slide), persistence baseline :472
Pinned: tests/test_oracle.py executes the reference's own StreamingMetrics class (its source lines, extracted from
the unmodified file) on the same inputs and demands identical accumulators.
"""
import numpy as np
import torch


class StreamingMetrics:
    def __init__(self, num_channels, exclude_channels=None):
        self.C = num_channels
        self.exclude_channels = set(exclude_channels or [])
        self.n = 0
        self.total_elem = 0
        self.sum_se = 0.0
        self.sum_ae = 0.0
        self.sum_se_per_ch = np.zeros(num_channels, dtype=np.float64)
        self.elem_per_ch = np.zeros(num_channels, dtype=np.int64)
        self.sum_acc = np.zeros(num_channels, dtype=np.float64)
        self.acc_count = np.zeros(num_channels, dtype=np.int64)

    def update(self, y_true, y_pred):
        """y_true, y_pred: [G, C*P] (one sample, as the reference's batch-1 loop feeds it)."""
        err = y_pred.float() - y_true.float()
        cp = y_true.shape[1]
        for c in range(cp):                                   # predict.py:75-88
            yt, yp = y_true[:, c].float(), y_pred[:, c].float()
            ch = c % self.C
            self.sum_se_per_ch[ch] += (yp - yt).pow(2).sum().item()
            self.elem_per_ch[ch] += yt.numel()
            yt_a, yp_a = yt - yt.mean(), yp - yp.mean()
            corr = (yt_a * yp_a).sum() / (yt_a.norm() * yp_a.norm() + 1e-8)
            self.sum_acc[ch] += corr.item()
            self.acc_count[ch] += 1
        dyn = [c for c in range(cp) if (c % self.C) not in self.exclude_channels]      # predict.py:91-97
        if dyn:
            e = err[:, dyn]
            self.sum_se += e.pow(2).sum().item()
            self.sum_ae += e.abs().sum().item()
            self.total_elem += e.numel()
        self.n += 1

    @property
    def mse(self):
        return self.sum_se / max(self.total_elem, 1)

    @property
    def rmse(self):
        return float(np.sqrt(self.mse))

    @property
    def mae(self):
        return self.sum_ae / max(self.total_elem, 1)

    @property
    def acc_per_channel(self):
        return self.sum_acc / np.maximum(self.acc_count, 1)

    @property
    def rmse_per_channel(self):
        return np.sqrt(self.sum_se_per_ch / np.maximum(self.elem_per_ch, 1))

    @property
    def acc(self):
        dyn = [c for c in range(self.C) if c not in self.exclude_channels]
        return float(self.acc_per_channel[dyn].mean()) if dyn else 0.0


def ar_rollout(model, X, ar_steps, num_channels, obs_window, y=None, static_ch=(), forcing_ch=(), residual=True):
    """One sample: X [1, G, OBS*C] -> [G, ar_steps*C]  (predict.py:534-580).  model(inp [1, G, OBS*C]) -> [G, C] or
    [1, G, C].  y [G, steps*C] supplies the forcing channels."""
    G, C = X.shape[1], num_channels
    curr = X.view(1, G, obs_window, C)
    outs = []
    y_f = y.view(y.shape[0], -1, C) if (forcing_ch and y is not None) else None
    for step in range(ar_steps):
        delta = model(curr.view(1, G, -1))
        if delta.dim() == 2:
            delta = delta.unsqueeze(0)
        step_out = curr[:, :, -1, :] + delta if residual else delta
        for ch in static_ch:
            step_out[:, :, ch] = curr[:, :, -1, ch]
        if y_f is not None and step < y_f.shape[1]:
            for ch in forcing_ch:
                step_out[:, :, ch] = y_f[:, step, ch].unsqueeze(0)
        outs.append(step_out)
        curr = torch.cat([curr[:, :, 1:, :], step_out.unsqueeze(2)], dim=2)
    return torch.cat(outs, dim=-1).squeeze(0)


def persistence(X, num_channels, horizons):
    """Baseline forecast: the last observed step repeated (predict.py:472).  X [1, G, OBS*C] -> [G, horizons*C]."""
    return X.squeeze(0)[:, -num_channels:].repeat(1, horizons)
