"""One training step of the forecast model, data-parallel over the GPUs of a box.

Mirrors the loop body of /root/reference/src/train.py:160-235 (AR rollout with BPTT, residual
prediction, latitude-weighted MSE, Adam) for B samples per GPU:

  * residual add + weighted MSE + its gradient are one kernel (gcl_wmse_f32);
  * all parameters live in ONE flat fp32 buffer (views handed to the modules), so the gradient
    all-reduce is a single NCCL call over NVLink and Adam a single kernel (gcl_adam_f32);
  * samples shard over ranks (rank r takes its own B samples); graphs and weights are replicated.
    The reference has no multi-GPU code at all; effective batch = world x B.

Static / forcing channel carry-forward (train.py:218-226), the channel / spatial loss masks (train.py:85-102) and
use_residual=False (train.py:203-207) are Trainer options; the whole per-step glue (residual add, masked weighted MSE,
carry-forward, window slide) is one kernel forward (gcl_ar_step_f32) and one backward.
"""
from typing import Optional

import torch
import torch.distributed as dist

from . import _cabi


def lat_weights(nlat: int, nlon: int, device) -> torch.Tensor:
    """cos(lat)/mean per grid node, [G] (train.py:53-72: expanded [lon, lat] then flattened)."""
    w = torch.cos(torch.deg2rad(torch.linspace(-90, 90, nlat)))
    w = w / w.mean()
    return w.view(1, -1).expand(nlon, nlat).reshape(-1).contiguous().to(device)


class _ARStep(torch.autograd.Function):
    """(delta [B,G,C], state [B,G,obs,C], y_step view) -> (scale * weighted MSE, new_state): residual add, masked
    latitude-weighted MSE, static / forcing carry-forward and the window slide of train.py:201-227 in one kernel;
    the backward (one kernel) returns d delta and d state."""

    @staticmethod
    def forward(ctx, delta, state, y, node_w, chan_w, carry, residual, inv_wsum, scale, want_state):
        lib = _cabi.load()
        d = delta.contiguous()
        st = state.contiguous()
        B, G, C = d.shape
        obs = st.shape[2]
        if st.shape != (B, G, obs, C):
            raise RuntimeError(f"gcl_b200: state {tuple(st.shape)} does not match delta {tuple(d.shape)}")
        if y.stride(2) != 1 or y.stride(0) != G * y.stride(1):
            raise RuntimeError("gcl_b200: y must be a [B,G,C] view with unit channel stride")
        new_state = torch.empty_like(st) if want_state else None
        g_loss = torch.empty_like(d)
        loss = torch.empty(1, dtype=torch.float32, device=d.device)
        nb = lib.gcl_wmse_workspace_bytes(B, G, C)
        ws = torch.empty(nb, dtype=torch.uint8, device=d.device)
        p = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(d.device):
            _cabi.check(lib.gcl_ar_step_f32(d.data_ptr(), st.data_ptr(), y.data_ptr(), y.stride(1), p(node_w), p(chan_w),
                                            p(carry), int(bool(residual)), float(inv_wsum), float(scale), p(new_state),
                                            g_loss.data_ptr(), loss.data_ptr(), 0, B, G, obs, C, ws.data_ptr(), nb,
                                            torch.cuda.current_stream().cuda_stream), "gcl_ar_step_f32")
        ctx.save_for_backward(g_loss, carry)
        ctx.residual, ctx.shape, ctx.state_needs = bool(residual), (B, G, obs, C), ctx.needs_input_grad[1]
        if want_state:
            return loss.squeeze(0), new_state
        return loss.squeeze(0), None

    @staticmethod
    def backward(ctx, dloss, dnew=None):
        g_loss, carry = ctx.saved_tensors
        B, G, obs, C = ctx.shape
        lib = _cabi.load()
        d_delta = torch.empty_like(g_loss)
        d_state = torch.empty((B, G, obs, C), dtype=torch.float32, device=g_loss.device) if ctx.state_needs else None
        dl = dloss.reshape(1).to(torch.float32).contiguous() if dloss is not None else None
        dn = dnew.contiguous() if dnew is not None else None
        p = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(g_loss.device):
            _cabi.check(lib.gcl_ar_step_bwd_f32(g_loss.data_ptr(), p(dl), p(dn), p(carry), int(ctx.residual),
                                                d_delta.data_ptr(), p(d_state), B, G, obs, C,
                                                torch.cuda.current_stream().cuda_stream), "gcl_ar_step_bwd_f32")
        return d_delta, d_state, None, None, None, None, None, None, None, None


class Trainer:
    """Owns the flat parameter / gradient / Adam-state buffers of `model` and runs training steps."""

    def __init__(self, model: torch.nn.Module, nlat: int, nlon: int, lr: float = 1e-3, ar_steps: int = 1,
                 use_latitude_weighting: bool = True, use_residual: bool = True, betas=(0.9, 0.999),
                 eps: float = 1e-8, process_group=None, static_channels=(), forcing_channels=(),
                 channel_mask: Optional[torch.Tensor] = None, spatial_mask: Optional[torch.Tensor] = None):
        """static_channels / forcing_channels: channel indices carried forward from the last input step / taken from
        the target during the rollout (train.py:218-226).  channel_mask [C], spatial_mask [G] (or the reference's
        [1,G,1]): loss weights of weighted_mse_loss (train.py:85-102)."""
        self.model = model
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        self.ar_steps, self.use_residual = int(ar_steps), use_residual
        self.pg = process_group
        for m in model.modules():            # SparseGATConv prunes on the attention averaged over THIS group
            if hasattr(m, "process_group"):
                m.process_group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        params, seen = [], set()
        for p in model.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        self.params = params
        dev = params[0].device
        # every parameter starts on a 128-byte boundary of the flat buffer: the kernels take their 128-bit
        # paths only for 16-byte aligned bias / attention vectors (padding stays 0 and gets 0 gradient)
        ALIGN = 32
        n = sum((p.numel() + ALIGN - 1) // ALIGN * ALIGN for p in params)
        self.flat_param = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        off = 0
        self._grad_views = []
        for p in params:
            k = p.numel()
            self.flat_param[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat_param[off:off + k].view_as(p)
            p.grad = self.flat_grad[off:off + k].view_as(p)
            self._grad_views.append(p.grad)
            off += (k + ALIGN - 1) // ALIGN * ALIGN
        self.num_params = sum(p.numel() for p in params)     # trainable scalars (the reference's count)
        self.flat_size = n                                   # incl. alignment padding
        self.lat_w = lat_weights(nlat, nlon, dev) if use_latitude_weighting else None
        self.G = nlat * nlon
        # per-node loss weight = latitude weight x spatial mask; per-channel weight = channel mask
        self.node_w = self.lat_w
        if spatial_mask is not None:
            sm = spatial_mask.to(dev, torch.float32).reshape(-1)
            self.node_w = sm if self.node_w is None else (self.node_w * sm).contiguous()
        self.chan_w = channel_mask.to(dev, torch.float32).reshape(-1).contiguous() if channel_mask is not None else None
        self._wsum = float(self.node_w.sum()) if self.node_w is not None else float(self.G)  # one-time sync
        self._csum = float(self.chan_w.sum()) if self.chan_w is not None else None
        self.static_channels, self.forcing_channels = tuple(static_channels), tuple(forcing_channels)
        self._carry = None
        if self.world > 1:   # identical replicas: take rank 0's initial weights
            dist.broadcast(self.flat_param, src=0, group=self.pg)

    # -- pieces, so a caller (bench.py) can graph-capture forward+backward and keep the collective eager
    def loss(self, X: torch.Tensor, y: torch.Tensor, attention_threshold: float = 0.0, **kwargs) -> torch.Tensor:
        """Mean over AR steps of the weighted MSE (train.py:173-231).  X [B,G,obs*C], y [B,G,steps*C]."""
        model = self.model
        B, G, _ = X.shape
        obs = model.obs_window
        C = X.shape[-1] // obs
        tsteps = y.shape[-1] // C
        steps = min(self.ar_steps, tsteps)
        if self.chan_w is not None and self.chan_w.numel() != C:
            raise ValueError(f"gcl_b200.Trainer: channel_mask has {self.chan_w.numel()} entries, the data {C} channels")
        wsum = max(self._wsum * B * (C if self._csum is None else self._csum), 1e-12)   # train.py:101
        if self._carry is None:
            carry = torch.zeros(C, dtype=torch.int32)
            for ch in self.static_channels:
                carry[ch] = 1
            for ch in self.forcing_channels:          # the reference applies forcing after static (train.py:222-226)
                carry[ch] = 2
            self._carry = carry.to(X.device) if (self.static_channels or self.forcing_channels) else False
        carry = self._carry if self._carry is not False else None
        ys = y.view(B, G, tsteps, C)
        state = X.view(B, G, obs, C)
        total = None
        for s in range(steps):
            delta = model(X=state.reshape(B, G, obs * C), attention_threshold=attention_threshold, **kwargs)
            if delta.dim() == 2:
                delta = delta.unsqueeze(0)
            last = s == steps - 1
            l, state_next = _ARStep.apply(delta, state, ys[:, :, s, :], self.node_w, self.chan_w, carry,
                                          self.use_residual, 1.0 / wsum, 1.0 / steps, not last)
            total = l if total is None else total + l
            if not last:
                state = state_next
        return total

    def zero_grad(self):
        self.flat_grad.zero_()

    def backward(self, loss: torch.Tensor):
        """loss.backward() with the gradients landing in the flat buffer through ONE multi-tensor copy: with
        .grad unset autograd hands each parameter its gradient tensor as is, instead of one `grad += g` kernel per
        parameter (40-odd tiny launches per step) on top of a zero fill."""
        for p in self.params:
            p.grad = None
        loss.backward()
        have = [(v, p.grad) for v, p in zip(self._grad_views, self.params) if p.grad is not None]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v, p in zip(self._grad_views, self.params):
            if p.grad is None:
                v.zero_()                 # parameter not reached by this loss
            p.grad = v

    def reduce_gradients(self):
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.pg)

    def optimizer_step(self):
        lib = _cabi.load()
        with torch.cuda.device(self.flat_param.device):
            _cabi.check(lib.gcl_adam_f32(self.flat_param.data_ptr(), self.flat_grad.data_ptr(),
                                         self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.flat_size,
                                         self.lr, self.betas[0], self.betas[1], self.eps, 1.0 / self.world,
                                         self.step_count.data_ptr(), torch.cuda.current_stream().cuda_stream),
                        "gcl_adam_f32")

    # -- CUDA-graph fast path: forward + backward of a fixed batch shape captured once, replayed per step
    def capture(self, batch: int, tf: int, yf: int, attention_threshold: float = 0.0, warmup: int = 2,
                whole_step: bool = True):
        """Allocate static input buffers [batch, G, tf] / [batch, G, yf] and capture the training step into one CUDA
        graph (every gcl_* entry point is enqueue-only).  whole_step: the gradient all-reduce (NCCL, one flat bucket)
        and the Adam kernel are captured too, so a step is ONE graph launch; if the collective cannot be captured on
        this build the graph ends after the backward and the tail runs eagerly (one NCCL call + two tiny kernels)."""
        dev = self.flat_param.device
        self.static_x = torch.zeros(batch, self.G, tf, dtype=torch.float32, device=dev)
        self.static_y = torch.zeros(batch, self.G, yf, dtype=torch.float32, device=dev)
        self.static_x.normal_()
        self.static_y.normal_()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):      # builds the CSR caches (they sync) before capture
                self.backward(self.loss(self.static_x, self.static_y, attention_threshold))
                if whole_step and self.world > 1:            # the communicator must exist before capture
                    dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.pg)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        lib = _cabi.load()

        def record(tail: bool):
            before = lib.gcl_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.static_loss = self.loss(self.static_x, self.static_y, attention_threshold)
                self.backward(self.static_loss)
                if tail:
                    self.reduce_gradients()
                    self.optimizer_step()
            return g, int(lib.gcl_launch_count() - before)

        saved = (self.flat_param.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.step_count.clone())
        self._graph_has_tail = False
        if whole_step:
            try:
                self.graph, self.launches_in_graph = record(True)
                self._graph_has_tail = True
            except RuntimeError:                              # e.g. a NCCL build that refuses stream capture
                torch.cuda.synchronize(dev)
        if not self._graph_has_tail:
            self.graph, self.launches_in_graph = record(False)
        # capture does not execute, but keep the optimiser state exactly as it was in any case
        for dst, src in zip((self.flat_param, self.exp_avg, self.exp_avg_sq, self.step_count), saved):
            dst.copy_(src)
        return self

    def step_captured(self) -> torch.Tensor:
        """One training step on whatever is in static_x / static_y (device resident)."""
        self.graph.replay()
        if not self._graph_has_tail:
            self.reduce_gradients()
            self.optimizer_step()
        return self.static_loss

    def prefetch(self, X_pinned: torch.Tensor, y_pinned: torch.Tensor):
        """Start the H2D copy of a batch into staging buffers on a side stream (overlaps whatever the compute
        stream is doing); the next step_from_host() consumes it with a device-to-device copy."""
        dev = self.flat_param.device
        if not hasattr(self, "_stage_x"):
            self._stage_x, self._stage_y = torch.empty_like(self.static_x), torch.empty_like(self.static_y)
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staged, self._consumed = torch.cuda.Event(), torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(dev))
        self._copy_stream.wait_event(self._consumed)          # the previous staged batch has left the buffers
        with torch.cuda.stream(self._copy_stream):
            self._stage_x.copy_(X_pinned, non_blocking=True)
            self._stage_y.copy_(y_pinned, non_blocking=True)
            self._staged.record(self._copy_stream)
        self._has_staged = True

    def step_from_host(self, X_pinned: torch.Tensor, y_pinned: torch.Tensor, next_batch=None) -> float:
        """End-to-end step from pinned host buffers: H2D copies, captured step, loss read back.  If the batch was
        prefetch()ed it is taken from the staging buffers; next_batch = (X, y) starts the following step's H2D
        copy while this step's kernels run."""
        cur = torch.cuda.current_stream(self.flat_param.device)
        if getattr(self, "_has_staged", False):
            cur.wait_event(self._staged)
            self.static_x.copy_(self._stage_x, non_blocking=True)
            self.static_y.copy_(self._stage_y, non_blocking=True)
            self._consumed.record(cur)
            self._has_staged = False
        else:
            self.static_x.copy_(X_pinned, non_blocking=True)
            self.static_y.copy_(y_pinned, non_blocking=True)
        loss = self.step_captured()
        if next_batch is not None:
            self.prefetch(*next_batch)
        return float(loss.item())

    def step_from_device(self, X: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """Captured step on device-resident tensors (what gcl_b200.data.ChunkedWindowLoader.batches() yields): two
        device-to-device copies into the graph's input buffers, then the replay.  Returns the loss tensor (no sync)."""
        self.static_x.copy_(X, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        return self.step_captured()

    def step(self, X: torch.Tensor, y: torch.Tensor, attention_threshold: float = 0.0, **kwargs) -> torch.Tensor:
        """forward + backward + gradient all-reduce + Adam; returns the (local) loss as a 0-d tensor."""
        loss = self.loss(X, y, attention_threshold, **kwargs)
        self.backward(loss)
        self.reduce_gradients()
        self.optimizer_step()
        return loss.detach()
