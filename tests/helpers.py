"""Shared helpers for the parity tests."""
import numpy as np
import torch

# tolerances stated by BASELINE.json north_star for fp32 layer outputs / gradients
RTOL_F32 = 1e-4


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| relative to the scale of b (max-norm): the 'rel 1e-4' of the north star."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)


def assert_close(a, b, tol=RTOL_F32, what="", atol=0.0):
    """atol: absolute floor for tensors that are analytically ~0 (e.g. d att_dst when every logit of a
    softmax row shares the LeakyReLU branch: both sides are rounding noise around 1e-12)."""
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    e = rel_err(a, b)
    if atol and float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max()) <= atol:
        return
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e}"


def random_graph(n, e, seed, self_loops=0, dups=0, isolated=0, heavy=None):
    """int64 [2, E] with optional existing self loops, duplicate edges, isolated nodes, one heavy row."""
    g = torch.Generator().manual_seed(seed)
    hi = max(n - isolated, 1)
    src = torch.randint(0, hi, (e,), generator=g)
    dst = torch.randint(0, hi, (e,), generator=g)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    parts = [torch.stack([src, dst])]
    if self_loops:
        v = torch.randint(0, hi, (self_loops,), generator=g)
        parts.append(torch.stack([v, v]))
    if dups and src.numel():
        idx = torch.randint(0, src.numel(), (dups,), generator=g)
        parts.append(torch.stack([src[idx], dst[idx]]))
    if heavy:
        s = torch.randint(0, hi, (heavy,), generator=g)
        s = s[s != 0]
        parts.append(torch.stack([s, torch.zeros_like(s)]))
    ei = torch.cat(parts, dim=1)
    perm = torch.randperm(ei.size(1), generator=g)
    return ei[:, perm].contiguous()


def np_csr_from_pyg(ei: np.ndarray, n: int, loops: bool):
    """Reference CSR for the bit-exact test: PyG order = kept edges then loops; stable sort by receiver."""
    src, dst = ei[0], ei[1]
    if loops:
        keep = src != dst
        src = np.concatenate([src[keep], np.arange(n)])
        dst = np.concatenate([dst[keep], np.arange(n)])
    order = np.argsort(dst, kind="stable")
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, dst + 1, 1)
    rowptr = np.cumsum(rowptr)
    order_t = np.argsort(src, kind="stable")
    rowptr_t = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr_t, src + 1, 1)
    rowptr_t = np.cumsum(rowptr_t)
    return dict(src=src, dst=dst, rowptr=rowptr, col=src[order], perm=order,
                rowptr_t=rowptr_t, col_t=dst[order_t], perm_t=order_t)
