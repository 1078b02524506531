"""f4 (inference rollout + streaming metrics): the product's rollout glue and metric reduction are CUDA kernels, so the
product-vs-oracle comparisons here are GPU tests; the oracle itself is pinned on the CPU against the reference's own
StreamingMetrics class."""
import os
import re

import numpy as np
import pytest
import torch

from gcl_b200 import predict as gp
from oracle import predict as op

REF = "/root/reference/scripts/predict.py"


def _data(B=3, G=40, C=5, P=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, G, C * P, generator=g), torch.randn(B, G, C * P, generator=g)


@pytest.mark.gpu
@pytest.mark.parametrize("exclude", [(), (1, 4)])
def test_streaming_metrics_match_oracle(exclude):
    yt, yp = _data()
    yp = 0.7 * yt + 0.3 * yp + 2.0                 # correlated forecasts with an offset: ACC away from 0, means matter
    mine, ref = gp.StreamingMetrics(5, exclude, device="cuda:0"), op.StreamingMetrics(5, list(exclude))
    for lo in (0, 2):                              # two updates of different batch sizes
        mine.update(yt[lo:lo + 2].cuda(), yp[lo:lo + 2].cuda())
        for b in range(lo, min(lo + 2, yt.shape[0])):
            ref.update(yt[b], yp[b])
    r = mine.result()
    assert r["n"] == ref.n == 3
    for k in ("mse", "rmse", "mae", "acc"):
        assert abs(r[k] - getattr(ref, k)) <= 1e-6 * max(1.0, abs(getattr(ref, k))), k
    np.testing.assert_allclose(r["acc_per_channel"], ref.acc_per_channel, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(r["rmse_per_channel"], ref.rmse_per_channel, rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("static_ch,forcing_ch,residual", [((), (), True), ((0,), (3,), True), ((), (2,), False)])
def test_rollout_matches_oracle(static_ch, forcing_ch, residual):
    B, G, C, OBS, AR = 3, 17, 4, 2, 3
    g = torch.Generator().manual_seed(1)
    X = torch.randn(B, G, OBS * C, generator=g)
    y = torch.randn(B, G, 2 * C, generator=g)     # ground truth for 2 of the 3 steps only (predict.py:564)
    W = torch.randn(OBS * C, C, generator=g) * 0.3
    model = lambda inp: torch.tanh(inp @ W)        # stands in for the forecast model: [.., G, OBS*C] -> [.., G, C]
    Wg = W.cuda()
    out = gp.rollout(lambda inp: torch.tanh(inp @ Wg), X.cuda(), AR, C, OBS, y=y.cuda(), static_ch=static_ch,
                     forcing_ch=forcing_ch, residual=residual).cpu()
    assert out.shape == (B, G, AR * C)
    for b in range(B):
        ref = op.ar_rollout(model, X[b:b + 1].clone(), AR, C, OBS, y=y[b], static_ch=static_ch, forcing_ch=forcing_ch,
                            residual=residual)
        assert torch.allclose(out[b], ref, rtol=0, atol=2e-6), (b, float((out[b] - ref).abs().max()))
    assert torch.equal(gp.persistence(X, C, 2)[1], op.persistence(X[1:2], C, 2))
    with pytest.raises(ValueError):
        gp.rollout(model, X[..., :-1], AR, C, OBS)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gp.rollout(model, X, AR, C, OBS)


@pytest.mark.reference
@pytest.mark.skipif(not os.path.exists(REF), reason="/root/reference not present")
def test_oracle_metrics_equal_the_reference_class():
    """The reference's own StreamingMetrics (source lines cut out of the unmodified scripts/predict.py, which cannot
    be imported as a module here) against the restatement: identical accumulators."""
    src = open(REF).read()
    m = re.search(r"^class StreamingMetrics:.*?(?=^def main\(\))", src, re.S | re.M)
    assert m, "StreamingMetrics not found in the reference"
    ns = {"np": np, "torch": torch}
    exec(compile(m.group(0), REF, "exec"), ns)
    theirs, ours = ns["StreamingMetrics"](5, [2]), op.StreamingMetrics(5, [2])
    yt, yp = _data(B=4, seed=3)
    for b in range(4):
        theirs.update(yt[b], yp[b])
        ours.update(yt[b], yp[b])
    assert theirs.n == ours.n and theirs.total_elem == ours.total_elem
    assert theirs.sum_se == ours.sum_se and theirs.sum_ae == ours.sum_ae
    for k in ("sum_se_per_ch", "elem_per_ch", "sum_acc", "acc_count"):
        assert np.array_equal(getattr(theirs, k), getattr(ours, k)), k
    assert theirs.rmse == ours.rmse and theirs.acc == ours.acc and theirs.mae == ours.mae
