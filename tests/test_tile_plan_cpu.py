"""Host-side tile planner (gcl_tile_plan_host): structural invariants and a numpy emulation of the tiled SpMM that
must reproduce the plain CSR product -- the plan is pure index data, so all of this runs without a GPU."""
import numpy as np
import pytest
import torch

from gcl_b200.graph import TilePlan


def _random_csr(n, avg_deg, seed, heavy_row=None, heavy_deg=0):
    rng = np.random.default_rng(seed)
    deg = rng.poisson(avg_deg, size=n)
    deg[rng.random(n) < 0.2] = 0                       # empty rows
    if heavy_row is not None:
        deg[heavy_row] = heavy_deg
    rowptr = np.zeros(n + 1, np.int32)
    rowptr[1:] = np.cumsum(deg)
    col = rng.integers(0, n, size=int(rowptr[-1])).astype(np.int32)   # duplicates allowed
    return rowptr, col


def _arrays(pl):
    a = {k: v.numpy() for k, v in pl.t.items()}
    a["lidx"] = a["lidx"].view(np.uint16)
    return a


@pytest.mark.parametrize("pad", [1, 2, 4])
@pytest.mark.parametrize("n,avg,mr,mu,me,order_seed", [(200, 4, 16, 24, 64, None), (300, 7, 64, 128, 1024, 3),
                                                       (50, 2, 8, 8, 16, 1), (1, 0, 4, 4, 8, None)])
def test_plan_invariants_and_emulated_spmm(n, avg, mr, mu, me, order_seed, pad):
    rowptr, col = _random_csr(n, avg, seed=n)
    order = None if order_seed is None else np.random.default_rng(order_seed).permutation(n).astype(np.int32)
    n_out, n_in = n, n
    pl = TilePlan(torch.from_numpy(rowptr), torch.from_numpy(col if col.size else np.zeros(1, np.int32)),
                  int(rowptr[-1]), n, n_out, n_in, order, mr, mu, me, pad=pad)
    a = _arrays(pl)
    T = pl.n_tiles
    assert pl.max_rows <= mr and pl.max_union <= mu and pl.max_entries <= me
    rows = a["rows"][: pl.n_plan_rows]
    heavy = a["heavy_rows"][: pl.n_heavy]
    assert sorted(np.concatenate([rows, heavy]).tolist()) == list(range(n_out))      # every row exactly once
    if order is not None and pl.n_heavy == 0:
        assert np.array_equal(rows, order)                                           # tiles follow the hint
    x = np.random.default_rng(0).normal(size=(n, 3))
    w = np.random.default_rng(1).normal(size=max(int(rowptr[-1]), 1))
    ref = np.zeros((n, 3))
    for r in range(n):
        for k in range(rowptr[r], rowptr[r + 1]):
            ref[r] += w[k] * x[col[k]]
    out = np.zeros((n, 3))
    for t in range(T):
        r0, r1 = a["tile_rowptr"][t], a["tile_rowptr"][t + 1]
        u0, u1 = a["tile_uptr"][t], a["tile_uptr"][t + 1]
        assert 0 < r1 - r0 <= mr and u1 - u0 <= mu
        usrc = a["usrc"][u0:u1]
        assert len(set(usrc.tolist())) == len(usrc)                                  # a union lists a row once
        xs = x[usrc]                                                                 # the "shared memory" stage
        assert a["eptr"][r1] - a["eptr"][r0] <= me
        for i in range(r0, r1):
            r = rows[i]
            e0, e1 = a["eptr"][i], a["eptr"][i + 1]
            ln = rowptr[r + 1] - rowptr[r]
            assert e1 - e0 == (ln + pad - 1) // pad * pad and e0 % pad == 0             # padded to a multiple of pad
            assert np.array_equal(a["ek"][e0:e0 + ln], np.arange(rowptr[r], rowptr[r + 1]))   # ascending CSR order
            assert (a["ek"][e0 + ln:e1] == -1).all() and (a["lidx"][e0 + ln:e1] == 0xFFFF).all()
            for e in range(e0, e0 + ln):
                out[r] += w[a["ek"][e]] * xs[a["lidx"][e]]
        d = a["tile_desc"][8 * t: 8 * t + 8]
        assert d.tolist() == [r0, r1 - r0, u0, u1 - u0, a["eptr"][r0], a["eptr"][r1] - a["eptr"][r0], 0, 0]
    for r in heavy:
        assert len(set(col[rowptr[r]:rowptr[r + 1]].tolist())) > mu or rowptr[r + 1] - rowptr[r] + pad - 1 > me
        out[r] = ref[r]
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12)
    # packed {lidx, weight bits} entries the SpMM kernel reads: pads carry {0xFFFF, 0}
    ent = pl.entries(torch.from_numpy(w.astype(np.float32)), key="k").numpy()
    assert ent.shape == (max(pl.n_entries, 1), 2) and pl.entries(None, key="k") is not None
    real = a["ek"][: pl.n_entries] >= 0
    li = a["lidx"][: pl.n_entries].astype(np.int32)
    assert np.array_equal(ent[: pl.n_entries, 0], np.where(li == 0xFFFF, pl.max_union, li))
    assert np.array_equal(ent[: pl.n_entries, 1].view(np.float32)[real], w.astype(np.float32)[a["ek"][: pl.n_entries][real]])
    assert (ent[: pl.n_entries, 1][~real] == 0).all()


def test_heavy_rows_prefix_and_masked_columns():
    n = 120
    rowptr, col = _random_csr(n, 3, seed=9, heavy_row=7, heavy_deg=90)
    n_out, n_in = 100, 80          # rows >= 100 are not produced; columns >= 80 are zero rows
    pl = TilePlan(torch.from_numpy(rowptr), torch.from_numpy(col), int(rowptr[-1]), n, n_out, n_in, None, 16, 32, 64, pad=1)
    a = _arrays(pl)
    assert 7 in a["heavy_rows"][: pl.n_heavy].tolist()
    rows = a["rows"][: pl.n_plan_rows]
    assert rows.max() < n_out and sorted(np.concatenate([rows, a["heavy_rows"][: pl.n_heavy]]).tolist()) == list(range(n_out))
    for t in range(pl.n_tiles):
        u0, u1 = a["tile_uptr"][t], a["tile_uptr"][t + 1]
        assert (a["usrc"][u0:u1] < n_in).all()
        for i in range(a["tile_rowptr"][t], a["tile_rowptr"][t + 1]):
            for e in range(a["eptr"][i], a["eptr"][i + 1]):
                c = col[a["ek"][e]]
                if c >= n_in:
                    assert a["lidx"][e] == 0xFFFF
                else:
                    assert a["usrc"][u0 + a["lidx"][e]] == c


def test_bad_order_is_rejected():
    rowptr, col = _random_csr(10, 2, seed=2)
    with pytest.raises(RuntimeError, match="out of range"):
        TilePlan(torch.from_numpy(rowptr), torch.from_numpy(col), int(rowptr[-1]), 10, 10, 10,
                 np.full(10, 11, np.int32))


def test_tiled_entry_points_reject_bad_arguments_without_a_gpu():
    import ctypes
    from gcl_b200 import _cabi
    lib = _cabi.load()
    st = _cabi.TilePlanStruct()
    assert lib.gcl_spmm_tiled_f32(ctypes.byref(st), None, None, None, None, None, None, 1, 1, 4, 4, 4, None, None, None,
                                  None) == -1
    assert "null" in _cabi.last_error()
    assert lib.gcl_spmm_tiled_f32(None, None, None, None, None, None, None, 1, 1, 4, 4, 4, None, None, None, None) == -1
    assert "plan" in _cabi.last_error()
