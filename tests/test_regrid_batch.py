"""mean_regrid (justplotit.py:31-63) and the batched thermal forward model (driver.py:176-245 per-sample
loop; BASELINE cfg5).  Rebinning is integer bookkeeping + sums in bincount order: bit-exact against scipy.
The thermal stack: fp64 rtol 1e-6 (north_star)."""
import numpy as np
import pytest
from scipy.stats import binned_statistic

from oracle import regrid as oreg
from picaso_b200 import synth
from picaso_b200.batch import shard
from picaso_b200.regrid import bin_edges, plan_ranges
from util import assert_close


def grids():
    rng = np.random.default_rng(11)
    x_up = np.sort(rng.uniform(300.0, 9000.0, size=4000))
    cases = {
        "ascending_newx": (x_up, dict(newx=np.linspace(500.0, 8000.0, 57))),
        "descending_x": (x_up[::-1].copy(), dict(newx=np.linspace(200.0, 9500.0, 33))),   # empty edge bins
        "constant_R": (x_up, dict(R=40.0)),
        "ragged_newx": (x_up, dict(newx=np.sort(rng.uniform(1000.0, 7000.0, size=21)))),
        "on_right_edge": (np.linspace(100.0, 200.0, 101), dict(newx=np.linspace(105.0, 195.0, 10))),  # x == last edge
        "empty_bins": (np.linspace(100.0, 200.0, 11), dict(newx=np.linspace(100.0, 200.0, 60))),
    }
    return cases


@pytest.mark.parametrize("name", sorted(grids()))
def test_plan_matches_scipy_bin_numbers(name):
    x, kw = grids()[name]
    edges = bin_edges(x, **kw)
    start, count = plan_ranges(x, edges)
    _, _, binnum = binned_statistic(x, x, bins=edges)
    nb = edges.size - 1
    for i in range(nb):
        idx = np.nonzero(binnum == i + 1)[0]
        assert idx.size == count[i]
        if idx.size:
            assert idx[0] == start[i] and idx[-1] == start[i] + count[i] - 1


def test_plan_matches_scipy_on_random_grids():
    """property test: random monotone model grids (both directions, duplicates allowed) against random data
    grids - bin membership identical to scipy's for every bin"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1), st.integers(5, 400), st.integers(2, 40), st.booleans())
    def check(seed, n, nb, descending):
        rng = np.random.default_rng(seed)
        x = np.sort(np.round(rng.uniform(0.0, 100.0, size=n), rng.integers(0, 4)))
        if descending:
            x = x[::-1].copy()
        lo, hi = sorted(rng.uniform(-10.0, 110.0, size=2))
        if hi - lo < 1e-3:
            hi = lo + 1.0
        edges = np.unique(np.concatenate([[lo, hi], rng.uniform(lo, hi, size=nb - 1)]))
        if edges.size < 2:
            return
        start, count = plan_ranges(x, edges)
        _, _, binnum = binned_statistic(x, x, bins=edges)
        for i in range(edges.size - 1):
            idx = np.nonzero(binnum == i + 1)[0]
            assert idx.size == count[i]
            if idx.size:
                assert idx[0] == start[i] and idx[-1] == start[i] + count[i] - 1

    check()


def test_plan_rejects_non_monotonic_x():
    from picaso_b200 import PicasoB200Error
    x = np.array([1.0, 5.0, 2.0, 6.0, 3.0])
    with pytest.raises(PicasoB200Error):
        plan_ranges(x, np.array([0.0, 4.0, 8.0]))


def test_shard_covers_batch():
    for n, world in ((1024, 8), (10, 3), (2, 4)):
        parts = [shard(n, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(grids()))
def test_gpu_mean_regrid_bit_exact(name):
    import picaso_b200 as pb
    x, kw = grids()[name]
    rng = np.random.default_rng(3)
    y = rng.lognormal(size=(3, x.size))
    nx, got = pb.mean_regrid(x, y, **kw)
    for b in range(3):
        ox, oy = oreg.mean_regrid(x, y[b], **kw)
        assert np.array_equal(nx, ox)
        assert np.array_equal(got[b], oy, equal_nan=True), name
    _, one = pb.mean_regrid(x, y[1], **kw)
    assert np.array_equal(one, got[1], equal_nan=True)


def _stack(B, L, W, seed):
    ds = [synth.thermal_inputs(L=L, W=W, seed=seed + b, t_range=(300.0 + 10 * b, 1500.0 + 40 * b)) for b in range(B)]
    d0 = ds[0]
    return d0, dict(wno=d0["wno"], tlevel=np.array([d["tlevel"] for d in ds]), plevel=np.array([d["plevel"] for d in ds]),
                    dtau=np.array([d["dtau"] for d in ds]), w0=np.array([d["w0"] for d in ds]),
                    cosb=np.array([d["cosb"] for d in ds]), ubar1=d0["ubar1"], gweight=d0["gweight"], tweight=d0["tweight"])


@pytest.mark.gpu
def test_gpu_thermal_batch_vs_oracle():
    import picaso_b200 as pb
    d0, kw = _stack(B=7, L=20, W=900, seed=500)
    newx = np.linspace(600.0, 9000.0, 40)
    x, y = pb.thermal_batch(**kw, newx=newx, scale=1e-8 * 0.37 ** 2)
    ox, oy = oreg.thermal_batch(**kw, newx=newx, scale=1e-8 * 0.37 ** 2, nthreads=4)
    assert np.array_equal(x, ox)
    assert_close(y, oy, 1e-6, "thermal_batch rebinned")
    x2, y2 = pb.thermal_batch(**kw)          # no rebinning: the compress_thermal stack itself
    ox2, oy2 = oreg.thermal_batch(**kw, nthreads=4)
    assert_close(y2, oy2, 1e-6, "thermal_batch native grid")
    # atmospheres are independent: any shard of the batch reproduces its rows bit for bit
    lo, hi = shard(7, 1, 3)
    sub = dict(kw, tlevel=kw["tlevel"][lo:hi], plevel=kw["plevel"][lo:hi], dtau=kw["dtau"][lo:hi],
               w0=kw["w0"][lo:hi], cosb=kw["cosb"][lo:hi])
    _, ys = pb.thermal_batch(**sub, newx=newx, scale=1e-8 * 0.37 ** 2)
    assert np.array_equal(ys, y[lo:hi])
