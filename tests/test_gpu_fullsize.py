"""GPU: every BASELINE.json configuration at its FULL size against the CPU oracle (on a strided wavelength /
atmosphere sample where the oracle would need minutes), rtol 1e-6, plus the SH4 tile kernel against the per-angle
kernel.  The headline (60 x 10 000 x 5 reflected) is in test_gpu_parity.py::test_full_size_properties."""
import numpy as np
import pytest

import cases as C
import oracle
import picaso_b200 as pb
from picaso_b200 import synth
from util import assert_close

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _subset(d, idx):
    ds = dict(d)
    for k, v in d.items():
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[-1] == d["nwno"]:
            ds[k] = np.ascontiguousarray(v[..., idx])
    ds["nwno"] = len(idx)
    return ds


def test_cfg2_thermal_full_size():
    """BASELINE config 2: thermal Toon, 90 layers x 10 000 waves x 5 angles, calc_type 0 and 1"""
    d = synth.thermal_inputs(L=90, W=10000, seed=1002)
    for ct in (0, 1):
        t = dict(d, calc_type=ct)
        ftop, none, th = pb.get_thermal_1d(*C.thermal_args(t), level_fluxes=False, gweight=d["gweight"],
                                           tweight=d["tweight"], return_thermal=True)
        oft, _ = oracle.get_thermal_1d(*C.thermal_args(t), nthreads=8, level_fluxes=False)
        assert_close(ftop, oft, RTOL, "cfg2 flux_at_top calc_type=%d" % ct)
        assert_close(th, oracle.compress_thermal(d["nwno"], oft, d["gweight"], d["tweight"]), RTOL, "cfg2 thermal")


@pytest.mark.parametrize("forms", [(1, 1, 1, 1, 1, 1), (0, 0, 0, 1, 1, 1)])
def test_cfg3_sh4_full_size(forms):
    """BASELINE config 3: SH4 reflected, 60 layers x 196 000 waves x 5 angles; OTHG forms (tile kernel) and the
    reference-default TTHG forms (per-angle kernel, Appendix-A1 drift); oracle on every 128th wavelength"""
    d = synth.reflected_inputs(L=60, W=196000, seed=1003, ngauss=5, stream=4)
    case = dict(forms=forms, stream=4, single_form=0)
    fd0 = d["f_deltaM"].copy()
    xint, flux, alb = pb.get_reflected_SH(*C.sh_args(d, case), gweight=d["gweight"], tweight=d["tweight"],
                                          return_albedo=True)
    assert np.isfinite(xint).all()
    idx = np.arange(0, 196000, 128)
    ds = _subset(dict(d, f_deltaM=fd0), idx)
    ox, _ = oracle.get_reflected_SH(*C.sh_args(ds, case), nthreads=8)
    assert_close(xint[..., idx], ox, RTOL, "cfg3 xint forms=%s" % (forms,))
    oalb = oracle.compress_disco(len(idx), ds["cos_theta"], ox, ds["gweight"], ds["tweight"], ds["F0PI"])
    assert_close(alb[idx], oalb, RTOL, "cfg3 albedo")


@pytest.mark.parametrize("L,W,G,nt,phase,surf", [(60, 300, 5, 1, 0.0, 0.0), (7, 77, 7, 1, 0.6, 0.3), (1, 33, 8, 1, 0.0, 0.2),
                                                  (23, 40, 6, 2, 1.1, 0.0)])
def test_sh_tile_vs_per_angle_kernel(L, W, G, nt, phase, surf, monkeypatch):
    """sh4_tile_kernel (angle-shared elimination, sh_reflected_tile.cuh) against sh_reflected_kernel<4> and the oracle;
    the (6, 2) geometry has 12 angles = two angle groups per wavelength tile (unfused disk integration)"""
    d = synth.reflected_inputs(L=L, W=W, seed=700 + L, ngauss=G, stream=4, phase=phase)
    d["surf_reflect"] = np.full(W, surf)
    if nt > 1:
        d["numt"] = nt
        d["ubar0"] = np.ascontiguousarray(np.hstack([d["ubar0"] * (1.0 - 0.07 * t) for t in range(nt)]))
        d["ubar1"] = np.ascontiguousarray(np.hstack([d["ubar1"] * (1.0 - 0.05 * t) for t in range(nt)]))
        d["tweight"] = np.full(nt, 1.0 / nt)
    for forms in ((1, 1, 1, 1, 1, 1), (1, 1, 0, 0, 0, 0), (1, 1, 1, 0, 1, 0)):
        case = dict(forms=forms, stream=4, single_form=0)
        monkeypatch.setenv("PB_SH_TILE", "1")
        x1, _, a1 = pb.get_reflected_SH(*C.sh_args(d, case), gweight=d["gweight"], tweight=d["tweight"], return_albedo=True)
        monkeypatch.setenv("PB_SH_TILE", "0")
        x0, _, a0 = pb.get_reflected_SH(*C.sh_args(d, case), gweight=d["gweight"], tweight=d["tweight"], return_albedo=True)
        assert_close(x1, x0, 1e-8, "tile vs per-angle xint %s" % (forms,))
        assert_close(a1, a0, 1e-8, "tile vs per-angle albedo")
        ox, _ = oracle.get_reflected_SH(*C.sh_args(d, case), nthreads=4)
        assert_close(x1, ox, RTOL, "tile vs oracle %s" % (forms,))


def test_cfg4_transit_full_size():
    """BASELINE config 4: 80 layers x 50 000 waves transit (chord kernel at full size; the opacity chain at this size is
    gated in bench.py --config cfg4 against the numpy port)"""
    d = synth.transit_inputs(L=80, W=50000, seed=1004)
    F = pb.get_transit_1d(*C.transit_args(d))
    assert_close(F, oracle.get_transit_1d(*C.transit_args(d), nthreads=8), RTOL, "cfg4 transit depth")


def test_cfg5_batch_sample():
    """BASELINE config 5: 1024 atmospheres x 60 x 2000 thermal batch - one GPU's share of the 8-way split (128
    atmospheres) through thermal_batch, 6 of them against the oracle; rows are independent of the batch they sit in"""
    from oracle import regrid as oreg
    B, L, W = 128, 60, 2000
    ds = [synth.thermal_inputs(L=L, W=W, seed=5000 + b) for b in range(B)]
    d0 = ds[0]
    kw = dict(wno=d0["wno"], tlevel=np.array([d["tlevel"] for d in ds]), plevel=np.array([d["plevel"] for d in ds]),
              ubar1=d0["ubar1"], gweight=d0["gweight"], tweight=d0["tweight"])
    arr = {k: np.array([d[k] for d in ds]) for k in ("dtau", "w0", "cosb")}
    newx = np.linspace(d0["wno"][5], d0["wno"][-5], 300)
    x, y = pb.thermal_batch(**kw, **arr, newx=newx, scale=1e-8)
    sel = np.array([0, 1, 37, 64, 100, 127])
    hk = dict(kw, tlevel=kw["tlevel"][sel], plevel=kw["plevel"][sel])
    xo, yo = oreg.thermal_batch(**hk, **{k: v[sel] for k, v in arr.items()}, newx=newx, scale=1e-8, nthreads=8)
    assert np.array_equal(x, xo)
    m = np.isfinite(yo)
    assert np.array_equal(np.isfinite(y[sel]), m)
    assert_close(y[sel][m], yo[m], RTOL, "cfg5 rebinned thermal spectra")
