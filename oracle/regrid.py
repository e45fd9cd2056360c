"""Oracle of mean_regrid and of the batched thermal forward model - TEST INFRASTRUCTURE ONLY.

mean_regrid restates picaso/justplotit.py:31-63 line by line and calls the same third-party routine the
reference calls there (scipy.stats.binned_statistic; scipy is pinned only as "scipy" in the reference's
pyproject.toml:26, 1.18.1 in this image); create_grid restates picaso/opacity_factory.py:712-739.
thermal_batch is the loop picaso/driver.py:214-232 runs per sample, over the C oracle."""
import numpy as np
from scipy.stats import binned_statistic

from . import compress_thermal, get_thermal_1d


def create_grid(min_wavelength, max_wavelength, constant_R):
    spacing = (2. * constant_R + 1.) / (2. * constant_R - 1.)
    npts = np.log(max_wavelength / min_wavelength) / np.log(spacing)
    wsize = int(np.ceil(npts)) + 1
    newwl = np.zeros(wsize)
    newwl[0] = min_wavelength
    for j in range(1, wsize):
        newwl[j] = newwl[j - 1] * spacing
    return 1e4 / newwl[::-1]


def mean_regrid(x, y, newx=None, R=None):
    if (isinstance(newx, type(None)) & (not isinstance(R, type(None)))):
        newx = create_grid(1e4 / max(x), 1e4 / min(x), R)
    elif (not isinstance(newx, type(None)) & (isinstance(R, type(None)))):
        d = np.diff(newx)
        binedges = np.array([newx[0] - d[0] / 2] + list(newx[0:-1] + d / 2.0) + [newx[-1] + d[-1] / 2])
        newx = binedges
    else:
        raise Exception('Please either enter a newx or a R')
    y, edges, binnum = binned_statistic(x, y, bins=newx)
    newx = (edges[0:-1] + edges[1:]) / 2.0
    return newx, y


def thermal_batch(wno, tlevel, plevel, dtau, w0, cosb, ubar1, gweight, tweight, surf_reflect=0.0, hard_surface=0,
                  dwno=None, calc_type=0, newx=None, R=None, scale=1.0, nthreads=1):
    B, V = np.shape(tlevel)
    W = len(wno)
    ng, nt = np.shape(ubar1)
    sr = np.broadcast_to(np.asarray(surf_reflect, dtype=np.float64), (B, W)) if np.ndim(surf_reflect) < 2 else surf_reflect
    rows = []
    x = np.asarray(wno)
    for b in range(B):
        ftop, _ = get_thermal_1d(V, wno, W, ng, nt, tlevel[b], dtau[b], w0[b], cosb[b], plevel[b], ubar1,
                                 np.ascontiguousarray(sr[b]), hard_surface, dwno if dwno is not None else np.ones(W),
                                 calc_type, nthreads=nthreads, level_fluxes=False)
        spec = compress_thermal(W, ftop, gweight, tweight) * scale
        if newx is not None or R is not None:
            x, spec = mean_regrid(wno, spec, newx=newx, R=R)
        rows.append(spec)
    return x, np.array(rows)
