"""Pre-mixed correlated-k opacities on the GPU: mirror of optics.RetrieveCKs (method
'preweighted', picaso/optics.py:654-1161, :1398-1512) + compute_opacity for ngauss > 1 (:257-262).

The ln(kappa) table [nP, nT, nwno, ngauss] is uploaded once; `get_opacities(atm)` records the
bilinear (1/T, log10 P) neighbours and the bracketing CIA temperatures; `compute_opacity` runs the
same fused kernel as the monochromatic path with (wavelength, gauss point) columns."""
import ctypes

import numpy as np

from . import _lib
from ._lib import PB_DEVICE, PB_HOST, CkMixArgs, OpacityArgs, addr

__all__ = ["DeviceCKs", "DeviceGasCKs"]


class DeviceCKs:
    """GPU-resident stand-in for optics.RetrieveCKs(..., method='preweighted').

    wno [W]; pressures [nP] (bar), temps [nT], nc_p [nT] (pressures available per temperature);
    kappa = ln(kappa) [nP, nT, W, K]; gauss_wts [K]; cia_temps [nTc] + continuum {pair: [nTc, W]};
    rayleigh_opa {molecule: [W]}."""

    def __init__(self, wno, pressures, temps, nc_p, kappa, gauss_wts, cia_temps, continuum, rayleigh_opa, ctx=None):
        self.ctx = ctx or _lib.default_context()
        self.wno = np.ascontiguousarray(wno, dtype=np.float64)
        self.wave = 1e4 / self.wno
        self.nwno = self.wno.size
        self.pressures = np.asarray(pressures, dtype=np.float64)
        self.temps = np.asarray(temps, dtype=np.float64)
        self.nc_p = np.asarray(nc_p)
        if kappa is not None:
            kappa = np.ascontiguousarray(kappa, dtype=np.float64)
            if kappa.ndim != 4 or kappa.shape[2] != self.nwno:
                raise ValueError("kappa must be ln(kappa) [nP, nT, nwno, ngauss]")
            self._np, self._nt, _, self.ngauss = kappa.shape
        self.gauss_wts = np.asarray(gauss_wts, dtype=np.float64)
        self.cia_temps = np.asarray(cia_temps, dtype=np.float64)
        self._cia_sorted = np.sort(self.cia_temps)
        self._cont_index = {k: i for i, k in enumerate(continuum)}
        self.avail_continuum = list(continuum.keys())
        self.rayleigh_molecules = list(rayleigh_opa.keys())
        self.rayleigh_opa = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in rayleigh_opa.items()}
        self._ray_index = {k: i for i, k in enumerate(rayleigh_opa)}
        self._mol_index = {}
        self.molecules = np.array([])
        lib, h = self.ctx.lib, self.ctx.h
        tab = ctypes.c_void_p()
        self.ctx.check(lib.pb_optab_create(h, self.nwno, 0, len(continuum), len(rayleigh_opa), ctypes.byref(tab)))
        self._tab = tab
        if kappa is not None:
            self.ctx.check(lib.pb_optab_set_ck(h, tab, addr(kappa), self._np, self._nt, self.ngauss))
        order = np.argsort(self.cia_temps)
        for k, t in continuum.items():
            t = np.ascontiguousarray(np.asarray(t, dtype=np.float64)[order])
            self.ctx.check(lib.pb_optab_set_continuum(h, tab, self._cont_index[k], addr(t), t.shape[0]))
        for k, s in self.rayleigh_opa.items():
            self.ctx.check(lib.pb_optab_set_rayleigh(h, tab, self._ray_index[k], addr(s)))
        self._plan = None
        self._ws = {}
        self.raman_stellar_shifts = None

    def _buffer(self, name, shape):
        from .optics import DeviceArray
        d = self._ws.get(name)
        if d is None or d.shape != tuple(shape) or d.ptr is None:
            if d is not None:
                d.free()
            d = DeviceArray(self.ctx, shape)
            self._ws[name] = d
        return d

    def _grid_neighbours(self, atmosphere, unique_temps):
        """the (1/T, log10 P) bracketing shared by get_pre_mix_ck (optics.py:1086-1149, unique temps) and
        get_mixing_indices (optics.py:1199-1277, temps as stored)"""
        t = np.asarray(atmosphere.layer["temperature"], dtype=np.float64)
        p = np.asarray(atmosphere.layer["pressure"], dtype=np.float64) / atmosphere.c.pconv
        t_inv, p_log = 1 / t, np.log10(p)
        pg = np.unique(self.pressures)
        p_log_grid = np.log10(pg[pg > 0])
        t_inv_grid = 1 / np.array(np.unique(self.temps) if unique_temps else self.temps)
        cnt = np.array([np.count_nonzero(t_inv_grid > x) for x in t_inv])
        t_low = np.where(cnt == 0, 0, cnt - 1)
        t_low = np.where(t_low == t_inv_grid.size - 1, t_inv_grid.size - 2, t_low)
        t_hi = t_low + 1
        p_low = np.array([(np.where(p_log_grid <= x)[0][-1] if np.any(p_log_grid <= x) else 0) for x in p_log])
        p_low = np.minimum(p_low, self.nc_p[t_hi] - 3)
        p_hi = p_low + 1
        ti = (t_inv - t_inv_grid[t_low]) / (t_inv_grid[t_hi] - t_inv_grid[t_low])
        pi = (p_log - p_log_grid[p_low]) / (p_log_grid[p_hi] - p_log_grid[p_low])
        return t, t_inv, p_low, p_hi, t_low, t_hi, ti, pi

    def _continuum_plan(self, t, t_inv):
        """get_continuum, optics.py:1410-1424: bracketing CIA temperatures, log-linear in 1/T"""
        st = self._cia_sorted
        L = t.size
        lo = np.zeros(L, dtype=np.int32)
        hi = np.zeros(L, dtype=np.int32)
        for i, tt in enumerate(t):
            if tt <= st[0]:
                lo[i], hi[i] = 0, 1
            elif tt >= st[-1]:
                lo[i], hi[i] = st.size - 2, st.size - 1
            else:
                lo[i] = np.where(st - tt <= 0)[0][-1]
                hi[i] = np.where(st - tt > 0)[0][0]
        ct = (t_inv - 1 / st[lo]) / (1 / st[hi] - 1 / st[lo])
        return lo, hi, np.ascontiguousarray(ct)

    def get_opacities(self, atmosphere, exclude_mol=1):
        """get_opacities_preweighted (optics.py:1500-1511) = get_continuum + get_pre_mix_ck, recorded as a
        plan of table rows and weights (nothing is fetched)."""
        t, t_inv, p_low, p_hi, t_low, t_hi, ti, pi = self._grid_neighbours(atmosphere, unique_temps=True)
        L = t.size
        idx = np.zeros((L, 4), dtype=np.int32)
        idx[:, 0], idx[:, 1] = p_low * self._nt + t_low, p_low * self._nt + t_hi
        idx[:, 2], idx[:, 3] = p_hi * self._nt + t_hi, p_hi * self._nt + t_low
        wts = np.stack([(1 - ti) * (1 - pi), ti * (1 - pi), ti * pi, (1 - ti) * pi], axis=1)
        lo, hi, ct = self._continuum_plan(t, t_inv)
        self._plan = dict(nlayer=L, idx=np.ascontiguousarray(idx), wts=np.ascontiguousarray(wts), lo=lo, hi=hi,
                          ct=np.ascontiguousarray(ct), fac={})
        self.molecular_opa = None
        self.continuum_opa = None

    def close(self):
        for d in getattr(self, "_ws", {}).values():
            d.free()
        self._ws = {}
        if getattr(self, "_tab", None) is not None and self.ctx.h is not None:
            self.ctx.lib.pb_optab_destroy(self.ctx.h, self._tab)
        self._tab = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceGasCKs(DeviceCKs):
    """GPU-resident stand-in for optics.RetrieveCKs(..., method='resortrebin') (optics.py:700-760):
    one ln(kappa) table [nP, nT, W, K] per gas kept in HBM and mixed on the fly per layer with the
    resort-rebin random-overlap rule (deq_chem.py:334-597).

    kappas = {molecule: ln(kappa)}; gauss_pts/gauss_wts [K] on (0, 1)."""

    def __init__(self, wno, pressures, temps, nc_p, kappas, gauss_pts, gauss_wts, cia_temps, continuum, rayleigh_opa,
                 ctx=None):
        super().__init__(wno, pressures, temps, nc_p, None, gauss_wts, cia_temps, continuum, rayleigh_opa, ctx=ctx)
        from .optics import DeviceArray
        self.gauss_pts = np.ascontiguousarray(gauss_pts, dtype=np.float64)
        self.gauss_wts = np.ascontiguousarray(gauss_wts, dtype=np.float64)
        self.ngauss = self.gauss_wts.size
        self.kappas = {}
        shape = None
        for m, k in kappas.items():
            k = np.ascontiguousarray(k, dtype=np.float64)
            if shape is None:
                shape = k.shape
            if k.shape != shape or k.ndim != 4 or k.shape[2] != self.nwno or k.shape[3] != self.ngauss:
                raise ValueError("kappas[%s] must be ln(kappa) [nP, nT, nwno, ngauss] on one grid" % m)
            d = DeviceArray(self.ctx, k.shape)
            self.ctx.check(self.ctx.lib.pb_memcpy_h2d(self.ctx.h, d.ptr, k.ctypes.data, k.nbytes))
            self.kappas[m] = d
        self._np, self._nt = shape[0], shape[1]
        self.ctx.sync()

    def get_mixing_indices(self, atmosphere):
        """optics.py:1199-1277: ([p_low, p_hi, t_low, t_hi], t_interp, p_interp)"""
        _, _, p_low, p_hi, t_low, t_hi, ti, pi = self._grid_neighbours(atmosphere, unique_temps=False)
        return np.array([p_low, p_hi, t_low, t_hi]), ti, pi

    def mix_my_opacities_gasesfly(self, atmosphere, exclude_mol=1, device_output=False, return_ln_mixed=False):
        """optics.py:1164-1197: sets self.molecular_opa [nlayer, nwno, ngauss] (a DeviceArray with
        device_output=True).  Gases are folded in the order of atmosphere.molecules, skipping those whose
        exclude_mol entry is not 1 (optics.py:1177-1181); a gas without a table raises KeyError as there."""
        indices, ti, pi = self.get_mixing_indices(atmosphere)
        mr = atmosphere.layer["mixingratios"]
        gases = [m for m in atmosphere.molecules if exclude_mol == 1 or exclude_mol[m] == 1]
        for m in gases:
            self.kappas[m]
        L = ti.size
        mixes = np.ascontiguousarray([np.asarray(mr[m], dtype=np.float64) for m in gases])
        ptrs = (ctypes.c_void_p * len(gases))(*[self.kappas[m].ptr for m in gases])
        ind = np.ascontiguousarray(indices, dtype=np.int32)
        ti, pi = np.ascontiguousarray(ti), np.ascontiguousarray(pi)
        a = CkMixArgs(nlayer=L, nwno=self.nwno, ngauss=self.ngauss, ngas=len(gases), np=self._np, nt=self._nt)
        a.kappas = ctypes.cast(ptrs, ctypes.c_void_p)
        a.mixes, a.indices, a.t_interp, a.p_interp = addr(mixes), addr(ind), addr(ti), addr(pi)
        a.gauss_pts, a.gauss_wts = addr(self.gauss_pts), addr(self.gauss_wts)
        shape = (L, self.nwno, self.ngauss)
        ln_mixed = None
        if device_output:
            out = self._buffer("molecular_opa", shape)
            a.molecular_opa = out.ptr
            if return_ln_mixed:
                ln_mixed = self._buffer("ln_mixed", shape + (4,))
                a.ln_mixed = ln_mixed.ptr
        else:
            out = np.zeros(shape)
            a.molecular_opa = addr(out)
            if return_ln_mixed:
                ln_mixed = np.zeros(shape + (4,))
                a.ln_mixed = addr(ln_mixed)
        self.ctx.check(self.ctx.lib.pb_ck_mix(self.ctx.h, ctypes.byref(a), PB_DEVICE if device_output else PB_HOST))
        self.molecular_opa = out
        return (out, ln_mixed) if return_ln_mixed else out

    def get_opacities(self, atmosphere, exclude_mol=1):
        """get_opacities_deq_onfly (optics.py:1513-1520) = get_continuum + mix_my_opacities_gasesfly; the mixed
        k-coefficients stay in HBM for compute_opacity."""
        t = np.asarray(atmosphere.layer["temperature"], dtype=np.float64)
        lo, hi, ct = self._continuum_plan(t, 1 / t)
        self.mix_my_opacities_gasesfly(atmosphere, exclude_mol=exclude_mol, device_output=True)
        self._plan = dict(nlayer=t.size, lo=lo, hi=hi, ct=ct, direct=self.molecular_opa)
        self.continuum_opa = None

    def close(self):
        for d in getattr(self, "kappas", {}).values():
            d.free()
        self.kappas = {}
        super().close()


def compute_opacity_ck(atm, opa, stream, delta_eddington, raman, fthin_cld, do_holes, device_outputs, outputs):
    """compute_opacity for a DeviceCKs connection; called by picaso_b200.optics.compute_opacity."""
    from .optics import OUTPUT_NAMES, _LEVEL, _layer_scalars
    if raman != 2:
        # the reference's compute_raman only fills gauss point 0 and copies it (optics.py:289-306);
        # supported through the monochromatic path only for now
        raise NotImplementedError("correlated-k opacities support raman=2 ('none') on the GPU path")
    ctx = opa.ctx
    if opa._plan is None or opa._plan["nlayer"] != atm.c.nlayer:
        raise RuntimeError("call opacityclass.get_opacities(atmosphere) first (justdoit.py:236)")
    L, W, K = atm.c.nlayer, opa.nwno, opa.ngauss
    _, cont, ray = _layer_scalars(atm, opa)
    colden = np.asarray(atm.layer["colden"], dtype=np.float64)
    mmw = np.asarray(atm.layer["mmw"], dtype=np.float64)
    ck_scale = np.ascontiguousarray(colden / mmw)
    pl = opa._plan
    a = OpacityArgs()
    a.nlayer, a.query, a.raman = L, 1, 2
    a.cont_index, a.cont_index_hi, a.cont_t, a.cont_mode = addr(pl["lo"]), addr(pl["hi"]), addr(pl["ct"]), 1
    a.cont_scale, a.ray_scale = addr(cont), addr(ray)
    a.ngauss, a.ck_scale = K, addr(ck_scale)
    if "direct" in pl:
        a.ck_direct = pl["direct"].ptr
    else:
        a.ck_index, a.ck_weights = addr(pl["idx"]), addr(pl["wts"])
    keep = [cont, ray, ck_scale]
    cloud = atm.layer.get("cloud") if isinstance(atm.layer, dict) else atm.layer["cloud"]
    if cloud is not None and np.any(np.asarray(cloud["opd"]) != 0):
        cl = [np.ascontiguousarray(np.broadcast_to(np.asarray(cloud[k], dtype=np.float64), (L, W)))
              for k in ("opd", "w0", "g0")]
        if device_outputs:
            dcl = [opa._buffer("cloud_" + k, (L, W)) for k in ("opd", "w0", "g0")]
            for d, c in zip(dcl, cl):
                ctx.check(ctx.lib.pb_memcpy_h2d(ctx.h, d.ptr, c.ctypes.data, c.nbytes))
            a.cloud_opd, a.cloud_w0, a.cloud_g0 = [d.ptr for d in dcl]
        else:
            a.cloud_opd, a.cloud_w0, a.cloud_g0 = [addr(c) for c in cl]
        keep += cl
        a.cloud_ld = W
    a.fthin_cld = float(fthin_cld) if fthin_cld is not None else 0.0
    a.do_holes = int(bool(do_holes))
    a.stream, a.delta_eddington = int(stream), int(bool(delta_eddington))
    want = set(OUTPUT_NAMES if outputs is None else outputs)
    res = {}
    for n in OUTPUT_NAMES:
        if n not in want:
            res[n] = None
            continue
        shape = ((L + 1) if n in _LEVEL else L, W, K)
        if device_outputs:
            res[n] = opa._buffer(n, shape)
            setattr(a, n, res[n].ptr)
        else:
            res[n] = np.zeros(shape)
            setattr(a, n, addr(res[n]))
    ctx.check(ctx.lib.pb_compute_opacity(ctx.h, opa._tab, ctypes.byref(a), PB_DEVICE if device_outputs else PB_HOST))
    if device_outputs:
        ctx.sync()
    return tuple(res[n] for n in OUTPUT_NAMES)
