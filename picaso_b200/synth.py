"""Seeded synthetic atmospheres of the BASELINE.json shapes.

Everything is float64, C-order, wavelength on the fastest axis - the layout the
reference passes to its flux solvers (picaso/justdoit.py:275-283, :337-342, :392-396).
Used by tests/, bench.py and tests/golden/make_golden.py; pure numpy.
"""
import numpy as np

__all__ = ["geometry_1d", "reflected_inputs", "thermal_inputs", "transit_inputs",
           "adversarial_reflected", "climate_inputs"]

# Abramowitz & Stegun 25.8 half-sphere Gauss points used by the reference for
# 1-D geometry (picaso/disco.py:67-87).
_GAUSS = {
    5: ([0.0985350858, 0.3045357266, 0.5620251898, 0.8019865821, 0.9601901429],
        [0.0157479145, 0.0739088701, 0.1463869871, 0.1671746381, 0.0967815902]),
    6: ([0.0730543287, 0.2307661380, 0.4413284812, 0.6630153097, 0.8519214003, 0.9706835728],
        [0.0087383018, 0.0439551656, 0.0986611509, 0.1407925538, 0.1355424972, 0.0723103307]),
    7: ([0.0562625605, 0.1802406917, 0.3526247171, 0.5471536263, 0.7342101772, 0.8853209468,
         0.9775206136],
        [0.0052143622, 0.0274083567, 0.0663846965, 0.1071250657, 0.1273908973, 0.1105092582,
         0.0559673634]),
    8: ([0.0446339553, 0.1443662570, 0.2868247571, 0.4548133152, 0.6280678354, 0.7856915206,
         0.9086763921, 0.9822200849],
        [0.0032951914, 0.0178429027, 0.0454393195, 0.0791995995, 0.1060473594, 0.1125057995,
         0.0911190236, 0.0445508044]),
}


def geometry_1d(ngauss=5, phase=0.0):
    """gangle, gweight, tangle, tweight, ubar0[ng,1], ubar1[ng,1], cos_theta.

    Same numbers as disco.get_angles_1d + disco.compute_disco (disco.py:36-50, :67-87).
    """
    g, w = _GAUSS[ngauss]
    gangle = np.array(g)
    gweight = np.array(w)
    tangle = np.array([0.0])
    tweight = np.array([1.0])
    cos_theta = np.cos(phase)
    lon = np.arcsin((gangle - (cos_theta - 1.0) / (cos_theta + 1.0)) / (2.0 / (cos_theta + 1.0)))
    if phase > np.pi:
        lon = -lon
    f = np.sin(np.arccos(tangle))
    ubar0 = np.outer(np.cos(lon - phase), f)
    ubar1 = np.outer(np.cos(lon), f)
    return gangle, gweight, tangle, tweight, ubar0, ubar1, float(cos_theta)


def _layer_fields(rng, L, W):
    lgrid = np.arange(L)[:, None] / max(L - 1, 1)
    dtau_og = 10.0 ** (-4.0 + 5.0 * lgrid + 0.5 * rng.standard_normal((L, W)))
    base = rng.uniform(0.1, 0.9, size=(1, W))
    w0_og = np.clip(base + 0.3 * (lgrid - 0.5) + 0.1 * rng.standard_normal((L, W)), 0.02, 0.98)
    cosb_og = rng.uniform(0.0, 0.85, size=(L, W))
    return dtau_og, w0_og, cosb_og


def reflected_inputs(L=60, W=300, seed=1001, stream=2, delta_eddington=True, ngauss=5, phase=0.0):
    """Inputs of get_reflected_1d / get_reflected_SH (fluxes.py:1010-1015, :2675-2679)."""
    rng = np.random.default_rng(seed)
    dtau_og, w0_og, cosb_og = _layer_fields(rng, L, W)
    ftau_cld = rng.uniform(0.05, 0.95, size=(L, W))
    ftau_ray = 1.0 - ftau_cld
    gcos2 = 0.5 * ftau_ray
    if delta_eddington:
        # optics.py:412-420
        f = cosb_og ** stream
        w0 = w0_og * (1.0 - f) / (1.0 - w0_og * f)
        cosb = (cosb_og - f) / (1.0 - f)
        dtau = dtau_og * (1.0 - w0_og * f)
    else:
        f = np.zeros_like(cosb_og)
        w0, cosb, dtau = w0_og.copy(), cosb_og.copy(), dtau_og.copy()
    tau = np.vstack([np.zeros((1, W)), np.cumsum(dtau, axis=0)])
    tau_og = np.vstack([np.zeros((1, W)), np.cumsum(dtau_og, axis=0)])
    gangle, gweight, tangle, tweight, ubar0, ubar1, cos_theta = geometry_1d(ngauss, phase)
    return dict(
        nlevel=L + 1, nwno=W, wno=np.linspace(1e4 / 1.0, 1e4 / 0.3, W),
        numg=ngauss, numt=1,
        dtau=dtau, tau=tau, w0=w0, cosb=cosb, gcos2=gcos2, ftau_cld=ftau_cld, ftau_ray=ftau_ray,
        dtau_og=dtau_og, tau_og=tau_og, w0_og=w0_og, cosb_og=cosb_og, f_deltaM=f,
        surf_reflect=np.zeros(W), ubar0=ubar0, ubar1=ubar1, cos_theta=cos_theta,
        F0PI=np.ones(W), gweight=gweight, tweight=tweight,
        frac_a=1.0, frac_b=-1.0, frac_c=2.0, constant_back=-0.5, constant_forward=1.0,
    )


def adversarial_reflected(L=12, seed=7):
    """Edge regimes of SURVEY.md section 8d run for parity only: one wavelength column each."""
    rng = np.random.default_rng(seed)
    cols = []
    for w0v in (1e-6, 0.5, 0.999999):
        for dt in (1e-10, 1e-3, 1.0, 50.0):
            for g in (0.0, 0.5, 0.95):
                cols.append((w0v, dt, g))
    W = len(cols)
    d = reflected_inputs(L=L, W=W, seed=seed, delta_eddington=False)
    for i, (w0v, dt, g) in enumerate(cols):
        d["w0"][:, i] = w0v
        d["w0_og"][:, i] = w0v
        d["dtau"][:, i] = dt * (1.0 + 0.01 * rng.standard_normal(L))
        d["dtau_og"][:, i] = d["dtau"][:, i]
        d["cosb"][:, i] = g
        d["cosb_og"][:, i] = g
    d["tau"] = np.vstack([np.zeros((1, W)), np.cumsum(d["dtau"], axis=0)])
    d["tau_og"] = d["tau"].copy()
    d["surf_reflect"] = np.full(W, 0.3)
    return d


def thermal_inputs(L=90, W=10000, seed=1002, ngauss=5, wno_range=(300.0, 10000.0),
                   t_range=(400.0, 2000.0)):
    """Inputs of get_thermal_1d (fluxes.py:1683-1684)."""
    rng = np.random.default_rng(seed)
    dtau, w0, cosb = _layer_fields(rng, L, W)
    tlevel = np.linspace(t_range[0], t_range[1], L + 1)
    plevel = np.logspace(-6, 2, L + 1) * 1e6
    wno = np.linspace(wno_range[0], wno_range[1], W)
    gangle, gweight, tangle, tweight, ubar0, ubar1, cos_theta = geometry_1d(ngauss, 0.0)
    return dict(nlevel=L + 1, nwno=W, wno=wno, numg=ngauss, numt=1, tlevel=tlevel, dtau=dtau,
                w0=w0, cosb=cosb, plevel=plevel, ubar1=ubar1, surf_reflect=np.zeros(W),
                hard_surface=0, dwno=np.gradient(wno) if W > 1 else np.ones(1),
                gweight=gweight, tweight=tweight)


def transit_inputs(L=80, W=50000, seed=1004):
    """Inputs of get_transit_1d (fluxes.py:2582-2583): isothermal hydrostatic H2/He envelope."""
    rng = np.random.default_rng(seed)
    k_b = 1.380649e-16
    amu = 1.66053906660e-24
    rjup = 7.1492e9
    mjup = 1.898e30
    G = 6.674e-8
    V = L + 1
    plevel = np.logspace(-6, 2, V) * 1e6                 # dyn/cm2, top -> bottom
    tlevel = np.full(V, 1000.0)
    mmw = np.full(L, 2.3)
    radius, mass = 1.2 * rjup, 1.0 * mjup
    # integrate upward from the bottom level at r=radius
    z = np.zeros(V)
    z[-1] = radius
    for i in range(V - 2, -1, -1):
        g = G * mass / z[i + 1] ** 2
        H = k_b * tlevel[i] / (mmw[min(i, L - 1)] * amu * g)
        z[i] = z[i + 1] + H * np.log(plevel[i + 1] / plevel[i])
    dz = np.zeros(V)
    dz[1:-1] = 0.5 * (z[:-2] - z[2:])
    dz[0] = dz[1]
    dz[-1] = dz[-2]
    gravity = G * mass / radius ** 2
    colden = (plevel[1:] - plevel[:-1]) / gravity
    dtau, _, _ = _layer_fields(rng, L, W)
    return dict(z=z, dz=dz, nlevel=V, nwno=W, rstar=6.957e10, mmw=mmw, k_b=k_b, amu=amu,
                player=plevel, tlayer=tlevel, colden=colden, DTAU=dtau)


# ---------------------------------------------------------------------------------------
# synthetic opacity "database" + atmosphere profile for the opacity path (a9-a11)
# ---------------------------------------------------------------------------------------
MOLECULES = ["H2O", "CH4", "CO", "CO2", "NH3", "Na", "K", "TiO", "VO", "H2S", "PH3", "FeH"]
CONTINUUM = [("H2", "H2"), ("H2", "He"), ("H2", "CH4"), ("H-", "bf"), ("H-", "ff"), ("H2-", "")]
RAYLEIGH = ["H2", "He", "CH4"]


def opacity_database(W=200, nmol=4, seed=2001, nT=12, nP=10, nTc=15, ragged=True,
                     wave_range=(0.3, 1.0)):
    """Synthetic stand-in for the reference's sqlite opacity DB (opacity_factory.py:622-668):
    per-molecule cross-section rows on a (T-major, P-minor) grid with `nc_p[it]` pressures per
    temperature (ragged like the real 1060/1460 grids when `ragged`), continuum tables on their
    own temperature grid, wavenumber grid ascending in cm^-1."""
    rng = np.random.default_rng(seed)
    wno = np.sort(1e4 / np.linspace(wave_range[1], wave_range[0], W))
    temps = np.round(np.geomspace(75.0, 4000.0, nT), 3)
    pressures = np.geomspace(1e-6, 3e3, nP)           # bar
    nc_p = np.full(nT, nP)
    if ragged:
        nc_p[-3:] = nP - 2                            # hottest temperatures lack the highest pressures
        nc_p[0] = nP - 1
    pt = []                                           # (ptid (1-based), pressure, temperature)
    for it, T in enumerate(temps):
        for ip in range(nc_p[it]):
            pt.append((len(pt) + 1, float(pressures[ip]), float(T)))
    npt = len(pt)
    mols = MOLECULES[:nmol]
    lgT = np.log10(np.array([p[2] for p in pt]))[:, None]
    lgP = np.log10(np.array([p[1] for p in pt]))[:, None]
    tables = {}
    for i, m in enumerate(mols):
        band = np.sin(np.linspace(0, 6 + i, W) + i)[None, :]
        lk = -24.0 + 2.5 * band + 0.8 * (lgT - 2.5) + 0.15 * lgP + 0.3 * rng.standard_normal((npt, W))
        k = 10.0 ** lk
        k[rng.random((npt, W)) < 0.01] = 0.0          # exercised by the reference's 1e-50 guard
        tables[m] = k
    cia_temps = np.round(np.geomspace(75.0, 3500.0, nTc), 2)
    cont = {}
    for a, b in CONTINUUM:
        lk = -7.0 + np.cos(np.linspace(0, 4, W))[None, :] + 0.5 * np.log10(cia_temps / 300.0)[:, None] \
            + 0.2 * rng.standard_normal((nTc, W))
        if a in ("H-", "H2-"):
            lk = lk - 18.0
        cont[a + b] = 10.0 ** lk
    return dict(wno=wno, nwno=W, temps=temps, pressures=pressures, nc_p=nc_p, pt_pairs=pt,
                molecules=mols, tables=tables, cia_temps=cia_temps, continuum=cont,
                continuum_molecules=list(CONTINUUM), rayleigh_molecules=list(RAYLEIGH))


def raman_table(seed=2002, n=56):
    """Synthetic H2 Raman transitions with the structure of Oklopcic+2016 table 2 (ji, C, deltanu);
    every fourth row is a Rayleigh (deltanu = 0) term."""
    rng = np.random.default_rng(seed)
    ji = np.repeat(np.arange(10), 6)[:n]
    c = 10.0 ** rng.uniform(-48, -44.5, n)
    dnu = rng.choice([354.6, 587.4, 814.9, 4161.2, 4500.2, -354.6, -587.4], n)
    dnu[::4] = 0.0
    return ji.astype(np.int64), c, dnu


def atmosphere_profile(db, L=12, seed=2003, cloudy=True):
    """Per-layer scalars compute_opacity multiplies by (atmsetup.py: level/layer T,P, mixing
    ratios, mmw, column density) for a hot-Jupiter-like profile, CGS units."""
    rng = np.random.default_rng(seed)
    W = db["nwno"]
    pconv = 1e6
    plevel = np.geomspace(3e-6, 80.0, L + 1) * pconv           # dyn/cm2
    tlevel = 300.0 + 1400.0 * (np.log10(plevel / pconv) + 6) / 8 + 30 * rng.standard_normal(L + 1)
    tlevel = np.clip(tlevel, 90.0, 3800.0)
    tlayer = 0.5 * (tlevel[1:] + tlevel[:-1])
    player = np.sqrt(plevel[1:] * plevel[:-1])
    gravity = 2500.0                                            # cm/s2
    mmw = np.full(L, 2.3) + 0.01 * rng.standard_normal(L)
    colden = (plevel[1:] - plevel[:-1]) / gravity
    species = sorted(set(db["molecules"]) | {"H2", "He", "CH4", "H", "H-"})
    mix = {s: 10.0 ** rng.uniform(-8, -3, L) for s in species}
    mix["H2"] = np.full(L, 0.84)
    mix["He"] = np.full(L, 0.155)
    mix["H"] = 10.0 ** rng.uniform(-9, -6, L)
    mix["H-"] = 10.0 ** rng.uniform(-14, -11, L)
    electrons = 10.0 ** rng.uniform(-12, -8, L)
    if cloudy:
        lgrid = np.arange(L)[:, None] / max(L - 1, 1)
        opd = 10.0 ** (-3 + 3 * np.exp(-((lgrid - 0.6) / 0.15) ** 2) + 0.1 * rng.standard_normal((L, W)))
        cw0 = np.clip(0.9 + 0.05 * rng.standard_normal((L, W)), 0.05, 0.999)
        cg0 = np.clip(0.6 + 0.1 * rng.standard_normal((L, W)), 0.0, 0.9)
    else:
        opd = np.zeros((L, W)); cw0 = np.zeros((L, W)); cg0 = np.zeros((L, W))
    return dict(nlayer=L, nlevel=L + 1, pconv=pconv, rgas=8.31446261815324e7, amu=1.66053906660e-24,
                k_b=1.380649e-16, plevel=plevel, tlevel=tlevel, player=player, tlayer=tlayer,
                gravity=gravity, mmw=mmw, colden=colden, mixingratios=mix, electrons=electrons,
                cloud_opd=opd, cloud_w0=cw0, cloud_g0=cg0)


def ck_database(W=40, K=8, seed=2101, nT=9, nP=8, nTc=12):
    """Synthetic pre-mixed correlated-k table in the layout RetrieveCKs keeps it (optics.py:737-753):
    ln(kappa) [nP, nT, W, K], K double-Gauss points per wavenumber bin, monotone in the gauss index."""
    rng = np.random.default_rng(seed)
    wno = np.sort(1e4 / np.linspace(30.0, 0.3, W))
    temps = np.round(np.geomspace(75.0, 4000.0, nT), 3)
    pressures = np.geomspace(1e-6, 3e3, nP)
    nc_p = np.full(nT, nP)
    base = -26.0 + 2.0 * np.sin(np.linspace(0, 7, W))[None, None, :, None] + \
        0.7 * (np.log10(temps) - 2.5)[None, :, None, None] + 0.2 * np.log10(pressures)[:, None, None, None]
    spread = np.sort(rng.uniform(0.0, 4.0, (nP, nT, W, K)), axis=3)
    lnk = (base + spread) * np.log(10.0)
    gauss_wts = np.array([0.16523105, 0.30976895, 0.30976895, 0.16523105, 0.00869637, 0.01630363, 0.01630363,
                          0.00869637])[:K]
    cia_temps = np.round(np.geomspace(75.0, 3500.0, nTc), 2)
    cont = {}
    for a, b in CONTINUUM:
        lk = -7.0 + np.cos(np.linspace(0, 4, W))[None, :] + 0.5 * np.log10(cia_temps / 300.0)[:, None] \
            + 0.2 * rng.standard_normal((nTc, W))
        if a in ("H-", "H2-"):
            lk = lk - 18.0
        cont[a + b] = 10.0 ** lk
    return dict(wno=wno, nwno=W, ngauss=K, temps=temps, pressures=pressures, nc_p=nc_p, kappa=lnk,
                gauss_wts=gauss_wts, cia_temps=cia_temps, continuum=cont, continuum_molecules=list(CONTINUUM),
                rayleigh_molecules=list(RAYLEIGH), molecules=[])


def climate_inputs(L=40, W=120, K=4, seed=3001, ng=1, nt=1, surf=0.0, clear=False):
    """The namedtuples picaso.climate.get_fluxes takes (climate.py:1962-1966, justdoit.py:5038),
    with [nlayer, nwno, ngauss] opacity arrays (gauss point fastest, optics.py:423-431).
    Gauss point k is optically thicker by ~10^k/2 (a k-distribution); clear=True builds the
    cloud-free column of the same atmosphere for the do_holes path."""
    from collections import namedtuple
    rng = np.random.default_rng(seed)
    wed = {k: np.zeros(((L + 1) if k == "TAU" else L, W, K)) for k in
           ("DTAU", "TAU", "W0", "COSB", "ftau_cld", "ftau_ray", "GCOS2", "W0_no_raman", "f_deltaM")}
    noed = {k: np.zeros(((L + 1) if k == "TAU" else L, W, K)) for k in ("DTAU", "TAU", "W0", "COSB")}
    for k in range(K):
        dtau_og, w0_og, cosb_og = _layer_fields(rng, L, W)
        dtau_og = dtau_og * 10.0 ** (0.5 * k - 1.0)
        ftau_cld = rng.uniform(0.05, 0.95, size=(L, W))
        if clear:
            ftau_cld = ftau_cld * 0.0
            cosb_og = cosb_og * 0.0
        ftau_ray = 1.0 - ftau_cld
        f = cosb_og ** 2
        wed["W0"][:, :, k] = w0_og * (1.0 - f) / (1.0 - w0_og * f)
        wed["COSB"][:, :, k] = (cosb_og - f) / (1.0 - f)
        wed["DTAU"][:, :, k] = dtau_og * (1.0 - w0_og * f)
        wed["TAU"][1:, :, k] = np.cumsum(wed["DTAU"][:, :, k], axis=0)
        wed["ftau_cld"][:, :, k] = ftau_cld
        wed["ftau_ray"][:, :, k] = ftau_ray
        wed["GCOS2"][:, :, k] = 0.5 * ftau_ray
        wed["W0_no_raman"][:, :, k] = np.clip(w0_og * rng.uniform(0.97, 1.0, size=(L, W)), 0.0, 0.999)
        wed["f_deltaM"][:, :, k] = f
        noed["DTAU"][:, :, k] = dtau_og
        noed["TAU"][1:, :, k] = np.cumsum(dtau_og, axis=0)
        noed["W0"][:, :, k] = w0_og
        noed["COSB"][:, :, k] = cosb_og
    wno = np.linspace(300.0, 30000.0, W)
    dwno = np.gradient(wno) if W > 1 else np.ones(1)
    gpts, gw = np.polynomial.legendre.leggauss(K)
    gauss_wts = 0.5 * gw
    Atmosphere = namedtuple("Atmosphere_Tuple", ["dtdp", "mmw_layer", "nlevel", "t_level", "p_level", "condensables",
                                                 "condensable_abundances", "condensable_weights", "scale_height"])
    OpacityWEd = namedtuple("OpacityWEd_Tuple", ["DTAU", "TAU", "W0", "COSB", "ftau_cld", "ftau_ray", "GCOS2",
                                                 "W0_no_raman", "f_deltaM"])
    OpacityNoEd = namedtuple("OpacityNoEd_Tuple", ["DTAU", "TAU", "W0", "COSB"])
    ScatteringPhase = namedtuple("ScatteringPhase_Tuple", ["surf_reflect", "single_phase", "multi_phase", "frac_a",
                                                           "frac_b", "frac_c", "constant_back", "constant_forward"])
    Disco = namedtuple("Disco_Tuple", ["ng", "nt", "gweight", "tweight", "ubar0", "ubar1", "cos_theta"])
    Opagrid = namedtuple("Opagrid", ["nwno", "delta_wno", "wno", "ngauss", "gauss_wts", "tmin", "tmax"])
    if nt == 1:
        gangle, gweight, tangle, tweight, ubar0, ubar1, cos_theta = geometry_1d(ng if ng in _GAUSS else 5, 0.0)
        if ng not in _GAUSS:  # ng = 1: the climate default single stream
            ubar0, ubar1 = np.full((ng, 1), 0.5), np.full((ng, 1), 0.5)
            gweight, tweight = np.full(ng, 1.0 / ng), np.array([1.0])
    else:
        ubar1 = rng.uniform(0.1, 1.0, size=(ng, nt))
        ubar0 = ubar1.copy()
        gweight, tweight, cos_theta = rng.uniform(0.1, 0.4, size=ng), rng.uniform(0.2, 1.0, size=nt), 1.0
    tlevel = np.linspace(150.0, 1800.0, L + 1) + 20.0 * np.sin(np.arange(L + 1))
    plevel = np.logspace(-6, 2, L + 1) * 1e6
    return dict(
        Atmosphere=Atmosphere(np.zeros(L), np.full(L, 2.3), L + 1, tlevel, plevel, (), np.zeros((1, 1)),
                              np.zeros(1), np.zeros(L + 1)),
        OpacityWEd=OpacityWEd(**wed), OpacityNoEd=OpacityNoEd(**noed),
        ScatteringPhase=ScatteringPhase(np.full(W, float(surf)), 3, 0, 1.0, -1.0, 2.0, -0.5, 1.0),
        Disco=Disco(ng, nt, gweight, tweight, ubar0, ubar1, float(cos_theta)),
        Opagrid=Opagrid(W, dwno, wno, K, gauss_wts, 100.0, 4000.0),
        F0PI=rng.uniform(0.5, 2.0, size=W))
