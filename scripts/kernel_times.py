"""Device-resident kernel timings for the BASELINE.json configurations (run on the GPU box).

    python scripts/kernel_times.py [--reps 50] [--only refl,thermal,transit,batch]

Inputs are uploaded once; each timed launch rotates over enough distinct input sets to
exceed the 126 MB L2.  Prints one JSON line per configuration with the achieved fraction
of the measured HBM roofline (algorithmic bytes).
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picaso_b200 as pb  # noqa: E402
from picaso_b200 import _lib, synth  # noqa: E402
from picaso_b200._lib import PB_DEVICE, ReflectedArgs, ShArgs, ThermalArgs, TransitArgs  # noqa: E402

L2_BYTES = 126e6


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(ctx, launch, nsets, reps):
    for i in range(3):
        launch(i % nsets)
    ctx.sync()
    ctx.timer_start()
    for i in range(reps):
        launch(i % nsets)
    return ctx.timer_stop() / reps


def report(name, ms, alg_bytes, units, unit_name, extra=None):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    out = {"config": name, "ms_per_launch": ms, "alg_bytes": alg_bytes, "GB/s": gbs,
           "hbm_frac": gbs / hbm_peak(), unit_name + "/s": units / (ms * 1e-3)}
    out.update(extra or {})
    print(json.dumps(out), flush=True)


def bench_reflected(ctx, L, W, G, reps, sp=3, batch=1, tag=""):
    per_set = (9 * L + 2 * (L + 1) + 2) * 8 * W * batch
    nsets = max(1, int(np.ceil(2 * L2_BYTES / per_set)))
    nsets = min(nsets, 6)
    lay = ("dtau", "w0", "cosb", "gcos2", "ftau_cld", "ftau_ray", "dtau_og", "w0_og", "cosb_og")
    lev = ("tau", "tau_og")
    wav = ("surf_reflect", "F0PI")
    args, keep = [], []
    d_x = ctx.dev_alloc(batch * G * W * 8)
    d_a = ctx.dev_alloc(batch * W * 8)
    for s in range(nsets):
        ds = [synth.reflected_inputs(L=L, W=W, seed=50 + s * batch + b, ngauss=G) for b in range(batch)]
        a = ReflectedArgs()
        a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = L, W, G, 1, batch, W
        for k in lay + lev + wav:
            setattr(a, k, ctx.to_device(np.stack([d[k] for d in ds])))
        d0 = ds[0]
        vec = [np.ascontiguousarray(d0[k]).reshape(-1) for k in ("ubar0", "ubar1", "gweight", "tweight")]
        keep.append(vec)
        a.ubar0, a.ubar1, a.gweight, a.tweight = [_lib.addr(v) for v in vec]
        a.cos_theta = d0["cos_theta"]
        a.single_phase, a.multi_phase, a.toon_coefficients = sp, 0, 0
        a.frac_a, a.frac_b, a.frac_c = d0["frac_a"], d0["frac_b"], d0["frac_c"]
        a.constant_back, a.constant_forward = d0["constant_back"], d0["constant_forward"]
        a.get_toa_intensity, a.get_lvl_flux = 1, 0
        a.xint_at_top, a.albedo = d_x, d_a
        args.append(a)
    fn = ctx.lib.pb_reflected_toon_1d
    ms = timeit(ctx, lambda i: ctx.check(fn(ctx.h, ctypes.byref(args[i]), PB_DEVICE)), nsets, reps)
    alg = ((9 * L + 2 * (L + 1) + 2) * 8 + G * 8) * W * batch
    report("reflected_toon L=%d W=%d G=%d batch=%d sp=%d%s" % (L, W, G, batch, sp, tag), ms, alg,
           W * batch, "wave-points", {"nsets": nsets})
    for a in args:
        for k in lay + lev + wav:
            ctx.dev_free(getattr(a, k))
    ctx.dev_free(d_x)
    ctx.dev_free(d_a)


def bench_sh(ctx, L, W, G, reps, stream=4, forms=(0, 0, 0, 1, 1, 1)):
    per_set = (9 * L + 2 * (L + 1) + 2) * 8 * W
    nsets = min(6, max(1, int(np.ceil(2 * L2_BYTES / per_set))))
    lay = ("dtau", "w0", "ftau_cld", "ftau_ray", "f_deltaM", "dtau_og", "w0_og", "cosb_og")
    lev = ("tau", "tau_og")
    wav = ("surf_reflect", "F0PI")
    args, keep = [], []
    d_x = ctx.dev_alloc(G * W * 8)
    d_a = ctx.dev_alloc(W * 8)
    for s_ in range(nsets):
        d = synth.reflected_inputs(L=L, W=W, seed=150 + s_, ngauss=G, stream=stream)
        a = ShArgs()
        a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = L, W, G, 1, 1, W
        for k in lay + lev + wav:
            setattr(a, k, ctx.to_device(d[k]))
        vec = [np.ascontiguousarray(d[k]).reshape(-1) for k in ("ubar0", "ubar1", "gweight", "tweight")]
        keep.append(vec)
        a.ubar0, a.ubar1, a.gweight, a.tweight = [_lib.addr(v) for v in vec]
        a.cos_theta = d["cos_theta"]
        (a.w_single_form, a.w_multi_form, a.psingle_form, a.w_single_rayleigh, a.w_multi_rayleigh,
         a.psingle_rayleigh) = forms
        a.frac_a, a.frac_b, a.frac_c = d["frac_a"], d["frac_b"], d["frac_c"]
        a.constant_back, a.constant_forward = d["constant_back"], d["constant_forward"]
        a.stream, a.flx, a.single_form = stream, 0, 0
        a.xint_at_top, a.albedo = d_x, d_a
        args.append(a)
    fn = ctx.lib.pb_reflected_sh
    ms = timeit(ctx, lambda i: ctx.check(fn(ctx.h, ctypes.byref(args[i]), PB_DEVICE)), nsets, reps)
    alg = ((9 * L + 2 * (L + 1) + 2) * 8 + G * 8) * W
    report("reflected_SH%d L=%d W=%d G=%d forms=%s" % (stream, L, W, G, "".join(map(str, forms))), ms, alg,
           W, "wave-points", {"nsets": nsets})
    for a in args:
        for k in lay + lev + wav:
            ctx.dev_free(getattr(a, k))
    ctx.dev_free(d_x)
    ctx.dev_free(d_a)


def bench_opacity(ctx, L, W, nmol, reps, query="linear", raman=2, outputs=None, tag=""):
    """cfg4: opacity-interpolation-dominated path.  Tables resident in HBM; per call only O(L) scalars."""
    import time
    import types
    from picaso_b200 import optics as po
    from picaso_b200._lib import OpacityArgs
    db = synth.opacity_database(W=W, nmol=nmol, seed=4001, nT=20, nP=18, nTc=30, wave_range=(0.3, 5.0))
    atm = synth.atmosphere_profile(db, L=L, seed=4003, cloudy=True)
    rng = np.random.default_rng(1)
    ray = {m: 10.0 ** rng.uniform(-27, -25, W) for m in db["rayleigh_molecules"]}
    ji, c, dnu = synth.raman_table()
    t0 = time.perf_counter()
    opa = pb.DeviceOpacities(db["wno"], db["pt_pairs"], db["tables"], db["cia_temps"], db["continuum"], ray,
                             raman_db=(c, ji, dnu), query_method=query, ctx=ctx)
    if raman == 0:
        opa.raman_stellar_shifts = 0.6 + 0.8 * rng.random((W, c.size))
    t_up = time.perf_counter() - t0
    a = types.SimpleNamespace()
    a.c = types.SimpleNamespace(nlayer=L, pconv=atm["pconv"], rgas=atm["rgas"], amu=atm["amu"], k_b=atm["k_b"])
    a.level = {"temperature": atm["tlevel"], "pressure": atm["plevel"]}
    a.layer = {"temperature": atm["tlayer"], "pressure": atm["player"], "colden": atm["colden"], "mmw": atm["mmw"],
               "mixingratios": atm["mixingratios"], "electrons": atm["electrons"],
               "cloud": {"opd": atm["cloud_opd"], "w0": atm["cloud_w0"], "g0": atm["cloud_g0"]}}
    a.planet = types.SimpleNamespace(gravity=atm["gravity"])
    a.molecules, a.rayleigh_molecules = list(db["molecules"]), list(db["rayleigh_molecules"])
    a.continuum_molecules = [list(x) for x in db["continuum_molecules"]]
    opa.get_opacities(a)
    names = po.OUTPUT_NAMES if outputs is None else outputs
    # wall-clock of the public call (host scalars + launch + sync), outputs staying in HBM
    res = pb.compute_opacity(a, opa, stream=2, delta_eddington=True, test_mode=None, raman=raman,
                             device_outputs=True, outputs=names)
    t0 = time.perf_counter()
    n_api = 10
    for _ in range(n_api):
        res = pb.compute_opacity(a, opa, stream=2, delta_eddington=True, test_mode=None, raman=raman,
                                 device_outputs=True, outputs=names)
    api_ms = (time.perf_counter() - t0) / n_api * 1e3
    # kernel-only: prebuilt argument block, device-resident cloud arrays and outputs
    mol, cont, rays = po._layer_scalars(a, opa)
    idx, wts, cia = opa._plan["idx"], opa._plan["wts"], opa._plan["cia"]
    oa = OpacityArgs()
    oa.nlayer, oa.query = L, (1 if query == "linear" else 0)
    oa.pt_index, oa.weights, oa.cont_index = _lib.addr(idx), _lib.addr(wts), _lib.addr(cia)
    oa.mol_scale, oa.cont_scale, oa.ray_scale = _lib.addr(mol), _lib.addr(cont), _lib.addr(rays)
    oa.raman = raman
    jf = np.ascontiguousarray([po.j_fraction(j, atm["tlayer"]) for j in range(10)])
    oa.jfrac = _lib.addr(jf)
    cl = [po.DeviceArray.from_numpy(ctx, atm[k]) for k in ("cloud_opd", "cloud_w0", "cloud_g0")]
    oa.cloud_opd, oa.cloud_w0, oa.cloud_g0 = [x.ptr for x in cl]
    oa.cloud_ld, oa.stream, oa.delta_eddington = W, 2, 1
    outs = {n: po.DeviceArray(ctx, (L + (1 if n in ("TAU", "TAU_OG") else 0), W)) for n in names}
    for n, dv in outs.items():
        setattr(oa, n, dv.ptr)
    fn = ctx.lib.pb_compute_opacity
    ms = timeit(ctx, lambda i: ctx.check(fn(ctx.h, opa._tab, ctypes.byref(oa), PB_DEVICE)), 1, reps)
    # algorithmic bytes: every distinct table row touched once + cloud arrays + requested outputs
    nrows = len(np.unique(idx if query == "linear" else idx[:, :1]))
    ncont_rows = len(np.unique(cia))
    alg = (nmol * nrows + len(db["continuum"]) * ncont_rows + len(ray) + 3 * L) * W * 8
    alg += sum((L + (1 if n in ("TAU", "TAU_OG") else 0)) * W * 8 for n in names)
    if raman == 0:
        alg += (c.size + 1) * W * 8
    report("compute_opacity L=%d W=%d nmol=%d query=%s raman=%d outputs=%d%s" % (L, W, nmol, query, raman, len(names), tag),
           ms, alg, W, "wave-points", {"distinct_rows_per_molecule": nrows, "api_ms": api_ms,
                                       "table_bytes": opa.device_bytes(), "table_upload_s": t_up})
    opa.close()


def bench_thermal(ctx, L, W, G, reps, batch=1, calc_type=0, levels=False):
    per_set = (3 * L + 3) * 8 * W * batch
    nsets = min(6, max(1, int(np.ceil(2 * L2_BYTES / per_set))))
    args, keep = [], []
    V = L + 1
    d_f = ctx.dev_alloc(batch * G * W * 8)
    d_t = ctx.dev_alloc(batch * W * 8)
    d_lv = [ctx.dev_alloc(batch * G * V * W * 8) for _ in range(4)] if levels else None
    for s in range(nsets):
        ds = [synth.thermal_inputs(L=L, W=W, seed=70 + s * batch + b, ngauss=G) for b in range(batch)]
        a = ThermalArgs()
        a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = L, W, G, 1, batch, W
        for k in ("dtau", "w0", "cosb"):
            setattr(a, k, ctx.to_device(np.stack([d[k] for d in ds])))
        d0 = ds[0]
        a.wno, a.dwno = ctx.to_device(d0["wno"]), ctx.to_device(d0["dwno"])
        a.surf_reflect = None
        vec = [np.ascontiguousarray(np.stack([d["tlevel"] for d in ds])),
               np.ascontiguousarray(np.stack([d["plevel"] for d in ds])),
               np.ascontiguousarray(d0["ubar1"]).reshape(-1), np.ascontiguousarray(d0["gweight"]),
               np.ascontiguousarray(d0["tweight"])]
        keep.append(vec)
        a.tlevel, a.plevel, a.ubar1, a.gweight, a.tweight = [_lib.addr(v) for v in vec]
        a.hard_surface, a.calc_type = 0, calc_type
        a.flux_at_top, a.thermal = d_f, d_t
        if levels:
            a.flux_minus, a.flux_plus, a.flux_minus_mdpt, a.flux_plus_mdpt = d_lv
        args.append(a)
    fn = ctx.lib.pb_thermal_toon_1d
    ms = timeit(ctx, lambda i: ctx.check(fn(ctx.h, ctypes.byref(args[i]), PB_DEVICE)), nsets, reps)
    alg = ((3 * L + 3) * 8 + G * 8 + (4 * G * V * 8 if levels else 0)) * W * batch
    report("thermal_toon L=%d W=%d G=%d batch=%d calc_type=%d levels=%d" % (L, W, G, batch, calc_type, levels),
           ms, alg, W * batch, "wave-points", {"nsets": nsets})
    for a in args:
        for k in ("dtau", "w0", "cosb", "wno", "dwno"):
            ctx.dev_free(getattr(a, k))
    for p in [d_f, d_t] + (d_lv or []):
        ctx.dev_free(p)


def bench_transit(ctx, L, W, reps):
    per_set = L * 8 * W
    nsets = min(8, max(1, int(np.ceil(2 * L2_BYTES / per_set))))
    args, keep = [], []
    d_F = ctx.dev_alloc(W * 8)
    for s in range(nsets):
        d = synth.transit_inputs(L=L, W=W, seed=90 + s)
        a = TransitArgs()
        a.nlevel, a.nwno, a.nbatch, a.ld = L + 1, W, 1, W
        a.DTAU = ctx.to_device(d["DTAU"])
        vec = [np.ascontiguousarray(d[k], dtype=np.float64) for k in ("z", "dz", "player", "tlayer", "mmw", "colden")]
        keep.append(vec)
        a.z, a.dz, a.player, a.tlayer, a.mmw, a.colden = [_lib.addr(v) for v in vec]
        a.rstar, a.k_b, a.amu, a.F = d["rstar"], d["k_b"], d["amu"], d_F
        args.append(a)
    fn = ctx.lib.pb_transit_1d
    ms = timeit(ctx, lambda i: ctx.check(fn(ctx.h, ctypes.byref(args[i]), PB_DEVICE)), nsets, reps)
    report("transit L=%d W=%d" % (L, W), ms, (L + 1) * 8 * W, W, "wave-points", {"nsets": nsets})
    for a in args:
        ctx.dev_free(a.DTAU)
    ctx.dev_free(d_F)


def bench_mix(ctx, L, W, K, ngas, reps):
    """resort-rebin mixing at the climate solver's shape (661 bins x 8 gauss points, ~a dozen gases)"""
    from picaso_b200._lib import CkMixArgs
    nP, nT = 8, 9
    rng = np.random.default_rng(5)
    dk = [ctx.to_device(np.sort(rng.uniform(-70.0, -45.0, (nP, nT, W, K)), axis=3)) for _ in range(ngas)]
    ptrs = (ctypes.c_void_p * ngas)(*dk)
    mixes = np.ascontiguousarray(10.0 ** rng.uniform(-8, -0.5, (ngas, L)))
    ind = np.ascontiguousarray(np.stack([rng.integers(0, nP - 1, L), rng.integers(1, nP, L), rng.integers(0, nT - 1, L),
                                         rng.integers(1, nT, L)]), dtype=np.int32)
    ti, pi = rng.uniform(0, 1, L), rng.uniform(0, 1, L)
    x, w = np.polynomial.legendre.leggauss(K)
    gp, gw = np.ascontiguousarray(0.5 * (x + 1)), np.ascontiguousarray(0.5 * w)
    out = ctx.dev_alloc(L * W * K * 8)
    a = CkMixArgs(nlayer=L, nwno=W, ngauss=K, ngas=ngas, np=nP, nt=nT)
    a.kappas = ctypes.cast(ptrs, ctypes.c_void_p)
    a.mixes, a.indices, a.t_interp, a.p_interp = _lib.addr(mixes), _lib.addr(ind), _lib.addr(ti), _lib.addr(pi)
    a.gauss_pts, a.gauss_wts, a.molecular_opa = _lib.addr(gp), _lib.addr(gw), out
    fn = ctx.lib.pb_ck_mix
    ms = timeit(ctx, lambda i: ctx.check(fn(ctx.h, ctypes.byref(a), PB_DEVICE)), 1, reps)
    nmix = L * W * 4 * (ngas - 1)
    report("ck_mix L=%d W=%d K=%d ngas=%d" % (L, W, K, ngas), ms, (4 * ngas + 1) * L * W * K * 8, nmix, "pair-mixes",
           {"ns_per_pair_mix": ms * 1e6 / nmix})
    for d in dk:
        ctx.dev_free(d)
    ctx.dev_free(out)


def bench_climate(ctx, L, W, K, ng, reps):
    """picaso.climate.get_fluxes at the climate solver's shape: wall clock of the public call (host
    tuples in, 8 host arrays out), the same call on device-resident opacities, and the CPU oracle."""
    import time
    from picaso_b200.optics import DeviceArray
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from oracle import climate as oclim
    d = synth.climate_inputs(L=L, W=W, K=K, seed=77, ng=ng)
    args = [d["Atmosphere"], d["OpacityWEd"], d["OpacityNoEd"], d["ScatteringPhase"], d["Disco"], d["Opagrid"],
            d["F0PI"], True, True]
    dd = list(args)
    dd[1] = type(d["OpacityWEd"])(*[DeviceArray.from_numpy(ctx, a) for a in d["OpacityWEd"]])
    dd[2] = type(d["OpacityNoEd"])(*[DeviceArray.from_numpy(ctx, a) for a in d["OpacityNoEd"]])

    def wall(fn, n):
        fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t0) / n * 1e3

    l0 = ctx.launch_count()
    pb.get_fluxes(*dd, ctx=ctx)
    launches = ctx.launch_count() - l0
    ms_host = wall(lambda: pb.get_fluxes(*args, ctx=ctx), reps)
    ms_dev = wall(lambda: pb.get_fluxes(*dd, ctx=ctx), reps)
    ms_net = wall(lambda: pb.get_fluxes(*dd, ctx=ctx, full_arrays=False), reps)
    ms_cpu = wall(lambda: oclim.get_fluxes(*args, nthreads=os.cpu_count() or 1), 3)
    ms_cpu1 = wall(lambda: oclim.get_fluxes(*args, nthreads=1), 1)
    # algorithmic bytes: 12 opacity arrays read once + the 4 level arrays written and read back by the reductions
    alg = (12 * L + 2) * W * K * 8 + 2 * 4 * (1 + ng) * (L + 1) * W * K * 8
    # the Jacobian of t_start in one call: nlevel + 1 temperature profiles over one set of opacities (thermal half)
    V = L + 1
    t0_ = np.asarray(d["Atmosphere"].t_level, dtype=np.float64)
    tls = np.tile(t0_, (V + 1, 1))
    for j in range(V):
        tls[j + 1, j] += 0.01 * t0_[j]
    ms_jac = wall(lambda: pb.get_fluxes_jacobian(*dd[:6], tls, ctx=ctx), max(2, reps // 4))
    ms_one_ir = wall(lambda: pb.get_fluxes(*dd[:7], False, True, ctx=ctx, full_arrays=False), reps)
    print(json.dumps({"config": "climate.get_fluxes_jacobian L=%d W=%d ngauss=%d ng=%d: %d temperature profiles, thermal half" % (
                          L, W, K, ng, V + 1),
                      "ms_per_call": ms_jac, "ms_per_profile": ms_jac / (V + 1),
                      "ms_one_profile_per_call_thermal_only": ms_one_ir}), flush=True)
    print(json.dumps({"config": "climate.get_fluxes L=%d W=%d ngauss=%d ng=%d (reflected + thermal)" % (L, W, K, ng),
                      "ms_per_call_host_arrays": ms_host, "ms_per_call_device_opacities": ms_dev,
                      "ms_per_call_device_opacities_net_fluxes_only": ms_net,
                      "kernel_launches_per_call": launches, "alg_bytes": alg,
                      "cpu_port_ms_all_cores": ms_cpu, "cpu_port_ms_1_thread": ms_cpu1, "cores": os.cpu_count(),
                      "rt_columns_per_s": K / (ms_dev * 1e-3)}), flush=True)


def bench_thermal_batch(ctx, B, L, W, nbins, reps):
    """BASELINE cfg5 per-GPU share: B atmospheres x (L x W) thermal -> disk integration -> rebin, device resident"""
    from picaso_b200.optics import DeviceArray
    ds = [synth.thermal_inputs(L=L, W=W, seed=900 + b) for b in range(B)]
    d0 = ds[0]
    kw = dict(wno=d0["wno"], tlevel=np.array([d["tlevel"] for d in ds]), plevel=np.array([d["plevel"] for d in ds]),
              ubar1=d0["ubar1"], gweight=d0["gweight"], tweight=d0["tweight"])
    for k in ("dtau", "w0", "cosb"):
        kw[k] = DeviceArray.from_numpy(ctx, np.array([d[k] for d in ds]))
    newx = np.linspace(d0["wno"][5], d0["wno"][-5], nbins)
    fn = lambda i=0: pb.thermal_batch(**kw, newx=newx, scale=1e-8, ctx=ctx)
    l0 = ctx.launch_count()
    fn()
    launches = ctx.launch_count() - l0
    for _ in range(2):
        fn()
    ctx.sync()
    import time
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ms = (time.perf_counter() - t0) / reps * 1e3
    alg = ((3 * L + 3) * 8 + 8) * W * B
    report("thermal_batch B=%d L=%d W=%d -> %d bins (cfg5 per-GPU share; wall clock of the public call, device inputs)"
           % (B, L, W, nbins), ms, alg, W * B, "wave-points", {"launches_per_call": launches, "atmospheres/s": B / (ms * 1e-3)})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--only", default="refl,sh,opacity,thermal,transit,mix,batch,climate,retrieval")
    a = ap.parse_args()
    only = set(a.only.split(","))
    ctx = pb.Context(0)
    print(json.dumps({"device": ctx.device_name(), "hbm_peak_gbs": hbm_peak()}), flush=True)
    if "refl" in only:
        bench_reflected(ctx, 60, 300, 5, a.reps, sp=1, tag=" (cfg1)")
        bench_reflected(ctx, 60, 10000, 5, a.reps, tag=" (headline)")
        bench_reflected(ctx, 60, 10000, 5, a.reps, sp=1, tag=" (headline OTHG)")
        bench_reflected(ctx, 60, 196000, 5, max(5, a.reps // 5), tag=" (cfg3 shape, Toon)")
    if "sh" in only:
        bench_sh(ctx, 60, 10000, 5, max(5, a.reps // 5), stream=2)
        bench_sh(ctx, 60, 10000, 5, max(5, a.reps // 5), stream=4)
        bench_sh(ctx, 60, 196000, 5, 5, stream=4)
        bench_sh(ctx, 60, 196000, 5, 5, stream=4, forms=(1, 1, 1, 1, 1, 1))
    if "shcfg3" in only:   # the one launch an ncu capture of cfg3 wants
        bench_sh(ctx, 60, 196000, 5, 1, stream=4, forms=(1, 1, 1, 1, 1, 1))
    if "opacity" in only:
        bench_opacity(ctx, 80, 50000, 12, a.reps, outputs=("DTAU_OG",), tag=" (cfg4: transit needs DTAU only)")
        bench_opacity(ctx, 80, 50000, 12, a.reps, tag=" (cfg4, all 13 outputs)")
        bench_opacity(ctx, 80, 50000, 12, a.reps, query="nearest", tag=" (reference default query)")
        bench_opacity(ctx, 60, 10000, 12, a.reps, raman=0, tag=" (headline shape, Raman on)")
    if "thermal" in only:
        bench_thermal(ctx, 90, 10000, 5, a.reps)
        bench_thermal(ctx, 90, 10000, 5, max(5, a.reps // 5), calc_type=1, levels=True)
        bench_thermal(ctx, 90, 100000, 5, max(5, a.reps // 5))
    if "transit" in only:
        bench_transit(ctx, 80, 50000, a.reps)
    if "mix" in only:
        bench_mix(ctx, 90, 661, 8, 12, max(5, a.reps // 5))
    if "batch" in only:
        bench_thermal(ctx, 60, 2000, 5, max(5, a.reps // 5), batch=128)
        bench_reflected(ctx, 60, 10000, 5, max(5, a.reps // 5), batch=8, tag=" (8 spectra/launch)")
    if "climate" in only:
        bench_climate(ctx, 90, 661, 8, 1, max(5, a.reps // 5))
    if "retrieval" in only:
        bench_thermal_batch(ctx, 128, 60, 2000, 300, max(5, a.reps // 5))
    ctx.close()


if __name__ == "__main__":
    main()
