"""bench_configs.py - the other BASELINE.json configurations behind `bench.py --config cfgN`.

    python bench.py --config cfg2|cfg3|cfg4|cfg5 [--gpus N --steps K --warmup W] [--impl reference]

Same contract as the headline line (bench.py): parity gate on the benchmarked configuration first, then W warm-up
steps, K steps timed with CUDA events on the launching stream (max over ranks), `e2e` through the public Python API
with host buffers, `roofline` for the dominant kernel, `cpu_baseline` (the C / numpy port in oracle/, all host
threads, bounded sample) on rank 0 at N = 1.

  cfg2  thermal Toon, 90 layers x 10 000 waves x 5 angles (get_thermal_1d + compress_thermal); N > 1: weak
        (every rank its own spectrum), no data-path collective.
  cfg3  SH4 reflected, 60 layers x 196 000 waves x 5 angles, OTHG forms + Rayleigh (get_reflected_SH + compress_disco);
        N > 1: STRONG scaling - the wave grid of one spectrum is split into N contiguous slabs, one ncclAllGather of the
        albedo vector at the end of every step (north_star's split).
  cfg4  transit, 80 layers x 50 000 waves: opacity query (bilinear, 12 molecules) + compute_opacity (DTAU only) +
        get_transit_1d, tables resident in HBM; N > 1: weak.
  cfg5  batched retrieval, 1024 atmospheres x 60 layers x 2000 waves thermal + rebin to 300 bins (thermal_batch);
        N > 1: STRONG scaling over atmospheres, one all-gather of the [1024, 300] rebinned spectra per step.
"""
import ctypes
import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
L2_BYTES = 126e6
UNIT = "wave-points/s"


def _dist(world, local_rank):
    if world <= 1:
        return None, None
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    return torch, dist


def _max_over_ranks(torch, dist, x):
    if dist is None:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def _cpu_rate(fn, units, budget_s=10.0, max_n=50):
    """units per second of fn() (one untimed call first), bounded wall time"""
    fn()
    t0 = time.perf_counter()
    n = 0
    while True:
        fn()
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= max_n:
            break
    return units * n / el, n, el


# ---------------------------------------------------------------------------------------------------------
class Cfg2:
    """thermal Toon 90 x 10 000 x 5"""
    name = "cfg2"
    L, W, G = 90, 10000, 5
    metric = "wave-points/sec (90-layer x 10k-wave thermal Toon spectrum)"
    workload = "thermal_toon_1d L=90 W=10000 G=5 calc_type=0 hard_surface=0 (BASELINE config 2)"
    scaling = "weak"
    kernel = "therm_toa_chain_kernel"
    dtype = "f64"

    def __init__(self, pb, ctx, rank, world, torch=None, dist=None):
        from picaso_b200 import _lib, synth
        from picaso_b200._lib import ThermalArgs
        self.pb, self.ctx, self.rank, self.world = pb, ctx, rank, world
        L, W, G = self.L, self.W, self.G
        self.alg_bytes = ((3 * L + 3) * 8 + G * 8) * W          # per launch (2 224 B per wave-point)
        self.nsets = 7                                           # 7 x 21.6 MB = 151 MB > L2
        self.sets = [synth.thermal_inputs(L=L, W=W, seed=1002 + 97 * rank + i) for i in range(self.nsets)]
        self.units_per_step = W * world
        self.d_f = ctx.dev_alloc(G * W * 8)
        self.d_t = ctx.dev_alloc(W * 8)
        self.keep, self.args = [], []
        for d in self.sets:
            a = ThermalArgs()
            a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = L, W, G, 1, 1, W
            for k in ("dtau", "w0", "cosb"):
                setattr(a, k, ctx.to_device(d[k]))
            a.wno = ctx.to_device(d["wno"])
            a.surf_reflect = None
            vec = [np.ascontiguousarray(d["tlevel"]), np.ascontiguousarray(d["plevel"]),
                   np.ascontiguousarray(d["ubar1"]).reshape(-1), np.ascontiguousarray(d["gweight"]),
                   np.ascontiguousarray(d["tweight"])]
            self.keep.append(vec)
            a.tlevel, a.plevel, a.ubar1, a.gweight, a.tweight = [_lib.addr(v) for v in vec]
            a.hard_surface, a.calc_type = 0, 0
            a.flux_at_top, a.thermal = self.d_f, self.d_t
            self.args.append(a)
        self.fn = ctx.lib.pb_thermal_toon_1d
        self.l2_policy = "inputs larger than L2: %d input sets x %.1f MB rotated" % (self.nsets, 3 * L * W * 8 / 1e6)
        self.collective = "none (independent spectra per rank)"

    def step(self, i):
        from picaso_b200._lib import PB_DEVICE
        self.ctx.check(self.fn(self.ctx.h, ctypes.byref(self.args[i % self.nsets]), PB_DEVICE))

    def finish(self):
        pass

    def _oracle(self, d, nthreads):
        import cases as C
        import oracle
        t = dict(d, calc_type=0)
        f, _ = oracle.get_thermal_1d(*C.thermal_args(t), nthreads=nthreads, level_fluxes=False)
        return f, oracle.compress_thermal(self.W, f, d["gweight"], d["tweight"])

    def parity(self):
        self.step(0)
        self.ctx.sync()
        got_f = self.ctx.from_device(self.d_f, (self.G, 1, self.W))
        got_t = self.ctx.from_device(self.d_t, (self.W,))
        f, th = self._oracle(self.sets[0], os.cpu_count() or 1)
        return max(_rel(got_f, f), _rel(got_t, th))

    def e2e_setup(self):
        import cases as C
        self.pinned = []
        for d in self.sets[:4]:
            pd = dict(d, calc_type=0)
            for k in ("dtau", "w0", "cosb"):
                buf = self.ctx.pinned_empty(d[k].shape)
                buf[...] = d[k]
                pd[k] = buf
            self.pinned.append(C.thermal_args(pd))
        self.h2d = 3 * self.L * self.W * 8 + 3 * self.W * 8
        self.d2h = (self.G + 1) * self.W * 8
        self.e2e_api = "picaso_b200.get_thermal_1d(..., level_fluxes=False, return_thermal=True), pinned host inputs"

    def e2e_step(self, i):
        d = self.sets[i % 4]
        return self.pb.get_thermal_1d(*self.pinned[i % 4], ctx=self.ctx, level_fluxes=False, gweight=d["gweight"],
                                      tweight=d["tweight"], return_thermal=True)

    def cpu_baseline(self):
        n = os.cpu_count() or 1
        v, k, el = _cpu_rate(lambda: self._oracle(self.sets[0], n), self.W)
        t0 = time.perf_counter()
        self._oracle(self.sets[0], 1)
        one = time.perf_counter() - t0
        return {"value": v, "unit": UNIT, "cores": n, "kind": "port",
                "sample": "%d full 90x10000x5 thermal spectra in %.1f s; C port of the reference algorithm (oracle/), OpenMP over wavelengths" % (k, el),
                "single_thread_value": self.W / one, "host_cpus": os.cpu_count()}

    def note(self):
        return ("issue / fp64-latency bound (DESIGN.md 4.3, profiles/r2_therm_chain_cfg2.summary.json): the angle-independent "
                "tridiagonal elimination runs once per wavelength on a chain warp, Planck per level once per CTA")


# ---------------------------------------------------------------------------------------------------------
class Cfg3:
    """SH4 reflected 60 x 196 000 x 5, strong scaling over the wave grid"""
    name = "cfg3"
    L, W, G = 60, 196000, 5
    metric = "wave-points/sec (60-layer x 196k-wave SH4 reflected spectrum)"
    scaling = "strong"
    kernel = "sh_reflected_kernel<4>"
    dtype = "f64"
    FORMS = (1, 1, 1, 1, 1, 1)   # OTHG single / multi / p_single forms, Rayleigh on (drift-free)

    def __init__(self, pb, ctx, rank, world, torch=None, dist=None):
        from picaso_b200 import _lib, synth
        from picaso_b200._lib import ShArgs
        from picaso_b200.sharded import wave_slice
        self.pb, self.ctx, self.rank, self.world, self.torch, self.dist = pb, ctx, rank, world, torch, dist
        L, G = self.L, self.G
        forms = os.environ.get("PB_BENCH_SH_FORMS", "othg")
        if forms == "tthg":
            self.FORMS = (0, 0, 0, 1, 1, 1)
        self.workload = ("reflected_SH stream=4 L=60 W=196000 G=5 forms=%s rayleigh=on delta-M (BASELINE config 3)"
                         % ("".join(map(str, self.FORMS))))
        sl = wave_slice(self.W, rank, world)
        self.w0, self.Wr = sl.start, sl.stop - sl.start
        per_set = (9 * L + 2 * (L + 1) + 2) * 8 * self.Wr
        self.nsets = 1 if per_set > 2 * L2_BYTES else 2
        self.alg_bytes = ((9 * L + 2 * (L + 1) + 2) * 8 + G * 8) * self.Wr
        self.units_per_step = self.W
        lay = ("dtau", "w0", "ftau_cld", "ftau_ray", "f_deltaM", "dtau_og", "w0_og", "cosb_og")
        lev = ("tau", "tau_og")
        wav = ("surf_reflect", "F0PI")
        self.keys = lay + lev + wav
        # the spectrum's wave slab of this rank (columns are independent: a slab with its own seed is a slab of
        # the full grid)
        self.sets = [synth.reflected_inputs(L=L, W=self.Wr, seed=1003 + 977 * rank + i, ngauss=G, stream=4)
                     for i in range(self.nsets)]
        self.d_x = ctx.dev_alloc(G * self.Wr * 8)
        if world > 1:
            self.t_mine = torch.empty((self.Wr,), dtype=torch.float64, device="cuda")
            self.t_all = torch.empty((world, self.Wr), dtype=torch.float64, device="cuda")
            if self.W % world:
                raise SystemExit("cfg3: 196000 waves must divide over the ranks")
            self.d_a = self.t_mine.data_ptr()
        else:
            self.d_a = ctx.dev_alloc(self.Wr * 8)
        self.keep, self.args = [], []
        for d in self.sets:
            a = ShArgs()
            a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = L, self.Wr, G, 1, 1, self.Wr
            for k in self.keys:
                setattr(a, k, ctx.to_device(d[k]))
            vec = [np.ascontiguousarray(d[k]).reshape(-1) for k in ("ubar0", "ubar1", "gweight", "tweight")]
            self.keep.append(vec)
            a.ubar0, a.ubar1, a.gweight, a.tweight = [_lib.addr(v) for v in vec]
            a.cos_theta = d["cos_theta"]
            (a.w_single_form, a.w_multi_form, a.psingle_form, a.w_single_rayleigh, a.w_multi_rayleigh,
             a.psingle_rayleigh) = self.FORMS
            a.frac_a, a.frac_b, a.frac_c = d["frac_a"], d["frac_b"], d["frac_c"]
            a.constant_back, a.constant_forward = d["constant_back"], d["constant_forward"]
            a.stream, a.flx, a.single_form = 4, 0, 0
            a.xint_at_top, a.albedo = self.d_x, self.d_a
            self.args.append(a)
        self.fn = ctx.lib.pb_reflected_sh
        self.l2_policy = "inputs larger than L2: %d set(s) x %.0f MB per rank" % (self.nsets, per_set / 1e6)
        self.collective = ("none" if world == 1 else
                           "one ncclAllGather of the albedo slab [W/N] -> [W] per step on the launching stream")

    def step(self, i):
        from picaso_b200._lib import PB_DEVICE
        self.ctx.check(self.fn(self.ctx.h, ctypes.byref(self.args[i % self.nsets]), PB_DEVICE))
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.t_all, self.t_mine)

    def finish(self):
        pass

    def _sample(self):
        return np.arange(0, self.Wr, max(1, self.Wr // 1536))

    def _oracle(self, d, idx, nthreads):
        import cases as C
        import oracle
        ds = dict(d)
        for k, v in d.items():
            if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[-1] == d["nwno"]:
                ds[k] = np.ascontiguousarray(v[..., idx])
        ds["nwno"] = len(idx)
        x, _ = oracle.get_reflected_SH(*C.sh_args(ds, dict(forms=self.FORMS, stream=4, single_form=0)), nthreads=nthreads)
        return x, oracle.compress_disco(len(idx), ds["cos_theta"], x, ds["gweight"], ds["tweight"], ds["F0PI"])

    def parity(self):
        self.step(0)
        self.ctx.sync()
        if self.world > 1:
            self.torch.cuda.synchronize()
            got_a = self.t_all[self.rank].cpu().numpy()
        else:
            got_a = self.ctx.from_device(self.d_a, (self.Wr,))
        got_x = self.ctx.from_device(self.d_x, (self.G, 1, self.Wr))
        idx = self._sample()
        x, alb = self._oracle(self.sets[0], idx, os.cpu_count() or 1)
        return max(_rel(got_x[..., idx], x), _rel(got_a[idx], alb))

    def e2e_setup(self):
        import cases as C
        d = self.sets[0]
        pd = dict(d)
        for k in self.keys:
            buf = self.ctx.pinned_empty(d[k].shape)
            buf[...] = d[k]
            pd[k] = buf
        self.e2e_args = C.sh_args(pd, dict(forms=self.FORMS, stream=4, single_form=0))
        self.h2d = sum(d[k].nbytes for k in self.keys) + self.Wr * 8
        self.d2h = (self.G + 1) * self.Wr * 8
        self.e2e_api = "picaso_b200.get_reflected_SH(..., stream=4, return_albedo=True) on the rank's wave slab, pinned host inputs"

    def e2e_step(self, i):
        d = self.sets[0]
        return self.pb.get_reflected_SH(*self.e2e_args, ctx=self.ctx, gweight=d["gweight"], tweight=d["tweight"],
                                        return_albedo=True, inplace_f_deltaM=False)

    def cpu_baseline(self):
        n = os.cpu_count() or 1
        idx = np.arange(0, self.Wr, self.Wr // 2000)[:2000]
        v, k, el = _cpu_rate(lambda: self._oracle(self.sets[0], idx, n), len(idx), budget_s=12.0, max_n=20)
        t0 = time.perf_counter()
        self._oracle(self.sets[0], idx[:200], 1)
        one = time.perf_counter() - t0
        return {"value": v, "unit": UNIT, "cores": n, "kind": "port",
                "sample": "%d x 2000 of the 196000 wavelengths (every 98th) in %.1f s; C port (dgbsv restated), OpenMP over wavelengths" % (k, el),
                "single_thread_value": 200 / one, "host_cpus": os.cpu_count()}

    def note(self):
        return "fp64-pipe bound: pivoted block elimination in registers, see DESIGN.md 4.4"


# ---------------------------------------------------------------------------------------------------------
class Cfg4:
    """transit 80 x 50 000: opacity query + mixing + chord integration"""
    name = "cfg4"
    L, W, NMOL = 80, 50000, 12
    metric = "wave-points/sec (80-layer x 50k-wave transit spectrum incl. opacity interpolation)"
    workload = ("transit L=80 W=50000: get_opacities(bilinear, 12 molecules) + compute_opacity(DTAU) + get_transit_1d, "
                "tables resident in HBM (BASELINE config 4)")
    scaling = "weak"
    kernel = "opacity_layer_kernel"
    dtype = "f64"
    NPROF = 4

    def __init__(self, pb, ctx, rank, world, torch=None, dist=None):
        from picaso_b200 import synth
        self.pb, self.ctx, self.rank, self.world = pb, ctx, rank, world
        L, W = self.L, self.W
        db = synth.opacity_database(W=W, nmol=self.NMOL, seed=4001 + rank, nT=20, nP=18, nTc=30, wave_range=(0.3, 5.0))
        rng = np.random.default_rng(1)
        ray = {m: 10.0 ** rng.uniform(-27, -25, W) for m in db["rayleigh_molecules"]}
        self.db, self.ray = db, ray
        t0 = time.perf_counter()
        self.opa = pb.DeviceOpacities(db["wno"], db["pt_pairs"], db["tables"], db["cia_temps"], db["continuum"], ray,
                                      query_method="linear", ctx=ctx)
        self.upload_s = time.perf_counter() - t0
        self.atms, self.ducks = [], []
        for i in range(self.NPROF):
            atm = synth.atmosphere_profile(db, L=L, seed=4100 + i, cloudy=True)
            a = types.SimpleNamespace()
            a.c = types.SimpleNamespace(nlayer=L, pconv=atm["pconv"], rgas=atm["rgas"], amu=atm["amu"], k_b=atm["k_b"])
            a.level = {"temperature": atm["tlevel"], "pressure": atm["plevel"]}
            a.layer = {"temperature": atm["tlayer"], "pressure": atm["player"], "colden": atm["colden"], "mmw": atm["mmw"],
                       "mixingratios": atm["mixingratios"], "electrons": atm["electrons"],
                       "cloud": {"opd": atm["cloud_opd"], "w0": atm["cloud_w0"], "g0": atm["cloud_g0"]}}
            a.planet = types.SimpleNamespace(gravity=atm["gravity"])
            a.molecules, a.rayleigh_molecules = list(db["molecules"]), list(db["rayleigh_molecules"])
            a.continuum_molecules = [list(x) for x in db["continuum_molecules"]]
            atm["cia_pairs"] = {x + y: (x, y) for x, y in db["continuum_molecules"]}
            self.atms.append(atm)
            self.ducks.append(a)
        self.tr = synth.transit_inputs(L=L, W=W, seed=1004)
        self.units_per_step = W * world
        # algorithmic bytes of the dominant kernel: every distinct table row touched once + cloud arrays + DTAU out
        self.opa.get_opacities(self.ducks[0])
        idx, cia = self.opa._plan["idx"], self.opa._plan["cia"]
        nrows, ncont = len(np.unique(idx)), len(np.unique(cia))
        self.alg_bytes = (self.NMOL * nrows + len(db["continuum"]) * ncont + len(ray) + 3 * L) * W * 8 + L * W * 8
        self.l2_policy = "tables (%.1f GB) >> L2; %d atmosphere profiles rotated" % (self.opa.device_bytes() / 1e9, self.NPROF)
        self.collective = "none (independent spectra per rank)"
        self._prepare_device_steps()

    def _prepare_device_steps(self):
        """timed region: inputs resident in HBM - per profile a prebuilt pb_opacity_args (table-row plan, layer scalars,
        cloud arrays on the device) and pb_transit_args reading the DTAU the opacity kernel just wrote"""
        from picaso_b200 import _lib, optics as po
        from picaso_b200._lib import OpacityArgs, TransitArgs
        ctx, L, W, t = self.ctx, self.L, self.W, self.tr
        self.d_dtau = po.DeviceArray(ctx, (L, W))
        self.d_F = ctx.dev_alloc(W * 8)
        self.dev_args, self.keep = [], []
        for a, atm in zip(self.ducks, self.atms):
            self.opa.get_opacities(a)
            mol, cont, rays = po._layer_scalars(a, self.opa)
            idx, wts, cia = (np.array(self.opa._plan[k]) for k in ("idx", "wts", "cia"))
            oa = OpacityArgs()
            oa.nlayer, oa.query = L, 1
            oa.pt_index, oa.weights, oa.cont_index = _lib.addr(idx), _lib.addr(wts), _lib.addr(cia)
            oa.mol_scale, oa.cont_scale, oa.ray_scale = _lib.addr(mol), _lib.addr(cont), _lib.addr(rays)
            oa.raman = 2
            cl = [po.DeviceArray.from_numpy(ctx, atm[k]) for k in ("cloud_opd", "cloud_w0", "cloud_g0")]
            oa.cloud_opd, oa.cloud_w0, oa.cloud_g0 = [x.ptr for x in cl]
            oa.cloud_ld, oa.stream, oa.delta_eddington = W, 2, 1
            oa.DTAU_OG = self.d_dtau.ptr
            ta = TransitArgs()
            ta.nlevel, ta.nwno, ta.nbatch, ta.ld = L + 1, W, 1, W
            ta.DTAU = self.d_dtau.ptr
            vec = [np.ascontiguousarray(x, dtype=np.float64) for x in
                   (t["z"], t["dz"], a.level["pressure"], a.level["temperature"], a.layer["mmw"], a.layer["colden"])]
            ta.z, ta.dz, ta.player, ta.tlayer, ta.mmw, ta.colden = [_lib.addr(v) for v in vec]
            ta.rstar, ta.k_b, ta.amu, ta.F = t["rstar"], t["k_b"], t["amu"], self.d_F
            self.dev_args.append((oa, ta))
            self.keep.append((idx, wts, cia, mol, cont, rays, cl, vec))

    def _one(self, i):
        a = self.ducks[i % self.NPROF]
        t = self.tr
        self.opa.get_opacities(a)
        dev = self.pb.compute_opacity(a, self.opa, ngauss=1, stream=2, delta_eddington=True, test_mode=None, raman=2,
                                      device_outputs=True, outputs=("DTAU_OG",))
        dt = dev[7]   # DTAU_OG (OUTPUT_NAMES order of the reference's 13-tuple)
        return self.pb.get_transit_1d(t["z"], t["dz"], t["nlevel"], t["nwno"], t["rstar"], a.layer["mmw"], t["k_b"], t["amu"],
                                      a.level["pressure"], a.level["temperature"], a.layer["colden"], dt[:, :, 0], ctx=self.ctx)

    def step(self, i):
        from picaso_b200._lib import PB_DEVICE
        oa, ta = self.dev_args[i % self.NPROF]
        self.ctx.check(self.ctx.lib.pb_compute_opacity(self.ctx.h, self.opa._tab, ctypes.byref(oa), PB_DEVICE))
        self.ctx.check(self.ctx.lib.pb_transit_1d(self.ctx.h, ctypes.byref(ta), PB_DEVICE))

    def finish(self):
        pass

    def _cpu(self, i, nthreads):
        import oracle
        from oracle import optics as oo
        db, atm, t = self.db, self.atms[i % self.NPROF], self.tr
        pbar = atm["player"] / atm["pconv"]
        ti, pi, ill, ihl, ilh, ihh = oo.find_needed_pts(db["temps"], db["pressures"], db["nc_p"], atm["tlayer"], pbar)
        mol = {m: oo.interp_molecular(db["tables"][m], ti, pi, ill, ihl, ilh, ihh) for m in db["molecules"]}
        ic = oo.nearest_cia_temp(db["cia_temps"], atm["tlayer"])
        cont = {k: db["continuum"][k][ic] for k in db["continuum"]}
        o = oo.compute_opacity(atm, mol, cont, self.ray, None, stream=2, delta_eddington=True)
        return oracle.get_transit_1d(t["z"], t["dz"], t["nlevel"], t["nwno"], t["rstar"], atm["mmw"], t["k_b"], t["amu"],
                                     atm["plevel"], atm["tlevel"], atm["colden"], o[7], nthreads=nthreads)

    def parity(self):
        want = self._cpu(0, os.cpu_count() or 1)
        self.step(0)
        self.ctx.sync()
        got_dev = self.ctx.from_device(self.d_F, (self.W,))
        got_api = self._one(0)
        return max(_rel(got_dev, want), _rel(got_api, want))

    def e2e_setup(self):
        nsc = self.L * (self.NMOL + len(self.db["continuum"]) + len(self.ray) + 12) * 8
        self.h2d = nsc + 3 * self.L * self.W * 8       # layer scalars + the three cloud arrays
        self.d2h = self.W * 8
        self.e2e_api = ("DeviceOpacities.get_opacities(linear) + compute_opacity(device_outputs=True, outputs=('DTAU_OG',)) + "
                        "get_transit_1d(DeviceArray): host profile in, transit depth [W] out")

    def e2e_step(self, i):
        return self._one(i)

    def cpu_baseline(self):
        n = os.cpu_count() or 1
        v, k, el = _cpu_rate(lambda: self._cpu(0, n), self.W, budget_s=10.0, max_n=20)
        return {"value": v, "unit": UNIT, "cores": n, "kind": "port",
                "sample": "%d full 80x50000 transit spectra in %.1f s; numpy port of the opacity chain + C port of get_transit_1d" % (k, el),
                "host_cpus": os.cpu_count()}

    def note(self):
        return ("2 launches per step (opacity_layer_kernel, transit_kernel) through pb_compute_opacity + pb_transit_1d on "
                "device-resident plans / cloud arrays; achieved = the opacity kernel's algorithmic bytes over the WHOLE step time; "
                "gathers of 4 table rows per molecule per layer, see DESIGN.md 4.5")


# ---------------------------------------------------------------------------------------------------------
class Cfg5:
    """batched retrieval: 1024 atmospheres x 60 x 2000 thermal + rebin, strong scaling over atmospheres"""
    name = "cfg5"
    B, L, W, G, NBINS = 1024, 60, 2000, 5, 300
    metric = "wave-points/sec (1024 atmospheres x 60 layers x 2000 waves thermal batch)"
    workload = ("thermal_batch B=1024 L=60 W=2000 G=5 -> compress_thermal -> mean_regrid to 300 bins, atmospheres sharded "
                "over ranks (BASELINE config 5)")
    scaling = "strong"
    kernel = "therm_toa_wave_kernel<5>"
    dtype = "f64"

    def __init__(self, pb, ctx, rank, world, torch=None, dist=None):
        from picaso_b200 import synth
        from picaso_b200.batch import shard
        from picaso_b200.optics import DeviceArray
        self.pb, self.ctx, self.rank, self.world, self.torch, self.dist = pb, ctx, rank, world, torch, dist
        self.b0, b1 = shard(self.B, rank, world)
        self.Br = b1 - self.b0
        if self.B % world:
            raise SystemExit("cfg5: 1024 atmospheres must divide over the ranks")
        L, W = self.L, self.W
        ds = [synth.thermal_inputs(L=L, W=W, seed=5000 + b) for b in range(self.b0, self.b0 + self.Br)]
        d0 = ds[0]
        self.host = dict(wno=d0["wno"], tlevel=np.array([d["tlevel"] for d in ds]), plevel=np.array([d["plevel"] for d in ds]),
                         ubar1=d0["ubar1"], gweight=d0["gweight"], tweight=d0["tweight"])
        self.harr = {k: np.array([d[k] for d in ds]) for k in ("dtau", "w0", "cosb")}
        self.kw = dict(self.host)
        for k in ("dtau", "w0", "cosb"):
            self.kw[k] = DeviceArray.from_numpy(ctx, self.harr[k])
        self.newx = np.linspace(d0["wno"][5], d0["wno"][-5], self.NBINS)
        self.ds = ds
        self.alg_bytes = ((3 * L + 3) * 8 + 8) * W * self.Br
        self.units_per_step = self.B * W
        per = 3 * L * W * 8 * self.Br
        self.l2_policy = "inputs larger than L2: %.0f MB per rank, one set" % (per / 1e6)
        if world > 1:
            self.t_all = torch.empty((world, self.Br, self.NBINS - 1), dtype=torch.float64, device="cuda")
        self.collective = ("none" if world == 1 else "one ncclAllGather of the rebinned spectra [B/N, 299] -> [B, 299] per step")
        self.last = None

    def step(self, i):
        x, y = self.pb.thermal_batch(**self.kw, newx=self.newx, scale=1e-8, ctx=self.ctx, device_output=True)
        self.last = (x, y)
        if self.world > 1:
            src = _as_tensor(self.torch, y.ptr, (self.Br, y.shape[1]))
            if self.t_all.shape[2] != y.shape[1]:
                self.t_all = self.torch.empty((self.world, self.Br, y.shape[1]), dtype=self.torch.float64, device="cuda")
            self.dist.all_gather_into_tensor(self.t_all, src)

    def finish(self):
        pass

    def parity(self):
        from oracle import regrid as oreg
        self.step(0)
        self.ctx.sync()
        x, y = self.last
        got = y.numpy()
        nb = min(self.Br, 6)
        sel = np.linspace(0, self.Br - 1, nb).astype(int)
        hk = dict(self.host, tlevel=self.host["tlevel"][sel], plevel=self.host["plevel"][sel])
        for k in ("dtau", "w0", "cosb"):
            hk[k] = self.harr[k][sel]
        xo, yo = oreg.thermal_batch(**hk, newx=self.newx, scale=1e-8)
        m = np.isfinite(yo)
        assert np.array_equal(np.isfinite(got[sel]), m)
        return _rel(got[sel][m], yo[m])

    def e2e_setup(self):
        self.hkw = dict(self.host)
        for k in ("dtau", "w0", "cosb"):
            buf = self.ctx.pinned_empty(self.harr[k].shape)
            buf[...] = self.harr[k]
            self.hkw[k] = buf
        self.h2d = 3 * self.L * self.W * 8 * self.Br + 2 * (self.L + 1) * 8 * self.Br
        self.d2h = self.Br * (self.NBINS - 1) * 8
        self.e2e_api = "picaso_b200.thermal_batch(host arrays [B/N, 60, 2000] x 3, newx=300 bins): pinned inputs in, [B/N, 299] out"

    def e2e_step(self, i):
        return self.pb.thermal_batch(**self.hkw, newx=self.newx, scale=1e-8, ctx=self.ctx)

    def cpu_baseline(self):
        from oracle import regrid as oreg
        n = os.cpu_count() or 1
        nb = 16
        hk = dict(self.host, tlevel=self.host["tlevel"][:nb], plevel=self.host["plevel"][:nb])
        for k in ("dtau", "w0", "cosb"):
            hk[k] = self.harr[k][:nb]
        v, k, el = _cpu_rate(lambda: oreg.thermal_batch(**hk, newx=self.newx, scale=1e-8, nthreads=n), nb * self.W, budget_s=10.0, max_n=40)
        return {"value": v, "unit": UNIT, "cores": n, "kind": "port",
                "sample": "%d x 16 of the 1024 atmospheres in %.1f s; C port of get_thermal_1d + scipy binned_statistic (the reference's own rebin call)" % (k, el),
                "host_cpus": os.cpu_count()}

    def note(self):
        return "one thread per (atmosphere, wavelength), all 5 angles in registers; 2 launches per step; see DESIGN.md 4.3 / 4.5d"


def _as_tensor(torch, ptr, shape):
    """zero-copy torch view of a device buffer owned by the picaso_b200 context (bench plumbing for NCCL only)"""
    n = int(np.prod(shape))

    class _Cai:
        pass
    o = _Cai()
    o.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(o, device="cuda").view(*shape)


CONFIGS = {"cfg2": Cfg2, "cfg3": Cfg3, "cfg4": Cfg4, "cfg5": Cfg5}


# ---------------------------------------------------------------------------------------------------------
def run(args, rank, local_rank, world):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import bench
    import picaso_b200 as pb
    if args.impl == "reference":
        return run_reference(args, rank)
    numa = bench.pin_to_gpu_numa(local_rank) if world > 1 else None
    torch, dist = _dist(world, local_rank)
    ctx = pb.Context(local_rank)
    if world > 1:
        side = torch.cuda.Stream()
        torch.cuda.set_stream(side)
        ctx.set_stream(side.cuda_stream)
    cfg = CONFIGS[args.config](pb, ctx, rank, world, torch, dist)
    warm = max(args.warmup, 3)

    def barrier():
        ctx.sync()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    parity = cfg.parity()
    if not parity < 1e-6:
        raise SystemExit("%s parity gate failed: max rel err %.3e" % (cfg.name, parity))
    sampler = bench.ClockSampler(local_rank)
    sampler.start()
    for i in range(warm):
        cfg.step(i)
    ctx.sync()
    sampler.active = True
    # clock ramp: keep the GPU busy ~0.3 s before the timed region (fixed count: every rank issues the same collectives)
    barrier()
    ctx.timer_start()
    for i in range(3):
        cfg.step(i)
    # the ramp length must be the SAME on every rank (steps may contain collectives): agree on the slowest probe
    probe_ms = _max_over_ranks(torch, dist, max(ctx.timer_stop() / 3, 1e-3))
    for i in range(int(min(300.0 / probe_ms, 3000))):
        cfg.step(i)
    barrier()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for i in range(args.steps):
        cfg.step(i)
    cfg.finish()
    ms = ctx.timer_stop()
    sampler.active = False
    launches = ctx.launch_count() - l0
    barrier()
    ms = _max_over_ranks(torch, dist, ms)
    ms_per_step = ms / args.steps
    value = cfg.units_per_step * args.steps / (ms * 1e-3)

    # ---- end to end through the public API: host buffers in, host result out ----
    cfg.e2e_setup()
    if world > 1:
        ctx.set_stream(None)
    ke = args.e2e_steps or min(args.steps, 10)
    for i in range(2):
        cfg.e2e_step(i)
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    for i in range(ke):
        cfg.e2e_step(i)
    e2e_dt = time.perf_counter() - t0
    sampler.active = False
    sampler.stop_flag = True
    e2e_dt = _max_over_ranks(torch, dist, e2e_dt)
    peak, peak_src = bench.peaks()
    achieved = cfg.alg_bytes / (ms_per_step * 1e-3) / 1e9
    out = {
        "metric": cfg.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": cfg.scaling, "vs_baseline": None,
        "dtype": cfg.dtype, "data": "synthetic",
        "config": {"workload": cfg.workload, "bench_config": cfg.name, "l2_policy": cfg.l2_policy,
                   "collective": cfg.collective, "parity_max_rel_err": parity},
        "gpu_launches": int(launches),
        "e2e": {"value": cfg.units_per_step * ke / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": int(cfg.h2d),
                "d2h_bytes_per_step": int(cfg.d2h), "steps": ke, "ms_per_step": 1e3 * e2e_dt / ke, "api": cfg.e2e_api,
                "host_affinity": numa},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": bench.traffic_for(cfg.kernel), "kernel": cfg.kernel,
                     "algorithmic_bytes_per_launch": int(cfg.alg_bytes), "peak_source": peak_src,
                     "note": cfg.note() + ("; step time includes every launch of the step" if launches > args.steps else "")},
        "clocks": sampler.summary(),
    }
    try:
        out["roofline"]["fp64_peak_tflops_measured"] = ctx.fp64_peak_tflops()
    except Exception:
        pass
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cfg.cpu_baseline()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args, rank):
    """reference arm of a configuration: the CPU port of the reference algorithm (oracle/) on all host cores"""
    if rank != 0:
        return
    import picaso_b200 as pb

    class _NoCtx:  # the reference arm never touches the GPU library
        def __getattr__(self, k):
            raise RuntimeError("reference arm must not use the CUDA context")
    cls = CONFIGS[args.config]
    cfg = cls.__new__(cls)
    # build only the host-side inputs
    from picaso_b200 import synth
    n = os.cpu_count() or 1
    if args.config == "cfg2":
        cfg.sets = [synth.thermal_inputs(L=cls.L, W=cls.W, seed=1002)]
        units, fn, sample = cls.W, (lambda: cfg._oracle(cfg.sets[0], n)), "one full 90x10000x5 thermal spectrum per step"
    elif args.config == "cfg3":
        cfg.Wr = 4000
        cfg.sets = [synth.reflected_inputs(L=cls.L, W=cfg.Wr, seed=1003, ngauss=cls.G, stream=4)]
        idx = np.arange(cfg.Wr)
        units, fn, sample = cfg.Wr, (lambda: cfg._oracle(cfg.sets[0], idx, n)), "4000 of the 196000 wavelengths per step (SH4, OTHG forms)"
    elif args.config == "cfg4":
        db = synth.opacity_database(W=cls.W, nmol=cls.NMOL, seed=4001, nT=20, nP=18, nTc=30, wave_range=(0.3, 5.0))
        rng = np.random.default_rng(1)
        cfg.db, cfg.ray = db, {m: 10.0 ** rng.uniform(-27, -25, cls.W) for m in db["rayleigh_molecules"]}
        atm = synth.atmosphere_profile(db, L=cls.L, seed=4100, cloudy=True)
        atm["cia_pairs"] = {x + y: (x, y) for x, y in db["continuum_molecules"]}
        cfg.atms, cfg.NPROF = [atm], 1
        cfg.tr = synth.transit_inputs(L=cls.L, W=cls.W, seed=1004)
        units, fn, sample = cls.W, (lambda: cfg._cpu(0, n)), "one full 80x50000 transit spectrum incl. the opacity chain per step"
    else:
        from oracle import regrid as oreg
        nb = 16
        ds = [synth.thermal_inputs(L=cls.L, W=cls.W, seed=5000 + b) for b in range(nb)]
        hk = dict(wno=ds[0]["wno"], tlevel=np.array([d["tlevel"] for d in ds]), plevel=np.array([d["plevel"] for d in ds]),
                  ubar1=ds[0]["ubar1"], gweight=ds[0]["gweight"], tweight=ds[0]["tweight"])
        for k in ("dtau", "w0", "cosb"):
            hk[k] = np.array([d[k] for d in ds])
        newx = np.linspace(ds[0]["wno"][5], ds[0]["wno"][-5], cls.NBINS)
        units, fn, sample = nb * cls.W, (lambda: oreg.thermal_batch(**hk, newx=newx, scale=1e-8, nthreads=n)), "16 of the 1024 atmospheres per step"
    for _ in range(max(args.warmup, 1)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    val = units * args.steps / dt
    wl = getattr(cls, "workload", None) or "reflected_SH stream=4 L=60 W=196000 G=5 forms=111111 rayleigh=on delta-M (BASELINE config 3)"
    print(json.dumps({
        "impl": "reference", "metric": cls.metric, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": cls.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": wl, "bench_config": args.config},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": n, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
