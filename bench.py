#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native PICASO hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json metric): reflected-light Toon89 spectrum, 60 layers x 10 000
wavelengths x 5 Gauss angles, TTHG_ray single scattering, N=2 multiple scattering,
delta-Eddington on (so the *_og arrays are distinct), fused disk integration.
One step = one such spectrum per GPU (get_reflected_1d + compress_disco): a single kernel
launch on device-resident inputs.  Steps rotate over NSETS distinct input sets whose total
size exceeds the 126 MB L2, so every step streams its inputs from HBM.
Multi-GPU: weak scaling - each rank owns its own 10 000-wavelength slab (wavelengths are
independent); the per-rank albedo vectors are all-gathered inside the timed step over NVLink peer
memory (pb_peer_gather; PB_BENCH_GATHER selects the delivery mode or NCCL).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

L, W, NG = 60, 10000, 5
NSETS = 4
KW = dict(single_phase=3, multi_phase=0, toon_coefficients=0, get_lvl_flux=0)
LAYER_KEYS = ("dtau", "w0", "cosb", "gcos2", "ftau_cld", "ftau_ray", "dtau_og", "w0_og", "cosb_og")
LEVEL_KEYS = ("tau", "tau_og")
WAVE_KEYS = ("surf_reflect", "F0PI")
# algorithmic bytes per wave-point (SURVEY.md section 8d / BASELINE.md section 4):
# 9 layer arrays + 2 level arrays + F0PI + surf_reflect read once, G intensities written
ALG_BYTES_PER_WAVE = (9 * L + 2 * (L + 1) + 2) * 8 + NG * 8
METRIC = "wave-points/sec (60-layer x 10k-wave reflected Toon spectrum)"
UNIT = "wave-points/s"
WORKLOAD = "reflected_toon_1d L=60 W=10000 G=5 TTHG_ray N=2 delta-eddington (BASELINE headline)"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_for(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu captures (profiles/traffic.json), or None"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return None
    key = "".join(ch if ch.isalnum() else "_" for ch in kernel).strip("_") + "_dram_bytes_per_launch"
    return t.get(key)


def numba_reference():
    """The UNMODIFIED reference functions (picaso/fluxes.py, picaso/disco.py loaded by file path, oracle/ref_loader.py)
    when the reference tree and numba are on this box - they are in the build container, not on the GPU box - else
    None."""
    if os.environ.get("PB_BENCH_REFERENCE") == "port":
        return None
    try:
        from oracle import ref_loader
        if not ref_loader.available():
            return None
        import numba  # noqa: F401
        return ref_loader.load("fluxes"), ref_loader.load("disco")
    except Exception:
        return None


def pin_to_gpu_numa(index):
    """Bind this process (and so its first-touch / pinned allocations) to the CPUs of the NUMA node GPU `index`
    hangs off.  With 8 ranks each staging ~50 MB per step through pinned host memory, un-pinned ranks pull half of
    their traffic across the socket interconnect (round 1: e2e 1.29 ms per rank at N = 1, 2.39 ms at N = 8).
    Returns a short description for the JSON line, or None when the topology cannot be read."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        try:
            # NVML knows the CPUs closest to the GPU: binds the calling (main, allocating) thread to them
            before = len(os.sched_getaffinity(0))
            pynvml.nvmlDeviceSetCpuAffinity(h)
            after = len(os.sched_getaffinity(0))
            if 0 < after < before or after == before:
                return "nvmlDeviceSetCpuAffinity: %d of %d cpus" % (after, before)
        except Exception:
            pass
        bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:          # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bdf = bdf[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return "numa node %d (%d cpus)" % (node, len(cpus))
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.sm_max = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
        self.active = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            if self.active:
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, nm in names.items():
                        if mask & bit:
                            self.reasons.add(nm)
                except Exception:
                    pass
            time.sleep(0.01)  # 100 Hz: NVML queries take driver locks, keep them sparse

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------
# Spectrum-level end to end (extra key `e2e_spectrum`): what picaso() does per spectrum - opacityclass.
# get_opacities(atm) + compute_opacity(atm, opa) + get_reflected_1d + compress_disco (justdoit.py:236-310) -
# with the cross-section tables resident (HBM here, host memory in the CPU arm), a fresh atmosphere profile per
# step.  Only O(nlayer) scalars go to the device and the [nwno] albedo (+ xint) comes back.
# ---------------------------------------------------------------------------------------------------------
NMOL_SPEC, NPROF = 12, 4


def _spectrum_setup():
    import types
    from picaso_b200 import synth
    db = synth.opacity_database(W=W, nmol=NMOL_SPEC, seed=4001, nT=20, nP=18, ragged=True)
    # the synthetic tables are rescaled to a planet-like regime (Rayleigh tau ~ 10, total tau ~ 1e3-1e4 at 80 bar,
    # single-scattering albedo from ~1 at the top to ~0 at depth): values only, the arithmetic is the same
    for k in db["continuum"]:
        db["continuum"][k] = db["continuum"][k] * 1e-6
    for m in db["tables"]:
        db["tables"][m] = db["tables"][m] * 1e-1
    ray = {m: 6e-4 * (db["wno"] / 1e4) ** 4 * (1.0 + 0.3 * i) for i, m in enumerate(db["rayleigh_molecules"])}
    atms, ducks = [], []
    for i in range(NPROF):
        atm = synth.atmosphere_profile(db, L=L, seed=4100 + i, cloudy=False)
        a = types.SimpleNamespace()
        a.c = types.SimpleNamespace(nlayer=atm["nlayer"], pconv=atm["pconv"], rgas=atm["rgas"], amu=atm["amu"], k_b=atm["k_b"])
        a.level = {"temperature": atm["tlevel"], "pressure": atm["plevel"]}
        a.layer = {"temperature": atm["tlayer"], "pressure": atm["player"], "colden": atm["colden"], "mmw": atm["mmw"],
                   "mixingratios": atm["mixingratios"], "electrons": atm["electrons"], "cloud": None}
        a.planet = types.SimpleNamespace(gravity=atm["gravity"])
        a.molecules = list(db["molecules"])
        a.continuum_molecules = [list(x) for x in db["continuum_molecules"]]
        a.rayleigh_molecules = list(db["rayleigh_molecules"])
        atm["cia_pairs"] = {x + y: (x, y) for x, y in db["continuum_molecules"]}
        atms.append(atm)
        ducks.append(a)
    return db, ray, atms, ducks


def _spectrum_geometry():
    from picaso_b200 import synth
    gangle, gweight, tangle, tweight, ubar0, ubar1, cos_theta = synth.geometry_1d(NG, 0.0)
    tail = (0, ubar0, ubar1, cos_theta, np.ones(W), KW["single_phase"], KW["multi_phase"], 1.0, -1.0, 2.0, -0.5, 1.0)
    return gweight, tweight, cos_theta, tail


def spectrum_cpu(db, ray, atm, nthreads):
    """CPU port of the same chain: oracle/optics.py (numpy) + the C oracle (OpenMP)."""
    import oracle
    from oracle import optics as oo
    pbar = atm["player"] / atm["pconv"]
    ti, pi, ill, ihl, ilh, ihh = oo.find_needed_pts(db["temps"], db["pressures"], db["nc_p"], atm["tlayer"], pbar)
    mol = {m: oo.interp_molecular(db["tables"][m], ti, pi, ill, ihl, ilh, ihh) for m in db["molecules"]}
    ic = oo.nearest_cia_temp(db["cia_temps"], atm["tlayer"])
    cont = {k: db["continuum"][k][ic] for k in db["continuum"]}
    o = oo.compute_opacity(atm, mol, cont, ray, None, stream=2, delta_eddington=True)
    DTAU, TAU, W0, COSB, fcld, fray, GCOS2, DTAU_OG, TAU_OG, W0_OG, COSB_OG = o[:11]
    gweight, tweight, cos_theta, tail = _spectrum_geometry()
    x, _ = oracle.get_reflected_1d(L + 1, db["wno"], W, NG, 1, DTAU, TAU, W0, COSB, GCOS2, fcld, fray, DTAU_OG, TAU_OG,
                                   W0_OG, COSB_OG, *tail, nthreads=nthreads)
    return oracle.compress_disco(W, cos_theta, x, gweight, tweight, np.ones(W))


def spectrum_gpu_factory(pb, ctx, db, ray, ducks):
    opa = pb.DeviceOpacities(db["wno"], db["pt_pairs"], db["tables"], db["cia_temps"], db["continuum"], ray,
                             query_method="linear", ctx=ctx)
    gweight, tweight, cos_theta, tail = _spectrum_geometry()

    ubar0, ubar1 = tail[1], tail[2]

    def one(i):
        a = ducks[i % NPROF]
        opa.get_opacities(a)
        # one C call: opacity kernel -> flux kernel (fused disk integration) -> one D2H copy (pb_spectrum_reflected)
        return pb.reflected_spectrum(a, opa, ubar0, ubar1, cos_theta, gweight, tweight, single_phase=KW["single_phase"],
                                     multi_phase=KW["multi_phase"], toon_coefficients=KW["toon_coefficients"],
                                     stream=2, delta_eddington=True, raman=2)

    def chain(i):
        """the same spectrum through the three public calls (compute_opacity -> get_reflected_1d + fused albedo)"""
        a = ducks[i % NPROF]
        opa.get_opacities(a)
        dev = pb.compute_opacity(a, opa, ngauss=1, stream=2, delta_eddington=True, test_mode=None, raman=2,
                                 device_outputs=True)
        DTAU, TAU, W0, COSB, fcld, fray, GCOS2, DTAU_OG, TAU_OG, W0_OG, COSB_OG = dev[:11]
        sl = lambda d: d[:, :, 0]
        _, _, alb = pb.get_reflected_1d(L + 1, db["wno"], W, NG, 1, sl(DTAU), sl(TAU), sl(W0), sl(COSB), sl(GCOS2),
                                        sl(fcld), sl(fray), sl(DTAU_OG), sl(TAU_OG), sl(W0_OG), sl(COSB_OG), *tail,
                                        gweight=gweight, tweight=tweight, return_albedo=True, ctx=ctx)
        return alb
    one.chain = chain
    return opa, one


def time_spectrum_cpu(db, ray, atms, nthreads, budget_s=8.0):
    spectrum_cpu(db, ray, atms[0], nthreads)
    t0 = time.perf_counter()
    n = 0
    while True:
        spectrum_cpu(db, ray, atms[n % NPROF], nthreads)
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= 40:
            break
    return W * n / el, n, el


def make_sets(rank):
    from picaso_b200 import synth
    return [synth.reflected_inputs(L=L, W=W, seed=1000 + 97 * rank + i) for i in range(NSETS)]


def run_reference(args, rank, world):
    """The reference arm.  The reference is single-threaded numba (no parallel=True / prange / nogil anywhere,
    SURVEY.md section 1), so the line's `value` is a 1-core figure: the reference's own get_reflected_1d +
    compress_disco (JIT excluded) where /root/reference and numba exist (`kind: reference`); on the GPU box the
    reference tree cannot travel, so the C port of the same algorithm (oracle/) on ONE thread stands in
    (`kind: port`), with its all-threads OpenMP figure beside it (`all_threads_value`)."""
    if rank != 0:
        return
    import cases as C
    import oracle
    nthreads = os.cpu_count() or 1
    d = make_sets(0)[0]
    a = C.reflected_args(d, KW)

    def port(n):
        x, _ = oracle.get_reflected_1d(*a, nthreads=n)
        return oracle.compress_disco(W, d["cos_theta"], x, d["gweight"], d["tweight"], d["F0PI"])

    ref = numba_reference()
    if ref is not None:
        rf, rd = ref

        def one():
            x, _ = rf.get_reflected_1d(*a)
            return rd.compress_disco(W, d["cos_theta"], x, d["gweight"], d["tweight"], d["F0PI"])
        kind, cores = "reference", 1
        sample = "%d full spectra (60x10000x5) per run: the unmodified picaso.fluxes.get_reflected_1d + disco.compress_disco (numba, JIT excluded), 1 core of %d" % (args.steps, nthreads)
    else:
        one = lambda: port(1)
        kind, cores = "port", 1
        sample = ("%d full spectra (60x10000x5) per run: C port of the reference algorithm (oracle/) on 1 thread - the reference "
                  "itself is single-threaded numba and /root/reference does not exist on this box; all-threads OpenMP figure in "
                  "all_threads_value" % args.steps)
    for _ in range(max(min(args.warmup, 2), 1)):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    val = W * args.steps / dt
    port(nthreads)
    t0 = time.perf_counter()
    n = 0
    while n < 50 and (n < 3 or time.perf_counter() - t0 < 5.0):
        port(nthreads)
        n += 1
    allv = W * n / (time.perf_counter() - t0)
    db, ray, atms, _ = _spectrum_setup()
    sv, sn, sel = time_spectrum_cpu(db, ray, atms, nthreads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "spectra_per_sec": args.steps / dt,
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "all_threads_value": allv, "all_threads_cores": nthreads, "all_threads_kind": "port (OpenMP over wavelengths)",
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "e2e_spectrum": {"value": sv, "unit": UNIT, "steps": sn,
                         "api": "numpy port of get_opacities(linear) + compute_opacity (%d molecules, clear, no Raman) + C port of "
                                "get_reflected_1d + compress_disco, %d threads" % (NMOL_SPEC, nthreads)},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 40)")
    ap.add_argument("--config", default="headline", choices=["headline", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json configuration; headline = 60 x 10 000 x 5 reflected Toon (the metric's own config)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.config != "headline":
        import bench_configs
        return bench_configs.run(args, rank, local_rank, world)
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import ctypes
    import cases as C
    import picaso_b200 as pb
    from picaso_b200 import _lib
    from picaso_b200._lib import PB_DEVICE, ReflectedArgs

    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    numa = pin_to_gpu_numa(local_rank) if world > 1 else None
    ctx = pb.Context(local_rank)
    if world > 1:
        # run our kernels on the stream NCCL's collectives are enqueued on: no host sync needed
        # (a dedicated non-default stream: handle 0 would mean "the context's own stream")
        side = torch.cuda.Stream()
        torch.cuda.set_stream(side)
        ctx.set_stream(side.cuda_stream)
    warm = max(args.warmup, 3)

    sets = make_sets(rank)
    G = NG
    dev_sets = []
    for d in sets:
        dd = {k: ctx.to_device(d[k]) for k in LAYER_KEYS + LEVEL_KEYS + WAVE_KEYS}
        dev_sets.append(dd)
    d0 = sets[0]
    u0 = np.ascontiguousarray(d0["ubar0"]).reshape(-1)
    u1 = np.ascontiguousarray(d0["ubar1"]).reshape(-1)
    gw = np.ascontiguousarray(d0["gweight"])
    tw = np.ascontiguousarray(d0["tweight"])
    d_xint = ctx.dev_alloc(G * W * 8)
    # world == 1 with PB_BENCH_GATHER=p2p: the rank gathers to itself (diagnostic: in-kernel cost of the epilogue)
    # p2p: own all-gather over peer memory, pushed by a side-stream copy kernel (default); p2p_fused: stores and
    # flags in the solver kernel's epilogue; nccl: ncclAllGather on a second stream
    gather_mode = os.environ.get("PB_BENCH_GATHER", "p2p" if world > 1 else "none")
    if world == 1 and gather_mode == "nccl":
        gather_mode = "none"
    # p2p_deferred (default): a courier CTA of the next launch pushes the slab and publishes it (measured at N = 2,
    # us per step over 200 steps: no exchange 61.9, deferred 62.5, lazy 65.0, push 67.0, fused 69.8 - profiles/);
    # p2p_lazy: stores in the solver's epilogue, flags published by the next launch; p2p: side-stream push kernel;
    # p2p_fused: stores + fence + flags in the solver's epilogue
    if world > 1 and "PB_BENCH_GATHER" not in os.environ:
        gather_mode = "p2p_deferred"
    push = {"p2p": True, "p2p_fused": False, "p2p_lazy": "lazy", "p2p_deferred": "deferred"}.get(gather_mode, True)
    if gather_mode in ("p2p_fused", "p2p_lazy", "p2p_deferred"):
        gather_mode = "p2p"
    NBUF = int(os.environ.get("PB_BENCH_NBUF", "3"))  # rotating gathered buffers: ranks may run NBUF - 2 steps apart
    if gather_mode == "nccl":
        # NCCL all-gather, double-buffered so that the gather of step i (stream `comm`) overlaps the kernel
        # of step i+1 (stream `side`)
        comm = torch.cuda.Stream()
        alb_all = [torch.empty((world, W), dtype=torch.float64, device="cuda") for _ in range(2)]
        alb_mine = [torch.empty((W,), dtype=torch.float64, device="cuda") for _ in range(2)]
        ev_kernel = [torch.cuda.Event() for _ in range(2)]
        ev_gather = [torch.cuda.Event() for _ in range(2)]
        for e in ev_gather:
            e.record(side)
        d_albs = [t.data_ptr() for t in alb_mine]
    elif gather_mode == "p2p":
        # all-gather fused into the kernel epilogue over peer memory (include/picaso_b200.h: pb_peer_gather):
        # every rank maps every rank's gathered buffers [NBUF][world][W] and arrival flags [world]
        from picaso_b200.sharded import PeerAllGather
        def exchange(obj):   # bench plumbing: the IPC handles travel over the process group that exists anyway
            parts = [None] * world
            dist.all_gather_object(parts, obj)
            return parts
        pag = PeerAllGather(ctx, rank, world, W, nbuf=NBUF, push=push, exchange=exchange if world > 1 else None)
        d_albs = [ctx.dev_alloc(W * 8)]
        if world > 1:
            dist.barrier()
    else:
        d_albs = [ctx.dev_alloc(W * 8)]
    d_alb = d_albs[0]

    def make_args(dd, d_alb):
        a = ReflectedArgs()
        a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = L, W, NG, 1, 1, W
        for k in LAYER_KEYS + LEVEL_KEYS + WAVE_KEYS:
            setattr(a, k, dd[k])
        a.b_top = None
        a.ubar0, a.ubar1, a.gweight, a.tweight = _lib.addr(u0), _lib.addr(u1), _lib.addr(gw), _lib.addr(tw)
        a.cos_theta = d0["cos_theta"]
        a.single_phase, a.multi_phase, a.toon_coefficients = KW["single_phase"], KW["multi_phase"], KW["toon_coefficients"]
        a.frac_a, a.frac_b, a.frac_c = d0["frac_a"], d0["frac_b"], d0["frac_c"]
        a.constant_back, a.constant_forward = d0["constant_back"], d0["constant_forward"]
        a.get_toa_intensity, a.get_lvl_flux = 1, 0
        a.xint_at_top, a.albedo = d_xint, d_alb
        return a

    cargs = [[make_args(dd, da) for da in d_albs] for dd in dev_sets]
    fn = ctx.lib.pb_reflected_toon_1d

    def step(i):
        if gather_mode == "none":
            ctx.check(fn(ctx.h, ctypes.byref(cargs[i % NSETS][0]), PB_DEVICE))
            return
        if gather_mode == "p2p":
            a = cargs[i % NSETS][0]
            a.gather = pag.next()   # step counter, rotating buffer, wait_step = step - (NBUF - 1)
            ctx.check(fn(ctx.h, ctypes.byref(a), PB_DEVICE))
            return
        bf = i & 1
        side.wait_event(ev_gather[bf])  # the gather of step i-2 has consumed alb_mine[bf]
        ctx.check(fn(ctx.h, ctypes.byref(cargs[i % NSETS][bf]), PB_DEVICE))
        ev_kernel[bf].record(side)
        comm.wait_event(ev_kernel[bf])
        with torch.cuda.stream(comm):
            dist.all_gather_into_tensor(alb_all[bf], alb_mine[bf])
            ev_gather[bf].record(comm)

    def drain():
        # the launching stream waits for the outstanding gathers: the stop event covers them
        if gather_mode == "nccl":
            side.wait_event(ev_gather[0])
            side.wait_event(ev_gather[1])
        elif gather_mode == "p2p":
            pag.wait()

    def barrier():
        drain()
        ctx.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- parity gate on this very configuration (cheap: one set, oracle on host threads) ----
    import oracle
    step(0)
    ctx.sync()
    drain()
    ctx.sync()
    if gather_mode == "nccl":
        got = alb_all[0][rank].cpu().numpy()
    elif gather_mode == "p2p":
        # the fused gather against NCCL's: every rank's buffer must hold every rank's slab, bit for bit
        gath = pag.gathered()
        mine = gath[rank].copy() if push is True else ctx.from_device(d_alb, (W,))
        if world > 1:
            ref_all = torch.empty((world, W), dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(ref_all, torch.from_numpy(mine).cuda())
            torch.cuda.synchronize()
            ref_np = ref_all.cpu().numpy()
        else:
            ref_np = mine[None, :]
        if not np.array_equal(gath, ref_np):
            raise SystemExit("rank %d: fused peer-memory all-gather differs from the NCCL all-gather" % rank)
        got = gath[rank]
    else:
        got = ctx.from_device(d_alb, (W,))
    ox, _ = oracle.get_reflected_1d(*C.reflected_args(sets[0], KW), nthreads=os.cpu_count() or 1)
    want = oracle.compress_disco(W, d0["cos_theta"], ox, gw, tw, sets[0]["F0PI"])
    parity = float(np.max(np.abs(got - want) / np.abs(want)))
    if not parity < 1e-6:
        raise SystemExit("parity gate failed: albedo max rel err %.3e" % parity)

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- warm-up: W steps, then keep going until clocks had ~0.3 s to ramp ----
    for i in range(warm):
        step(i)
    ctx.sync()
    sampler.active = True  # the ramp runs the same kernel back to back: samples are under load
    # fixed step count (not wall-clock) so that every rank issues the same collectives
    i = warm
    for _ in range(150):
        for _ in range(20):
            step(i)
            i += 1
        ctx.sync()
    # ---- diagnostic (multi-GPU): every rank alone, same kernel, no exchange - the slowest GPU bounds the
    # coupled rate, so weak-scaling loss can be attributed to GPU-to-GPU variation vs the collective ----
    uncoupled = None
    if world > 1:
        barrier()
        for a_ in (c[0] for c in cargs):
            a_.gather = None
        ctx.timer_start()
        for i in range(100):
            ctx.check(fn(ctx.h, ctypes.byref(cargs[i % NSETS][0]), PB_DEVICE))
        own = ctx.timer_stop() / 100 * 1e3
        t = torch.tensor([own], dtype=torch.float64, device="cuda")
        allt = torch.empty((world,), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allt, t)
        uncoupled = [round(float(x), 2) for x in allt.cpu()]
    # ---- timed region: exactly K steps, CUDA events on the launching stream ----
    barrier()
    if gather_mode == "p2p" and world > 1:
        # device-side start barrier over the peer mappings: the ranks' GPUs leave it within a microsecond of each
        # other, so host-side skew after the NCCL barrier is not billed to the first coupled steps
        pag.barrier()
    l0 = ctx.launch_count()
    sampler.active = True
    ctx.timer_start()
    for i in range(args.steps):
        step(i)
    drain()
    ms = ctx.timer_stop()
    sampler.active = False
    launches = ctx.launch_count() - l0
    courier_us = None
    if gather_mode == "p2p":
        if pag.timed_out():
            raise SystemExit("rank %d: peer all-gather timed out waiting for a flag" % rank)
        if push == "deferred":
            courier_us = pag.courier_us()
    barrier()
    if world > 1:
        if courier_us is not None:
            t = torch.tensor([courier_us], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            courier_us = round(float(t.item()), 2)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * W * args.steps / (ms * 1e-3)

    if os.environ.get("PB_BENCH_SWEEP") and gather_mode == "p2p":
        # diagnostic (stderr, not part of the bench line): where the coupled step's extra time goes - every delivery mode
        # of pb_peer_gather and no exchange at all, over short and long timed regions (fixed vs per-step cost)
        sweep = []
        for mode, pushv in (("none", None), ("deferred", 3), ("lazy", 2), ("push", 1), ("fused", 0), ("deferred", 3)):
            for k in (20, 200):
                if pushv is not None:
                    pag.push = pushv
                barrier()
                if world > 1:
                    pag.barrier()
                ctx.timer_start()
                th0 = time.perf_counter()
                for i in range(k):
                    a_ = cargs[i % NSETS][0]
                    a_.gather = pag.next() if pushv is not None else None
                    ctx.check(fn(ctx.h, ctypes.byref(a_), PB_DEVICE))
                host_us = (time.perf_counter() - th0) / k * 1e6
                if pushv is not None:
                    pag.wait()
                t_ms = ctx.timer_stop()
                if world > 1:
                    t = torch.tensor([t_ms, host_us], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    t_ms, host_us = float(t[0]), float(t[1])
                sweep.append({"mode": mode, "steps": k, "us_per_step": round(1e3 * t_ms / k, 2),
                              "host_enqueue_us_per_step": round(host_us, 2)})
        pag.push = 2 if push == "lazy" else 3 if push == "deferred" else int(bool(push))
        barrier()
        if rank == 0:
            sys.stderr.write("PB_BENCH_SWEEP " + json.dumps({"n_gpus": world, "sweep": sweep}) + "\n")

    # ---- end to end: public API, pinned host inputs, H2D + kernel + D2H per step ----
    ke = args.e2e_steps or min(args.steps, 40)
    pinned_sets = []
    for d in sets:
        pd = dict(d)
        for k in LAYER_KEYS + LEVEL_KEYS + WAVE_KEYS:
            buf = ctx.pinned_empty(d[k].shape)
            buf[...] = d[k]
            pd[k] = buf
        pinned_sets.append(pd)
    if world > 1:
        ctx.set_stream(None)

    def e2e_step(i):
        d = pinned_sets[i % NSETS]
        return pb.get_reflected_1d(*C.reflected_args(d, KW), ctx=ctx, gweight=gw, tweight=tw,
                                   return_albedo=True)

    for i in range(3):
        xint_h, _, alb_h = e2e_step(i)
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    per = []
    for i in range(ke):
        t1 = time.perf_counter()
        xint_h, _, alb_h = e2e_step(i)
        per.append(time.perf_counter() - t1)
    e2e_dt = time.perf_counter() - t0
    if os.environ.get("PB_BENCH_DEBUG"):
        sys.stderr.write("e2e per-step ms: " + " ".join("%.2f" % (1e3 * x) for x in per) + "\n")
    sampler.active = False
    sampler.stop_flag = True
    if world > 1:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_val = world * W * ke / e2e_dt
    h2d = sum(sets[0][k].nbytes for k in LAYER_KEYS + LEVEL_KEYS + WAVE_KEYS) + W * 8  # + b_top
    d2h = (G + 1) * W * 8

    peak, peak_src = peaks()
    alg_bytes = ALG_BYTES_PER_WAVE * W
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    kernel = "refl_toa_kernel5"
    traffic = traffic_for(kernel)
    fp64_pct = None
    for fpath in ("r2_refl_toa_v5.summary.json", "r1_refl_toa_v4gen.summary.json"):
        try:
            fp64_pct = json.load(open(os.path.join(ROOT, "profiles", fpath)))["launches"][0][
                "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]["value"]
            break
        except Exception:
            continue
    # second roof: the fp64 pipe (there is no fp64 figure in MEASURED_PEAKS.json; measured here with a DFMA loop)
    try:
        fp64_peak = ctx.fp64_peak_tflops()
    except Exception:
        fp64_peak = None
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "spectra_per_sec": world * args.steps / (ms * 1e-3),
        "config": {"workload": WORKLOAD, "waves_per_gpu": W, "layers": L, "angles": G,
                   "l2_policy": "inputs larger than L2: %d input sets x %.1f MB rotated" % (NSETS, alg_bytes / 1e6),
                   "collective": {"none": "none",
                                  "p2p": "all-gather of the per-rank albedo [W] every step over NVLink peer memory "
                                         "(pb_peer_gather): " + ("the solver writes its slab into the local gathered buffer, a "
                                         "side-stream copy kernel pushes it to every peer and publishes release flags while the "
                                         "next step computes" if push is True else ("P2P stores in the solver kernel's epilogue, the flags "
                                         "of step s published by the first CTA of launch s + 1 (no fence on any launch's critical path)"
                                         if push == "lazy" else ("the solver CTAs fill the local row only; one extra (courier) CTA of the launch "
                                         "of step s + 1 pushes the slab of step s to every peer over NVLink and publishes its flags "
                                         "while the solver CTAs compute" if push == "deferred" else "P2P stores + release flags in the "
                                         "solver kernel's epilogue"))) + "; device-side start barrier; %d rotating buffers; checked bit-for-bit against ncclAllGather before "
                                         "timing; the timed region ends after pb_gather_wait saw the last step of every rank" % NBUF,
                                  "nccl": "ncclAllGather of the per-rank albedo [W] every step, double-buffered on a second "
                                          "stream (PB_BENCH_GATHER=nccl)"}[gather_mode],
                   "parity_albedo_max_rel_err": parity,
                   "kernel_us_per_rank_without_exchange": uncoupled,
                   "courier_cta_us_max_over_ranks": courier_us},
        "gpu_launches": int(launches),
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": ke, "ms_per_step": 1e3 * e2e_dt / ke,
                "api": "picaso_b200.get_reflected_1d(..., return_albedo=True), pinned host inputs",
                "host_affinity": numa},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "kernel": kernel,
                     "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                     "fp64_pipe_pct_ncu": fp64_pct, "fp64_peak_tflops_measured": fp64_peak,
                     "fp64_roof_note": "second roof: DFMA loop measured in this run (pb_microbench); the kernel issues ~14 M fp64 "
                                       "warp-instructions per launch = 24 us at that rate, against 8.2 us of HBM time",
                     "note": "not HBM-bound: ~45 fp64 flop/B; 435 CTAs of 5 warps = 15 warps per SM (the problem has only "
                             "10.6 consumer warps per SM), dependent fp64 chains (ncu stall_wait 2.4 of 7.3 cycles per issue); "
                             "DRAM traffic = algorithmic bytes; see DESIGN.md 4.1, profiles/r2_refl_toa_v5.summary.json and "
                             "profiles/r2_refl_toa_v5_sass.txt"},
        "clocks": sampler.summary(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nthreads = os.cpu_count() or 1
        a = C.reflected_args(sets[0], KW)
        oracle.get_reflected_1d(*a, nthreads=nthreads)
        t0 = time.perf_counter()
        n = 0
        while True:
            x, _ = oracle.get_reflected_1d(*a, nthreads=nthreads)
            oracle.compress_disco(W, d0["cos_theta"], x, gw, tw, sets[0]["F0PI"])
            n += 1
            el = time.perf_counter() - t0
            if el > 10.0 or n >= 200:
                break
        t1 = time.perf_counter()
        oracle.get_reflected_1d(*a, nthreads=1)
        one = time.perf_counter() - t1
        out["cpu_baseline"] = {"value": W / one, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": "one full 60x10000x5 spectrum on 1 thread (the reference is single-threaded numba); C port of the "
                                         "reference algorithm (oracle/); all-threads OpenMP figure: %d spectra in %.1f s" % (n, el),
                               "all_threads_value": W * n / el, "all_threads_cores": nthreads, "host_cpus": os.cpu_count()}
    if rank == 0 and world == 1:
        # ---- spectrum-level end to end: profile in, albedo out, opacity tables resident in HBM ----
        # (an extra key: a failure here is reported IN the line instead of taking the headline numbers down with it)
        try:
            db, ray, atms, ducks = _spectrum_setup()
            opa, one = spectrum_gpu_factory(pb, ctx, db, ray, ducks)
            got = one(0)
            want = spectrum_cpu(db, ray, atms[0], os.cpu_count() or 1)
            sp_par = float(np.max(np.abs(got - want) / np.abs(want)))
            if not sp_par < 1e-6:
                raise SystemExit("spectrum-level parity gate failed: %.3e" % sp_par)
            if not np.array_equal(got, one.chain(0)):
                raise SystemExit("pb_spectrum_reflected differs from the compute_opacity -> get_reflected_1d chain")
            for i in range(3):
                one(i)
            ns = max(ke, 40)
            l0 = ctx.launch_count()
            t0 = time.perf_counter()
            for i in range(ns):
                one(i)
            sdt = time.perf_counter() - t0
            sp = {"value": W * ns / sdt, "unit": UNIT, "steps": ns, "ms_per_step": 1e3 * sdt / ns,
                  "gpu_launches_per_step": (ctx.launch_count() - l0) / ns,
                  "h2d_bytes_per_step": int(L * (NMOL_SPEC + len(db["continuum"]) + len(ray) + 12) * 8),
                  "d2h_bytes_per_step": int((NG + 1) * W * 8), "parity_albedo_max_rel_err": sp_par,
                  "api": "DeviceOpacities.get_opacities(linear) + picaso_b200.reflected_spectrum (pb_spectrum_reflected: "
                         "compute_opacity, %d molecules, clear, no Raman -> get_reflected_1d -> compress_disco in one C call, "
                         "one D2H copy); tables resident in HBM" % NMOL_SPEC}
            for i in range(3):
                one.chain(i)
            t0 = time.perf_counter()
            for i in range(ns):
                one.chain(i)
            sp["three_call_chain_ms_per_step"] = 1e3 * (time.perf_counter() - t0) / ns
            if not args.no_cpu_baseline:
                cv, cn, cel = time_spectrum_cpu(db, ray, atms, os.cpu_count() or 1)
                sp["cpu_port_value"] = cv
                sp["cpu_port_sample"] = "%d spectra in %.1f s on %d threads" % (cn, cel, os.cpu_count() or 1)
            out["e2e_spectrum"] = sp
            opa.close()
        except (Exception, SystemExit) as exc:
            sys.stderr.write("e2e_spectrum failed: %r\n" % (exc,))
            out["e2e_spectrum"] = {"error": str(exc)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
