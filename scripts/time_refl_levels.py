"""Device time of the reflected level-flux launch at the climate shape (8 gauss points x 90 layers x 661 waves,
one mu = 0.5 stream) for the three implementations (PB_REFL_LEVELS = fused | rec | scan).  GPU box."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import picaso_b200 as pb
from picaso_b200 import _lib, synth
from picaso_b200._lib import PB_DEVICE, ReflectedArgs

ctx = pb.Context(0)
B, L, W = 8, 90, 661
ds = [synth.reflected_inputs(L=L, W=W, seed=300 + b) for b in range(B)]
a = ReflectedArgs()
a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = L, W, 1, 1, B, W
for k in ("dtau", "w0", "cosb", "gcos2", "ftau_cld", "ftau_ray", "dtau_og", "w0_og", "cosb_og", "tau", "tau_og"):
    setattr(a, k, ctx.to_device(np.stack([d[k] for d in ds])))
a.surf_reflect = ctx.to_device(np.zeros((B, W)))
a.F0PI = ctx.to_device(np.ones((B, W)))
half, one = np.array([0.5]), np.array([1.0])
a.ubar0, a.ubar1, a.gweight, a.tweight = _lib.addr(half), _lib.addr(half), _lib.addr(one), _lib.addr(one)
a.cos_theta = 1.0
a.single_phase, a.multi_phase, a.toon_coefficients = 3, 0, 0
a.frac_a, a.frac_b, a.frac_c, a.constant_back, a.constant_forward = 1.0, -1.0, 2.0, -0.5, 1.0
a.get_toa_intensity, a.get_lvl_flux = 0, 1
outs = [ctx.dev_alloc(B * (L + 1) * W * 8) for _ in range(4)]
a.flux_minus, a.flux_plus, a.flux_minus_mdpt, a.flux_plus_mdpt = outs
fn = ctx.lib.pb_reflected_toon_1d
res = {}
for mode in ("fused", "rec", "scan"):
    os.environ["PB_REFL_LEVELS"] = mode
    for _ in range(3):
        ctx.check(fn(ctx.h, ctypes.byref(a), PB_DEVICE))
    ctx.sync()
    ctx.timer_start()
    for _ in range(50):
        ctx.check(fn(ctx.h, ctypes.byref(a), PB_DEVICE))
    ms = ctx.timer_stop() / 50
    res[mode] = [ctx.from_device(o, (B, L + 1, W)) for o in outs]
    print("%-6s %8.1f us per launch" % (mode, 1e3 * ms))
for mode in ("rec", "scan"):
    worst = 0.0
    for x, y in zip(res[mode], res["fused"]):
        colmax = np.max(np.abs(y), axis=1, keepdims=True)
        worst = max(worst, float(np.max(np.abs(x - y) / (1e-300 + colmax))))
    print("%-6s max |diff| / column max vs fused: %.2e" % (mode, worst))
