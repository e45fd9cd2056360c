"""cProfile + wall clock split of the strict end-to-end call (pinned host arrays -> get_reflected_1d(return_albedo)); GPU box."""
import cProfile, pstats, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import cases as C, picaso_b200 as pb
from picaso_b200 import synth
KW = dict(single_phase=3, multi_phase=0, toon_coefficients=0, get_lvl_flux=0)
ctx = pb.Context(0)
keys = ("dtau","w0","cosb","gcos2","ftau_cld","ftau_ray","dtau_og","w0_og","cosb_og","tau","tau_og","surf_reflect","F0PI")
sets = []
for i in range(4):
    d = synth.reflected_inputs(L=60, W=10000, seed=10 + i)
    q = dict(d)
    for k in keys:
        b = ctx.pinned_empty(d[k].shape); b[...] = d[k]; q[k] = b
    sets.append(q)
gw, tw = sets[0]["gweight"], sets[0]["tweight"]
def one(i):
    return pb.get_reflected_1d(*C.reflected_args(sets[i % 4], KW), ctx=ctx, gweight=gw, tweight=tw, return_albedo=True)
for i in range(5): one(i)
t0 = time.perf_counter()
for i in range(100): one(i)
print("wall per call: %.3f ms" % ((time.perf_counter() - t0) * 10))
pr = cProfile.Profile(); pr.enable()
for i in range(100): one(i)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
