"""Where does the end-to-end (host buffers) time go?  Run on the GPU box."""
import sys, os, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import cases as C, picaso_b200 as pb
from picaso_b200 import synth, _lib
KW = dict(single_phase=3, multi_phase=0, toon_coefficients=0, get_lvl_flux=0)
ctx = pb.Context(0)
d = synth.reflected_inputs(L=60, W=10000, seed=1)
keys = ("dtau","w0","cosb","gcos2","ftau_cld","ftau_ray","dtau_og","w0_og","cosb_og","tau","tau_og","surf_reflect","F0PI")
pd = dict(d)
for k in keys:
    b = ctx.pinned_empty(d[k].shape); b[...] = d[k]; pd[k] = b
def t(f, n=20):
    f(); t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
print("pageable inputs  : %.3f ms" % t(lambda: pb.get_reflected_1d(*C.reflected_args(d, KW), ctx=ctx)))
print("pinned inputs    : %.3f ms" % t(lambda: pb.get_reflected_1d(*C.reflected_args(pd, KW), ctx=ctx)))
# raw H2D copies of the same buffers
dev = {k: ctx.dev_alloc(d[k].nbytes) for k in keys}
def h2d(src):
    for k in keys:
        ctx.check(ctx.lib.pb_memcpy_h2d(ctx.h, dev[k], src[k].ctypes.data, src[k].nbytes))
    ctx.sync()
nb = sum(d[k].nbytes for k in keys)
tp = t(lambda: h2d(pd)); print("raw pinned H2D   : %.3f ms  (%.1f GB/s)" % (tp, nb / tp / 1e6))
tq = t(lambda: h2d(d)); print("raw pageable H2D : %.3f ms  (%.1f GB/s)" % (tq, nb / tq / 1e6))
big = ctx.pinned_empty((nb // 8,)); dbig = ctx.dev_alloc(nb)
def one():
    ctx.check(ctx.lib.pb_memcpy_h2d(ctx.h, dbig, big.ctypes.data, nb)); ctx.sync()
tb = t(one); print("single 53MB pinned H2D: %.3f ms (%.1f GB/s)" % (tb, nb / tb / 1e6))
print("np.zeros x4 level arrays: %.3f ms" % t(lambda: [np.zeros((5, 1, 61, 10000)) for _ in range(4)]))
# ---- rotation over 4 pinned sets, as bench.py does ----
sets = [synth.reflected_inputs(L=60, W=10000, seed=10 + i) for i in range(4)]
psets = []
for s in sets:
    q = dict(s)
    for k in keys:
        b = ctx.pinned_empty(s[k].shape); b[...] = s[k]; q[k] = b
    psets.append(q)
cnt = [0]
def rot():
    q = psets[cnt[0] % 4]; cnt[0] += 1
    return pb.get_reflected_1d(*C.reflected_args(q, KW), ctx=ctx, gweight=d["gweight"], tweight=d["tweight"], return_albedo=True)
print("rotating 4 pinned sets + albedo: %.3f ms" % t(rot, 40))
import oracle
oracle.get_reflected_1d(*C.reflected_args(sets[0], KW), nthreads=os.cpu_count())
print("after an OpenMP oracle call    : %.3f ms" % t(rot, 40))
time.sleep(1.0)
print("1 s later                      : %.3f ms" % t(rot, 40))
# ---- does an NVML polling thread slow the CUDA calls down? ----
import threading, pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
stop = [False]; durs = []
def poll(interval):
    while not stop[0]:
        t0 = time.perf_counter()
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        durs.append(time.perf_counter() - t0)
        time.sleep(interval)
for interval in (0.0005, 0.01, 0.05):
    stop[0] = False; durs.clear()
    th = threading.Thread(target=poll, args=(interval,), daemon=True); th.start()
    r = t(rot, 40)
    stop[0] = True; th.join()
    print("with NVML poll every %.1f ms: %.3f ms/step; nvml query pair takes %.3f ms median (%d polls)" % (interval * 1e3, r, 1e3 * np.median(durs), len(durs)))
