"""The all-gather over peer memory (pb_peer_gather, sharded.PeerAllGather) with the ranks simulated inside
one process on one GPU (one Context = one stream per rank; plain device pointers instead of IPC mappings).
All delivery modes: in the solver kernel's epilogue (flags published at its end, or lazily by the next launch),
pushed by the side-stream copy kernel, and deferred to a courier CTA of the next launch.  The
multi-process / multi-GPU path (CUDA IPC over NVLink) is exercised by bench.py, which checks the gathered
buffers bit-for-bit against ncclAllGather before timing."""
import ctypes

import numpy as np
import pytest

import cases as C
import picaso_b200 as pb
from picaso_b200 import _lib, sharded, synth
from picaso_b200._lib import PB_DEVICE, ReflectedArgs

pytestmark = pytest.mark.gpu
KW = dict(single_phase=3, multi_phase=0, toon_coefficients=0, get_lvl_flux=0)
LAYER = ("dtau", "w0", "cosb", "gcos2", "ftau_cld", "ftau_ray", "dtau_og", "w0_og", "cosb_og")
LEVEL = ("tau", "tau_og")
WAVE = ("surf_reflect", "F0PI")


def _args(ctx, d, keep):
    a = ReflectedArgs()
    W = d["nwno"]
    a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = d["nlevel"] - 1, W, d["numg"], d["numt"], 1, W
    for k in LAYER + LEVEL + WAVE:
        setattr(a, k, ctx.to_device(np.broadcast_to(np.asarray(d[k], dtype=np.float64), (W,)) if k in WAVE else d[k]))
    vecs = [np.ascontiguousarray(d[k], dtype=np.float64).reshape(-1) for k in ("ubar0", "ubar1", "gweight", "tweight")]
    keep.append(vecs)
    a.ubar0, a.ubar1, a.gweight, a.tweight = [_lib.addr(v) for v in vecs]
    a.cos_theta = d["cos_theta"]
    a.single_phase, a.multi_phase, a.toon_coefficients = KW["single_phase"], KW["multi_phase"], KW["toon_coefficients"]
    a.frac_a, a.frac_b, a.frac_c = d["frac_a"], d["frac_b"], d["frac_c"]
    a.constant_back, a.constant_forward = d["constant_back"], d["constant_forward"]
    a.get_toa_intensity, a.get_lvl_flux = 1, 0
    a.albedo = ctx.dev_alloc(W * 8)
    return a


@pytest.mark.parametrize("push", [False, True, "lazy", "deferred"])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_peer_all_gather_in_process(world, push):
    W, nsteps = 333, 7
    ctxs = [pb.Context(0) for _ in range(world)]
    group = sharded.PeerAllGather.local_group(ctxs, W, nbuf=3, push=push)
    keep = []
    # every rank owns a different slab per step (different seeds); the expected rows come from plain calls
    data = [[synth.reflected_inputs(L=9, W=W, seed=100 * r + s) for s in range(nsteps)] for r in range(world)]
    want = [[pb.get_reflected_1d(*C.reflected_args(d, KW), gweight=d["gweight"], tweight=d["tweight"],
                                 return_albedo=True, ctx=ctxs[0])[2] for d in row] for row in data]
    args = [[_args(ctxs[r], d, keep) for d in data[r]] for r in range(world)]
    fn = ctxs[0].lib.pb_reflected_toon_1d
    for s in range(nsteps):
        for r in range(world):      # interleaved enqueue: no rank ever waits on the host
            a = args[r][s]
            a.gather = group[r].next()
            ctxs[r].check(fn(ctxs[r].h, ctypes.byref(a), PB_DEVICE))
        if s in (2, nsteps - 1):
            for r in range(world):      # every rank enqueues its wait (lazy mode: publishes its last step) ...
                group[r].wait()
            for r in range(world):      # ... before any rank blocks the host (the ranks share this process)
                ctxs[r].sync()
                got = group[r].gathered()
                for q in range(world):
                    assert np.array_equal(got[q], want[q][s]), (world, push, s, r, q)
    # device-side barrier over the same peer mappings: every rank signals, every rank waits (stream-ordered)
    for _ in range(2):
        for r in range(world):
            group[r].barrier()
    for r in range(world):
        ctxs[r].sync()
    for g in group:
        assert not g.timed_out()
        g.close()
    for c in ctxs:
        c.close()


def test_gather_argument_checks():
    ctx = pb.default_context()
    d = synth.reflected_inputs(L=5, W=64, seed=3)
    keep = []
    a = _args(ctx, d, keep)
    g = sharded.PeerAllGather(ctx, 0, 1, 64)
    a.gather = g.next()
    a.albedo = None                      # the gather needs the fused albedo
    assert ctx.lib.pb_reflected_toon_1d(ctx.h, ctypes.byref(a), PB_DEVICE) != 0
    g.close()
