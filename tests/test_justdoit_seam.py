"""The drop-in seam under the REAL justdoit.picaso() (justdoit.py:64-618).

/root/reference does not exist on the GPU box and there is no GPU in the build container, so the check is split:

* build container (CPU, needs /root/reference): load the unmodified justdoit module (tests/support/ref_justdoit.py),
  `picaso_b200.patch()` it and check what gets rebound, and that every replacement accepts the reference function's
  positional parameters under the same names;
* anywhere (CPU): every call the real picaso() made while tests/golden/make_golden_justdoit.py drove
  `inputs().spectrum()` binds to the replacement's signature;
* GPU: replay those recorded calls - arguments exactly as picaso() formed them - through the replacements and compare
  with what the reference functions returned (rtol 1e-6; level arrays by the level-flux criterion).
"""
import inspect
import os
import sys

import numpy as np
import pytest

import oracle
import picaso_b200 as pb
from util import assert_close, assert_level_close_yardstick, golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "support"))
import ref_justdoit  # noqa: E402

RTOL = 1e-6


def recorded_calls():
    g = golden("justdoit_calls")
    keys = sorted({k.rsplit("/", 1)[0] for k in g.files if "/nargs" in k})
    out = []
    for key in keys:
        fname = key.split("/")[1].split("_", 1)[1]
        args = [g[f"{key}/a{j:02d}"] for j in range(int(g[key + "/nargs"]))]
        args = [a.item() if a.ndim == 0 else a for a in args]
        kw = {str(n): g[f"{key}/k_{n}"] for n in g[key + "/kwnames"]}
        kw = {n: (v.item() if v.ndim == 0 else v) for n, v in kw.items()}
        res = []
        for j in range(int(g[key + "/nres"])):
            shape = tuple(int(x) for x in g[f"{key}/r{j:02d}_shape"])
            r = g[f"{key}/r{j:02d}"]
            res.append(r if r.shape == shape else np.zeros(shape))   # all-zero outputs are stored as their shape
        out.append((key, fname, args, kw, res))
    return out


CALLS = recorded_calls()


@pytest.mark.skipif(not ref_justdoit.available(), reason="needs the reference checkout (build container)")
def test_patch_rebinds_the_real_justdoit_module():
    jdi = ref_justdoit.load()
    before = {n: getattr(jdi, n) for n in pb._PATCHED if hasattr(jdi, n)}
    old = pb.patch(jdi)
    try:
        # every flux / disk-integration name justdoit imports (justdoit.py:2,9) is now ours
        for n in ("get_reflected_1d", "get_reflected_3d", "get_thermal_1d", "get_thermal_3d", "get_reflected_SH",
                  "get_thermal_SH", "get_transit_1d", "compress_disco", "compress_thermal", "mean_regrid"):
            assert getattr(jdi, n) is getattr(pb, n), n
            assert old[n] is before[n]
        # justdoit binds no get_fluxes (the climate loop is jitted, climate.py:804): patch() says so instead of implying it
        assert "get_fluxes" in old.missing
    finally:
        pb.unpatch(jdi, old)
    for n, fn in before.items():
        assert getattr(jdi, n) is fn


@pytest.mark.skipif(not ref_justdoit.available(), reason="needs the reference checkout (build container)")
@pytest.mark.parametrize("name", ["get_reflected_1d", "get_reflected_3d", "get_thermal_1d", "get_thermal_3d",
                                  "get_reflected_SH", "get_thermal_SH", "get_transit_1d", "compress_disco",
                                  "compress_thermal", "compute_opacity"])
def test_replacement_signature_covers_the_reference(name):
    """same positional parameters, same names, same order, same defaults; extras are keyword-only"""
    jdi = ref_justdoit.load()
    ref = getattr(jdi, name)
    ref = getattr(ref, "py_func", ref)          # numba dispatcher -> the Python function
    rp = list(inspect.signature(ref).parameters.values())
    mp = list(inspect.signature(getattr(pb, name)).parameters.values())
    pos = [p for p in mp if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    assert [p.name for p in pos] == [p.name for p in rp], name
    for a, b in zip(pos, rp):
        if b.default is not inspect.Parameter.empty:
            assert a.default == b.default, (name, a.name)
    assert all(p.kind == p.KEYWORD_ONLY for p in mp[len(pos):]), name


@pytest.mark.parametrize("key,fname,args,kw,res", CALLS, ids=[c[0] for c in CALLS])
def test_recorded_picaso_calls_bind(key, fname, args, kw, res):
    """CPU: the argument lists picaso() really forms are accepted by the replacement's signature"""
    inspect.signature(getattr(pb, fname)).bind(*args, **kw)


@pytest.mark.parametrize("key,fname,args,kw,res", CALLS, ids=[c[0] for c in CALLS])
def test_recorded_picaso_calls_oracle(key, fname, args, kw, res):
    """CPU: the oracle on the arguments of the real call sites against what the reference returned there"""
    fn = getattr(oracle, fname, None)
    if fn is None:
        pytest.skip("not restated in oracle/")
    got = fn(*args, **kw)
    flat = []
    for r in (got if isinstance(got, tuple) else (got,)):
        flat.extend(r if isinstance(r, tuple) else (r,))
    for j, (a, b) in enumerate(zip(flat, res)):
        if a is None:
            continue
        a = np.asarray(a)
        if a.shape != b.shape or (fname in ("get_reflected_1d", "get_thermal_1d") and j >= 1):
            continue      # level arrays: tests/test_oracle_golden.py holds them to the level-flux criterion
        assert_close(a, b, 1e-9, f"{key} oracle output {j}")


def _exact_levels(fname, args, kw):
    """binary128 evaluation of the reference formulas for the level arrays (yardstick of tests/util.py)"""
    if fname == "get_reflected_1d":
        return oracle.get_reflected_1d(*args, **kw, quad=True)[1]
    return oracle.get_thermal_1d(*args, **kw, quad=True)[1]


@pytest.mark.gpu
@pytest.mark.parametrize("key,fname,args,kw,res", CALLS, ids=[c[0] for c in CALLS])
def test_recorded_picaso_calls_replayed_on_gpu(key, fname, args, kw, res):
    got = getattr(pb, fname)(*args, **kw)
    flat = []
    for r in (got if isinstance(got, tuple) else (got,)):
        flat.extend(r if isinstance(r, tuple) else (r,))
    assert len(flat) == len(res), key
    levels = None
    for j, (a, b) in enumerate(zip(flat, res)):
        a = np.asarray(a)
        assert a.shape == b.shape, (key, j, a.shape, b.shape)
        if fname in ("get_reflected_1d", "get_thermal_1d") and j >= 1:
            if not b.any():
                assert not a.any(), (key, j)      # TOA-only call: the reference returns zero level arrays
                continue
            if levels is None:
                levels = _exact_levels(fname, args, kw)
            assert_level_close_yardstick(a, b, np.asarray(levels[j - 1]), what=f"{key} level array {j - 1}")
        elif fname == "get_reflected_SH" and j == 1:
            assert not a.any() and not b.any()   # flx = 0 in this run
        else:
            assert_close(a, b, RTOL, f"{key} output {j}")
