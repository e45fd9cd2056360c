"""Record what the UNMODIFIED reference picaso() (justdoit.py:64-618) passes to, and gets back from, every function
picaso_b200.patch() replaces - tests/golden/justdoit_calls.npz.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_justdoit.py

The real justdoit module is loaded under stub packages for its plotting / stellar-spectrum imports
(tests/support/ref_justdoit.py); atmsetup, optics (RetrieveOpacities on a synthetic sqlite file with the reference
schema), fluxes and disco are the reference's own files.  `inputs().spectrum(opa, calculation=...)` then runs with
recording wrappers bound to the module globals picaso() resolves at call time (justdoit.py:2,8,9) - the same seam
patch() uses.  Stored per call: the positional arguments and keywords exactly as picaso() formed them (the [:, :, ig]
slices, wno*0, int(get_lvl_flux), ...) and the reference's return values.  The GPU test replays the recorded calls
through picaso_b200's functions (tests/test_justdoit_seam.py): a drop-in check at the real call sites that can run
on a box where /root/reference does not exist.
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, HERE, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "support")):
    sys.path.insert(0, p)

os.environ["NUMBA_CACHE_DIR"] = os.environ.get("PB_JUSTDOIT_NUMBA_CACHE", "/tmp/numba_cache_justdoit")  # before numba loads
import ref_justdoit  # noqa: E402
from picaso_b200 import synth  # noqa: E402
from test_opacity_db import write_db  # noqa: E402

REF_COMMIT = "0369089"
RECORDED = ("get_reflected_1d", "get_thermal_1d", "get_transit_1d", "get_reflected_SH", "get_thermal_SH",
            "compress_disco", "compress_thermal")


def atmosphere_frame(db, L, seed):
    """pressure-temperature-composition table (levels) inside the synthetic opacity grid"""
    rng = np.random.default_rng(seed)
    nlevel = L + 1
    P = np.logspace(-5.5, 1.5, nlevel)                              # bar
    T = np.linspace(np.min(db["temps"]) * 1.08, np.max(db["temps"]) * 0.92, nlevel)
    df = pd.DataFrame({"pressure": P, "temperature": T})
    df["H2"] = 0.84
    df["He"] = 0.15
    for i, m in enumerate(db["molecules"]):
        df[m] = 10.0 ** (-3.5 - 0.4 * i) * (1.0 + 0.2 * rng.random(nlevel))
    return df


def run_case(jdi, opa, name, calculation, out, *, L=14, seed=9001, rt_method="toon", stream=2, lvl=False, raman="none"):
    case = jdi.inputs()
    case.phase_angle(0, num_gangle=6, num_tangle=1)
    # gravity() / star() go through astropy units and stellar grids (stubbed here): fill the dictionary entries those
    # methods set (justdoit.py:1586-1680, :1756-1913) by hand, in cgs
    case.inputs["planet"].update(gravity=2479.0, gravity_unit="cm/(s**2)", radius=7.1492e9, radius_unit="cm",
                                 mass=1.898e30, mass_unit="g")
    case.inputs["star"].update(database="nostar", radius=6.957e10, radius_unit="cm", semi_major=7.78e13,
                               semi_major_unit="cm", temp="nostar", logg="nostar", metal="nostar")
    opa.unshifted_stellar_spec = np.ones(opa.nwno)
    opa.relative_flux = np.ones(opa.nwno)
    case.atmosphere(df=atmosphere_frame(opa_db, L, seed))
    case.approx(raman=raman, rt_method=rt_method, stream=stream, get_lvl_flux=lvl)
    calls = []
    originals = {}

    def wrap(fname):
        fn = getattr(jdi, fname)

        def rec(*a, **k):
            res = fn(*a, **k)
            calls.append((fname, a, k, res))
            return res
        return fn, rec

    for fname in RECORDED:
        originals[fname], w = wrap(fname)
        setattr(jdi, fname, w)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ret = case.spectrum(opa, calculation=calculation)
    finally:
        for fname, fn in originals.items():
            setattr(jdi, fname, fn)
    # the level-flux run compresses every level separately (justdoit.py:546-550): three of those calls are enough
    kept, seen = [], {}
    for c in calls:
        seen[c[0]] = seen.get(c[0], 0) + 1
        if seen[c[0]] <= 3:
            kept.append(c)
    for i, (fname, a, k, res) in enumerate(kept):
        key = f"{name}/{i:02d}_{fname}"
        out[key + "/nargs"] = np.array(len(a))
        for j, v in enumerate(a):
            out[f"{key}/a{j:02d}"] = np.asarray(v)
        out[key + "/kwnames"] = np.array(sorted(k), dtype="U32")
        for kk, v in k.items():
            out[f"{key}/k_{kk}"] = np.asarray(v)
        res_t = res if isinstance(res, tuple) else (res,)
        flat = []
        for r in res_t:
            flat.extend(r if isinstance(r, tuple) else (r,))
        out[key + "/nres"] = np.array(len(flat))
        for j, v in enumerate(flat):
            v = np.asarray(v)
            # level arrays of a TOA-only call are all zero in the reference: keep the shape, not the megabytes
            out[f"{key}/r{j:02d}"] = v if v.any() or v.size < 64 else np.zeros(v.shape[:0] + (0,)) + 0
            out[f"{key}/r{j:02d}_shape"] = np.array(v.shape)
    for kk in ("albedo", "thermal", "transit_depth"):
        if kk in ret and not isinstance(ret[kk], list):
            out[f"{name}/returns/{kk}"] = np.asarray(ret[kk])
    print(name, calculation, [c[0] for c in calls])


if __name__ == "__main__":
    jdi = ref_justdoit.load()
    opa_db = synth.opacity_database(W=96, nmol=4, seed=77)
    path = "/tmp/pb_justdoit_opa.db"
    if os.path.exists(path):
        os.remove(path)
    write_db(path, opa_db)
    opa = jdi.opannection(filename_db=path)
    out = {}
    run_case(jdi, opa, "toon_all", "reflected+thermal+transmission", out)
    run_case(jdi, opa, "toon_lvl", "reflected+thermal", out, L=10, seed=9002, lvl=True)
    run_case(jdi, opa, "sh4_reflected", "reflected", out, L=12, seed=9003, rt_method="SH", stream=4)
    np.savez_compressed(os.path.join(HERE, "justdoit_calls.npz"), ref_commit=REF_COMMIT, **out)
    print("justdoit_calls:", len(out), "arrays")
