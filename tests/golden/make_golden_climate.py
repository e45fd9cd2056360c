"""Golden vectors for climate.get_fluxes from the unmodified reference (numba nopython).

picaso/climate.py is loaded under the synthetic `refpicaso` package of oracle/ref_loader.py with stub
modules for astropy.units / virga (module-level imports the function never touches)."""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import ref_loader  # noqa: E402
import cases as C  # noqa: E402


def load_climate():
    ref_loader.load_optics()

    def stub(name):
        if name in sys.modules:
            return
        m = types.ModuleType(name)
        sys.modules[name] = m
        parent, _, leaf = name.rpartition(".")
        if parent:
            stub(parent)
            setattr(sys.modules[parent], leaf, m)

    for n in ("astropy.units", "astropy.constants", "virga", "virga.justdoit", "virga.justplotit"):
        try:
            importlib.import_module(n)
        except Exception:
            stub(n)
    return importlib.import_module("refpicaso.climate")


def main():
    clim = load_climate()
    out = {}
    for name, case in C.climate_cases().items():
        d = C.build_climate(case)
        # the reference's tuples (numba types namedtuples by class): rebuild with its own classes
        A = clim.Atmosphere_Tuple(*d["Atmosphere"])
        Wd = clim.OpacityWEd_Tuple(*d["OpacityWEd"])
        Nd = clim.OpacityNoEd_Tuple(*d["OpacityNoEd"])
        S = clim.ScatteringPhase_Tuple(*d["ScatteringPhase"])
        D = clim.Disco_Tuple(*d["Disco"])
        args = [A, Wd, Nd, S, D, d["Opagrid"], d["F0PI"], case["reflected"], case["thermal"]]
        if "fhole" in case:
            args += [True, case["fhole"], clim.OpacityWEd_Tuple(*d["hole_OpacityWEd"]),
                     clim.OpacityNoEd_Tuple(*d["hole_OpacityNoEd"])]
        res = clim.get_fluxes(*args)
        for k, v in zip(C.CLIMATE_OUT, res):
            out[name + "/" + k] = np.asarray(v)
        print(name, "done", flush=True)
    np.savez_compressed(os.path.join(HERE, "climate.npz"), ref_commit="0369089", **out)
    print("climate:", len(out), "arrays")


if __name__ == "__main__":
    main()
