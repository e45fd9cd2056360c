"""Generate tests/golden/sh_flux.npz: get_reflected_SH(flx=1) of the UNMODIFIED reference (fluxes.py:2675-2976).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_sh_flux.py
Stores the reference's `flux` output [ng, nt, stream*nlevel, nwno] (calculate_flux(F, G, X), fluxes.py:2889-2890)
and xint_at_top for the cases of cases.sh_flux_cases(); inputs are regenerated from seeds.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_loader  # noqa: E402
import cases as C  # noqa: E402

REF_COMMIT = "0369089"


def main():
    F = ref_loader.load("fluxes")
    out = {}
    for name, case in C.sh_flux_cases().items():
        d = C.build_sh(case)
        xint, flux = F.get_reflected_SH(*C.sh_args(d, case, flx=1))
        out[name + "/xint"] = xint
        out[name + "/flux"] = flux
    np.savez_compressed(os.path.join(HERE, "sh_flux.npz"), ref_commit=REF_COMMIT, **out)
    print("sh_flux:", len(out), "arrays")


if __name__ == "__main__":
    main()
