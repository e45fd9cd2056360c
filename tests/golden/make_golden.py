"""Generate tests/golden/*.npz by running the UNMODIFIED reference here.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference modules are imported by file path (oracle/ref_loader.py); inputs come
from picaso_b200.synth seeds (tests/golden/cases.py) and only reference outputs are
stored, tagged with the reference commit.
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
    D = ref_loader.load("disco")
    out = {}
    for name, case in C.reflected_cases().items():
        d = C.build_reflected(case)
        xint, lv = F.get_reflected_1d(*C.reflected_args(d, case["kw"]))
        out[name + "/xint"] = xint
        out[name + "/albedo"] = D.compress_disco(d["nwno"], d["cos_theta"], xint, d["gweight"],
                                                 d["tweight"], d["F0PI"])
        if case["kw"]["get_lvl_flux"]:
            for k, a in zip(("fm", "fp", "fmm", "fpm"), lv):
                out[name + "/" + k] = a
    np.savez_compressed(os.path.join(HERE, "reflected.npz"), ref_commit=REF_COMMIT, **out)
    print("reflected:", len(out), "arrays")

    out = {}
    for name, case in C.thermal_cases().items():
        d = C.build_thermal(case)
        ftop, lv = F.get_thermal_1d(*C.thermal_args(d))
        out[name + "/ftop"] = ftop
        out[name + "/thermal"] = D.compress_thermal(d["nwno"], ftop, d["gweight"], d["tweight"])
        if d["nwno"] <= 64:
            for k, a in zip(("fm", "fp", "fmm", "fpm"), lv):
                out[name + "/" + k] = a
    np.savez_compressed(os.path.join(HERE, "thermal.npz"), ref_commit=REF_COMMIT, **out)
    print("thermal:", len(out), "arrays")

    out = {}
    for name, case in C.sh_cases().items():
        d = C.build_sh(case)
        a = C.sh_args(d, case)
        xint, _ = F.get_reflected_SH(*a)
        out[name + "/xint"] = xint
        out[name + "/albedo"] = D.compress_disco(d["nwno"], d["cos_theta"], xint, d["gweight"],
                                                 d["tweight"], d["F0PI"])
        out[name + "/f_deltaM_after"] = a[10]   # the reference scaled it in place (Appendix A1)
    np.savez_compressed(os.path.join(HERE, "sh.npz"), ref_commit=REF_COMMIT, **out)
    print("sh:", len(out), "arrays")

    out = {}
    from picaso_b200 import synth
    for name, kw in C.transit_cases().items():
        d = synth.transit_inputs(**kw)
        out[name + "/F"] = F.get_transit_1d(*C.transit_args(d))
    np.savez_compressed(os.path.join(HERE, "transit.npz"), ref_commit=REF_COMMIT, **out)
    print("transit:", len(out), "arrays")


if __name__ == "__main__":
    main()
