"""Shared set-up for the opacity-path tests: rebuild the synthetic DB / atmosphere of a case and
fetch the reference-derived inputs stored beside the golden outputs."""
import numpy as np

import cases as C
from picaso_b200 import synth
from util import golden

OUT_NAMES = ("DTAU", "TAU", "W0", "COSB", "ftau_cld", "ftau_ray", "GCOS2", "DTAU_OG", "TAU_OG", "W0_OG",
             "COSB_OG", "W0_no_raman", "f_deltaM")


def load_case(name):
    case = C.optics_cases()[name]
    g = golden("optics")
    db = synth.opacity_database(**case["db"])
    atm = synth.atmosphere_profile(db, **case["atm"])
    atm["cia_pairs"] = {a + b: (a, b) for a, b in db["continuum_molecules"]}
    ins = dict(rayleigh={m: g[f"{name}/in/rayleigh/{m}"] for m in db["rayleigh_molecules"]},
               raman_c=g[f"{name}/in/raman_c"], raman_ji=g[f"{name}/in/raman_ji"],
               raman_deltanu=g[f"{name}/in/raman_deltanu"], stellar_shifts=g[f"{name}/in/stellar_shifts"])
    return case, g, db, atm, ins
