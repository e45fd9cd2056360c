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


def duck_atmosphere(db, atm):
    """what compute_opacity / get_opacities read from the reference's ATMSETUP object"""
    import types
    a = types.SimpleNamespace()
    a.c = types.SimpleNamespace(nlayer=atm["nlayer"], pconv=atm["pconv"], rgas=atm["rgas"], amu=atm["amu"],
                                k_b=atm["k_b"])
    a.level = {"temperature": atm["tlevel"], "pressure": atm["plevel"]}
    a.layer = {"temperature": atm["tlayer"], "pressure": atm["player"], "colden": atm["colden"],
               "mmw": atm["mmw"], "mixingratios": atm["mixingratios"], "electrons": atm["electrons"],
               "cloud": {"opd": atm["cloud_opd"], "w0": atm["cloud_w0"], "g0": atm["cloud_g0"]}}
    a.planet = types.SimpleNamespace(gravity=atm["gravity"])
    a.molecules = list(db["molecules"])
    a.continuum_molecules = [list(x) for x in db["continuum_molecules"]]
    a.rayleigh_molecules = list(db["rayleigh_molecules"])
    return a


def device_opacities(pb, case, db, ins):
    opa = pb.DeviceOpacities(db["wno"], db["pt_pairs"], db["tables"], db["cia_temps"], db["continuum"],
                             ins["rayleigh"], raman_db=(ins["raman_c"], ins["raman_ji"], ins["raman_deltanu"]),
                             query_method=case["query"])
    if case["raman"] == 0:
        opa.raman_stellar_shifts = ins["stellar_shifts"]
    return opa
