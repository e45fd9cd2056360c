"""Load the UNMODIFIED reference modules by file path (test infrastructure only).

Only usable where /root/reference exists (the build container); nothing that runs
on the GPU box may import this.  Recipe from SURVEY.md section 8c: the numba
functions are cache=True and /root/reference is read-only, so NUMBA_CACHE_DIR must
point somewhere writable, and the module must be registered in sys.modules or a
warm numba cache fails to unpickle.
"""
import importlib.util
import os
import sys

REF_ROOT = os.environ.get("PICASO_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "picaso", "fluxes.py"))


def load(name):
    """name in {'fluxes','disco','deq_chem'} -> module object of the reference file."""
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    modname = "ref_" + name
    if modname in sys.modules:
        return sys.modules[modname]
    path = os.path.join(REF_ROOT, "picaso", name + ".py")
    spec = importlib.util.spec_from_file_location(modname, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[modname] = m
    spec.loader.exec_module(m)
    return m


def load_optics():
    """The reference's picaso/optics.py under a synthetic parent package with stub modules for
    its plotting / file-format imports (bokeh, astropy.io.fits, h5py) - none of which the math
    in compute_opacity / compute_raman / interp_matrix touches."""
    import types
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    os.environ.setdefault("picaso_refdata", os.path.join(REF_ROOT, "reference"))
    if "refpicaso.optics" in sys.modules:
        return sys.modules["refpicaso.optics"]
    for name in ("bokeh", "bokeh.plotting", "bokeh.palettes", "astropy", "astropy.io", "astropy.io.fits",
                 "h5py"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = types.ModuleType(name)
                m.__dict__.update(figure=None, show=None, output_file=None, inferno=None)
                sys.modules[name] = m
    if "astropy.io" in sys.modules and not hasattr(sys.modules["astropy.io"], "fits"):
        sys.modules["astropy.io"].fits = sys.modules["astropy.io.fits"]
    pkg = types.ModuleType("refpicaso")
    pkg.__path__ = [os.path.join(REF_ROOT, "picaso")]
    sys.modules["refpicaso"] = pkg
    return importlib.import_module("refpicaso.optics")
