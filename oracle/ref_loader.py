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
