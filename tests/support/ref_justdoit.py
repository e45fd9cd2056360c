"""Load the UNMODIFIED reference picaso/justdoit.py (test infrastructure, build container only).

justdoit imports plotting, stellar-spectrum and file-format packages that are not in this image (virga, synphot,
stsynphot, astropy, xarray, bokeh, matplotlib ...).  None of them is touched by the name binding that
picaso_b200.patch() relies on (justdoit.py:2,8,9) nor by picaso() itself (justdoit.py:49-640), so they are served
by catch-all stub modules; the reference's own modules (atmsetup, fluxes, climate, optics, disco, ...) are the
real files, loaded under a synthetic parent package like oracle/ref_loader.load_optics does.
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("PICASO_REFERENCE", "/root/reference")


class _Anything:
    """stands for any attribute of a stubbed package: callable, subscriptable, usable as a decorator"""
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Anything()

    def __getattr__(self, n):
        return _Anything()

    def __getitem__(self, k):
        return _Anything()

    def __iter__(self):
        return iter(())

    def _same(self, other):
        return self
    __mul__ = __rmul__ = __truediv__ = __rtruediv__ = __pow__ = __add__ = __radd__ = __sub__ = __rsub__ = _same


class _StubModule(types.ModuleType):
    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Anything()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ("virga", "synphot", "stsynphot", "astropy", "xarray", "h5py", "bokeh", "photutils", "matplotlib",
             "mpl_toolkits", "holoviews", "hvplot", "dynesty", "ultranest", "corner", "pysynphot", "seaborn", "colorcet",
             "IPython", "PyMieScatt", "sklearn_extra", "numba_progress")

    def find_spec(self, name, path, target=None):
        root = name.split(".")[0]
        if root in self.ROOTS:
            try:   # the real package wins when it is installed
                for f in sys.meta_path:
                    if f is not self and getattr(f, "find_spec", None):
                        s = f.find_spec(name, path, target)
                        if s is not None:
                            return None
            except Exception:
                pass
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, m):
        pass


class _Const:
    """astropy.constants entry as atmsetup.get_constants uses it (atmsetup.py:53-56): .value (SI) and .to(cgs unit).value"""
    def __init__(self, si, cgs):
        self.value, self._cgs = si, cgs

    def to(self, *a, **k):
        return _Const(self._cgs, self._cgs)


def _install_constants():
    """CODATA 2018 (astropy's default set) for the four constants ATMSETUP reads; only needed when astropy is a stub"""
    m = sys.modules.get("astropy.constants")
    if isinstance(m, _StubModule):
        m.k_B = _Const(1.380649e-23, 1.380649e-16)
        m.G = _Const(6.6743e-11, 6.6743e-08)
        m.u = _Const(1.66053906660e-27, 1.66053906660e-24)
        m.R = _Const(8.31446261815324, 8.31446261815324e7)


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "picaso", "justdoit.py"))


def load():
    """-> the reference's justdoit module object (its sibling modules are the real reference files)"""
    if "refpicaso.justdoit" in sys.modules:
        return sys.modules["refpicaso.justdoit"]
    # own numba cache: oracle/ref_loader.py caches the same source files under other module names
    os.environ["NUMBA_CACHE_DIR"] = os.environ.get("PB_JUSTDOIT_NUMBA_CACHE", "/tmp/numba_cache_justdoit")
    os.environ.setdefault("picaso_refdata", os.path.join(REF_ROOT, "reference"))
    cdbs = os.environ.setdefault("PYSYN_CDBS", "/tmp/picaso_b200_cdbs")
    os.makedirs(cdbs, exist_ok=True)
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _StubFinder())
    if "refpicaso" not in sys.modules:
        pkg = types.ModuleType("refpicaso")
        pkg.__path__ = [os.path.join(REF_ROOT, "picaso")]
        sys.modules["refpicaso"] = pkg
    importlib.import_module("astropy.constants")
    _install_constants()
    return importlib.import_module("refpicaso.justdoit")
