"""picaso_b200 - B200-native (sm_100a CUDA behind a C ABI) implementation of PICASO's
per-wavelength radiative-transfer hot path: Toon89 reflected / thermal solvers, transit
chord integration and disk integration, behind the reference's Python signatures.

    import picaso_b200
    picaso_b200.patch(picaso.justdoit)     # drop-in behind jdi.inputs().spectrum()

No PyTorch, no CPU fallback: importing is cheap, the first call creates the CUDA context
and raises if the library or a device is missing.
"""
from ._lib import PicasoB200Error, Context, default_context, load_library  # noqa: F401
from .fluxes import get_reflected_1d, get_reflected_SH, get_thermal_1d, get_transit_1d  # noqa: F401
from .fluxes_sh_thermal import get_thermal_SH  # noqa: F401
from .fluxes_3d import get_reflected_3d, get_thermal_3d  # noqa: F401
from .optics import (DeviceArray, DeviceOpacities, compute_opacity, reflected_spectrum, thermal_spectrum,  # noqa: F401
                     transit_spectrum)
from .optics_ck import DeviceCKs, DeviceGasCKs  # noqa: F401
from .opacity_db import opannection, read_opacity_db  # noqa: F401
from .climate import get_fluxes, get_fluxes_jacobian, BoundFluxes  # noqa: F401
from .regrid import mean_regrid, RegridPlan  # noqa: F401
from .batch import thermal_batch  # noqa: F401
from .disco import compress_disco, compress_thermal, get_angles_1d, get_angles_3d, compute_disco  # noqa: F401

__version__ = "0.1.0"

_PATCHED = ("get_reflected_1d", "get_reflected_3d", "get_reflected_SH", "get_thermal_1d", "get_thermal_3d",
            "get_thermal_SH", "get_transit_1d", "compress_disco",
            "compress_thermal", "get_fluxes", "mean_regrid")


class Patched(dict):
    """{name: original} of the names `patch` rebound; `.missing` lists the names the module does not bind
    (picaso.justdoit binds neither `get_fluxes` nor - in 4.0.1 - anything of the climate solver's jitted loop:
    that path is reached through picaso_b200.get_fluxes / BoundFluxes)"""
    missing = ()


def patch(module):
    """Rebind the hot-path names of a loaded `picaso.justdoit` module (justdoit.py:2,9) to
    the CUDA implementations.  Returns the replaced originals (a dict, for `unpatch`); its `.missing`
    attribute names what the module does not bind, so a caller can see what was NOT replaced."""
    import sys
    me = sys.modules[__name__]
    old = Patched()
    missing = []
    for name in _PATCHED:
        if hasattr(module, name):
            old[name] = getattr(module, name)
            setattr(module, name, getattr(me, name))
        else:
            missing.append(name)
    old.missing = tuple(missing)
    return old


def unpatch(module, old):
    for name, fn in old.items():
        setattr(module, name, fn)
