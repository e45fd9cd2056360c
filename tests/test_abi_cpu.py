"""CPU: the C-ABI library loads and exports every symbol include/picaso_b200.h declares;
without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import picaso_b200
from picaso_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load_library()


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "picaso_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export " + n
        assert n in _lib.SYMBOLS, "no ctypes prototype for " + n
    assert sorted(_lib.SYMBOLS) == names


def test_version(lib):
    assert lib.pb_version() >= 100


def test_struct_sizes_match_header():
    # 5 ints + pad, i64, 18 pointers, double, 3 ints (+pad), 5 doubles, 2 ints, 6 pointers
    # ... variant (+pad), gather pointer
    assert ctypes.sizeof(_lib.ReflectedArgs) == 24 + 8 + 18 * 8 + 8 + 16 + 40 + 8 + 48 + 8 + 8
    assert ctypes.sizeof(_lib.ThermalArgs) == 24 + 8 + 11 * 8 + 8 + 6 * 8 + 8
    assert ctypes.sizeof(_lib.TransitArgs) == 16 + 8 + 7 * 8 + 24 + 8


def test_ctypes_structs_match_the_compiled_header(tmp_path):
    """sizeof of every argument struct as gcc lays out include/picaso_b200.h == the ctypes mirror"""
    import subprocess
    pairs = {"pb_reflected_args": _lib.ReflectedArgs, "pb_thermal_args": _lib.ThermalArgs,
             "pb_transit_args": _lib.TransitArgs, "pb_sh_args": _lib.ShArgs, "pb_thermal_sh_args": _lib.ThermalShArgs,
             "pb_opacity_args": _lib.OpacityArgs, "pb_ck_mix_args": _lib.CkMixArgs, "pb_climate_args": _lib.ClimateArgs,
             "pb_peer_gather": _lib.PeerGather, "pb_spectrum_args": _lib.SpectrumArgs,
             "pb_spectrum_thermal_args": _lib.SpectrumThermalArgs, "pb_spectrum_transit_args": _lib.SpectrumTransitArgs}
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "picaso_b200.h"\nint main(void){\n' +
                   "".join('printf("%s %%zu\\n", sizeof(%s));\n' % (n, n) for n in pairs) + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    sizes = dict(zip(out[0::2], map(int, out[1::2])))
    for n, cls in pairs.items():
        assert sizes[n] == ctypes.sizeof(cls), (n, sizes[n], ctypes.sizeof(cls))


def test_no_cpu_fallback(lib):
    n = ctypes.c_int(-1)
    lib.pb_device_count(ctypes.byref(n))
    if n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(picaso_b200.PicasoB200Error):
        _lib.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "picaso_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and \
                    "picaso_oracle" not in src, f


def test_patch_rebinds_and_restores_names():
    """picaso_b200.patch(module): the drop-in seam of INTEGRATION.md - names the module has are rebound to the
    CUDA mirrors, names it does not have are left alone, unpatch restores the originals"""
    import types
    mod = types.ModuleType("fake_justdoit")
    sentinel = object()
    for n in ("get_reflected_1d", "get_thermal_1d", "compress_disco", "get_fluxes", "mean_regrid"):
        setattr(mod, n, sentinel)
    mod.unrelated = 7
    old = picaso_b200.patch(mod)
    assert set(old) == {"get_reflected_1d", "get_thermal_1d", "compress_disco", "get_fluxes", "mean_regrid"}
    assert mod.get_reflected_1d is picaso_b200.get_reflected_1d and mod.get_fluxes is picaso_b200.get_fluxes
    assert mod.mean_regrid is picaso_b200.mean_regrid and mod.unrelated == 7
    assert not hasattr(mod, "get_transit_1d")
    picaso_b200.unpatch(mod, old)
    assert mod.get_reflected_1d is sentinel and mod.get_fluxes is sentinel
