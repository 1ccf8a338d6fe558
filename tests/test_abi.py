"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/ndspmhd_b200.h declares, the
ctypes mirror matches the C structs byte for byte, and compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from ndspmhd_b200 import abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ndspmhd_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ndspmhd_b200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ndspmhd_b200.h but not exported"
    assert sorted(lib.EXPORTS) == syms, "lib.EXPORTS out of date with the header"


def test_header_is_plain_c_and_struct_layout_matches_ctypes():
    """Compile the header as C (gcc) and compare sizeof/offsetof with the ctypes mirror."""
    fields = {
        "nd_options": ["iener", "iavlim", "ibound", "device_ghosts", "idustevol", "hfact", "gamma", "xmin", "Bconst", "hhmax", "reserved_d"],
        "nd_arrays": ["x", "rho_in", "hh", "dens", "force", "del2u", "x_out", "dustevol", "dustfrac_in", "dustfrac", "ddeltavdt"],
        "nd_scalars": ["dtcourant", "fmean", "itsdensity", "ncellsx", "nrelink", "ncalctotal", "reserved_i"],
        "nd_step_opts": ["C_cour", "C_force", "dtfixed", "reserved"],
        "nd_state_out": ["x", "rho", "dustevol", "deltav"],
        "nd_evwrite": ["ekin", "rhomin", "emagp", "divBtot", "fluxtotmag", "ekiny", "totmassdust", "mom", "fluxtot", "reserved"],
    }
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for st, fl in fields.items():
        prog.append(f'printf("{st} %zu\\n", sizeof({st}));')
        for f in fl:
            prog.append(f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    prog.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        cfile = os.path.join(d, "t.c")
        open(cfile, "w").write("\n".join(prog))
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-o", exe, cfile])
        out = subprocess.check_output([exe]).decode().split("\n")
    got = dict(l.split() for l in out if l.strip())
    mirror = {"nd_options": abi.NdOptions, "nd_arrays": abi.NdArrays, "nd_scalars": abi.NdScalars, "nd_step_opts": abi.NdStepOpts,
              "nd_state_out": abi.NdStateOut, "nd_evwrite": abi.NdEvwrite}
    for st, cls in mirror.items():
        assert int(got[st]) == C.sizeof(cls), st
        for f in fields[st]:
            assert int(got[f"{st}.{f}"]) == getattr(cls, f).offset, f"{st}.{f}"


def test_default_options_match_reference_defaults():
    """src/defaults.f90:47-118"""
    L = lib.load()
    o = abi.NdOptions()
    assert L.ndspmhd_b200_default_options(C.byref(o)) == 0
    py = abi.default_options()
    for name, _ in abi.NdOptions._fields_:
        if name.startswith("reserved") or name in ("device_ghosts",):
            continue
        a, b = getattr(o, name), getattr(py, name)
        if hasattr(a, "__len__"):
            assert list(a) == list(b), name
        else:
            assert a == b, name
    assert (o.iener, o.iav, o.ikernav, o.ihvar, o.maxdensits) == (2, 2, 3, 2, 250)
    assert (o.hfact, o.tolh, o.alphamin, o.alphaBmin, o.beta, o.psidecayfact) == (1.2, 1e-3, 0.1, 1.0, 2.0, 0.1)
    assert list(o.iavlim) == [2, 1, 0]


def test_unsupported_option_is_an_error_code_not_a_fallback():
    o = abi.default_options()
    o.iprterm = 2
    with pytest.raises(lib.NdError) as e:
        lib.Hotpath(o, 3)
    assert e.value.code == abi.ND_ERR_UNSUPPORTED_OPTION
    o = abi.default_options()
    with pytest.raises(lib.NdError) as e:
        lib.Hotpath(o, 4)
    assert e.value.code == abi.ND_ERR_INVALID_ARG


def test_no_cpu_fallback_without_a_device():
    if lib.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(lib.NdError) as e:
        lib.Hotpath(abi.default_options(), 3)
    assert e.value.code == abi.ND_ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    """The package and the C-ABI sources must not import, link or mention the oracle."""
    pkg = os.path.join(ROOT, "ndspmhd_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile", ".f90")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "nd_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, os.path.join(base, f)
    out = subprocess.check_output(["ldd", lib.LIB_PATH]).decode()
    assert "oracle" not in out
