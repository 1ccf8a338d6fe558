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
        "nd_arrays": ["x", "rho_in", "hh", "dens", "force", "del2u", "x_out", "dustevol", "dustfrac_in", "dustfrac", "ddeltavdt", "alpha_out"],
        "nd_scalars": ["dtcourant", "fmean", "itsdensity", "ncellsx", "nrelink", "ncalctotal", "lmax", "list_overflows", "rate_chunks", "reserved_i", "npairs_rates", "ntrips_rates"],
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


def test_fortran_module_types_match_the_c_abi():
    """ndspmhd_b200/fortran/ndspmhd_b200_mod.f90 cannot be compiled here (no Fortran compiler in the image): its `type, bind(C)` blocks are
    parsed with numpy.f2py's crackfortran instead and compared with the ctypes mirror of include/ndspmhd_b200.h -- member names, order,
    base type and array extents -- and no source line of the shims may pass column 132 (the reference builds with -std=f2008)."""
    import contextlib
    import ctypes as C
    import glob
    import io

    from numpy.f2py import crackfortran

    fdir = os.path.join(ROOT, "ndspmhd_b200", "fortran")
    for f in glob.glob(os.path.join(fdir, "*.f90")):
        for i, line in enumerate(open(f), 1):
            assert len(line.rstrip("\n")) <= 132, (os.path.basename(f), i)
    src = os.path.join(fdir, "ndspmhd_b200_mod.f90")
    crackfortran.verbose = 0
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        blocks = crackfortran.crackfortran([src])
    types = {b["name"]: b for b in blocks[0]["body"] if b["block"] == "type"}
    text = open(src).read().lower()
    pairs = {"nd_options": abi.NdOptions, "nd_arrays": abi.NdArrays, "nd_scalars": abi.NdScalars, "nd_step_opts": abi.NdStepOpts,
             "nd_state_out": abi.NdStateOut}
    kinds = {"c_double": ("real", "c_double"), "c_int": ("integer", "c_int"), "c_longlong": ("integer", "c_long_long"),
             "c_long": ("integer", "c_long_long")}   # ctypes aliases c_longlong to c_long on LP64
    for fname, ct in pairs.items():
        t = types[fname]
        forder = [v.lower() for v in (t.get("sortvars") or list(t["vars"]))]
        assert forder == [n.lower() for n, _ in ct._fields_], fname
        for n, ctyp in ct._fields_:
            v = t["vars"][n] if n in t["vars"] else t["vars"][n.lower()]
            base, length = ctyp, 1
            while hasattr(base, "_length_"):
                length *= base._length_
                base = base._type_
            if base.__name__ in kinds:
                ftype, fkind = kinds[base.__name__]
                kind = (v.get("kindselector") or {}).get("kind")
                assert v.get("typespec") == ftype, (fname, n)
                assert str(kind) == fkind or f"{ftype}({fkind}) :: {n.lower()}" in text, (fname, n, kind)
                flen = 1
                for d in v.get("dimension") or []:
                    flen *= int(d)
                assert flen == length, (fname, n, flen, length)
            else:                                   # pointers travel as type(c_ptr)
                assert v.get("typespec") == "type" and "c_ptr" in str(v.get("typename", "")).lower(), (fname, n)
    # the interface block: every bound function exists in the header with the same number of arguments; a C pointer argument is either a
    # by-reference dummy or a type(c_ptr) passed by value, every other C argument is a `value` dummy
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ndspmhd_b200.h")).read(), flags=re.S)
    protos = {m.group(2): [a.strip() for a in m.group(3).split(",")]
              for m in re.finditer(r"\b(int|void|const char \*|void \*)\s*(ndspmhd_b200_\w+)\s*\(([^)]*)\)\s*;", hdr)}
    iface = [b for b in blocks[0]["body"] if b["block"] == "interface"][0]
    assert len(iface["body"]) >= 15
    for f in iface["body"]:
        cargs = protos[f["name"]]
        assert len(cargs) == len(f["args"]), f["name"]
        for fa, ca in zip(f["args"], cargs):
            v = f["vars"][fa]
            byval = "value" in (v.get("attrspec") or [])
            is_cptr = v.get("typespec") == "type" and "c_ptr" in str(v.get("typename", "")).lower()
            if "*" in ca or "[" in ca:                  # a C array parameter is a pointer
                assert (not byval) or is_cptr, (f["name"], fa, ca)
            else:
                assert byval and not is_cptr, (f["name"], fa, ca)
