"""ctypes binding of the CPU oracle (oracle/libnd_oracle.so).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from ndspmhd_b200.abi import NdEvwrite, NdOptions, NdScalars, Particles

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NDO_GHOSTS, NDO_LINK, NDO_DENSITY, NDO_C2P, NDO_RATES, NDO_ALL = 1, 2, 4, 8, 16, 31

_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)

_D_NAMES = ["x", "vel", "pmass", "hh", "en", "Bevol", "alpha", "psi", "rho"]
_I_NAMES = ["itype", "ireal"]
_D2_NAMES = ["gradh", "gradhn", "gradsoft", "gradgradh", "rhoalt", "drhodt", "dhdt"]
_D3_NAMES = ["dens", "uu", "pr", "spsound", "Bfield", "sqrtg", "force", "dudt", "dendt", "dBevoldt", "daldt", "dpsidt",
             "gradpsi", "fmag", "divB", "curlB", "graddivv", "del2u", "xsphterm",
             "dustevol", "dustfrac", "deltav", "rhogas", "rhodust", "ddustevoldt", "ddeltavdt"]


class NdoArrays(C.Structure):
    _fields_ = ([(n, _DP) for n in _D_NAMES] + [(n, _IP) for n in _I_NAMES] + [(n, _DP) for n in _D2_NAMES]
                + [("numneigh", _IP)] + [(n, _DP) for n in _D3_NAMES])


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libnd_oracle.so")
    src = os.path.join(_HERE, "nd_oracle.cpp")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ndo_derivs.restype = C.c_int
        _LIB.ndo_derivs.argtypes = [C.POINTER(NdOptions), C.c_int, C.POINTER(NdoArrays), C.c_int, _IP, C.c_int, C.c_int,
                                    C.POINTER(NdScalars), _DP]
        _LIB.ndo_step.restype = C.c_int
        _LIB.ndo_step.argtypes = [C.POINTER(NdOptions), C.c_int, C.POINTER(NdoArrays), C.c_int, _IP, C.c_int, _DP, C.c_double, C.c_double, C.c_int,
                                  C.POINTER(NdScalars)]
        _LIB.ndo_evwrite.restype = C.c_int
        _LIB.ndo_evwrite.argtypes = [C.POINTER(NdOptions), C.c_int, C.POINTER(NdoArrays), C.c_int, C.POINTER(NdEvwrite)]
        _LIB.ndo_kernel_tables.restype = C.c_int
        _LIB.ndo_kernel_tables.argtypes = [C.c_int, C.c_int, C.c_int, _DP, _DP, _DP, _DP, _DP, _DP]
        _LIB.ndo_interpolate.restype = C.c_int
        _LIB.ndo_interpolate.argtypes = [C.c_int, C.c_int, C.c_double, _DP, _DP, _DP]
        _LIB.ndo_ran1.restype = C.c_double
        _LIB.ndo_ran1.argtypes = [_IP]
        _LIB.ndo_bruteforce_pairs.restype = C.c_longlong
        _LIB.ndo_bruteforce_pairs.argtypes = [C.c_int, _DP, _DP, C.c_int, C.c_int, C.c_double, _IP, _IP, C.c_longlong]
        _LIB.ndo_linklist_pairs.restype = C.c_longlong
        _LIB.ndo_linklist_pairs.argtypes = [C.POINTER(NdOptions), C.c_int, C.POINTER(NdoArrays), C.c_int, C.c_int, C.c_int,
                                            _IP, _IP, C.c_longlong]
        _LIB.ndo_get_curl.restype = C.c_int
        _LIB.ndo_get_curl.argtypes = [C.POINTER(NdOptions), C.c_int, C.POINTER(NdoArrays), C.c_int, C.c_int, C.c_int, C.c_int, _DP, _DP, _DP]
        _LIB.ndo_last_error.restype = C.c_char_p
    return _LIB


def _arrays(p: Particles) -> NdoArrays:
    a = NdoArrays()
    for n, _ in NdoArrays._fields_:
        setattr(a, n, p.ptr(n))
    return a


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def derivs(opts: NdOptions, p: Particles, phases: int = NDO_ALL):
    """One `derivs` on the host arrays of `p` (in place).  Returns (scalars dict, ms[5])."""
    L = lib()
    a = _arrays(p)
    nt = C.c_int(p.ntotal)
    s = NdScalars()
    ms = (C.c_double * 5)()
    e = L.ndo_derivs(C.byref(opts), p.ndim, C.byref(a), p.npart, C.byref(nt), p.idim, phases, C.byref(s), ms)
    p.ntotal = nt.value
    if e != 0:
        raise OracleError(e, L.ndo_last_error().decode())
    return s.as_dict(), list(ms)


def step(opts: NdOptions, p: Particles, dt: float, C_cour: float = 0.3, C_force: float = 0.25, dtfixed: bool = False):
    """One leapfrog `step` (src/stepND_leapfrog_mhd.f90:39-300) on the host arrays of `p`, which must hold the rates of a previous
    derivs.  Returns (dt for the next step, scalars of the inner derivs)."""
    L = lib()
    a = _arrays(p)
    nt = C.c_int(p.ntotal)
    s = NdScalars()
    d = C.c_double(dt)
    e = L.ndo_step(C.byref(opts), p.ndim, C.byref(a), p.npart, C.byref(nt), p.idim, C.byref(d), C_cour, C_force, int(dtfixed), C.byref(s))
    p.ntotal = nt.value
    if e != 0:
        raise OracleError(e, L.ndo_last_error().decode())
    return d.value, s.as_dict()


def evwrite(opts: NdOptions, p: Particles) -> dict:
    """The sums of `evwrite` (src/evwrite_mhd.f90:124-284) over the host arrays of `p` (after a derivs or a step)."""
    L = lib()
    a = _arrays(p)
    ev = NdEvwrite()
    e = L.ndo_evwrite(C.byref(opts), p.ndim, C.byref(a), p.npart, C.byref(ev))
    if e != 0:
        raise OracleError(e, "evwrite")
    return ev.as_dict()


def get_curl(opts: NdOptions, p: Particles, Bvec: np.ndarray, icurltype: int = 1, want_gradB: bool = False, hhmax: float | None = None):
    """`get_curl` (src/get_curl.f90:64-287) of Bvec[(idim,3)] on the arrays of `p` as a previous derivs left them (rho, hh, gradh, ghosts).
    Returns (curlB[(idim,3)], gradB[(idim,3,3)] or None); gradB[i, k, l] = gradB(l,k,i) of the reference = d B_k / d x_l."""
    L = lib()
    o2 = NdOptions.from_buffer_copy(opts)   # bound:hhmax as set_ghost_particles left it (the cell size of the re-link)
    o2.hhmax = float(np.max(p.hh[: p.npart])) if hhmax is None else hhmax
    opts = o2
    a = _arrays(p)
    Bvec = np.ascontiguousarray(Bvec, dtype=np.float64)
    assert Bvec.shape == (p.idim, 3)
    curlB = np.zeros((p.idim, 3))
    gradB = np.zeros((p.idim, 3, 3)) if want_gradB else None
    e = L.ndo_get_curl(C.byref(opts), p.ndim, C.byref(a), p.npart, p.ntotal, p.idim, icurltype, Bvec.ctypes.data_as(_DP), curlB.ctypes.data_as(_DP),
                       gradB.ctypes.data_as(_DP) if want_gradB else None)
    if e != 0:
        raise OracleError(e, L.ndo_last_error().decode())
    return curlB, gradB


def kernel_tables(ikernel: int, ikerneldrag: int, ndim: int):
    L = lib()
    n = 4001
    w, gw, ggw, wd = (np.zeros(n) for _ in range(4))
    r2, dq2 = C.c_double(), C.c_double()
    e = L.ndo_kernel_tables(ikernel, ikerneldrag, ndim, w.ctypes.data_as(_DP), gw.ctypes.data_as(_DP), ggw.ctypes.data_as(_DP),
                            wd.ctypes.data_as(_DP), C.byref(r2), C.byref(dq2))
    if e:
        raise OracleError(e, "kernel tables")
    return w, gw, ggw, wd, r2.value, dq2.value


def interpolate(ikernel: int, ndim: int, q2: float):
    L = lib()
    w, gw, ggw = C.c_double(), C.c_double(), C.c_double()
    e = L.ndo_interpolate(ikernel, ndim, q2, C.byref(w), C.byref(gw), C.byref(ggw))
    if e:
        raise OracleError(e, "interpolate")
    return w.value, gw.value, ggw.value


def ran1(iseed: int):
    s = C.c_int(iseed)
    v = lib().ndo_ran1(C.byref(s))
    return v, s.value


def bruteforce_pairs(p: Particles, radkern2: float = 4.0):
    """All unordered pairs (i<j, 1-based) within range of either particle, at least one real."""
    L = lib()
    x = np.ascontiguousarray(p.x[: p.ntotal])
    h = np.ascontiguousarray(p.hh[: p.ntotal])
    cap = max(1024, 200 * p.ntotal)
    pi = np.zeros(cap, np.int32)
    pj = np.zeros(cap, np.int32)
    n = L.ndo_bruteforce_pairs(p.ndim, x.ctypes.data_as(_DP), h.ctypes.data_as(_DP), p.npart, p.ntotal, radkern2,
                               pi.ctypes.data_as(_IP), pj.ctypes.data_as(_IP), cap)
    assert n <= cap, "pair buffer too small"
    return pi[:n].copy(), pj[:n].copy()


def linklist_pairs(opts: NdOptions, p: Particles, hhmax: float | None = None):
    """Pairs visited by the reference's rates loop (src/ratesND_mhd.f90:304-467).  Overwrites rates outputs of `p`.

    `hhmax` is bound:hhmax as left by set_ghost_particles (the cell size, src/linkND.f90:72); default max h of the real rows."""
    L = lib()
    o2 = NdOptions.from_buffer_copy(opts)
    o2.hhmax = float(np.max(p.hh[: p.npart])) if hhmax is None else hhmax
    opts = o2
    a = _arrays(p)
    cap = max(1024, 200 * p.ntotal)
    pi = np.zeros(cap, np.int32)
    pj = np.zeros(cap, np.int32)
    n = L.ndo_linklist_pairs(C.byref(opts), p.ndim, C.byref(a), p.npart, p.ntotal, p.idim, pi.ctypes.data_as(_IP),
                             pj.ctypes.data_as(_IP), cap)
    if n < 0:
        raise OracleError(-1, L.ndo_last_error().decode())
    assert n <= cap
    return pi[:n].copy(), pj[:n].copy()
