"""ctypes mirror of include/ndspmhd_b200.h (structs, enums) and the host-side particle container.

The container keeps every array in the reference's native layout: Fortran `x(ndim,idim)` is a C-order
numpy array of shape (idim, ndim) (src/allocateND.f90:317-405), so pointers can be handed to the C-ABI
exactly as an ISO_C_BINDING shim would hand over the module arrays.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

# error codes -----------------------------------------------------------------------------------------
ND_OK = 0
ND_ERR_INVALID_ARG = 1
ND_ERR_UNSUPPORTED_OPTION = 2
ND_ERR_H_NONPOSITIVE = 3
ND_ERR_RHO_NONPOSITIVE = 4
ND_ERR_DENSITY_NOT_CONVERGED = 5
ND_ERR_VSIG_DET = 6
ND_ERR_CUDA = 7
ND_ERR_NO_DEVICE = 8
ND_ERR_LINK = 9
ND_ERR_NEIGHBOUR_OVERFLOW = 10
ND_ERR_STATE = 11
ND_NEED_RELINK = 100

ITYPE_GAS, ITYPE_BND, ITYPE_DUST, ITYPE_GAS1, ITYPE_GAS2, ITYPE_BND2, ITYPE_BNDDUST = 0, 1, 2, 3, 4, 11, 12

DL_DENSITY, DL_PRIM, DL_RATES, DL_GHOSTS, DL_ALL = 1, 2, 4, 8, 15
DL_REAL_ROWS = 16   # modifier: rows [0,npart) only


class NdOptions(C.Structure):
    _fields_ = [
        ("iener", C.c_int), ("icty", C.c_int), ("iav", C.c_int), ("ikernav", C.c_int), ("ihvar", C.c_int), ("iprterm", C.c_int),
        ("imhd", C.c_int), ("imagforce", C.c_int), ("idivbzero", C.c_int), ("iresist", C.c_int),
        ("idust", C.c_int), ("idrag_nature", C.c_int),
        ("ixsph", C.c_int), ("igravity", C.c_int), ("iexternal_force", C.c_int),
        ("ikernel", C.c_int), ("ikernelalt", C.c_int),
        ("maxdensits", C.c_int),
        ("iavlim", C.c_int * 3),
        ("ibound", C.c_int * 3),
        ("usenumdens", C.c_int), ("ibiascorrection", C.c_int), ("onef_dust", C.c_int), ("use_smoothed_rhodust", C.c_int),
        ("islope_limiter", C.c_int), ("iuse_exact_derivs", C.c_int), ("iambipolar", C.c_int), ("ivisc", C.c_int),
        ("iquantum", C.c_int), ("ind_timesteps", C.c_int),
        ("nsubsteps_divB", C.c_int),
        ("device_ghosts", C.c_int), ("want_aux", C.c_int),
        ("idustevol", C.c_int),
        ("reserved_i", C.c_int * 5),
        ("hfact", C.c_double), ("psep", C.c_double), ("tolh", C.c_double),
        ("gamma", C.c_double), ("polyk", C.c_double),
        ("alphamin", C.c_double), ("alphaumin", C.c_double), ("alphaBmin", C.c_double), ("beta", C.c_double),
        ("avdecayconst", C.c_double), ("avfact", C.c_double),
        ("psidecayfact", C.c_double), ("etamhd", C.c_double), ("Kdrag", C.c_double), ("damp", C.c_double), ("pext", C.c_double),
        ("xmin", C.c_double * 3), ("xmax", C.c_double * 3),
        ("Bconst", C.c_double * 3),
        ("hhmax", C.c_double),
        ("reserved_d", C.c_double * 8),
    ]


_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)


class NdArrays(C.Structure):
    _fields_ = [
        ("x", _DP), ("vel", _DP), ("pmass", _DP), ("hh_in", _DP), ("itype", _IP), ("ireal", _IP),
        ("en", _DP), ("Bevol", _DP), ("alpha", _DP), ("psi", _DP), ("rho_in", _DP),
        ("hh", _DP), ("rho", _DP), ("gradh", _DP), ("drhodt", _DP), ("dhdt", _DP), ("numneigh", _IP),
        ("rhoalt", _DP), ("gradhn", _DP), ("gradsoft", _DP), ("gradgradh", _DP),
        ("dens", _DP), ("uu", _DP), ("pr", _DP), ("spsound", _DP), ("Bfield", _DP),
        ("force", _DP), ("dudt", _DP), ("dendt", _DP), ("dBevoldt", _DP), ("daldt", _DP), ("dpsidt", _DP),
        ("gradpsi", _DP), ("divB", _DP), ("curlB", _DP), ("graddivv", _DP), ("del2u", _DP),
        ("x_out", _DP), ("vel_out", _DP), ("ireal_out", _IP), ("itype_out", _IP),
        ("dustevol", _DP), ("dustfrac_in", _DP), ("deltav", _DP),
        ("dustfrac", _DP), ("rhogas", _DP), ("rhodust", _DP), ("ddustevoldt", _DP), ("ddeltavdt", _DP),
        ("alpha_out", _DP),
    ]


class NdStepOpts(C.Structure):
    """module timestep: C_cour, C_force, dtfixed (src/variablesND.f90:255-273; defaults src/defaults.f90: C_cour=0.3, C_force=0.25)."""
    _fields_ = [("C_cour", C.c_double), ("C_force", C.c_double), ("dtfixed", C.c_int), ("reserved", C.c_int)]


class NdStateOut(C.Structure):
    _fields_ = [(n, _DP) for n in ("x", "vel", "hh", "en", "Bevol", "alpha", "psi", "rho", "dustevol", "deltav")]


class NdEvwrite(C.Structure):
    """The columns of the reference's .ev line (src/evwrite_mhd.f90:296-320) plus the vector sums behind them."""
    _fields_ = ([(n, C.c_double) for n in ("ekin", "etherm", "emag", "epot", "etot", "momtot", "angtot", "rhomax", "rhomean", "rhomin",
                                            "emagp", "crosshel", "betamhdmin", "betamhdav", "divBav", "divBmax", "divBtot",
                                            "omegamhdav", "omegamhdmax", "fracdivBok", "fluxtotmag",
                                            "ekiny", "dmomtot", "totmassgas", "totmassdust")]
                + [(n, C.c_double * 3) for n in ("mom", "dmom", "ang", "fluxtot")] + [("reserved", C.c_double * 8)])

    def as_dict(self):
        out = {}
        for name, typ in self._fields_:
            if name == "reserved":
                continue
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


class NdScalars(C.Structure):
    _fields_ = [
        ("dtcourant", C.c_double), ("dtforce", C.c_double), ("dtav", C.c_double), ("dtdrag", C.c_double), ("dtvisc", C.c_double),
        ("vsig2max", C.c_double), ("vsigmax", C.c_double),
        ("stressmax", C.c_double), ("ts_min", C.c_double), ("h_on_csts_max", C.c_double), ("fhmax", C.c_double),
        ("hhmax", C.c_double), ("dxcell", C.c_double),
        ("fmean", C.c_double * 3),
        ("itsdensity", C.c_int), ("nneigh_min", C.c_int), ("nneigh_max", C.c_int), ("nclumped", C.c_int),
        ("ntotal", C.c_int), ("ncells", C.c_int), ("ncellsx", C.c_int * 3), ("nrelink", C.c_int),
        ("ncalctotal", C.c_longlong),
        ("lmax", C.c_int), ("list_overflows", C.c_int), ("rate_chunks", C.c_int),
        ("reserved_i", C.c_int * 5),
        ("npairs_rates", C.c_longlong), ("ntrips_rates", C.c_longlong),
    ]

    def as_dict(self):
        out = {}
        for name, typ in self._fields_:
            if name.startswith("reserved"):
                continue
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


def default_options(ndim: int = 3) -> NdOptions:
    """src/defaults.f90:47-118; avfact as src/initialiseND_mhd.f90:168-172 for gamma=5/3."""
    o = NdOptions()
    o.psep = 0.01
    o.gamma = 5.0 / 3.0
    o.iener = 2
    o.polyk = 1.0
    o.icty = 0
    o.maxdensits = 250
    o.iprterm = 0
    o.iav = 2
    o.alphamin, o.alphaumin, o.alphaBmin, o.beta = 0.1, 0.0, 1.0, 2.0
    o.iavlim[0], o.iavlim[1], o.iavlim[2] = 2, 1, 0
    o.avdecayconst = 0.1
    o.ikernav = 3
    o.ihvar = 2
    o.hfact = 1.2
    o.tolh = 1.0e-3
    o.imhd = 0
    o.imagforce = 2
    o.idivbzero = 0
    o.psidecayfact = 0.1
    o.iresist = 0
    o.etamhd = 0.0
    o.ixsph = 0
    o.igravity = 0
    o.damp = 0.0
    o.iexternal_force = 0
    o.ikernel = 0
    o.ikernelalt = 0
    o.usenumdens = 0
    o.use_smoothed_rhodust = 1
    o.iuse_exact_derivs = 0
    o.idust = 0
    o.idustevol = 0
    o.idrag_nature = 0
    o.Kdrag = 0.0
    o.ibiascorrection = 0
    o.iambipolar = 0
    o.iquantum = 0
    o.islope_limiter = -1
    o.ivisc = 0
    o.ind_timesteps = 0
    o.nsubsteps_divB = 0
    o.pext = 0.0
    set_gamma(o, o.gamma)
    for d in range(3):
        o.ibound[d] = 0
        o.xmin[d] = 0.0
        o.xmax[d] = 0.0
        o.Bconst[d] = 0.0
    o.device_ghosts = 0
    o.want_aux = 1
    return o


def set_gamma(o: NdOptions, gamma: float) -> None:
    """gamma plus the derived avfact (src/initialiseND_mhd.f90:168-172)."""
    import math

    o.gamma = gamma
    if abs(gamma - 1.0) > 1.0e-3:
        o.avfact = math.log(4.0) / (math.log((gamma + 1.0) / (gamma - 1.0)))
    else:
        o.avfact = 1.0


# array name -> (ncomp or 'ndim', dtype)
_ARRAY_SPEC = {
    "x": ("ndim", np.float64), "vel": (3, np.float64), "pmass": (1, np.float64), "hh": (1, np.float64),
    "itype": (1, np.int32), "ireal": (1, np.int32), "en": (1, np.float64), "Bevol": (3, np.float64),
    "alpha": (3, np.float64), "psi": (1, np.float64), "rho": (1, np.float64),
    "gradh": (1, np.float64), "gradhn": (1, np.float64), "gradsoft": (1, np.float64), "gradgradh": (1, np.float64),
    "rhoalt": (1, np.float64), "drhodt": (1, np.float64), "dhdt": (1, np.float64), "numneigh": (1, np.int32),
    "dens": (1, np.float64), "uu": (1, np.float64), "pr": (1, np.float64), "spsound": (1, np.float64),
    "Bfield": (3, np.float64), "sqrtg": (1, np.float64),
    "force": (3, np.float64), "dudt": (1, np.float64), "dendt": (1, np.float64), "dBevoldt": (3, np.float64),
    "daldt": (3, np.float64), "dpsidt": (1, np.float64), "gradpsi": (3, np.float64), "fmag": (3, np.float64),
    "divB": (1, np.float64), "curlB": (3, np.float64), "graddivv": (3, np.float64), "del2u": (1, np.float64),
    "xsphterm": (3, np.float64),
    # one-fluid dust (ndust = 1, src/variablesND.f90:172)
    "dustevol": (1, np.float64), "dustfrac": (1, np.float64), "deltav": (3, np.float64), "rhogas": (1, np.float64),
    "rhodust": (1, np.float64), "ddustevoldt": (1, np.float64), "ddeltavdt": (3, np.float64),
}


@dataclass
class Particles:
    """Host copy of the reference's module arrays `part`, `rates`, `hterms`, `derivB`, `bound:ireal`."""

    ndim: int
    npart: int
    idim: int
    ntotal: int = 0
    arrays: dict = field(default_factory=dict)

    def __post_init__(self):
        if self.ntotal == 0:
            self.ntotal = self.npart
        for name, (nc, dt) in _ARRAY_SPEC.items():
            ncomp = self.ndim if nc == "ndim" else nc
            shape = (self.idim,) if (ncomp == 1 and nc != "ndim") else (self.idim, ncomp)
            self.arrays[name] = np.zeros(shape, dtype=dt)
        self.arrays["sqrtg"][:] = 1.0

    def __getattr__(self, name):
        arrs = self.__dict__.get("arrays")
        if arrs is not None and name in arrs:
            return arrs[name]
        raise AttributeError(name)

    def copy(self) -> "Particles":
        p = Particles(self.ndim, self.npart, self.idim, self.ntotal)
        for k, v in self.arrays.items():
            p.arrays[k][...] = v
        return p

    def ptr(self, name):
        a = self.arrays[name]
        if a.dtype == np.int32:
            return a.ctypes.data_as(_IP)
        return a.ctypes.data_as(_DP)
