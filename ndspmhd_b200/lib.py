"""Host-side mirror of the reference's hot-path call sites over the C-ABI (include/ndspmhd_b200.h).

The reference's `derivs` (src/derivs.f90:74-156) calls, in order, `set_linklist`, `iterate_density`,
`conservative2primitive`, `get_rates`; those are the method names here, with the same meaning and the same error
behaviour (the reference prints and calls `quit`; here an `NdError` carries the C-ABI code and message).

There is NO CPU fallback: without the built CUDA library (ndspmhd_b200/libndspmhd_b200.so) or without a GPU every
compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi
from .abi import NdArrays, NdEvwrite, NdOptions, NdScalars, NdStateOut, NdStepOpts, Particles

_HERE = os.path.dirname(os.path.abspath(__file__))
# NDSPMHD_B200_LIB: development override used by tools/ to time differently tuned builds of the same sources
LIB_PATH = os.environ.get("NDSPMHD_B200_LIB") or os.path.join(_HERE, "libndspmhd_b200.so")
_LIB = None

_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)

# every symbol include/ndspmhd_b200.h declares
EXPORTS = [
    "ndspmhd_b200_default_options", "ndspmhd_b200_version", "ndspmhd_b200_device_count", "ndspmhd_b200_create",
    "ndspmhd_b200_set_options", "ndspmhd_b200_destroy", "ndspmhd_b200_last_error", "ndspmhd_b200_get_kernel_tables",
    "ndspmhd_b200_upload", "ndspmhd_b200_update_ghosts", "ndspmhd_b200_link", "ndspmhd_b200_iterate_density",
    "ndspmhd_b200_cons2prim", "ndspmhd_b200_get_rates", "ndspmhd_b200_derivs", "ndspmhd_b200_download",
    "ndspmhd_b200_host_alloc", "ndspmhd_b200_host_free", "ndspmhd_b200_last_timings", "ndspmhd_b200_launch_count",
    "ndspmhd_b200_stream", "ndspmhd_b200_rates_pairs", "ndspmhd_b200_rewind", "ndspmhd_b200_set_comm", "ndspmhd_b200_row_counts", "ndspmhd_b200_selftest_math", "ndspmhd_b200_derivs_host",
    "ndspmhd_b200_step", "ndspmhd_b200_download_state", "ndspmhd_b200_evwrite", "ndspmhd_b200_get_curl",
    "ndspmhd_b200_nccl_unique_id", "ndspmhd_b200_set_comm_nccl", "ndspmhd_b200_comm_stats",
    "ndspmhd_b200_set_row_ids", "ndspmhd_b200_get_row_ids", "ndspmhd_b200_migration_stats",
]


class NdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ndspmhd_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


def load():
    """dlopen the CUDA library.  Raises if it has not been built: the product never falls back to a CPU path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  ndspmhd_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.ndspmhd_b200_default_options.argtypes = [C.POINTER(NdOptions)]
    L.ndspmhd_b200_create.argtypes = [C.POINTER(NdOptions), C.c_int, C.c_int, C.POINTER(vp)]
    L.ndspmhd_b200_set_options.argtypes = [vp, C.POINTER(NdOptions)]
    L.ndspmhd_b200_destroy.argtypes = [vp]
    L.ndspmhd_b200_last_error.argtypes = [vp]
    L.ndspmhd_b200_last_error.restype = C.c_char_p
    L.ndspmhd_b200_get_kernel_tables.argtypes = [vp, _DP, _DP, _DP, _DP, _DP, _DP]
    L.ndspmhd_b200_upload.argtypes = [vp, C.POINTER(NdArrays), C.c_int, C.c_int, C.c_int]
    L.ndspmhd_b200_update_ghosts.argtypes = [vp, C.POINTER(NdArrays), C.c_int, C.c_int, C.c_double]
    L.ndspmhd_b200_link.argtypes = [vp]
    L.ndspmhd_b200_iterate_density.argtypes = [vp, C.c_int, C.POINTER(NdScalars)]
    L.ndspmhd_b200_cons2prim.argtypes = [vp]
    L.ndspmhd_b200_get_rates.argtypes = [vp, C.POINTER(NdScalars)]
    L.ndspmhd_b200_derivs.argtypes = [vp, C.POINTER(NdScalars)]
    L.ndspmhd_b200_download.argtypes = [vp, C.POINTER(NdArrays), C.c_uint, C.c_int]
    L.ndspmhd_b200_derivs_host.argtypes = [vp, C.POINTER(NdArrays), C.c_int, C.c_int, C.c_int, C.c_uint, C.POINTER(NdScalars)]
    L.ndspmhd_b200_host_alloc.argtypes = [C.c_size_t]
    L.ndspmhd_b200_host_alloc.restype = vp
    L.ndspmhd_b200_host_free.argtypes = [vp]
    L.ndspmhd_b200_host_free.restype = None
    L.ndspmhd_b200_last_timings.argtypes = [vp, _DP]
    L.ndspmhd_b200_launch_count.argtypes = [vp]
    L.ndspmhd_b200_launch_count.restype = C.c_longlong
    L.ndspmhd_b200_rewind.argtypes = [vp]
    L.ndspmhd_b200_stream.argtypes = [vp]
    L.ndspmhd_b200_stream.restype = vp
    L.ndspmhd_b200_rates_pairs.argtypes = [vp, _IP, _IP, C.c_longlong, C.POINTER(C.c_longlong)]
    L.ndspmhd_b200_step.argtypes = [vp, C.POINTER(NdStepOpts), _DP, C.POINTER(NdScalars)]
    L.ndspmhd_b200_download_state.argtypes = [vp, C.POINTER(NdStateOut), C.c_int]
    L.ndspmhd_b200_evwrite.argtypes = [vp, C.POINTER(NdEvwrite)]
    L.ndspmhd_b200_get_curl.argtypes = [vp, C.c_int, _DP, _DP, _DP, C.c_int]
    L.ndspmhd_b200_nccl_unique_id.argtypes = [C.c_char_p]
    L.ndspmhd_b200_set_comm_nccl.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_longlong]
    L.ndspmhd_b200_comm_stats.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.ndspmhd_b200_set_row_ids.argtypes = [vp, C.POINTER(C.c_longlong), C.c_int]
    L.ndspmhd_b200_get_row_ids.argtypes = [vp, C.POINTER(C.c_longlong), C.c_int]
    L.ndspmhd_b200_migration_stats.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    _LIB = L
    return L


def device_count() -> int:
    return int(load().ndspmhd_b200_device_count())


def pinned_particles(ndim: int, npart: int, idim: int) -> Particles:
    """A Particles container whose arrays live in page-locked host memory (for the end-to-end bench)."""
    L = load()
    p = Particles(ndim, npart, idim)
    keep = []
    for name, arr in list(p.arrays.items()):
        nbytes = max(arr.nbytes, 8)
        ptr = L.ndspmhd_b200_host_alloc(nbytes)
        if not ptr:
            raise MemoryError("cudaMallocHost failed")
        buf = (C.c_char * nbytes).from_address(ptr)
        new = np.frombuffer(buf, dtype=arr.dtype, count=arr.size).reshape(arr.shape)
        new[...] = arr
        p.arrays[name] = new
        keep.append(ptr)
    p.__dict__["_pinned"] = keep
    return p


def free_pinned(p: Particles) -> None:
    L = load()
    for ptr in p.__dict__.get("_pinned", []):
        L.ndspmhd_b200_host_free(ptr)
    p.__dict__["_pinned"] = []


def arrays_struct(p: Particles) -> NdArrays:
    """Hand over the module arrays exactly as an ISO_C_BINDING shim would."""
    a = NdArrays()
    a.x, a.vel, a.pmass = p.ptr("x"), p.ptr("vel"), p.ptr("pmass")
    # the reference iterates hh in place; a caller that wants to keep its guess (bench.py repeats the same step) may park it
    # in a separate `hh_guess` array
    a.hh_in = p.ptr("hh_guess") if "hh_guess" in p.arrays else p.ptr("hh")
    a.itype, a.ireal = p.ptr("itype"), p.ptr("ireal")
    a.en, a.Bevol, a.alpha, a.psi, a.rho_in = p.ptr("en"), p.ptr("Bevol"), p.ptr("alpha"), p.ptr("psi"), p.ptr("rho")
    for n in ("hh", "rho", "gradh", "drhodt", "dhdt", "numneigh", "rhoalt", "gradhn", "gradsoft", "gradgradh", "dens", "uu", "pr",
              "spsound", "Bfield", "force", "dudt", "dendt", "dBevoldt", "daldt", "dpsidt", "gradpsi", "divB", "curlB", "graddivv",
              "del2u"):
        setattr(a, n, p.ptr(n))
    a.x_out, a.vel_out, a.ireal_out, a.itype_out = p.ptr("x"), p.ptr("vel"), p.ptr("ireal"), p.ptr("itype")
    # one-fluid dust: part:dustfrac is read on entry (density sums) and rewritten by conservative2primitive, like the module array
    a.dustevol, a.dustfrac_in, a.deltav = p.ptr("dustevol"), p.ptr("dustfrac"), p.ptr("deltav")
    for n in ("dustfrac", "rhogas", "rhodust", "ddustevoldt", "ddeltavdt"):
        setattr(a, n, p.ptr(n))
    a.alpha_out = p.ptr("alpha")   # iavlim(3) = 2: conservative2primitive rewrites alpha(3,:) in place, like the module array
    return a


class Hotpath:
    """One context per host thread and GPU (the reference is single-threaded with module globals)."""

    def __init__(self, opts: NdOptions, ndim: int, device: int = 0):
        self.L = load()
        self.opts = opts
        self.ndim = ndim
        self.ctx = C.c_void_p()
        e = self.L.ndspmhd_b200_create(C.byref(opts), ndim, device, C.byref(self.ctx))
        if e:
            msg = self.L.ndspmhd_b200_last_error(self.ctx).decode() if self.ctx else "create failed"
            self.L.ndspmhd_b200_destroy(self.ctx)
            self.ctx = C.c_void_p()
            raise NdError(e, msg)

    # ---- plumbing ----
    def _chk(self, e: int):
        if e:
            raise NdError(e, self.L.ndspmhd_b200_last_error(self.ctx).decode())

    def close(self):
        if self.ctx:
            self.L.ndspmhd_b200_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_options(self, opts: NdOptions):
        self._chk(self.L.ndspmhd_b200_set_options(self.ctx, C.byref(opts)))
        self.opts = opts

    def upload(self, p: Particles):
        a = arrays_struct(p)
        self._chk(self.L.ndspmhd_b200_upload(self.ctx, C.byref(a), p.npart, p.ntotal, p.idim))

    def update_ghosts(self, p: Particles, hhmax: float):
        a = arrays_struct(p)
        self._chk(self.L.ndspmhd_b200_update_ghosts(self.ctx, C.byref(a), p.ntotal, p.idim, hhmax))

    def download(self, p: Particles, mask: int = abi.DL_ALL):
        a = arrays_struct(p)
        self._chk(self.L.ndspmhd_b200_download(self.ctx, C.byref(a), mask, p.idim))

    # ---- the reference's call sites ----
    def set_linklist(self) -> None:
        """`call set_linklist` (src/derivs.f90:82)."""
        self._chk(self.L.ndspmhd_b200_link(self.ctx))

    def iterate_density(self, resume: int = 0) -> dict:
        """`call iterate_density` (src/derivs.f90:92).  Raises NdError(ND_NEED_RELINK) in host-ghost mode when h outgrew hhmax."""
        s = NdScalars()
        self._chk(self.L.ndspmhd_b200_iterate_density(self.ctx, resume, C.byref(s)))
        return s.as_dict()

    def conservative2primitive(self) -> None:
        """`call conservative2primitive` (src/derivs.f90:98)."""
        self._chk(self.L.ndspmhd_b200_cons2prim(self.ctx))

    def get_rates(self) -> dict:
        """`call get_rates` (src/derivs.f90:156)."""
        s = NdScalars()
        self._chk(self.L.ndspmhd_b200_get_rates(self.ctx, C.byref(s)))
        return s.as_dict()

    def derivs(self) -> dict:
        """One `derivs` on the resident state: link + density iteration + cons2prim + rates."""
        s = NdScalars()
        self._chk(self.L.ndspmhd_b200_derivs(self.ctx, C.byref(s)))
        return s.as_dict()

    def step(self, dt: float, C_cour: float = 0.3, C_force: float = 0.25, dtfixed: bool = False):
        """`call step` (src/stepND_leapfrog_mhd.f90:39) on the resident state: predictor, derivs, corrector, periodic wrap.
        Returns (dt for the next step, scalars of the inner derivs)."""
        so = NdStepOpts(C_cour, C_force, int(dtfixed), 0)
        d = C.c_double(dt)
        s = NdScalars()
        self._chk(self.L.ndspmhd_b200_step(self.ctx, C.byref(so), C.byref(d), C.byref(s)))
        return d.value, s.as_dict()

    def download_state(self, p: Particles):
        """x, vel, hh, en, Bevol, alpha, psi, rho (+ dustevol, deltav) of rows [0,npart) into the host arrays of `p`."""
        st = NdStateOut()
        for n in ("x", "vel", "hh", "en", "Bevol", "alpha", "psi", "rho", "dustevol", "deltav"):
            setattr(st, n, p.ptr(n))
        self._chk(self.L.ndspmhd_b200_download_state(self.ctx, C.byref(st), p.idim))

    def set_row_ids(self, ids: np.ndarray) -> None:
        """Global ids of the uploaded rows (they travel with a row when it migrates to another slab in `step`)."""
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        self._chk(self.L.ndspmhd_b200_set_row_ids(self.ctx, ids.ctypes.data_as(C.POINTER(C.c_longlong)), ids.size))

    def get_row_ids(self, cap: int) -> np.ndarray:
        ids = np.zeros(cap, np.int64)
        self._chk(self.L.ndspmhd_b200_get_row_ids(self.ctx, ids.ctypes.data_as(C.POINTER(C.c_longlong)), cap))
        return ids

    def migration_stats(self):
        """(rows that left this rank, rows that arrived, payload bytes sent) since the context was created."""
        a, b, c_ = C.c_longlong(), C.c_longlong(), C.c_longlong()
        self.L.ndspmhd_b200_migration_stats(self.ctx, C.byref(a), C.byref(b), C.byref(c_))
        return a.value, b.value, c_.value

    def evwrite(self) -> dict:
        """The sums of `evwrite` (src/evwrite_mhd.f90:27) over the resident state, as device reductions."""
        ev = NdEvwrite()
        self._chk(self.L.ndspmhd_b200_evwrite(self.ctx, C.byref(ev)))
        return ev.as_dict()

    def get_curl(self, Bvec: np.ndarray, icurltype: int = 1, want_gradB: bool = False):
        """`get_curl` (src/get_curl.f90:64) of Bvec[(idim,3)] on the resident state after iterate_density.  Returns (curlB[(idim,3)],
        gradB[(idim,3,3)] or None) with rows [0,npart) filled; gradB[i, k, l] = d Bvec_k / d x_l."""
        Bvec = np.ascontiguousarray(Bvec, dtype=np.float64)
        idim = Bvec.shape[0]
        curlB = np.zeros((idim, 3))
        gradB = np.zeros((idim, 3, 3)) if want_gradB else None
        self._chk(self.L.ndspmhd_b200_get_curl(self.ctx, icurltype, Bvec.ctypes.data_as(_DP), curlB.ctypes.data_as(_DP),
                                               gradB.ctypes.data_as(_DP) if want_gradB else None, idim))
        return curlB, gradB

    def selftest_math(self, x: np.ndarray):
        """sqrt_nr / rsqrt_nr of the pair kernels evaluated on the device for the given arguments."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        s, r = np.empty_like(x), np.empty_like(x)
        self.L.ndspmhd_b200_selftest_math.argtypes = [C.c_void_p, _DP, _DP, _DP, C.c_int]
        self._chk(self.L.ndspmhd_b200_selftest_math(self.ctx, x.ctypes.data_as(_DP), s.ctypes.data_as(_DP), r.ctypes.data_as(_DP), x.size))
        return s, r

    def derivs_host(self, p: Particles, mask: int = abi.DL_ALL, skip=()) -> dict:
        """upload + derivs + download in one call with the copies overlapped with the kernels (host arrays in, host arrays out).
        `skip`: output arrays handed over as NULL pointers -- the library leaves them on the device (include/ndspmhd_b200.h, nd_arrays)."""
        a = arrays_struct(p)
        for nm in skip:
            setattr(a, nm, None)
        s = NdScalars()
        self._chk(self.L.ndspmhd_b200_derivs_host(self.ctx, C.byref(a), p.npart, p.ntotal, p.idim, mask, C.byref(s)))
        p.ntotal = s.ntotal if not getattr(self, "_comm", None) else p.ntotal
        return s.as_dict()

    def rewind(self) -> None:
        """Restore the smoothing-length guess of the last upload (bench hook: makes repeated derivs() do identical work)."""
        self._chk(self.L.ndspmhd_b200_rewind(self.ctx))

    # ---- diagnostics ----
    def timings(self):
        ms = (C.c_double * 8)()
        self.L.ndspmhd_b200_last_timings(self.ctx, ms)
        # phases between events on the library's stream; "rates_pair" = list build + pair kernel, "rates_pair_kernel" = the kernel alone
        return dict(zip(["link", "density", "c2p_gather", "rates_pair", "rates_final", "rates_pair_kernel"], list(ms)[:6]))

    def launch_count(self) -> int:
        return int(self.L.ndspmhd_b200_launch_count(self.ctx))

    def stream(self) -> int:
        return int(self.L.ndspmhd_b200_stream(self.ctx) or 0)

    def kernel_tables(self):
        n = 4001
        w, gw, ggw, wd = (np.zeros(n) for _ in range(4))
        r2, dq2 = C.c_double(), C.c_double()
        self._chk(self.L.ndspmhd_b200_get_kernel_tables(self.ctx, w.ctypes.data_as(_DP), gw.ctypes.data_as(_DP), ggw.ctypes.data_as(_DP),
                                                        wd.ctypes.data_as(_DP), C.byref(r2), C.byref(dq2)))
        return w, gw, ggw, wd, r2.value, dq2.value

    def rates_pairs(self, cap: int):
        pi = np.zeros(cap, np.int32)
        pj = np.zeros(cap, np.int32)
        n = C.c_longlong()
        self._chk(self.L.ndspmhd_b200_rates_pairs(self.ctx, pi.ctypes.data_as(_IP), pj.ctypes.data_as(_IP), cap, C.byref(n)))
        if n.value > cap:
            raise NdError(abi.ND_ERR_NEIGHBOUR_OVERFLOW, f"pair buffer too small: {n.value} > {cap}")
        return pi[: n.value].copy(), pj[: n.value].copy()


def derivs_host(opts: NdOptions, p: Particles, device: int = 0, hot: Hotpath | None = None, pipelined: bool = False) -> dict:
    """The call a user of the reference makes: host arrays in, host arrays out (upload + derivs + download)."""
    own = hot is None
    if own:
        hot = Hotpath(opts, p.ndim, device)
    try:
        if pipelined:
            return hot.derivs_host(p, abi.DL_ALL)
        hot.upload(p)
        s = hot.derivs()
        p.ntotal = s["ntotal"]
        hot.download(p, abi.DL_ALL)
        return s
    finally:
        if own:
            hot.close()
