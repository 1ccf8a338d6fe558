"""Synthetic particle sets for the BASELINE.json configs (host-side, numpy).

These restate the reference's initial-condition generators far enough to give the hot path the same kind of
input: `set_uniform_cartesian` cubic lattice (src/set_uniform_distributionND.f90:434-508), close-packed lattice
(:122-257), `perturb` (:513-529), and the field set-ups of src/setup_shock1D_mhd.f90, src/setup_orszagtang2D_mhd.f90,
src/setup_turbulenceND_mhd.f90, src/setup_dustybox.f90.  After the reference's `primitive2conservative`
(src/conservative2primitive.f90:478-560): h = hfact (m/rho)^(1/ndim), en/Bevol from the primitive variables,
alpha = (alphamin, alphaumin, alphaBmin) (src/initialiseND_mhd.f90:243-245).

The random displacement uses numpy's PCG64 with a fixed seed (the reference's `ran1` is sequential and too slow
in Python for 16M particles; the oracle restates it for cross-checks).  Everything is deterministic.
"""
from __future__ import annotations

import math

import numpy as np

from .abi import ITYPE_BND, ITYPE_DUST, ITYPE_GAS, NdOptions, Particles, default_options, set_gamma

PI = 3.1415926536  # src/variablesND.f90 setup_params:pi


def ghost_capacity(ndim: int, npart: int, opts: NdOptions, hmax: float, slack: float = 1.6) -> int:
    """Upper bound on ntotal for periodic/reflecting ghosts (rows the caller must allocate)."""
    frac = 1.0
    for d in range(ndim):
        if opts.ibound[d] >= 2:
            L = opts.xmax[d] - opts.xmin[d]
            frac *= (L + 2.0 * slack * 2.0 * hmax) / L
    return int(npart * frac) + 1024


def _finish(p: Particles, opts: NdOptions, dens: np.ndarray, uu: np.ndarray, B: np.ndarray | None):
    """primitive2conservative: rho, h, en, Bevol, alpha, psi (src/conservative2primitive.f90:505-520, :600-760)."""
    n = p.npart
    p.rho[:n] = dens
    p.dens[:n] = dens
    p.hh[:n] = opts.hfact * (p.pmass[:n] / dens) ** (1.0 / p.ndim)
    p.uu[:n] = uu
    if B is not None:
        p.Bfield[:n] = B
        if opts.imhd >= 11:
            p.Bevol[:n] = B
        elif opts.imhd >= 1:
            p.Bevol[:n] = B / dens[:, None]
    if opts.iener == 3:
        v2 = (p.vel[:n] ** 2).sum(axis=1)
        b2 = (p.Bfield[:n] ** 2).sum(axis=1) / dens
        p.en[:n] = uu + 0.5 * v2 + 0.5 * b2
    else:
        p.en[:n] = uu
    p.alpha[:n, 0] = opts.alphamin
    p.alpha[:n, 1] = opts.alphaumin
    p.alpha[:n, 2] = opts.alphaBmin
    p.psi[:n] = 0.0
    p.ntotal = n


def cubic_lattice(xmin, xmax, psep):
    """set_uniform_cartesian(1,...,fill=.true.): j (y) fastest, then i (x), then k (z) (:491-506)."""
    ndim = len(xmin)
    npx = [max(1, int(round((xmax[d] - xmin[d]) / psep))) for d in range(ndim)]
    delta = [(xmax[d] - xmin[d]) / npx[d] for d in range(ndim)]
    ax = [xmin[d] + np.arange(npx[d]) * delta[d] + 0.5 * delta[d] for d in range(ndim)]
    if ndim == 1:
        x = ax[0][:, None]
    elif ndim == 2:
        X, Y = np.meshgrid(ax[0], ax[1], indexing="ij")  # y fastest
        x = np.stack([X.ravel(), Y.ravel()], axis=1)
    else:
        Z, X, Y = np.meshgrid(ax[2], ax[0], ax[1], indexing="ij")  # z slowest, x, y fastest
        x = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    return np.ascontiguousarray(x), npx


def closepacked_lattice(xmin, xmax, psep, periodic=True):
    """set_uniform_cartesian(2,...): hexagonal close packing stretched to fill the box (:122-257)."""
    ndim = len(xmin)
    assert ndim >= 2
    deltax = psep
    deltay = 0.5 * math.sqrt(3.0) * psep
    deltaz = math.sqrt(6.0) / 3.0 * psep
    fac = 1.0 - np.finfo(float).eps
    npartx = int(fac * (xmax[0] - xmin[0]) / deltax) + 1
    nparty = int(fac * (xmax[1] - xmin[1]) / deltay) + 1
    npartz = int(fac * (xmax[2] - xmin[2]) / deltaz) + 1 if ndim >= 3 else 1
    if periodic:
        nparty = 2 * (nparty // 2)
        if ndim == 3:
            npartz = 3 * (npartz // 3)
    deltax = (xmax[0] - xmin[0]) / float(npartx)
    deltay = (xmax[1] - xmin[1]) / float(nparty)
    if ndim >= 3:
        deltaz = (xmax[2] - xmin[2]) / float(npartz)
    i = np.arange(1, npartx + 1)
    out = []
    for k in range(1, npartz + 1):
        for j in range(1, nparty + 1):
            ystart = deltay / 6.0
            zstart = 0.5 * deltaz
            xstart = 0.25 * deltax
            if k % 3 == 0:
                ystart += 2.0 / 3.0 * deltay
                if j % 2 == 0:
                    xstart += 0.5 * deltax
            elif k % 3 == 2:
                ystart += 1.0 / 3.0 * deltay
                if j % 2 == 1:
                    xstart += 0.5 * deltax
            elif j % 2 == 0:
                xstart += 0.5 * deltax
            row = np.empty((npartx, ndim))
            row[:, 0] = xmin[0] + (i - 1) * deltax + xstart
            row[:, 1] = xmin[1] + (j - 1) * deltay + ystart
            if ndim >= 3:
                row[:, 2] = xmin[2] + (k - 1) * deltaz + zstart
            out.append(row)
    return np.ascontiguousarray(np.concatenate(out, axis=0)), (npartx, nparty, npartz)


def perturb(x, psep, amplitude, seed=268):
    """:513-529: x += perturb*psep*(ran-0.5) per component."""
    if amplitude <= 0:
        return x
    rng = np.random.default_rng(seed)
    return x + amplitude * psep * (rng.random(x.shape) - 0.5)


def wrap_periodic(x, opts: NdOptions):
    """src/boundaryND.f90:65-93 periodic wrap."""
    for d in range(x.shape[1]):
        if opts.ibound[d] == 3:
            lo, hi = opts.xmin[d], opts.xmax[d]
            over = x[:, d] > hi
            x[over, d] = lo + x[over, d] - hi
            under = x[:, d] < lo
            x[under, d] = hi - (lo - x[under, d])
    return x


def _alloc(ndim, x, opts, hmax_guess, extra=0):
    n = x.shape[0]
    idim = ghost_capacity(ndim, n, opts, hmax_guess) + extra
    p = Particles(ndim, n, idim)
    p.x[:n] = x
    return p


# ---------------------------------------------------------------------------------------------------------
# C2 / C3 / C5: Orszag-Tang vortex, 2D or 3D thin slab / cube (src/setup_orszagtang2D_mhd.f90)
# ---------------------------------------------------------------------------------------------------------
def orszag_tang(ndim=2, nx=64, lattice="cubic", zfrac=0.125, perturb_amp=0.0, imhd=11, idivbzero=2, iener=2,
                evolved=True, seed=268, cube=False, slab=None, weak=False):
    """Periodic box [-0.5,0.5]^2 (x [-zfrac/2, zfrac/2] in 3D, or a unit cube with cube=True).

    slab=(rank, nranks): build only the rows of one x-slab of the same global particle set (multi-GPU runs); the return
    value is then (options, particles, info) with info = {edges, nglobal, rows (global indices of the local rows)}.

    weak=True (with slab): the box is `nranks` periods long in x -- [-0.5, -0.5 + nranks] -- and every rank makes only its own period (the
    fields have period 1 in x; lattice points sit at cell centres and the perturbation is < psep/2, so no particle leaves its period):
    per-GPU work stays fixed as ranks are added.

    evolved=True puts non-trivial psi / alpha / energy perturbations on the particles so every term of the
    rates is exercised (the t=0 state has psi=0 and uniform alpha, u).
    """
    o = default_options(ndim)
    o.imhd, o.idivbzero, o.iener = imhd, idivbzero, iener
    psep = 1.0 / nx
    o.psep = psep
    for d in range(ndim):
        o.ibound[d] = 3
    xmin = [-0.5, -0.5] + ([-0.5 * (1.0 if cube else zfrac)] if ndim == 3 else [])
    xmax = [-v for v in xmin]
    for d in range(ndim):
        o.xmin[d], o.xmax[d] = xmin[d], xmax[d]
    if lattice == "cubic":
        x, _ = cubic_lattice(xmin, xmax, psep)
    else:
        x, _ = closepacked_lattice(xmin, xmax, psep)
    if weak and slab is not None:
        rank, nranks = slab
        x = perturb(x, psep, min(perturb_amp, 0.9), seed + 1000 * rank)       # |dx| < psep/2: stays inside its lattice cell, hence its period
        x[:, 0] += float(rank)
        o.xmax[0] = xmax[0] = xmin[0] + float(nranks)
        nglobal = x.shape[0] * nranks
        info = {"edges": xmin[0] + np.arange(nranks + 1, dtype=np.float64), "nglobal": nglobal, "rows": np.arange(x.shape[0]) + rank * x.shape[0]}
        slab_done = True
    else:
        x = wrap_periodic(perturb(x, psep, perturb_amp, seed), o)
        nglobal = x.shape[0]
        info = None
        slab_done = False
    if slab is not None and not slab_done:
        from .slab import owner_of, slab_edges

        rank, nranks = slab
        edges = slab_edges(x[:, 0], nranks, xmin[0], xmax[0])
        rows = np.nonzero(owner_of(x[:, 0], edges) == rank)[0]
        x = np.ascontiguousarray(x[rows])
        info = {"edges": edges, "nglobal": nglobal, "rows": rows}
    n = x.shape[0]
    const = 4.0 * PI
    betazero, machzero, vzero = 10.0 / 3.0, 1.0, 1.0
    bzero = 1.0 / math.sqrt(const)
    przero = 0.5 * bzero**2 * betazero
    denszero = o.gamma * przero * machzero
    uuzero = przero / ((o.gamma - 1.0) * denszero)
    vol = np.prod([xmax[d] - xmin[d] for d in range(ndim)])
    massp = denszero * vol / nglobal
    hguess = o.hfact * (massp / denszero) ** (1.0 / ndim)
    p = _alloc(ndim, x, o, hguess, extra=(n // 4 + 4096) if slab is not None else 0)
    p.pmass[:n] = massp
    p.vel[:n, 0] = -vzero * np.sin(2.0 * PI * (x[:, 1] - xmin[1]))
    p.vel[:n, 1] = vzero * np.sin(2.0 * PI * (x[:, 0] - xmin[0]))
    B = np.zeros((n, 3))
    B[:, 0] = -bzero * np.sin(2.0 * PI * (x[:, 1] - xmin[1]))
    B[:, 1] = bzero * np.sin(4.0 * PI * (x[:, 0] - xmin[0]))
    dens = np.full(n, denszero)
    uu = np.full(n, uuzero)
    if ndim == 3 and evolved:
        p.vel[:n, 2] = 0.1 * vzero * np.sin(2.0 * PI * (x[:, 0] - xmin[0])) * np.cos(2.0 * PI * (x[:, 1] - xmin[1]))
        B[:, 2] = 0.2 * bzero * np.cos(2.0 * PI * (x[:, 1] - xmin[1]))
    if evolved:
        uu = uu * (1.0 + 0.2 * np.sin(2.0 * PI * x[:, 0]) * np.cos(4.0 * PI * x[:, 1]))
    if iener == 0:
        set_gamma(o, 1.0)
        o.polyk = 2.0 / 3.0 * uuzero
    _finish(p, o, dens, uu, B if imhd != 0 else None)
    if evolved:
        p.psi[:n] = 0.05 * bzero * np.sin(2.0 * PI * x[:, 0]) * np.sin(2.0 * PI * x[:, 1])
        p.alpha[:n, 0] = 0.1 + 0.9 * (0.5 + 0.5 * np.sin(4.0 * PI * x[:, 1])) ** 2
        p.alpha[:n, 1] = 0.5 * (0.5 + 0.5 * np.cos(2.0 * PI * x[:, 0]))
        p.alpha[:n, 2] = 1.0 - 0.5 * (0.5 + 0.5 * np.cos(4.0 * PI * x[:, 0])) ** 2
    if slab is not None:
        return o, p, info
    return o, p


# ---------------------------------------------------------------------------------------------------------
# hydro variant for the non-MHD tuple
# ---------------------------------------------------------------------------------------------------------
def hydro_box(ndim=3, nx=16, perturb_amp=0.1, seed=1):
    o, p = orszag_tang(ndim=ndim, nx=nx, perturb_amp=perturb_amp, imhd=0, idivbzero=0, iener=2, evolved=True, seed=seed,
                       cube=(ndim == 3))
    return o, p


# ---------------------------------------------------------------------------------------------------------
# C1: 1D MHD shock tube with fixed end particles (src/setup_shock1D_mhd.f90, multi/multi_shock.f90:65-87 Brio-Wu)
# ---------------------------------------------------------------------------------------------------------
def shock1d(nright=125, mhd=True, iener=2):
    o = default_options(1)
    o.ibound[0] = 1
    o.xmin[0], o.xmax[0] = -0.5, 0.5
    set_gamma(o, 2.0 if mhd else 5.0 / 3.0)
    o.imhd = 1 if mhd else 0
    o.iener = iener
    nbpts = 6
    densl, densr, prl, prr = 1.0, 0.125, 1.0, 0.1
    vleft = np.zeros(3)
    vright = np.zeros(3)
    Bx, Byl, Byr, Bzl, Bzr = (0.75, 1.0, -1.0, 0.0, 0.0) if mhd else (0.0,) * 5
    psep = 0.5 / nright  # particle spacing on the low-density side
    o.psep = psep
    gam1 = o.gamma - 1.0
    uul, uur = prl / (gam1 * densl), prr / (gam1 * densr)
    massp = densr * psep
    dsmooth = 20.0
    xs, dens, uu, vel, By, Bz = [], [], [], [], [], []
    xs.append(o.xmin[0] + 0.5 * massp / densl)
    xs.append(o.xmin[0] + psep * densr / densl + 0.5 * massp / densl)
    for _ in range(2):
        dens.append(densl); uu.append(uul); vel.append(vleft.copy()); By.append(Byl); Bz.append(Bzl)
    i = 1
    while xs[i] < o.xmax[0]:
        i += 1
        delta = 2.0 * (xs[i - 1] - 0.0) / psep
        if delta > dsmooth:
            d, u, v, by, bz = densr, uur, vright, Byr, Bzr
        elif delta < -dsmooth:
            d, u, v, by, bz = densl, uul, vleft, Byl, Bzl
        else:
            exx = math.exp(delta)
            d = (densl + densr * exx) / (1.0 + exx)
            u = (prl + prr * exx) / ((1.0 + exx) * gam1 * d)
            v = vright if delta > 0 else vleft
            by = (Byl + Byr * exx) / (1.0 + exx)
            bz = (Bzl + Bzr * exx) / (1.0 + exx)
        dens.append(d); uu.append(u); vel.append(np.array(v)); By.append(by); Bz.append(bz)
        xs.append(xs[i - 2] + 2.0 * massp / dens[i - 1])
    n = i  # npart = i-1 in 1-based Fortran == first i entries here
    x = np.array(xs[:n])[:, None]
    p = Particles(1, n, n + 16)
    p.x[:n] = x
    p.pmass[:n] = massp
    p.vel[:n] = np.array(vel[:n])
    B = np.stack([np.full(n, Bx), np.array(By[:n]), np.array(Bz[:n])], axis=1)
    _finish(p, o, np.array(dens[:n]), np.array(uu[:n]), B if mhd else None)
    # set_fixedbound (src/set_fixedbound.f90:44-59): first/last nbpts fixed, copies of the adjacent free particle
    p.itype[:nbpts] = ITYPE_BND
    p.itype[n - nbpts:n] = ITYPE_BND
    p.ireal[:nbpts] = nbpts + 1
    p.ireal[n - nbpts:n] = n - nbpts
    for arr in (p.rho, p.hh):
        arr[:nbpts] = arr[nbpts]
        arr[n - nbpts:n] = arr[n - nbpts - 1]
    return o, p


# ---------------------------------------------------------------------------------------------------------
# C4: two-fluid dust + gas in a periodic box (src/setup_dustybox.f90 geometry, fat-box variant of SURVEY 8d)
# ---------------------------------------------------------------------------------------------------------
def dustybox(ndim=3, nx=16, perturb_amp=0.05, Kdrag=1.0, idrag_nature=1, seed=7, coincident=False):
    o = default_options(ndim)
    o.idust, o.idrag_nature, o.Kdrag = 2, idrag_nature, Kdrag
    o.iener = 2
    psep = 1.0 / nx
    o.psep = psep
    xmin, xmax = [0.0] * ndim, [1.0] * ndim
    for d in range(ndim):
        o.ibound[d] = 3
        o.xmin[d], o.xmax[d] = xmin[d], xmax[d]
    xg, _ = cubic_lattice(xmin, xmax, psep)
    xg = perturb(xg, psep, perturb_amp, seed)
    if coincident:
        xd = xg.copy()  # the reference puts dust on top of gas (setup_dustybox.f90:58-69)
    else:
        xd = perturb(cubic_lattice(xmin, xmax, psep)[0] + 0.37 * psep, psep, perturb_amp, seed + 1)
    x = wrap_periodic(np.concatenate([xg, xd], axis=0), o)
    ngas = xg.shape[0]
    n = x.shape[0]
    massp = 1.0 / ngas
    hguess = o.hfact * psep
    p = _alloc(ndim, x, o, hguess)
    p.itype[:ngas] = ITYPE_GAS
    p.itype[ngas:n] = ITYPE_DUST
    p.pmass[:n] = massp
    p.vel[:ngas, 0] = 1.0 + 0.1 * np.sin(2.0 * PI * x[:ngas, 0])
    p.vel[ngas:n, 0] = 0.05 * np.cos(2.0 * PI * x[ngas:n, 0])
    dens = np.ones(n)
    uu = np.where(np.arange(n) < ngas, 1.0, 0.0)
    _finish(p, o, dens, uu, None)
    return o, p


# ---------------------------------------------------------------------------------------------------------
# C4 variant: one-fluid dust (idust=1), the dusty wave of src/setup_wave_x_ND_dust.f90:95-150 on a periodic box, with a
# perturbed dust fraction and a non-zero differential velocity so that every one-fluid term is live
# ---------------------------------------------------------------------------------------------------------
def dustywave_onefluid(ndim=3, nx=12, perturb_amp=0.1, dust_to_gas=1.0, Kdrag=1.0, idrag_nature=1, iav=2, mhd=False, seed=3,
                       use_smoothed_rhodust=True):
    o, p = orszag_tang(ndim=ndim, nx=nx, perturb_amp=perturb_amp, imhd=11 if mhd else 0, idivbzero=2 if mhd else 0, iener=2,
                       evolved=True, seed=seed, cube=(ndim == 3))
    o.idust, o.onef_dust, o.idustevol = 1, 1, 0
    o.idrag_nature, o.Kdrag = idrag_nature, Kdrag
    o.use_smoothed_rhodust = 1 if use_smoothed_rhodust else 0
    o.iav = iav
    n = p.npart
    x0 = p.x[:n, 0]
    eps0 = dust_to_gas / (1.0 + dust_to_gas)                                   # :148
    eps = eps0 * (1.0 + 0.1 * np.sin(2.0 * PI * x0) * np.cos(2.0 * PI * p.x[:n, min(1, ndim - 1)]))
    p.dustfrac[:n] = eps                                                        # the value the previous c2p left behind
    p.dustevol[:n] = eps * (1.0 + 0.02 * np.cos(4.0 * PI * x0))               # idustevol = 0: the evolved variable has moved on
    p.deltav[:n, 0] = 0.05 * np.sin(2.0 * PI * x0)
    p.deltav[:n, 1] = 0.02 * np.cos(2.0 * PI * x0)
    p.deltav[:n, 2] = -0.01 * np.sin(4.0 * PI * x0)
    p.pmass[:n] = p.pmass[:n] / (1.0 - eps0)                                   # :150 total (gas + dust) mass
    p.rho[:n] = p.rho[:n] / (1.0 - eps0)
    p.hh[:n] = o.hfact * (p.pmass[:n] / p.rho[:n]) ** (1.0 / ndim)
    return o, p


# ---------------------------------------------------------------------------------------------------------
# C4 as the reference sets it up: DUSTYBOX in the thin periodic box 1 x 11dp x 11dp (src/setup_dustybox.f90:46-107), gas lattice
# with the dust lattice on top of it (:58-69), gas moving at v_x = 1 through dust at rest.  In this box the periodic y/z ghosts
# outnumber the real particles.
# ---------------------------------------------------------------------------------------------------------
def dustybox_thin(nx=64, ny=11, Kdrag=1.0, idrag_nature=1, perturb_amp=0.0, seed=11):
    o = default_options(3)
    o.idust, o.idrag_nature, o.Kdrag = 2, idrag_nature, Kdrag
    o.iener = 2
    psep = 1.0 / nx
    o.psep = psep
    xmin, xmax = [0.0, 0.0, 0.0], [1.0, ny * psep, ny * psep]
    for d in range(3):
        o.ibound[d] = 3
        o.xmin[d], o.xmax[d] = xmin[d], xmax[d]
    xg, _ = cubic_lattice(xmin, xmax, psep)
    xg = perturb(xg, psep, perturb_amp, seed)
    xd = xg.copy()                                                             # :58-69: dust on top of gas
    x = wrap_periodic(np.concatenate([xg, xd], axis=0), o)
    ngas, n = xg.shape[0], 2 * xg.shape[0]
    massp = 1.0 * np.prod([xmax[d] - xmin[d] for d in range(3)]) / ngas        # rho_gas = rho_dust = 1
    p = _alloc(3, x, o, o.hfact * psep)
    p.itype[:ngas] = ITYPE_GAS
    p.itype[ngas:n] = ITYPE_DUST
    p.pmass[:n] = massp
    p.vel[:ngas, 0] = 1.0
    dens = np.ones(n)
    uu = np.where(np.arange(n) < ngas, 1.0, 0.0)
    _finish(p, o, dens, uu, None)
    return o, p


# ---------------------------------------------------------------------------------------------------------
# reflecting walls (ibound = 2, src/ghostND_mhd.f90:173, :226-228: ghost at xbound - (x - xbound), normal velocity flipped),
# optionally mixed with periodic directions
# ---------------------------------------------------------------------------------------------------------
def reflecting_box(ndim=3, nx=12, perturb_amp=0.2, mhd=True, ibound=None, seed=5):
    o, p = orszag_tang(ndim=ndim, nx=nx, perturb_amp=perturb_amp, imhd=11 if mhd else 0, idivbzero=2 if mhd else 0, iener=2,
                       evolved=True, seed=seed, cube=(ndim == 3)) if ndim > 1 else shock1d(nright=40, mhd=mhd)
    ib = ibound or [2] * ndim
    n = p.npart
    if ndim == 1:
        # the shock tube without its fixed end particles: the walls reflect instead
        p.itype[:n] = ITYPE_GAS
        p.ireal[:n] = 0
    for d in range(ndim):
        o.ibound[d] = ib[d]
    return o, p
