"""CPU tests that pin the oracle (oracle/nd_oracle.cpp) without the reference.

The reference ships no tests or golden data and cannot be compiled here (no Fortran compiler), so the oracle is
pinned by (1) kernel known answers computed independently from the published spline formulas the reference
tabulates (src/kernelND.f90:4143-4193), (2) the reference's own validation logic -- the O(N^2) neighbour check of
src/check_neighbourlist.f90:149-173, the momentum diagnostic of src/ratesND_mhd.f90:678,933 -- (3) independent
numpy restatements of the density sum, and (4) symmetry/conservation properties of the SPH equations.
"""
import math

import numpy as np
import pytest

import parity
from ndspmhd_b200 import abi, setups
from oracle import oracle

PI_K = 3.141592653589  # src/kernelND.f90:41


def cubic_analytic(q, ndim):
    cn = {1: 0.66666666666, 2: 10.0 / (7.0 * PI_K), 3: 1.0 / PI_K}[ndim]  # src/kernelND.f90:4152-4159
    q = np.asarray(q, float)
    w = np.where(q < 1, 1 - 1.5 * q**2 + 0.75 * q**3, np.where(q < 2, 0.25 * (2 - q) ** 3, 0.0))
    g = np.where(q < 1, -3 * q + 2.25 * q**2, np.where(q < 2, -0.75 * (2 - q) ** 2, 0.0))
    gg = np.where(q < 1, -3 + 4.5 * q, np.where(q < 2, 1.5 * (2 - q), 0.0))
    return cn * w, cn * g, cn * gg


@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_cubic_kernel_table_known_answers(ndim):
    w, gw, ggw, wd, radkern2, dq2 = oracle.kernel_tables(0, 41, ndim)
    assert radkern2 == 4.0 and dq2 == 4.0 / 4000.0
    q = np.sqrt(np.arange(4001) * dq2)
    wa, ga, gga = cubic_analytic(q, ndim)
    assert np.max(np.abs(w - wa)) < 2e-15 and np.max(np.abs(gw - ga)) < 4e-15 and np.max(np.abs(ggw - gga)) < 4e-15
    cn = {1: 0.66666666666, 2: 10.0 / (7.0 * PI_K), 3: 1.0 / PI_K}[ndim]
    assert w[0] == cn and w[4000] == 0.0 and gw[4000] == 0.0
    assert abs(w[1000] - 0.25 * cn) < 1e-16 and abs(gw[1000] + 0.75 * cn) < 1e-16     # q = 1


@pytest.mark.parametrize("ikernel,radkern", [(2, 2.5), (3, 3.0)])
def test_quartic_quintic_tables_are_normalised(ikernel, radkern):
    """int W dV = 1 in 3D for every tabulated kernel (the tables include cnormk, src/kernelND.f90:4279-4289)."""
    w, gw, ggw, _, radkern2, dq2 = oracle.kernel_tables(ikernel, 42 if ikernel == 2 else 0, 3)
    assert radkern2 == radkern * radkern
    q2 = np.arange(4001) * dq2
    q = np.sqrt(q2)
    # dV = 4 pi q^2 dq = 2 pi q d(q^2): trapezoid in q^2
    f = 2.0 * math.pi * q * w
    integral = np.sum(0.5 * (f[1:] + f[:-1]) * dq2)
    assert abs(integral - 1.0) < 5e-5   # sqrt end-point error of the trapezoid rule in q^2
    # grw is dW/dq: compare with centred differences of the table
    qm = 0.5 * (q[2:] + q[:-2])
    dwdq = (w[2:] - w[:-2]) / (q[2:] - q[:-2])
    k = slice(50, 3900)
    assert np.max(np.abs(dwdq[k] - np.interp(qm[k], q, gw))) < 2e-4 * np.max(np.abs(gw))


def test_interpolation_is_linear_in_q2():
    """src/kernelND.f90:4435-4455: index=int(q2*ddq2table), linear between table nodes, clamped at ikern."""
    w, gw, ggw, _, radkern2, dq2 = oracle.kernel_tables(0, 41, 3)
    for q2 in [0.0, 0.3333, 1.0, 1.00049, 2.71828, 3.9999, 4.0, 7.0]:
        wi, gi, ggi = oracle.interpolate(0, 3, q2)
        idx = min(int(q2 * (1.0 / dq2)), 4000)
        idx1 = min(idx + 1, 4000)
        dxx = q2 - idx * dq2
        ddq2 = 1.0 / dq2
        assert wi == w[idx] + (w[idx1] - w[idx]) * ddq2 * dxx
        assert gi == gw[idx] + (gw[idx1] - gw[idx]) * ddq2 * dxx
        assert ggi == ggw[idx] + (ggw[idx1] - ggw[idx]) * ddq2 * dxx
    assert oracle.interpolate(0, 3, 4.5)[0] == 0.0


def test_ran1_is_the_numerical_recipes_generator():
    """src/random.f90:61-96 (Park-Miller + Bays-Durham shuffle).  Independent restatement in Python integers."""
    IA, IM, IQ, IR, NTAB = 16807, 2147483647, 127773, 2836, 32
    NDIV = 1 + (IM - 1) // NTAB
    AM, RNMX = 1.0 / IM, 1.0 - 1.2e-7

    def ran1_py(state):
        idum, iv, iy = state
        if idum <= 0 or iy == 0:
            idum = max(-idum, 1)
            iv = [0] * NTAB
            for j in range(NTAB + 8, 0, -1):
                k = idum // IQ
                idum = IA * (idum - k * IQ) - IR * k
                if idum < 0:
                    idum += IM
                if j <= NTAB:
                    iv[j - 1] = idum
            iy = iv[0]
        k = idum // IQ
        idum = IA * (idum - k * IQ) - IR * k
        if idum < 0:
            idum += IM
        j = iy // NDIV
        iy = iv[j]
        iv[j] = idum
        return min(AM * iy, RNMX), (idum, iv, iy)

    st = (-268, None, 0)
    seed = -268
    for _ in range(200):
        ref, st = ran1_py(st)
        got, seed = oracle.ran1(seed)
        assert got == ref


CONFIGS = {
    "ot3d": lambda: setups.orszag_tang(ndim=3, nx=12, zfrac=0.5, perturb_amp=0.3, evolved=True),
    "ot2d_cp": lambda: setups.orszag_tang(ndim=2, nx=32, lattice="cp", perturb_amp=0.3, evolved=True),
    "hydro3d": lambda: setups.hydro_box(ndim=3, nx=10, perturb_amp=0.3),
    "shock1d": lambda: setups.shock1d(nright=40),
    "dusty3d": lambda: setups.dustybox(ndim=3, nx=8),
}


def run_oracle(name, **kw):
    o, p = CONFIGS[name]()
    o.device_ghosts = 1
    for k, v in kw.items():
        setattr(o, k, v)
    s, ms = oracle.derivs(o, p)
    return o, p, s


@pytest.mark.parametrize("name", ["ot3d", "ot2d_cp", "shock1d", "dusty3d"])
def test_linklist_pairs_equal_bruteforce_pairs(name):
    """src/check_neighbourlist.f90:149-173: every pair with q2i<radkern2 .or. q2j<radkern2 is found by the link list."""
    o, p, s = run_oracle(name)
    li, lj = oracle.linklist_pairs(o, p, s["hhmax"])
    bi, bj = oracle.bruteforce_pairs(p, 4.0)
    # the link-list loop skips ghost-ghost pairs, as does the brute-force finder (i or j real)
    assert np.array_equal(parity.pair_set(li, lj), parity.pair_set(bi, bj))
    assert len(li) == len(np.unique(parity.pair_set(li, lj))), "a pair was visited twice"


@pytest.mark.parametrize("name", ["ot3d", "ot2d_cp", "hydro3d"])
def test_momentum_is_conserved_by_the_pair_sums(name):
    """sum m f = 0 (the reference's `fmean` diagnostic, src/ratesND_mhd.f90:678,933), relative to sum |m f|."""
    o, p, s = run_oracle(name)
    n = p.npart
    mf = p.pmass[:n, None] * p.force[:n]
    # the tensor MHD force with the stressmax correction is still pairwise antisymmetric
    assert np.all(np.abs(mf.sum(axis=0)) <= 1e-12 * np.abs(mf).sum())
    assert np.allclose(s["fmean"], mf.sum(axis=0), rtol=0, atol=1e-12 * np.abs(mf).sum())


def test_hydro_energy_is_conserved():
    """sum m (v.f + du/dt) = 0 for iener=2 hydro with AV + conductivity and grad-h terms (exact pairwise cancellation)."""
    o, p, s = run_oracle("hydro3d")
    n = p.npart
    e = p.pmass[:n] * ((p.vel[:n] * p.force[:n]).sum(axis=1) + p.dudt[:n])
    assert abs(e.sum()) <= 1e-11 * np.abs(e).sum()


def test_total_energy_form_follows_the_reference_quirk():
    """iener=3: the reference accumulates the dissipative pair terms into dendt (src/ratesND_mhd.f90:1829-1830) and then
    OVERWRITES dendt with v.f + P/rho^2 drho/dt in the finalisation loop (:820-826).  Replicated, not fixed."""
    o3, p3 = CONFIGS["hydro3d"]()
    o3.device_ghosts = 1
    o3.iener = 3
    n = p3.npart
    p3.en[:n] = p3.uu[:n] + 0.5 * (p3.vel[:n] ** 2).sum(axis=1)
    oracle.derivs(o3, p3)
    pdv = p3.pr[:n] / p3.rho[:n] ** 2 * p3.drhodt[:n]
    assert np.max(np.abs(p3.dudt[:n] - pdv)) <= 1e-13 * np.max(np.abs(pdv))
    rhs = (p3.vel[:n] * p3.force[:n]).sum(axis=1) + p3.dudt[:n]
    assert np.max(np.abs(p3.dendt[:n] - rhs)) <= 1e-13 * np.max(np.abs(rhs))
    # and the forces are those of the thermal-energy run (the energy choice does not feed back into dv/dt)
    o2, p2, _ = run_oracle("hydro3d")
    assert np.max(np.abs(p3.force[:n] - p2.force[:n])) <= 1e-12 * np.max(np.abs(p2.force[:n]))


@pytest.mark.parametrize("ndim,nx", [(1, 64), (2, 24), (3, 10)])
def test_uniform_lattice_is_uniform(ndim, nx):
    """On an unperturbed periodic lattice with uniform fields every particle sees the same neighbourhood: identical rho,
    gradh, numneigh; zero force; h converges in one pass to hfact*(m/rho)^(1/ndim) within tolh."""
    o, p = setups.orszag_tang(ndim=ndim, nx=nx, perturb_amp=0.0, evolved=False, imhd=0, idivbzero=0, cube=True) if ndim > 1 else (None, None)
    if ndim == 1:
        o = abi.default_options(1)
        o.ibound[0] = 3
        o.xmin[0], o.xmax[0] = 0.0, 1.0
        o.psep = 1.0 / nx
        x, _ = setups.cubic_lattice([0.0], [1.0], o.psep)
        p = abi.Particles(1, nx, nx + 64)
        p.x[:nx] = x
        p.pmass[:nx] = 1.0 / nx
        setups._finish(p, o, np.ones(nx), np.ones(nx), None)
    n = p.npart
    p.vel[:n] = 0.0
    o.device_ghosts = 1
    s, _ = oracle.derivs(o, p)
    assert s["itsdensity"] == (1 if ndim > 1 else 2)   # 1D: the lattice sum with h=1.2dx misses rho=1 by >tolh, one more round
    assert np.ptp(p.rho[:n]) <= 4e-14 * p.rho[0]
    assert np.ptp(p.gradh[:n]) <= 1e-12
    assert np.ptp(p.numneigh[:n]) == 0
    # independent count of lattice points with r < 2h (plus self), h = 1.2 psep
    r = np.arange(-3, 4)
    g = np.stack(np.meshgrid(*([r] * ndim), indexing="ij"), axis=-1).reshape(-1, ndim)
    expected = int(np.sum((g**2).sum(axis=1) < (2 * 1.2) ** 2))
    assert p.numneigh[0] == expected
    scale = float(np.max(p.pr[:n] / (p.rho[:n] * p.hh[:n])))
    assert np.max(np.abs(p.force[:n])) <= 1e-12 * scale
    assert np.max(np.abs(p.hh[:n] - o.hfact * (p.pmass[:n] / p.rho[:n]) ** (1.0 / ndim)) / p.hh[:n]) < 2 * o.tolh


def test_density_sum_against_independent_numpy_bruteforce():
    """rho_i = sum_j m_j W(|r_ij|, h_i) over ALL rows incl. ghosts with the analytic cubic spline; the table
    interpolation error is ~1e-7 relative, far below any algorithmic mistake (wrong h, wrong norm, missed ghosts)."""
    o, p, s = run_oracle("ot3d")
    n, nt = p.npart, s["ntotal"]
    x, h, m = p.x[:nt], p.hh[:nt], p.pmass[:nt]
    rho = np.zeros(n)
    dwdh = np.zeros(n)
    for i in range(n):
        d = np.sqrt(((x[i] - x) ** 2).sum(axis=1))
        q = d / h[i]
        w, g, _ = cubic_analytic(q, 3)
        rho[i] = np.sum(m * w) / h[i] ** 3
        dwdh[i] = np.sum(m * (-q * g - 3 * w)) / h[i] ** 4
    assert np.max(np.abs(rho - p.rho[:n]) / rho) < 1e-6
    omega = 1.0 + h[:n] / (3.0 * rho) * dwdh                       # src/iterate_density.f90:193-201
    assert np.max(np.abs(1.0 / omega - p.gradh[:n])) < 1e-5
    # converged: |h - hfact (m/rho)^(1/3)| within a few tolh
    assert np.max(np.abs(h[:n] - o.hfact * (m[:n] / p.rho[:n]) ** (1 / 3.0)) / h[:n]) < 3 * o.tolh


def test_ghosts_are_periodic_images():
    """src/ghostND_mhd.f90:205-225: every ghost is its parent shifted by whole box lengths and lies outside the box
    within radkern*hhmax of it; every real particle within that distance of a face has its image."""
    o, p, s = run_oracle("ot3d")
    n, nt = p.npart, s["ntotal"]
    L = np.array([o.xmax[d] - o.xmin[d] for d in range(3)])
    par = p.ireal[n:nt] - 1
    assert np.all((par >= 0) & (par < n))
    shift = (p.x[n:nt] - p.x[par]) / L
    assert np.max(np.abs(shift - np.round(shift))) < 1e-12
    assert np.all(np.abs(np.round(shift)).sum(axis=1) >= 1)
    lo = np.array([o.xmin[d] for d in range(3)])
    hi = np.array([o.xmax[d] for d in range(3)])
    reach = 2.0 * s["hhmax"]
    assert np.all(p.x[n:nt] > lo - reach - 1e-12) and np.all(p.x[n:nt] < hi + reach + 1e-12)
    # count: product over dims of (1 + near-lo + near-hi) - 1 per particle
    near = ((p.x[:n] - lo < reach) & (p.x[:n] - lo > 0)).astype(int) + ((hi - p.x[:n] < reach) & (hi - p.x[:n] > 0)).astype(int)
    assert nt - n == int(np.sum(np.prod(1 + near, axis=1) - 1))
    assert np.array_equal(p.vel[n:nt], p.vel[par]) and np.array_equal(p.itype[n:nt], p.itype[par])


def test_divB_and_curlB_of_a_linear_field():
    """B = (a y, b x, 0) has div B = 0 and curl B = (0,0,b-a); the SPH difference operators (src/ratesND_mhd.f90:2552,
    :2601-2603, divided by rho at :640-643) must recover them away from the periodic seams to O(h^2) error."""
    o, p = setups.orszag_tang(ndim=2, nx=40, lattice="cubic", perturb_amp=0.0, evolved=False)
    n = p.npart
    a, b = 0.3, -0.7
    B = np.zeros((n, 3))
    B[:, 0] = a * p.x[:n, 1]
    B[:, 1] = b * p.x[:n, 0]
    p.Bevol[:n] = B
    p.vel[:n] = 0.0
    o.device_ghosts = 1
    s, _ = oracle.derivs(o, p)
    inner = np.all(np.abs(p.x[:n]) < 0.5 - 2.0 * s["hhmax"] - 1e-9, axis=1)
    assert inner.sum() > 100
    assert np.max(np.abs(p.divB[:n][inner])) < 1e-10
    assert np.max(np.abs(p.curlB[:n, 2][inner] - (b - a))) < 2e-3 * abs(b - a)
    assert np.max(np.abs(p.curlB[:n, :2][inner])) < 1e-12


def test_dust_gas_drag_conserves_momentum_and_heats_gas():
    """src/ratesND_mhd.f90:1156-1167: equal and opposite drag force, heating of the gas only."""
    o, p, s = run_oracle("dusty3d")
    n = p.npart
    mf = p.pmass[:n, None] * p.force[:n]
    assert np.all(np.abs(mf.sum(axis=0)) <= 1e-12 * np.abs(mf).sum())
    dust = p.itype[:n] == abi.ITYPE_DUST
    assert np.all(p.dudt[:n][dust] == 0.0) and np.all(p.dendt[:n][dust] == 0.0)
    assert s["dtdrag"] < 1e300 and s["ts_min"] > 0
    # total energy: sum m (v.f + dudt) = 0 -- drag dissipation goes into the gas
    e = p.pmass[:n] * ((p.vel[:n] * p.force[:n]).sum(axis=1) + p.dudt[:n])
    assert abs(e.sum()) <= 1e-11 * np.abs(e).sum()


def test_fixed_particles_keep_their_state_and_have_zero_rates():
    """1D shock tube with nbpts fixed particles each end (src/set_fixedbound.f90, src/ratesND_mhd.f90:949-965)."""
    o, p, s = run_oracle("shock1d")
    n = p.npart
    fixed = p.itype[:n] == abi.ITYPE_BND
    assert fixed.sum() == 12
    assert np.all(p.force[:n][fixed] == 0) and np.all(p.dudt[:n][fixed] == 0) and np.all(p.dBevoldt[:n][fixed] == 0)
    par = p.ireal[:n][fixed] - 1
    assert np.array_equal(p.rho[:n][fixed], p.rho[par]) and np.array_equal(p.hh[:n][fixed], p.hh[par])
    # plateau densities of the Brio-Wu set-up away from the interface and the ends
    xl = (p.x[:n, 0] > -0.4) & (p.x[:n, 0] < -0.1)
    xr = (p.x[:n, 0] > 0.1) & (p.x[:n, 0] < 0.4)
    assert np.max(np.abs(p.rho[:n][xl] - 1.0)) < 2e-3 and np.max(np.abs(p.rho[:n][xr] - 0.125)) < 2e-3 * 0.125 * 8


# ---------------------------------------------------------------------------------------------------------
# one-fluid dust (idust=1): the reference's own "should be zero" debug sums
# ---------------------------------------------------------------------------------------------------------
def _onefluid_energy_sum(p, smoothed):
    """esum of src/ratesND_mhd.f90:913-917: total energy rate of the one-fluid mixture."""
    n = p.npart
    m, rho, eps = p.pmass[:n], p.rho[:n], p.dustfrac[:n]
    rg, rd = (p.rhogas[:n], p.rhodust[:n]) if smoothed else (rho * (1 - eps), rho * eps)
    t = m * ((p.vel[:n] * p.force[:n]).sum(1) + rg * rd / rho ** 2 * (p.deltav[:n] * p.ddeltavdt[:n]).sum(1)
             + ((1 - 2 * eps) * 0.5 * (p.deltav[:n] ** 2).sum(1) - p.uu[:n]) * p.ddustevoldt[:n] + rg / rho * p.dudt[:n])
    return t.sum(), np.abs(t).sum()


@pytest.mark.parametrize("ndim,nx", [(3, 8), (2, 24)])
def test_onefluid_dust_conserves_momentum_and_dust_mass(ndim, nx):
    o, p = setups.dustywave_onefluid(ndim=ndim, nx=nx)
    oracle.derivs(o, p)
    n = p.npart
    m = p.pmass[:n]
    f = m[:, None] * p.force[:n]
    assert np.all(np.abs(f.sum(0)) <= 1e-13 * np.abs(f).sum())                       # sum m f = 0 (:678)
    d = m * p.ddustevoldt[:n]
    assert abs(d.sum()) <= 1e-13 * np.abs(d).sum()                                   # dust mass: sum m d(eps)/dt = 0 (derivs.f90:225-232)
    assert np.all(np.abs(p.rhogas[:n] + p.rhodust[:n] - p.rho[:n]) <= 1e-14 * p.rho[:n])   # the two density sums partition rho
    assert np.array_equal(p.dustfrac[:n], p.dustevol[:n])                            # idustevol = 0 (conservative2primitive.f90:91)
    assert np.allclose(p.dens[:n], p.rho[:n] * (1 - p.dustfrac[:n]), rtol=1e-15)     # dens is the gas density (:112)


@pytest.mark.parametrize("iav", [1, 2])
def test_onefluid_dust_energy_identity(iav):
    """With a dust fraction consistent with rho_g, rho_d (unsmoothed) and no dust-fraction diffusion (alpha_B = 0) the
    reference's debug sum (ratesND_mhd.f90:908-917, 'should be zero if conserving energy') vanishes to round-off with
    viscosity, conductivity, drag heating and the deltav dissipation all active."""
    o, p = setups.dustywave_onefluid(ndim=3, nx=8, iav=iav, use_smoothed_rhodust=False)
    n = p.npart
    p.dustevol[:n] = p.dustfrac[:n]
    p.alpha[:n, 0], p.alpha[:n, 1], p.alpha[:n, 2] = 0.7, 0.4, 0.0
    oracle.derivs(o, p)
    e, scale = _onefluid_energy_sum(p, smoothed=False)
    assert abs(e) <= 1e-13 * scale, (e, scale)


def test_onefluid_dust_drag_terms():
    """Uniform lattice, uniform dust fraction and deltav: every pair sum cancels, leaving the local drag terms of
    ratesND_mhd.f90:561-582: d(deltav)/dt = -deltav/ts, du/dt = rho_d/rho deltav^2/ts, dtdrag = ts = rho_d rho_g/(K rho)."""
    o, p = setups.dustywave_onefluid(ndim=3, nx=8, perturb_amp=0.0, Kdrag=2.0)
    n = p.npart
    eps = 0.5
    p.dustfrac[:n] = eps; p.dustevol[:n] = eps
    p.deltav[:n] = np.array([0.03, -0.01, 0.02]); p.vel[:n] = 0.0
    p.en[:n] = 1.0; p.alpha[:n] = 0.0
    s, _ = oracle.derivs(o, p)
    rg, rd, rho = p.rhogas[:n], p.rhodust[:n], p.rho[:n]
    ts = rd * rg / (2.0 * rho)
    assert np.allclose(p.ddeltavdt[:n], -p.deltav[:n] / ts[:, None], rtol=1e-9, atol=1e-12)
    assert np.allclose(p.dudt[:n], rd / rho * (p.deltav[:n] ** 2).sum(1) / ts, rtol=1e-9)
    assert np.isclose(s["dtdrag"], ts.min(), rtol=1e-14)


def test_ideal_spmhd_rates_against_independent_numpy_bruteforce():
    """The non-dissipative SPMHD equations as PUBLISHED (Price & Monaghan 2005; Price 2012, J. Comp. Phys. 231, eqs. for variable-h SPMHD
    with the stress S = -(P + B^2/2) I + B B and the constant `stressmax` subtracted against the tensile instability), summed over ALL
    rows by brute force with the analytic cubic spline -- no cells, no link list, no tables, no pair symmetry -- against the oracle's
    restatement of src/ratesND_mhd.f90 with alpha = alpha_u = alpha_B = 0 (no artificial dissipation), imhd = 1, idivbzero = 0:

      drho_i/dt   =  sum_j m_j (v_ij . r^) F_i                      F_i = W'(r_ij, h_i) / Omega_i,  r^ = (x_i - x_j)/r_ij
      dv_i/dt     = -sum_j m_j [ (P_i + B_i^2/2)/rho_i^2 F_i + (P_j + B_j^2/2)/rho_j^2 F_j ] r^
                    +sum_j m_j [ (B_i (B_i.r^) - S r^)/rho_i^2 F_i + (B_j (B_j.r^) - S r^)/rho_j^2 F_j ]
      du_i/dt     =  P_i/rho_i^2 sum_j m_j (v_ij . r^) F_i
      d(B/rho)/dt = -1/rho_i^2 sum_j m_j v_ij (B_i . r^) F_i

    Agreement is limited by the linear interpolation of the kernel tables (~1e-7): a wrong factor, sign, index or a missed/duplicated
    neighbour shows at the 1e-2..1 level."""
    o, p = setups.orszag_tang(ndim=3, nx=10, cube=True, perturb_amp=0.25, evolved=True, imhd=1, idivbzero=0)
    o.device_ghosts = 1
    p.alpha[:] = 0.0
    # a field strong enough for a positive stressmax = max(B^2/2 - P) so that the correction term is exercised
    p.Bevol[: p.npart] *= 4.0
    s, _ = oracle.derivs(o, p)
    n, nt = p.npart, s["ntotal"]
    S = s["stressmax"]
    assert S > 0.0
    x, v, m, h, rho, om1, P, B = p.x[:nt], p.vel[:nt], p.pmass[:nt], p.hh[:nt], p.rho[:nt], p.gradh[:nt], p.pr[:nt], p.Bfield[:nt]
    assert np.all(rho[n:nt] == rho[p.ireal[n:nt] - 1]) and np.all(B[n:nt] == B[p.ireal[n:nt] - 1])   # ghosts carry their parents' state
    drho, acc, dudt, dBrho = np.zeros(n), np.zeros((n, 3)), np.zeros(n), np.zeros((n, 3))
    for i in range(n):
        dx = x[i] - x
        r = np.sqrt((dx**2).sum(axis=1))
        k = np.where((r > 0) & ((r < 2 * h[i]) | (r < 2 * h)))[0]
        rh = dx[k] / r[k, None]
        Fi = cubic_analytic(r[k] / h[i], 3)[1] / h[i] ** 4 * om1[i]
        Fj = cubic_analytic(r[k] / h[k], 3)[1] / h[k] ** 4 * om1[k]
        vr = ((v[i] - v[k]) * rh).sum(axis=1)
        drho[i] = np.sum(m[k] * vr * Fi)
        ci, cj = Fi / rho[i] ** 2, Fj / rho[k] ** 2
        iso = (P[i] + 0.5 * (B[i] ** 2).sum()) * ci + (P[k] + 0.5 * (B[k] ** 2).sum(axis=1)) * cj
        Bir, Bjr = (B[i] * rh).sum(axis=1), (B[k] * rh).sum(axis=1)
        aniso = (B[i][None, :] * Bir[:, None] - S * rh) * ci[:, None] + (B[k] * Bjr[:, None] - S * rh) * cj[:, None]
        acc[i] = (m[k, None] * (aniso - iso[:, None] * rh)).sum(axis=0)
        dudt[i] = P[i] / rho[i] ** 2 * drho[i]
        dBrho[i] = -(m[k, None] * (v[i] - v[k]) * (Bir * Fi)[:, None]).sum(axis=0) / rho[i] ** 2

    def err(a, b):
        return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))

    assert err(p.drhodt[:n], drho) < 1e-5
    assert err(p.force[:n], acc) < 1e-5
    assert err(p.dudt[:n], dudt) < 1e-5
    assert err(p.dBevoldt[:n], dBrho) < 1e-5


@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_ghosts_of_reflecting_walls_are_mirror_images(ndim):
    """src/ghostND_mhd.f90:173, :205, :226-228 with ibound = 2: a particle closer than radkern*h_i (its OWN h, not hhmax) to a wall has a
    mirror image x' = xbound - (x - xbound); a particle near two or three walls also has the images across the edge / corner.  Checked
    against numpy: the multiset of ghost positions is the multiset of mirror images, every ghost points at its parent, and a ghost across
    ONE wall carries the parent's velocity with that component reversed.  (The reference resets / reverses whole velocity vectors on edge
    and corner ghosts, :283-284, :307-309 -- restated as it is and not judged here.)"""
    o, p = setups.reflecting_box(ndim=ndim, nx=10 if ndim == 3 else 24, perturb_amp=0.2, mhd=(ndim > 1))
    o.device_ghosts = 1
    n = p.npart
    x0, v0, h0 = p.x[:n].copy(), p.vel[:n].copy(), p.hh[:n].copy()
    s, _ = oracle.derivs(o, p, phases=oracle.NDO_GHOSTS | oracle.NDO_LINK)     # ghosts of the input h, before the iteration moves it
    nt = s["ntotal"]
    lo = np.array([o.xmin[d] for d in range(ndim)])
    hi = np.array([o.xmax[d] for d in range(ndim)])
    reach = 2.0 * h0[:, None]
    near_lo = (x0 - lo < reach) & (x0 - lo > 0)
    near_hi = (hi - x0 < reach) & (hi - x0 > 0)
    # every non-empty choice of {keep, mirror in lo, mirror in hi} per dimension
    images, parents, nrefl = [], [], []
    import itertools
    for choice in itertools.product((0, 1, 2), repeat=ndim):
        if not any(choice):
            continue
        ok = np.ones(n, bool)
        xi = x0.copy()
        for d, c in enumerate(choice):
            if c == 1:
                ok &= near_lo[:, d]; xi[:, d] = lo[d] - (x0[:, d] - lo[d])
            elif c == 2:
                ok &= near_hi[:, d]; xi[:, d] = hi[d] - (x0[:, d] - hi[d])
        images.append(xi[ok]); parents.append(np.nonzero(ok)[0]); nrefl.append(np.full(int(ok.sum()), sum(c != 0 for c in choice)))
    images, parents, nrefl = np.concatenate(images), np.concatenate(parents), np.concatenate(nrefl)
    assert nt - n == images.shape[0] and nt - n > 0
    gx, gpar = p.x[n:nt], p.ireal[n:nt] - 1
    key = lambda xs, par: sorted(zip(par.tolist(), *[np.round(xs[:, d], 12).tolist() for d in range(ndim)]))
    assert key(gx, gpar) == key(images, parents)
    # one-wall ghosts: the normal velocity component is reversed, the others kept
    moved = np.abs(gx - x0[gpar]) > 0
    one = moved.sum(axis=1) == 1
    assert one.any()
    vexp = v0[gpar[one]].copy()
    d_of = np.argmax(moved[one], axis=1)
    vexp[np.arange(vexp.shape[0]), d_of] *= -1.0
    assert np.array_equal(p.vel[n:nt][one], vexp)


@pytest.mark.parametrize("name, iener, imhd", [("ot3d", 2, 11), ("ot3d", 0, 11), ("ot3d", 2, 1), ("shock1d", 3, 1), ("dusty3d", 2, 0)])
def test_cons2prim_and_eos_formulas(name, iener, imhd):
    """conservative2primitive.f90:190-193, :341-373 and eos.f90:74-100 evaluated independently in numpy on the converged densities:
    dens = rho; B = Bevol (imhd >= 11) or Bevol*rho (B/rho evolved); u = en (thermal energy), en - v^2/2 - B^2/(2 rho) (total energy),
    P/((gamma-1) rho) (polytropic, gamma /= 1); P = (gamma-1) u rho or polyk rho^gamma; c_s = sqrt(gamma P/rho); dust particles carry
    no pressure and no thermal energy."""
    o, p = CONFIGS[name]() if name != "shock1d" else setups.shock1d(nright=60, iener=iener)   # en = total energy there
    o.device_ghosts = 1
    if name != "dusty3d":
        o.iener = iener
        if name == "ot3d":
            o.imhd = imhd
            if iener == 0:
                setups.set_gamma(o, 1.0)
                o.polyk = 0.4
    en0, Bev0, v0 = p.en.copy(), p.Bevol.copy(), p.vel.copy()
    s, _ = oracle.derivs(o, p)
    n = p.npart
    rho = p.rho[:n]
    gas = p.itype[:n] != 2
    assert np.array_equal(p.dens[:n], rho)
    if o.imhd >= 11:
        B = Bev0[:n]
    elif o.imhd > 0:
        B = Bev0[:n] * rho[:, None]
    else:
        B = np.zeros((n, 3))
    if o.imhd != 0:
        assert np.max(np.abs(p.Bfield[:n] - B)) <= 1e-15 * max(np.max(np.abs(B)), 1e-300)
    if o.iener == 3:
        u = en0[:n] - 0.5 * (v0[:n] ** 2).sum(1) - 0.5 * (B ** 2).sum(1) / rho
    elif o.iener == 0:
        P = o.polyk * rho ** o.gamma
        u = P / ((o.gamma - 1.0) * rho) if abs(o.gamma - 1.0) > 1e-3 else p.uu[:n]
    else:
        u = en0[:n]
    if o.iener != 0:
        P = (o.gamma - 1.0) * u * rho
    P = np.where(gas, P, 0.0)
    tol = 1e-13
    assert np.max(np.abs(p.uu[:n][gas] - u[gas])) <= tol * np.max(np.abs(u[gas]))
    assert np.max(np.abs(p.pr[:n] - P)) <= tol * np.max(np.abs(P))
    cs = np.sqrt(o.gamma * P[gas] / rho[gas])
    assert np.max(np.abs(p.spsound[:n][gas] - cs)) <= tol * np.max(cs)
