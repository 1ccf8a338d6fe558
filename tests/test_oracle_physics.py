"""Physics known-answer tests of the whole restated path (third pin): the oracle's link -> density iteration -> cons2prim ->
rates -> leapfrog step, run on the reference's own shock-tube setups (src/setup_shock1D_mhd.f90 with the states of
multi/multi_shock.f90:65-87 and src/setup_shockND.f90:108-123), must land on answers obtained WITHOUT any SPH:

* Sod tube (gamma = 5/3): the exact Riemann solution (iterative pressure solve, Toro ch. 4) -- both plateaus, the contact speed, the shock
  position and the velocity profile of the rarefaction fan;
* Brio-Wu MHD tube (gamma = 2): a first-order HLL finite-volume solution of 1-D ideal MHD on 3000 cells, written here in numpy;
* DUSTYBOX (two-fluid drag, src/setup_dustybox.f90): the analytic exponential relaxation of the differential velocity.

The reference ships no expected outputs; these are the published test problems its documentation shows it on.  Tolerances are those of
SPH at this resolution (AV-broadened shocks, ~500-1100 particles), not round-off: what they exclude is a wrong term, sign or factor in
the force, energy, induction or dissipation sums, which shifts a plateau by tens of per cent.
"""
import numpy as np

from ndspmhd_b200 import setups
from oracle import oracle


def _dt0(s, C_cour=0.3, C_force=0.25):
    return min(C_force * s["dtforce"], C_cour * s["dtcourant"], 0.9 * s["dtdrag"], C_force * s["dtvisc"])


def _evolve(o, p, tmax):
    s, _ = oracle.derivs(o, p)
    dt, t, nsteps = _dt0(s), 0.0, 0
    while t < tmax:
        dt = min(dt, tmax - t)
        dtnew, s = oracle.step(o, p, dt)
        t += dt
        dt = dtnew
        nsteps += 1
        assert nsteps < 20000
    return nsteps


def exact_sod(rl, pl, rr, pr, g):
    """Exact Riemann solution for zero initial velocities with a left rarefaction and a right shock."""
    cl, cr = np.sqrt(g * pl / rl), np.sqrt(g * pr / rr)
    A, B = 2.0 / ((g + 1.0) * rr), (g - 1.0) / (g + 1.0) * pr

    def f(p):
        return 2.0 * cl / (g - 1.0) * ((p / pl) ** ((g - 1.0) / (2.0 * g)) - 1.0) + (p - pr) * np.sqrt(A / (p + B))

    lo, hi = pr, pl
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        lo, hi = (lo, mid) if f(mid) > 0 else (mid, hi)
    ps = 0.5 * (lo + hi)
    vs = (ps - pr) * np.sqrt(A / (ps + B))
    mu = (g - 1.0) / (g + 1.0)
    return dict(p=ps, v=vs, rho_l=rl * (ps / pl) ** (1.0 / g), rho_r=rr * (ps / pr + mu) / (mu * ps / pr + 1.0),
                s_shock=cr * np.sqrt((g + 1.0) / (2.0 * g) * ps / pr + (g - 1.0) / (2.0 * g)), s_head=-cl,
                s_tail=vs - cl * (ps / pl) ** ((g - 1.0) / (2.0 * g)), cl=cl)


def hll_mhd_1d(ncell, tend, g, left, right, bx, cfl=0.4):
    """First-order HLL finite-volume solution of 1-D ideal MHD on [-0.5, 0.5]; left/right = (rho, p, vx, vy, vz, by, bz), mu0 = 1."""
    x = (np.arange(ncell) + 0.5) / ncell - 0.5

    def cons(r, p, vx, vy, vz, by, bz):
        return np.array([r, r * vx, r * vy, r * vz, by, bz, p / (g - 1) + 0.5 * r * (vx * vx + vy * vy + vz * vz) + 0.5 * (bx * bx + by * by + bz * bz)])

    U = np.where(x[None, :] < 0, cons(*left)[:, None], cons(*right)[:, None]).astype(float)

    def prim(U):
        r, mx, my, mz, by, bz, e = U
        vx, vy, vz = mx / r, my / r, mz / r
        b2 = bx * bx + by * by + bz * bz
        return r, vx, vy, vz, by, bz, e, b2, (g - 1) * (e - 0.5 * r * (vx * vx + vy * vy + vz * vz) - 0.5 * b2)

    t = 0.0
    while t < tend:
        r, vx, vy, vz, by, bz, e, b2, p = prim(U)
        pt = p + 0.5 * b2
        F = np.array([r * vx, r * vx * vx + pt - bx * bx, r * vy * vx - bx * by, r * vz * vx - bx * bz, by * vx - bx * vy, bz * vx - bx * vz,
                      (e + pt) * vx - bx * (vx * bx + vy * by + vz * bz)])
        a2 = g * p / r
        cf = np.sqrt(0.5 * (a2 + b2 / r + np.sqrt(np.maximum((a2 + b2 / r) ** 2 - 4 * a2 * bx * bx / r, 0.0))))
        dt = min(cfl / ncell / np.max(np.abs(vx) + cf), tend - t)
        sl = np.minimum(np.minimum(vx[:-1] - cf[:-1], vx[1:] - cf[1:]), 0.0)
        sr = np.maximum(np.maximum(vx[:-1] + cf[:-1], vx[1:] + cf[1:]), 0.0)
        Fh = (sr * F[:, :-1] - sl * F[:, 1:] + sl * sr * (U[:, 1:] - U[:, :-1])) / (sr - sl)
        U[:, 1:-1] -= dt * ncell * (Fh[:, 1:] - Fh[:, :-1])
        t += dt
    r, vx, vy, vz, by, bz, e, b2, p = prim(U)
    return dict(x=x, rho=r, vx=vx, vy=vy, by=by, p=p)


def test_sod_shock_tube_lands_on_the_exact_riemann_solution():
    tmax = 0.15
    o, p = setups.shock1d(nright=60, mhd=False, iener=2)
    _evolve(o, p, tmax)
    n = p.npart
    x, rho, v, pr = p.x[:n, 0], p.rho[:n], p.vel[:n, 0], p.pr[:n]
    e = exact_sod(1.0, 1.0, 0.125, 0.1, o.gamma)

    def middle(xa, xb):
        return (x > xa + 0.25 * (xb - xa)) & (x < xb - 0.25 * (xb - xa))

    right = middle(e["v"] * tmax, e["s_shock"] * tmax)      # contact .. shock
    left = middle(e["s_tail"] * tmax, e["v"] * tmax)        # tail of the fan .. contact
    assert right.sum() >= 8 and left.sum() >= 16
    for m, rho_exact in ((right, e["rho_r"]), (left, e["rho_l"])):
        assert abs(np.median(rho[m]) / rho_exact - 1.0) < 0.01
        assert abs(np.median(v[m]) / e["v"] - 1.0) < 0.01
        assert abs(np.median(pr[m]) / e["p"] - 1.0) < 0.01
    # shock front: where rho falls through the mean of the pre- and post-shock densities
    mid = 0.5 * (0.125 + e["rho_r"])
    k = np.where((rho[:-1] > mid) & (rho[1:] <= mid) & (x[:-1] > e["v"] * tmax))[0]
    assert len(k) == 1 and abs(x[k[0]] - e["s_shock"] * tmax) < 0.01
    # simple-wave invariants from the undisturbed left state through the fan up to the contact: J+ = v + 2c/(gamma-1) and the entropy
    # P/rho^gamma (they hold for the smoothed initial jump of the setup as well, unlike the self-similar profile)
    wave = (x > -0.45) & (x < e["v"] * tmax - 0.03)
    assert wave.sum() > 300
    g = o.gamma
    J = v[wave] + 2.0 * np.sqrt(g * pr[wave] / rho[wave]) / (g - 1.0)
    assert np.max(np.abs(J / (2.0 * e["cl"] / (g - 1.0)) - 1.0)) < 0.01
    assert np.max(np.abs(pr[wave] / rho[wave] ** g - 1.0)) < 0.03
    # undisturbed states outside the waves
    assert np.allclose(rho[x < e["s_head"] * tmax - 0.05], 1.0, rtol=5e-3) and np.allclose(rho[(x > e["s_shock"] * tmax + 0.05) & (x < 0.45)], 0.125, rtol=5e-3)


def test_brio_wu_tube_lands_on_an_independent_finite_volume_solution():
    tmax = 0.1
    o, p = setups.shock1d(nright=125, mhd=True, iener=2)
    _evolve(o, p, tmax)
    n = p.npart
    x = p.x[:n, 0]
    sph = dict(rho=p.rho[:n], vx=p.vel[:n, 0], vy=p.vel[:n, 1], by=p.Bfield[:n, 1], p=p.pr[:n])
    assert np.max(np.abs(p.Bfield[:n, 0][np.abs(x) < 0.4] / 0.75 - 1.0)) < 0.05   # Bx = (B/rho)_x rho stays the constant it is in 1-D
    fv = hll_mhd_1d(3000, tmax, o.gamma, (1.0, 1.0, 0, 0, 0, 1.0, 0), (0.125, 0.1, 0, 0, 0, -1.0, 0), 0.75)
    # plateaus of the Brio-Wu solution at t = 0.1: behind the left fast rarefaction, between slow compound wave and contact,
    # between contact and slow shock, behind the right fast rarefaction
    windows = {"post_fast_rarefaction_left": (-0.065, -0.045, 0.03), "compound_to_contact": (0.0, 0.03, 0.05),
               "contact_to_slow_shock": (0.08, 0.13, 0.07), "post_fast_rarefaction_right": (0.17, 0.30, 0.04)}
    for name, (xa, xb, tol) in windows.items():
        ms, mf = (x > xa) & (x < xb), (fv["x"] > xa) & (fv["x"] < xb)
        assert ms.sum() >= 10, name
        for f in ("rho", "by", "p"):
            a, b = np.median(sph[f][ms]), np.median(fv[f][mf])
            assert abs(a / b - 1.0) < tol, (name, f, a, b)
        for f in ("vx", "vy"):                       # velocities against the largest speed of the problem (some plateaus are near rest)
            a, b = np.median(sph[f][ms]), np.median(fv[f][mf])
            assert abs(a - b) < tol * 1.6, (name, f, a, b)


import pytest  # noqa: E402


@pytest.mark.parametrize("ndim,nx", [(2, 16), (3, 8)])
def test_dustybox_relaxes_at_the_analytic_rate(ndim, nx):
    """DUSTYBOX (Laibe & Price 2011, MNRAS 418, 1491; the reference's src/setup_dustybox.f90): uniform gas streaming through uniform dust with
    a constant drag coefficient K.  The differential velocity decays as exp(-K (1/rho_g + 1/rho_d) t), the barycentre keeps its velocity
    and the lost kinetic energy heats the gas (src/ratesND_mhd.f90:1074-1169 + the leapfrog)."""
    K = 1.0
    o, p = setups.dustybox(ndim=ndim, nx=nx, perturb_amp=0.0, Kdrag=K)
    n = p.npart
    gas = p.itype[:n] == 0
    dust = ~gas
    p.vel[:n] = 0.0
    p.vel[:n, 0][gas] = 1.0
    m = p.pmass[:n]
    e0 = float(np.sum(m * (0.5 * (p.vel[:n] ** 2).sum(axis=1) + p.en[:n])))
    tmax = 0.5
    _evolve(o, p, tmax)
    vg, vd = p.vel[:n, 0][gas], p.vel[:n, 0][dust]
    rg, rd = float(p.rho[:n][gas].mean()), float(p.rho[:n][dust].mean())
    assert np.ptp(vg) < 1e-12 and np.ptp(vd) < 1e-12                       # the uniform state stays uniform
    assert abs((vg.mean() - vd.mean()) / np.exp(-K * (1.0 / rg + 1.0 / rd) * tmax) - 1.0) < 0.01
    assert abs(0.5 * (vg.mean() + vd.mean()) - 0.5) < 1e-12                # equal masses: barycentre velocity 1/2
    assert np.max(np.abs(p.vel[:n, 1:])) < 1e-4                             # pairwise drag acts along r^: transverse parts cancel to lattice order
    e1 = float(np.sum(m * (0.5 * (p.vel[:n] ** 2).sum(axis=1) + p.en[:n])))
    assert abs(e1 / e0 - 1.0) < 2e-3                                       # drag heating = kinetic energy lost (second-order in dt)
    assert np.all(p.en[:n][gas] > 1.2) and np.all(p.en[:n][dust] == 0.0)   # the heat goes to the gas only


def _wave1d(nx, kind, amp):
    """1-D periodic box [0,1], rho = 1 (the geometry of src/setup_wave_x_ND.f90).  `sound`: linear acoustic wave, c_s = 1;
    `alfven`: circularly polarised Alfven wave on Bx = 1 -- an exact nonlinear solution of ideal MHD travelling at v_A = 1."""
    from ndspmhd_b200.setups import _alloc, _finish, cubic_lattice, default_options
    o = default_options(1)
    o.ibound[0] = 3
    o.xmin[0], o.xmax[0] = 0.0, 1.0
    o.psep = 1.0 / nx
    o.imhd = 0 if kind == "sound" else 1
    o.iener = 2
    x, _ = cubic_lattice([0.0], [1.0], o.psep)
    n = x.shape[0]
    p = _alloc(1, x, o, o.hfact * o.psep)
    p.pmass[:n] = 1.0 / n
    k, g = 2.0 * np.pi, o.gamma
    xx = x[:, 0]
    if kind == "sound":
        xx = xx + amp / k * np.cos(k * xx)        # displaced lattice: the SPH density itself carries rho0 (1 + amp sin kx)
        p.x[:n, 0] = xx
        p.vel[:n, 0] = amp * np.sin(k * xx)
        dens, uu, B = 1.0 + amp * np.sin(k * xx), (1.0 + (g - 1.0) * amp * np.sin(k * xx)) / (g * (g - 1.0)), None
    elif kind == "fast":
        # fast magnetosonic wave across B = (0, 0.75, 0): v_f^2 = c_s^2 + v_A^2 = 1 + 0.5625; d rho/rho = dBy/By = dvx/v_f = amp sin kx
        xx = xx + amp / k * np.cos(k * xx)
        p.x[:n, 0] = xx
        p.vel[:n, 0] = 1.25 * amp * np.sin(k * xx)
        dens, uu = 1.0 + amp * np.sin(k * xx), (1.0 + (g - 1.0) * amp * np.sin(k * xx)) / (g * (g - 1.0))
        B = np.stack([np.zeros(n), 0.75 * (1.0 + amp * np.sin(k * xx)), np.zeros(n)], axis=1)
    else:
        dens, uu = np.ones(n), np.full(n, 0.15)
        B = np.stack([np.ones(n), amp * np.sin(k * xx), amp * np.cos(k * xx)], axis=1)
        p.vel[:n, 1], p.vel[:n, 2] = -amp * np.sin(k * xx), -amp * np.cos(k * xx)   # right-going: v_perp = -B_perp / sqrt(rho)
    _finish(p, o, dens, uu, B)
    return o, p


def _fourier(x, f):
    c, s = np.sum(f * np.cos(2.0 * np.pi * x)), np.sum(f * np.sin(2.0 * np.pi * x))
    return np.arctan2(c, s), 2.0 * np.hypot(c, s) / len(x)       # f = A sin(2 pi x + phase)


@pytest.mark.parametrize("kind,nx,amp,tol,speed_exact", [("sound", 64, 1e-3, 0.01, 1.0), ("alfven", 64, 0.1, 0.003, 1.0), ("alfven", 128, 0.1, 0.001, 1.0),
                                                         ("fast", 64, 1e-3, 0.01, 1.25)])
def test_waves_travel_at_the_sound_alfven_and_fast_speeds(kind, nx, amp, tol, speed_exact):
    o, p = _wave1d(nx, kind, amp)
    n = p.npart
    field = (lambda q: q.vel[:n, 0]) if kind != "alfven" else (lambda q: q.Bevol[:n, 1] * q.rho[:n])   # imhd = 1 evolves B/rho
    oracle.derivs(o, p)
    ph0, a0 = _fourier(p.x[:n, 0], field(p))
    tmax = 0.5 / speed_exact
    _evolve(o, p, tmax)
    ph1, a1 = _fourier(p.x[:n, 0], field(p))
    speed = ((ph0 - ph1) % (2.0 * np.pi)) / (2.0 * np.pi * tmax)
    assert abs(speed / speed_exact - 1.0) < tol, speed
    assert 0.98 < a1 / a0 < 1.001


@pytest.mark.parametrize("ndim,nx,tol", [(2, 16, 0.015), (3, 8, 0.04)])
def test_onefluid_dustybox_relaxes_at_the_analytic_rate(ndim, nx, tol):
    """The same relaxation in the one-fluid formulation (idust = 1; Laibe & Price 2014; src/ratesND_mhd.f90:548-582, :2726-2807): a uniform
    mixture with dust fraction 1/2 and differential velocity 1 obeys d(dv)/dt = -dv/t_s, t_s = rho_g rho_d / (K rho), the barycentre stays
    at rest and E = sum m [v^2/2 + eps (1 - eps) dv^2/2 + (1 - eps) u] is conserved."""
    K, tmax = 1.0, 0.15
    o, p = setups.dustywave_onefluid(ndim=ndim, nx=nx, perturb_amp=0.0, Kdrag=K)
    n = p.npart
    p.vel[:n] = 0.0
    p.dustfrac[:n] = p.dustevol[:n] = 0.5
    p.deltav[:n] = 0.0
    p.deltav[:n, 0] = 1.0
    p.en[:n] = 0.9
    p.alpha[:n] = [o.alphamin, o.alphaumin, o.alphaBmin]

    def energy(q):
        eps = q.dustfrac[:n]
        return float(np.sum(q.pmass[:n] * (0.5 * (q.vel[:n] ** 2).sum(axis=1) + 0.5 * eps * (1 - eps) * (q.deltav[:n] ** 2).sum(axis=1) + (1 - eps) * q.en[:n])))

    e0 = energy(p)
    _evolve(o, p, tmax)
    rho = float(p.rho[:n].mean())
    eps = float(p.dustfrac[:n].mean())
    assert abs(eps - 0.5) < 1e-12 and np.ptp(p.deltav[:n, 0]) < 1e-12            # uniform stays uniform, no dust-fraction evolution
    rate = -np.log(p.deltav[:n, 0].mean()) / tmax
    assert abs(rate / (K / (rho * eps * (1 - eps))) - 1.0) < tol, rate
    assert np.max(np.abs(p.vel[:n])) < 1e-12 and np.max(np.abs(p.deltav[:n, 1:])) < 1e-12
    assert abs(energy(p) / e0 - 1.0) < 2e-3


def test_hyperbolic_cleaning_carries_a_divergence_blob_away():
    """Dedner's static divergence-advection test in the Tricco & Price (2012) form: B_x = [(r/r0)^8 - 2 (r/r0)^4 + 1]/sqrt(4 pi) inside
    r0 = 1/sqrt(8) of a fluid at rest.  Without cleaning (idivbzero = 0) the divergence error sits there; with the psi terms (idivbzero = 2:
    gradpsi in dB/dt, dpsi/dt = -c_h^2 div B - psi/tau; src/ratesND_mhd.f90:2712-2717, :902) it is radiated away and damped -- a wrong sign in
    either term makes it grow instead."""
    def run(idivbzero):
        o, p = setups.orszag_tang(ndim=2, nx=32, lattice="cubic", perturb_amp=0.0, evolved=False, imhd=11, idivbzero=idivbzero)
        n = p.npart
        r = np.sqrt((p.x[:n] ** 2).sum(axis=1)) * np.sqrt(8.0)
        B = np.zeros((n, 3))
        B[:, 0] = np.where(r < 1.0, (r**8 - 2.0 * r**4 + 1.0) / np.sqrt(4.0 * np.pi), 0.0)
        B[:, 2] = 1.0 / np.sqrt(4.0 * np.pi)
        p.vel[:n] = 0.0
        p.Bfield[:n] = p.Bevol[:n] = B
        p.psi[:n] = 0.0
        oracle.derivs(o, p)
        before = np.abs(p.divB[:n]).copy()
        _evolve(o, p, 0.3)
        return before, np.abs(p.divB[:n])

    b0, b1 = run(0)
    assert abs(b1.max() / b0.max() - 1.0) < 0.05 and abs(b1.mean() / b0.mean() - 1.0) < 0.05
    c0, c1 = run(2)
    assert np.array_equal(b0, c0)
    assert c1.max() < 0.6 * c0.max() and c1.mean() < 0.8 * c0.mean()
