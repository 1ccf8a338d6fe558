"""CPU checks of the oracle's leapfrog restatement (oracle/nd_oracle.cpp: ndo_step, after src/stepND_leapfrog_mhd.f90:39-300 and
src/boundaryND.f90:65-93).  The reference ships no integrator tests; these are the algorithm-independent properties of the scheme."""
import numpy as np

from ndspmhd_b200 import setups
from oracle import oracle


def _dt0(s, C_cour=0.3, C_force=0.25):
    return min(C_force * s["dtforce"], C_cour * s["dtcourant"], 0.9 * s["dtdrag"], C_force * s["dtvisc"])


def test_zero_dt_leaves_the_state():
    o, p = setups.orszag_tang(ndim=2, nx=24, lattice="cp", perturb_amp=0.2, evolved=True)
    oracle.derivs(o, p)
    q = p.copy()
    oracle.step(o, q, 0.0, dtfixed=True)
    n = p.npart
    for f in ("x", "vel", "en", "Bevol", "alpha", "psi"):
        assert np.array_equal(getattr(p, f)[:n], getattr(q, f)[:n]), f


def test_momentum_conserved_and_particles_wrapped():
    o, p = setups.hydro_box(ndim=3, nx=12, perturb_amp=0.3)
    s, _ = oracle.derivs(o, p)
    n = p.npart
    mom0 = (p.pmass[:n, None] * p.vel[:n]).sum(0)
    dt = _dt0(s)
    # enough steps at an inflated dt for some particles to leave the box
    for _ in range(4):
        dt, s = oracle.step(o, p, dt)
    mom1 = (p.pmass[:n, None] * p.vel[:n]).sum(0)
    scale = float(np.sum(p.pmass[:n]) * np.max(np.abs(p.vel[:n])))
    assert np.max(np.abs(mom1 - mom0)) <= 1e-12 * scale        # pairwise antisymmetric forces, both half-kicks
    for d in range(3):
        assert np.all(p.x[:n, d] >= o.xmin[d]) and np.all(p.x[:n, d] <= o.xmax[d])
    assert dt > 0 and np.isfinite(dt)


def test_energy_drift_is_second_order():
    """E = sum m (v^2/2 + u) for iener=2 hydro: the drift over a few steps shrinks ~4x when dt halves (second-order scheme)."""
    def drift(fac):
        o, p = setups.hydro_box(ndim=2, nx=24, perturb_amp=0.3)
        s, _ = oracle.derivs(o, p)
        n = p.npart
        e0 = float(np.sum(p.pmass[:n] * (0.5 * np.sum(p.vel[:n] ** 2, 1) + p.en[:n])))
        dt = fac * _dt0(s)
        for _ in range(int(round(4 / fac))):
            oracle.step(o, p, dt, dtfixed=True)
        e1 = float(np.sum(p.pmass[:n] * (0.5 * np.sum(p.vel[:n] ** 2, 1) + p.en[:n])))
        return abs(e1 - e0) / abs(e0)
    d1, d2 = drift(1.0), drift(0.5)
    assert d1 < 1e-3
    assert d2 < 0.6 * d1 or d1 < 1e-9


def test_fixed_particles_keep_their_properties():
    o, p = setups.shock1d(nright=60)
    s, _ = oracle.derivs(o, p)
    n = p.npart
    fixed = p.itype[:n] == 1
    assert fixed.any()
    before = {f: np.array(getattr(p, f)[:n][fixed]) for f in ("vel", "rho", "hh", "en", "Bevol", "alpha", "psi")}
    dt = _dt0(s)
    for _ in range(3):
        dt, s = oracle.step(o, p, dt)
    for f, v in before.items():
        assert np.array_equal(getattr(p, f)[:n][fixed], v), f


def test_evwrite_sums_against_numpy():
    """ndo_evwrite (src/evwrite_mhd.f90:124-284) against the same sums written independently with numpy."""
    o, p = setups.orszag_tang(ndim=3, nx=12, zfrac=0.5, perturb_amp=0.2, evolved=True)
    oracle.derivs(o, p)
    ev = oracle.evwrite(o, p)
    n = p.npart
    m, v, B, rho = p.pmass[:n], p.vel[:n], p.Bfield[:n], p.rho[:n]
    ref = {
        "ekin": 0.5 * np.sum(m * np.sum(v * v, 1)), "etherm": np.sum(m * p.uu[:n]), "emag": 0.5 * np.sum(m * np.sum(B * B, 1) / rho),
        "emagp": 0.5 * np.sum(m * (B[:, 0] ** 2 + B[:, 1] ** 2) / rho), "rhomean": np.mean(rho), "rhomax": np.max(rho), "rhomin": np.min(rho),
        "divBmax": np.max(np.abs(p.divB[:n])), "divBav": np.mean(np.abs(p.divB[:n])), "divBtot": np.sum(m * np.abs(p.divB[:n]) / rho),
        "crosshel": np.sum(m * np.sum(v * B, 1) / rho), "ekiny": 0.5 * np.sum(m * v[:, 0] ** 2),
        "betamhdav": np.mean(p.pr[:n] / (0.5 * np.sum(B * B, 1))),
        "omegamhdmax": np.max(np.abs(p.divB[:n]) * p.hh[:n] / np.sqrt(np.sum(B * B, 1))),
    }
    for k, r in ref.items():
        assert abs(ev[k] - r) <= 1e-12 * max(abs(r), 1e-300), (k, ev[k], r)
    assert abs(ev["etot"] - (ev["ekin"] + ev["emag"] + ev["etherm"])) <= 1e-15
    mom = np.sum(m[:, None] * v, 0)
    assert np.allclose(ev["mom"], mom, rtol=0, atol=1e-14 * np.sum(m) * np.max(np.abs(v)))
    ang = np.sum(m[:, None] * np.cross(p.x[:n], v), 0)
    assert np.allclose(ev["ang"], ang, rtol=0, atol=1e-13 * np.sum(m))
