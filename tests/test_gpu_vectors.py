"""GPU tests added after the round's last GPU session (they sort after the suites that have run on a B200, so that `pytest -x` reaches those
first): parity against the frozen golden fixtures without executing the oracle, the two routes of drho/dt of the fast option tuple, row chunks
on the LIGHT route, 1-D periodic waves (ghosts in one dimension), and the known-answer runs of tests/test_oracle_physics.py stepped on the
device."""
import numpy as np
import pytest

import parity
from ndspmhd_b200 import abi, lib, setups
from oracle import oracle
from test_golden import NAMES, load_case
from test_gpu_parity import CASES, run_both
from test_gpu_step import _dt0

# none of these has run on a GPU yet: a stuck kernel must end the pytest process (and free the device), not hold the box
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]


@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_matches_the_golden_outputs(name):
    o, pin, pout, scal, aux = load_case(name)
    pg = pin.copy()
    sg = lib.derivs_host(o, pg)
    errs = parity.assert_parity(pg, pout, sg, scal, o, aux=bool(aux))
    assert max(errs.values()) <= parity.RTOL


def test_fast_tuple_fused_derivs_agree_with_phase_by_phase_calls():
    """Fast option tuple (want_aux=0): a fused derivs runs the density rounds LIGHT and takes drho/dt from the pair sums of
    get_rates; the phase-by-phase calls keep drho/dt in the density sums (src/density_sums.f90:297-303).  Same pairs, same
    grad W, different order of summation: everything made before the rates is bit-equal, drho/dt and what is built on it
    agree to the parity tolerance."""
    o, p = setups.orszag_tang(ndim=3, nx=16, zfrac=0.5, perturb_amp=0.2, evolved=True)
    o.device_ghosts = 1
    o.want_aux = 0
    p1, p2 = p.copy(), p.copy()
    s1 = lib.derivs_host(o, p1)
    hot = lib.Hotpath(o, 3)
    try:
        hot.upload(p2)
        hot.set_linklist()
        sd = hot.iterate_density()
        hot.conservative2primitive()
        s2 = hot.get_rates()
        p2.ntotal = sd["ntotal"]
        hot.download(p2)
    finally:
        hot.close()
    n = p.npart
    for f in ["hh", "rho", "gradh"] + parity.PRIM_FIELDS + ["force", "gradpsi", "divB", "curlB"]:
        assert np.array_equal(getattr(p1, f)[:n], getattr(p2, f)[:n]), f
    scales = parity.natural_scales(p2, n)
    for f in ["drhodt", "dhdt", "dudt", "dendt", "dBevoldt", "daldt", "dpsidt"]:
        err = parity.field_error(getattr(p1, f)[:n], getattr(p2, f)[:n], scales[f])
        assert err <= parity.RTOL, (f, err)
    for k in ("dtcourant", "dtforce", "vsigmax", "itsdensity", "ntotal"):
        assert s1[k] == s2[k]


@pytest.mark.parametrize("name,chunks", [("ot3d_glass_noaux", 3), ("briowu1d_noaux", 4)])
def test_row_chunked_rates_on_the_light_route_equal_the_single_launch(name, chunks, monkeypatch):
    """test_row_chunked_rates_equal_the_single_launch for the fast tuple: LIGHT density rounds, drho/dt and dh/dt made by the finalisation of
    each row chunk."""
    o, p = CASES[name][0]()
    o.device_ghosts = 1
    o.want_aux = 0
    a, b = p.copy(), p.copy()
    hot = lib.Hotpath(o, p.ndim)
    try:
        monkeypatch.setenv("NDSPMHD_B200_RATE_CHUNKS", "1")
        sa = lib.derivs_host(o, a, hot=hot, pipelined=True)
        monkeypatch.setenv("NDSPMHD_B200_RATE_CHUNKS", str(chunks))
        sb = lib.derivs_host(o, b, hot=hot, pipelined=True)
    finally:
        hot.close()
    for f in parity.DENSITY_FIELDS + parity.PRIM_FIELDS + parity.RATES_FIELDS:
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    for k in ("dtcourant", "dtforce", "dtav", "vsigmax", "fhmax", "itsdensity", "ntotal", "nclumped"):
        assert sa[k] == sb[k], k


@pytest.mark.parametrize("kind,aux", [("sound", 1), ("alfven", 1), ("alfven", 0)])
def test_periodic_1d_wave_parity(kind, aux):
    """One dimension with periodic ghosts (src/setup_wave_x_ND.f90 geometry): the shock tubes of the parity suite have fixed ends."""
    from test_oracle_physics import _wave1d
    o, pg, po, sg, so = run_both(lambda: _wave1d(64, kind, 0.05), aux)
    errs = parity.assert_parity(pg, po, sg, so, o, aux=bool(aux))
    assert max(errs.values()) <= parity.RTOL


def test_sod_tube_on_the_device_lands_on_the_exact_riemann_solution():
    """The same known-answer run as tests/test_oracle_physics.py, but stepped on the GPU from upload to t = 0.15 (~500 leapfrog steps on
    the resident state, no particle traffic in between): plateaus of the exact Riemann solution to 1 %, and the oracle's end state."""
    from test_oracle_physics import _evolve, exact_sod
    tmax = 0.15
    o, p = setups.shock1d(nright=60, mhd=False, iener=2)
    o.device_ghosts = 1
    o.want_aux = 0
    po, pg = p.copy(), p.copy()
    nsteps_oracle = _evolve(o, po, tmax)
    hot = lib.Hotpath(o, 1, 0)
    try:
        hot.upload(pg)
        sg = hot.derivs()
        dt, t, nsteps = _dt0(sg), 0.0, 0
        while t < tmax:
            dt = min(dt, tmax - t)
            dtnew, sg = hot.step(dt)
            t += dt
            dt = dtnew
            nsteps += 1
            assert nsteps < 20000
        hot.download_state(pg)
        pg.ntotal = sg["ntotal"]
        hot.download(pg, abi.DL_DENSITY | abi.DL_PRIM | abi.DL_RATES)
    finally:
        hot.close()
    assert abs(nsteps - nsteps_oracle) <= 1
    n = p.npart
    x, rho, v, pr = pg.x[:n, 0], pg.rho[:n], pg.vel[:n, 0], pg.pr[:n]
    e = exact_sod(1.0, 1.0, 0.125, 0.1, o.gamma)
    for xa, xb, rho_exact in ((e["v"] * tmax, e["s_shock"] * tmax, e["rho_r"]), (e["s_tail"] * tmax, e["v"] * tmax, e["rho_l"])):
        m = (x > xa + 0.25 * (xb - xa)) & (x < xb - 0.25 * (xb - xa))
        assert m.sum() >= 8
        assert abs(np.median(rho[m]) / rho_exact - 1.0) < 0.01
        assert abs(np.median(v[m]) / e["v"] - 1.0) < 0.01
        assert abs(np.median(pr[m]) / e["p"] - 1.0) < 0.01
    for f in ("x", "vel", "rho", "en"):
        a, b = np.asarray(getattr(pg, f)[:n]), np.asarray(getattr(po, f)[:n])
        assert float(np.max(np.abs(a - b))) <= 1e-6 * max(float(np.max(np.abs(b))), 1e-300), f


def test_dustybox_on_the_device_relaxes_at_the_analytic_rate():
    """DUSTYBOX (tests/test_oracle_physics.py) stepped on the device: exp(-K (1/rho_g + 1/rho_d) t)."""
    K, tmax = 1.0, 0.5
    o, p = setups.dustybox(ndim=3, nx=8, perturb_amp=0.0, Kdrag=K)
    o.device_ghosts = 1
    o.want_aux = 0
    n = p.npart
    gas = p.itype[:n] == 0
    p.vel[:n] = 0.0
    p.vel[:n, 0][gas] = 1.0
    hot = lib.Hotpath(o, 3, 0)
    try:
        hot.upload(p)
        sg = hot.derivs()
        dt, t, nsteps = _dt0(sg), 0.0, 0
        while t < tmax:
            dt = min(dt, tmax - t)
            dtnew, sg = hot.step(dt)
            t += dt
            dt = dtnew
            nsteps += 1
            assert nsteps < 2000
        hot.download_state(p)
    finally:
        hot.close()
    vg, vd = p.vel[:n, 0][gas], p.vel[:n, 0][~gas]
    rg, rd = float(p.rho[:n][gas].mean()), float(p.rho[:n][~gas].mean())
    assert abs((vg.mean() - vd.mean()) / np.exp(-K * (1.0 / rg + 1.0 / rd) * tmax) - 1.0) < 0.01
    assert abs(0.5 * (vg.mean() + vd.mean()) - 0.5) < 1e-12


@pytest.mark.parametrize("pipelined", [False, True])
def test_real_rows_modifier_leaves_ghost_rows_of_the_outputs_alone(pipelined):
    """ND_DL_REAL_ROWS: rows [0,npart) of the output arrays equal the full download bit for bit; rows [npart,ntotal) are not written."""
    o, p = setups.orszag_tang(ndim=3, nx=16, zfrac=0.5, perturb_amp=0.2, evolved=True)
    o.device_ghosts = 1
    o.want_aux = 0
    mask = abi.DL_DENSITY | abi.DL_PRIM | abi.DL_RATES
    a, b = p.copy(), p.copy()
    sentinel = -7.25
    for f in ("rho", "pr", "force", "divB", "drhodt"):
        getattr(b, f)[p.npart:] = sentinel
    hot = lib.Hotpath(o, 3)
    try:
        if pipelined:
            sa = hot.derivs_host(a, mask)
            sb = hot.derivs_host(b, mask | abi.DL_REAL_ROWS)
        else:
            hot.upload(a)
            sa = hot.derivs()
            a.ntotal = sa["ntotal"]
            hot.download(a, mask)
            b.ntotal = sa["ntotal"]
            hot.download(b, mask | abi.DL_REAL_ROWS)
            sb = sa
    finally:
        hot.close()
    n, nt = p.npart, sa["ntotal"]
    assert nt > n and sb["ntotal"] == nt
    for f in parity.DENSITY_FIELDS + parity.PRIM_FIELDS + [x for x in parity.RATES_FIELDS if x not in ("graddivv", "del2u")]:
        assert np.array_equal(getattr(a, f)[:n], getattr(b, f)[:n]), f
    for f in ("rho", "pr", "force", "divB", "drhodt"):
        assert np.all(getattr(b, f)[n:nt] == sentinel), f
