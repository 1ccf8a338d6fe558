"""GPU parity of the leapfrog integrator on the resident state: ndspmhd_b200_step against the oracle's
restatement of `step` (src/stepND_leapfrog_mhd.f90:39-300) and the periodic wrap of `boundary` (src/boundaryND.f90:65-93).

Both sides start from the same upload + derivs and take the same number of steps, each with the dt the previous step returned.
The evolved state must agree within RTOL = 1e-12 of each field's magnitude (the rates agree to ~1e-14 per derivs and enter the
state multiplied by dt); integer outputs of the inner derivs are not compared here because the two states differ in the last
bits after the first step (tests/test_gpu_parity.py pins them on bit-identical inputs)."""
import os

import numpy as np
import pytest

from ndspmhd_b200 import abi, lib, setups
from oracle import oracle

pytestmark = pytest.mark.gpu

RTOL = 1.0e-12
STATE = ["x", "vel", "hh", "en", "Bevol", "alpha", "psi", "rho"]

CASES = {
    "ot3d_glass": lambda: setups.orszag_tang(ndim=3, nx=16, zfrac=0.5, perturb_amp=0.2, evolved=True),
    "ot3d_isothermal": lambda: setups.orszag_tang(ndim=3, nx=16, zfrac=0.5, perturb_amp=0.2, evolved=True, iener=0),
    "ot2d_closepacked": lambda: setups.orszag_tang(ndim=2, nx=48, lattice="cp", perturb_amp=0.2, evolved=True),
    "hydro3d": lambda: setups.hydro_box(ndim=3, nx=16, perturb_amp=0.2),
    "briowu1d_fixed_ends": lambda: setups.shock1d(nright=60),
    "briowu1d_totalenergy": lambda: setups.shock1d(nright=60, iener=3),
    "dustybox3d": lambda: setups.dustybox(ndim=3, nx=10),
    "onefluid_dust3d": lambda: setups.dustywave_onefluid(ndim=3, nx=10),
}


def _dt0(s, C_cour=0.3, C_force=0.25):
    return min(C_force * s["dtforce"], C_cour * s["dtcourant"], 0.9 * s["dtdrag"], C_force * s["dtvisc"])


@pytest.mark.parametrize("name", list(CASES))
def test_step_matches_oracle(name):
    o, p = CASES[name]()
    o.device_ghosts = 1
    o.want_aux = 0
    nsteps = 3
    po, pg = p.copy(), p.copy()
    so, _ = oracle.derivs(o, po)
    dto = _dt0(so)
    assert np.isfinite(dto) and dto > 0
    for _ in range(nsteps):
        dto, so = oracle.step(o, po, dto)
    hot = lib.Hotpath(o, p.ndim, 0)
    try:
        hot.upload(pg)
        sg = hot.derivs()
        dtg = _dt0(sg)
        dts = []
        for _ in range(nsteps):
            dtg, sg = hot.step(dtg)
            dts.append(dtg)
        hot.download_state(pg)
        pg.ntotal = sg["ntotal"]
        hot.download(pg, abi.DL_DENSITY | abi.DL_PRIM | abi.DL_RATES)
    finally:
        hot.close()
    n = p.npart
    assert abs(dtg - dto) <= 1e-11 * abs(dto), (dtg, dto)
    fields = STATE + (["dustevol", "deltav"] if o.idust == 1 else [])
    errs = {}
    for f in fields:
        a, b = np.asarray(getattr(pg, f)[:n]), np.asarray(getattr(po, f)[:n])
        scale = max(float(np.max(np.abs(b))), 1e-300)
        errs[f] = float(np.max(np.abs(a - b))) / scale
    bad = {k: v for k, v in errs.items() if not v <= RTOL}
    assert not bad, f"state beyond {RTOL:g} after {nsteps} steps: {bad} (all {errs})"
    # the rates left behind by the last inner derivs feed the next step: same bar, against the force scale
    for f in ("force", "dendt"):
        a, b = np.asarray(getattr(pg, f)[:n]), np.asarray(getattr(po, f)[:n])
        scale = max(float(np.max(np.abs(b))), 1e-300)
        assert float(np.max(np.abs(a - b))) / scale <= 1e-10, f
    # particles stay inside a periodic domain (boundaryND.f90:65-93)
    for d in range(p.ndim):
        if o.ibound[d] == 3:
            assert np.all(pg.x[:n, d] >= o.xmin[d]) and np.all(pg.x[:n, d] <= o.xmax[d])


def test_step_needs_prior_derivs():
    o, p = setups.hydro_box(ndim=2, nx=16, perturb_amp=0.1)
    o.device_ghosts = 1
    hot = lib.Hotpath(o, 2, 0)
    try:
        hot.upload(p)
        with pytest.raises(lib.NdError) as ei:
            hot.step(1e-3)
        assert ei.value.code == abi.ND_ERR_STATE
    finally:
        hot.close()


def test_step_zero_dt_is_identity():
    o, p = setups.orszag_tang(ndim=2, nx=32, lattice="cp", perturb_amp=0.2, evolved=True)
    o.device_ghosts = 1
    hot = lib.Hotpath(o, 2, 0)
    try:
        hot.upload(p)
        hot.derivs()
        before = p.copy()
        hot.download_state(before)
        hot.step(0.0, dtfixed=True)
        after = p.copy()
        hot.download_state(after)
    finally:
        hot.close()
    n = p.npart
    for f in ("x", "vel", "en", "Bevol", "alpha", "psi"):
        assert np.array_equal(getattr(before, f)[:n], getattr(after, f)[:n]), f
    # h is re-converged by the inner derivs from the same positions: same answer to the iteration tolerance's last bits
    assert np.max(np.abs(before.hh[:n] - after.hh[:n]) / before.hh[:n]) < 1e-3


@pytest.mark.parametrize("name", ["ot3d_glass", "ot2d_closepacked", "hydro3d", "briowu1d_fixed_ends", "onefluid_dust3d"])
def test_evwrite_matches_oracle(name):
    """ndspmhd_b200_evwrite: the sums of `evwrite` (src/evwrite_mhd.f90:124-284) as fixed-order device reductions, against the
    oracle's serial loop on the same (bit-identical) inputs: 1e-12 of each sum's scale (sum of |terms|), extrema exact."""
    o, p = CASES[name]()
    o.device_ghosts = 1
    o.want_aux = 0
    po, pg = p.copy(), p.copy()
    oracle.derivs(o, po)
    evo = oracle.evwrite(o, po)
    hot = lib.Hotpath(o, p.ndim, 0)
    try:
        hot.upload(pg)
        hot.derivs()
        ev1 = hot.evwrite()
        ev2 = hot.evwrite()
        pg.ntotal = po.ntotal
        hot.download(pg, abi.DL_DENSITY | abi.DL_PRIM | abi.DL_RATES)
    finally:
        hot.close()
    assert ev1 == ev2                                     # deterministic reduction order
    # the oracle's sums over the GPU's own outputs isolate the reduction from the 1e-14 differences of the inputs
    for f in ("pmass", "vel", "x", "dustfrac", "deltav"):
        pg.arrays[f][...] = po.arrays[f]
    evg_in = oracle.evwrite(o, pg)
    n = p.npart
    mtot = float(np.sum(p.pmass[:n]))
    vmax = float(np.max(np.abs(po.vel[:n]))) + 1e-300
    brho = mtot * float(np.max(np.abs(po.Bfield[:n]))) / float(np.min(po.rho[:n])) + 1e-300      # scale of sum m |B|/rho: the flux sums cancel
    scale = {"mom": mtot * vmax, "dmom": mtot * (float(np.max(np.abs(po.force[:n]))) + 1e-300), "ang": mtot * vmax, "fluxtot": brho}
    for k, v in evg_in.items():
        g = ev1[k]
        if isinstance(v, list):
            s = scale[k]
            assert np.max(np.abs(np.array(g) - np.array(v))) <= 1e-12 * s, (k, g, v)
        elif k in ("rhomax", "rhomin", "divBmax"):
            assert g == v, (k, g, v)                                  # extrema of stored values: exact
        elif k in ("omegamhdmax", "betamhdmin"):
            assert abs(g - v) <= 1e-15 * abs(v), (k, g, v)            # extrema of derived values: FMA contraction in |B|^2, a few ulp
        elif k in ("momtot", "dmomtot", "angtot", "fluxtotmag"):
            assert abs(g - v) <= 1e-12 * scale[{"momtot": "mom", "dmomtot": "dmom", "angtot": "ang", "fluxtotmag": "fluxtot"}[k]], (k, g, v)
        elif k == "crosshel":
            assert abs(g - v) <= 1e-12 * brho * vmax, (k, g, v)
        else:
            assert abs(g - v) <= 1e-12 * max(abs(v), 1e-300), (k, g, v)


def test_context_teardown_after_every_api_mix():
    """Create / use / destroy contexts through every mix of entry points (pipelined host call with and without row chunks, resident
    derivs, steps, diagnostics): a regression test for a double free in the teardown that only showed after derivs_host + step."""
    o, p0 = setups.orszag_tang(ndim=3, nx=16, zfrac=0.5, perturb_amp=0.2, evolved=True)
    o.device_ghosts = 1
    o.want_aux = 0
    mask = abi.DL_DENSITY | abi.DL_PRIM | abi.DL_RATES
    for rep in range(6):
        p = p0.copy()
        hot = lib.Hotpath(o, 3, 0)
        try:
            hot.derivs_host(p, mask)
            os.environ["NDSPMHD_B200_RATE_CHUNKS"] = "3"
            try:
                hot.derivs_host(p, mask)
            finally:
                os.environ.pop("NDSPMHD_B200_RATE_CHUNKS")
            p.ntotal = p.npart
            hot.upload(p)
            s = hot.derivs()
            dt = min(0.25 * s["dtforce"], 0.3 * s["dtcourant"])
            for _ in range(2):
                dt, s = hot.step(dt)
            ev = hot.evwrite()
            assert np.isfinite(ev["etot"]) and ev["etot"] > 0
        finally:
            hot.close()
