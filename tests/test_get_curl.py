"""`get_curl` (src/get_curl.f90:64-287) and its call site on the path, the Tricco & Price resistivity switch inside
conservative2primitive (iavlim(3) = 2, src/conservative2primitive.f90:299-311).

CPU: the oracle's restatement is pinned by answers that involve no SPH code of ours -- the curl and gradient of a linear field, and an
independent numpy brute force of the published operator (analytic spline, no cells/lists/tables).  GPU: the CUDA pair-engine operator and
the switch against the oracle, 1e-12 relative to |B|/h.
"""
import numpy as np
import pytest

import parity
from ndspmhd_b200 import abi, lib, setups
from oracle import oracle


def _state(ndim=3, nx=12, **kw):
    o, p = setups.orszag_tang(ndim=ndim, nx=nx, zfrac=0.5, perturb_amp=0.15, evolved=True, **kw)
    o.device_ghosts = 1
    s, _ = oracle.derivs(o, p)
    return o, p, s


def _linear_field(p, nt):
    x = np.zeros((p.idim, 3))
    x[:nt, : p.ndim] = p.x[:nt].reshape(nt, p.ndim)
    B = np.zeros((p.idim, 3))
    B[:, 0], B[:, 1], B[:, 2] = 0.3 * x[:, 1], -0.2 * x[:, 2] + 0.1 * x[:, 0], 0.5 * x[:, 0]
    return B   # curl = (0.2, -0.5, 0.1 - 0.3); ghost rows carry the shifted coordinates, so the field is linear across the periodic faces


def test_oracle_get_curl_recovers_curl_and_gradient_of_a_linear_field():
    o, p, s = _state()
    n, nt = p.npart, p.ntotal
    B = _linear_field(p, nt)
    for ict, tol in ((1, 2e-2), (2, 0.3), (3, 8e-2), (4, 8e-2)):
        c, g = oracle.get_curl(o, p, B, ict, want_gradB=(ict == 1), hhmax=s["hhmax"])
        assert np.allclose(c[:n].mean(0), [0.2, -0.5, -0.2], atol=tol), (ict, c[:n].mean(0))
        if ict == 1:
            want = np.zeros((3, 3))
            want[0, 1], want[1, 2], want[1, 0], want[2, 0] = 0.3, -0.2, 0.1, 0.5   # gradB[i, k, l] = d B_k / d x_l
            assert np.allclose(g[:n].mean(0), want, atol=2e-3)
            # curl from the gradient: the two outputs are consistent particle by particle
            cg = np.stack([g[:n, 2, 1] - g[:n, 1, 2], g[:n, 0, 2] - g[:n, 2, 0], g[:n, 1, 0] - g[:n, 0, 1]], 1)
            assert np.allclose(cg, c[:n], atol=1e-12)


def test_oracle_get_curl_against_independent_numpy_bruteforce():
    """icurltype 1: curl_i = gradh_i/rho_i sum_j m_j (B_i - B_j) x r_ij/|r_ij| F(|r_ij|, h_i) over ALL rows j, with the analytic cubic spline
    derivative F -- no cells, lists, tables or pair symmetry.  Agreement to the interpolation error of the 4001-point table."""
    o, p, s = _state(nx=10)
    n, nt = p.npart, p.ntotal
    rng = np.random.default_rng(3)
    B = np.zeros((p.idim, 3))
    B[:n] = rng.normal(size=(n, 3))
    B[n:nt] = B[p.ireal[n:nt] - 1]
    c, g = oracle.get_curl(o, p, B, 1, want_gradB=True, hhmax=s["hhmax"])
    x = p.x[:nt]
    pi_ref = 3.141592653589                                          # src/kernelND.f90:41
    for i in rng.choice(n, 40, replace=False):
        d = x[i] - x
        r = np.sqrt((d * d).sum(1))
        q = r / p.hh[i]
        dw = np.where(q < 1, -3 * q + 2.25 * q * q, np.where(q < 2, -0.75 * (2 - q) ** 2, 0.0)) / pi_ref / p.hh[i] ** 4
        dr = np.where(r[:, None] > 0, d / np.maximum(r, 1e-300)[:, None], 0.0)
        dB = B[i] - B[:nt]
        ci = (p.pmass[:nt, None] * np.cross(dB, dr) * dw[:, None]).sum(0) * p.gradh[i] / p.rho[i]
        assert np.allclose(ci, c[i], rtol=0, atol=3e-6 * np.abs(c[:n]).max())
        gi = -(p.pmass[:nt, None, None] * dB[:, :, None] * dr[:, None, :] * dw[:, None, None]).sum(0) * p.gradh[i] / p.rho[i]
        assert np.allclose(gi, g[i], rtol=0, atol=3e-6 * np.abs(g[:n]).max())


def test_oracle_resistivity_switch_follows_its_formula():
    o, p, s = _state()
    o.iavlim[2] = 2
    q = setups.orszag_tang(ndim=3, nx=12, zfrac=0.5, perturb_amp=0.15, evolved=True)[1]
    s2, _ = oracle.derivs(o, q)
    n = q.npart
    a3 = q.alpha[:n, 2]
    assert np.all((a3 >= 0) & (a3 <= 1)) and a3.std() > 0.05
    # recompute from the operator on the converged state: alpha_B = min(h |grad B|_F / sqrt(B^2 + eps), 1)
    B = np.zeros((q.idim, 3))
    B[: q.ntotal] = q.Bfield[: q.ntotal]
    _, g = oracle.get_curl(o, q, B, 1, want_gradB=True, hhmax=s2["hhmax"])
    B2 = (q.Bfield[:n] ** 2).sum(1)
    want = np.minimum(q.hh[:n] * np.sqrt((g[:n] ** 2).sum((1, 2))) / np.sqrt(B2 + np.finfo(float).eps), 1.0)
    assert np.allclose(a3, np.where(B2 > 1e-8, want, 0.0), rtol=1e-13, atol=0)
    # the rates then use vsigB = 0.5 (vsig_i + vsig_j) + |dv.dr| (src/ratesND_mhd.f90:1436-1441): the run differs from the default switch
    assert not np.allclose(q.dBevoldt[:n], p.dBevoldt[:n])


@pytest.mark.gpu
@pytest.mark.parametrize("ndim,nx", [(3, 14), (2, 40)])
@pytest.mark.parametrize("icurltype", [1, 2, 3, 4])
def test_gpu_get_curl_operator_parity(ndim, nx, icurltype):
    o, p = setups.orszag_tang(ndim=ndim, nx=nx, zfrac=0.5, perturb_amp=0.2, evolved=True)
    o.device_ghosts = 1
    po = p.copy()
    so, _ = oracle.derivs(o, po)
    n, nt = po.npart, po.ntotal
    rng = np.random.default_rng(11)
    B = np.zeros((po.idim, 3))
    B[:n] = po.Bfield[:n] + 0.3 * rng.normal(size=(n, 3))
    B[n:nt] = B[po.ireal[n:nt] - 1]
    co, go = oracle.get_curl(o, po, B, icurltype, want_gradB=(icurltype == 1), hhmax=so["hhmax"])
    hot = lib.Hotpath(o, ndim)
    try:
        hot.upload(p)
        hot.set_linklist()
        hot.iterate_density()
        cg, gg = hot.get_curl(B, icurltype, want_gradB=(icurltype == 1))
        # the operator leaves the context usable: the rest of the derivs still matches the oracle
        hot.conservative2primitive()
        sg = hot.get_rates()
        p.ntotal = sg["ntotal"]
        hot.download(p)
    finally:
        hot.close()
    scale = np.abs(B[:n]).max() / po.hh[:n].min()
    assert np.abs(cg[:n] - co[:n]).max() <= 1e-12 * max(scale, np.abs(co[:n]).max())
    if icurltype == 1:
        assert np.abs(gg[:n] - go[:n]).max() <= 1e-12 * max(scale, np.abs(go[:n]).max())
    parity.assert_parity(p, po, sg, so, o, aux=True)


def _switch_case(name):
    if name == "periodic3d":
        return setups.orszag_tang(ndim=3, nx=14, zfrac=0.5, perturb_amp=0.2, evolved=True)
    if name == "periodic2d":
        return setups.orszag_tang(ndim=2, nx=48, lattice="cp", perturb_amp=0.2, evolved=True)
    if name == "walls3d":        # ghosts that do not get copy_particle keep the alpha_B they were created with
        return setups.reflecting_box(ndim=3, nx=12, ibound=[3, 2, 3])
    if name == "walls2d":
        return setups.reflecting_box(ndim=2, nx=40, ibound=[2, 2])
    if name == "fixed1d":        # Brio-Wu with B evolved: the fixed end particles copy alpha from their partner after the switch
        o, p = setups.shock1d(nright=60)
        o.imhd = 11
        p.Bevol[: p.npart] = p.Bfield[: p.npart]
        return o, p
    raise KeyError(name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["periodic3d", "periodic2d", "walls3d", "walls2d", "fixed1d"])
def test_gpu_resistivity_switch_parity(name):
    """iavlim(3) = 2 through the whole derivs: alpha(3,:) rewritten by conservative2primitive, the rates with the switch's vsigB."""
    o, p = _switch_case(name)
    o.device_ghosts = 1
    o.iavlim[2] = 2
    p.alpha[: p.npart, 2] = 0.3 + 0.4 * np.sin(5.0 * p.x[: p.npart].reshape(p.npart, -1)[:, 0]) ** 2    # a non-trivial alpha_B on entry
    po, pg = p.copy(), p.copy()
    so, _ = oracle.derivs(o, po)
    for pipelined in (False, True):
        q = pg.copy()
        sg = lib.derivs_host(o, q, pipelined=pipelined)
        n = q.npart
        assert np.abs(q.alpha[:n] - po.alpha[:n]).max() <= 1e-12
        assert q.alpha[:n, 2].std() > 0.01 and not np.allclose(q.alpha[:n, 2], p.alpha[:n, 2])
        parity.assert_parity(q, po, sg, so, o, aux=True)


@pytest.mark.gpu
def test_gpu_resistivity_switch_needs_evolved_B():
    o, p = setups.orszag_tang(ndim=3, nx=10, zfrac=0.5, perturb_amp=0.2, evolved=True, imhd=1, idivbzero=0)
    o.device_ghosts = 1
    o.iavlim[2] = 2
    with pytest.raises(lib.NdError) as e:
        lib.derivs_host(o, p)
    assert e.value.code == abi.ND_ERR_UNSUPPORTED_OPTION
