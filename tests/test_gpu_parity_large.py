"""GPU parity at the sizes BASELINE.json's configs state: the CUDA path through the C-ABI against the CPU oracle on the
same seeded inputs, at sizes the serial oracle still finishes in tens of seconds on the GPU box's host (it does ~0.1 M particle-updates/s).

    C2  2-D Orszag-Tang, close-packed, dp = 1/512 -> 512 x 592 = 303 104 particles (src/setup_orszagtang2D_mhd.f90:63-96)
    C3  3-D MHD box 128^3 = 2 097 152: cubic lattice and glass (lattice + 0.2 dp)
    C4  two-fluid dust + gas: fat box 100^3 + 100^3 = 2 000 000 and the reference's thin box 1 x 11dp x 11dp (src/setup_dustybox.f90:46-107)
    C4' one-fluid dust 100^3 = 1 000 000

These also take the paths only large runs take: rates in four row chunks through ndspmhd_b200_derivs_host (>= 1 Mi rows), 32-bit slot
arithmetic past 2 M rows, the neighbour-list capacity retry (forced with NDSPMHD_B200_LMAX0), and reflecting walls (ibound = 2) in 1-3-D.
Tolerances as everywhere (tests/parity.py): integers and ghost rows bit-exact, FP64 fields within 1e-12 of the summed pair terms.
"""
import os

import numpy as np
import pytest

import parity
from ndspmhd_b200 import abi, lib, setups
from oracle import oracle

pytestmark = pytest.mark.gpu


def _both(make, aux=0, pipelined=True, env=None):
    o, p = make()
    o.device_ghosts = 1
    o.want_aux = aux
    po, pg = p.copy(), p.copy()
    del p
    so, _ = oracle.derivs(o, po)
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        sg = lib.derivs_host(o, pg, pipelined=pipelined)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return o, pg, po, sg, so


def test_c2_orszag_tang_2d_512x592_closepacked():
    o, pg, po, sg, so = _both(lambda: setups.orszag_tang(ndim=2, nx=512, lattice="cp", perturb_amp=0.0, evolved=False))
    assert pg.npart == 512 * 592
    errs = parity.assert_parity(pg, po, sg, so, o, aux=False)
    assert max(errs.values()) <= parity.RTOL


@pytest.mark.parametrize("perturb_amp", [0.0, 0.2], ids=["lattice", "glass"])
def test_c3_mhd_cube_128(perturb_amp):
    o, pg, po, sg, so = _both(lambda: setups.orszag_tang(ndim=3, nx=128, cube=True, perturb_amp=perturb_amp, evolved=True))
    assert pg.npart == 128 ** 3
    assert sg["rate_chunks"] == 4                     # >= 1 Mi rows: derivs_host ran the rates in four row chunks
    errs = parity.assert_parity(pg, po, sg, so, o, aux=False)
    assert max(errs.values()) <= parity.RTOL


def test_c3_list_capacity_retry_at_full_size():
    """Same glass with the list capacity started at 40 entries a target (mean neighbour number 58): every first build overflows, the
    library grows lmax and repeats (nd_host.cuh build_lists), and the results do not change."""
    o, pg, po, sg, so = _both(lambda: setups.orszag_tang(ndim=3, nx=128, zfrac=0.25, perturb_amp=0.2, evolved=True),
                              env={"NDSPMHD_B200_LMAX0": "40"})
    assert sg["list_overflows"] >= 1 and sg["lmax"] > max(40, sg["nneigh_max"] - 1)
    parity.assert_parity(pg, po, sg, so, o, aux=False)


@pytest.mark.parametrize("aux", [1, 0])   # 0: the bench's tuple -- FAST + DRAG instantiation, LIGHT rounds, kind-split lists
def test_c4_two_fluid_fat_box_1e6_gas_1e6_dust(aux):
    o, pg, po, sg, so = _both(lambda: setups.dustybox(ndim=3, nx=100, perturb_amp=0.05), aux=aux)
    assert pg.npart == 2_000_000
    assert sg["rate_chunks"] == 4
    errs = parity.assert_parity(pg, po, sg, so, o, aux=bool(aux))
    assert max(errs.values()) <= parity.RTOL


def test_c4_two_fluid_thin_box_of_the_reference():
    """setup_dustybox's own geometry: 1 x 11dp x 11dp, dust on top of gas (coincident cross-type pairs), ghosts outnumber particles."""
    o, pg, po, sg, so = _both(lambda: setups.dustybox_thin(nx=8192), aux=1)
    assert pg.npart == 2 * 8192 * 11 * 11 and sg["ntotal"] > 1.8 * pg.npart   # 86 % ghost rows
    parity.assert_parity(pg, po, sg, so, o, aux=True)


def test_c4_one_fluid_dust_1e6():
    o, pg, po, sg, so = _both(lambda: setups.dustywave_onefluid(ndim=3, nx=100), aux=1, pipelined=False)
    assert pg.npart == 1_000_000
    parity.assert_parity(pg, po, sg, so, o, aux=True)


@pytest.mark.parametrize("ndim,nx,ibound", [(1, 0, [2]), (2, 48, [2, 2]), (2, 48, [2, 3]), (3, 14, [2, 2, 2]), (3, 14, [3, 2, 3]), (3, 14, [2, 3, 2])],
                         ids=["1d", "2d", "2d_mixed", "3d", "3d_y_walls", "3d_xz_walls"])
@pytest.mark.parametrize("mhd", [True, False], ids=["mhd", "hydro"])
def test_reflecting_walls(ndim, nx, ibound, mhd):
    """ibound = 2 (src/ghostND_mhd.f90:173, :226-228): ghost at xbound - (x - xbound) within radkern*h_i of the wall, normal velocity
    flipped; edges and corners through makeghost's recursion (:256-335).  Ghost rows are compared bit for bit."""
    o, pg, po, sg, so = _both(lambda: setups.reflecting_box(ndim=ndim, nx=nx, ibound=ibound, mhd=mhd), aux=1, pipelined=False)
    assert sg["ntotal"] == so["ntotal"] > pg.npart
    n, nt = pg.npart, sg["ntotal"]
    assert np.array_equal(pg.ireal[n:nt], po.ireal[n:nt]) and np.array_equal(pg.x[n:nt], po.x[n:nt]) and np.array_equal(pg.vel[n:nt], po.vel[n:nt])
    # at least one ghost has a flipped normal velocity component
    par = pg.ireal[n:nt] - 1
    d = ibound.index(2)
    assert np.any(pg.vel[n:nt, d] == -pg.vel[par, d])
    parity.assert_parity(pg, po, sg, so, o, aux=True)
