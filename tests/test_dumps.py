"""Dump reader / writer: the file layout of src/readwrite_dumps.f90:32-197 and the golden-vector route of
tools/check_against_dump.py.  Host-side only."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from ndspmhd_b200 import dumps, setups
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _state(make):
    o, p = make()
    oracle.derivs(o, p)
    return o, p


@pytest.mark.parametrize("make,imhd,ndim", [
    (lambda: setups.orszag_tang(ndim=3, nx=10, zfrac=0.5, perturb_amp=0.2, evolved=True), 11, 3),
    (lambda: setups.orszag_tang(ndim=2, nx=16, lattice="cp", perturb_amp=0.2, evolved=True), 11, 2),
    (lambda: setups.hydro_box(ndim=3, nx=8, perturb_amp=0.2), 0, 3),
    (lambda: setups.shock1d(nright=40), 1, 1),
])
def test_dump_roundtrip_and_layout(tmp_path, make, imhd, ndim):
    o, p = _state(make)
    path = str(tmp_path / "test_00000.dat")
    n = dumps.write_dump(path, 0.25, o, p)
    raw = open(path, "rb").read()
    # Fortran sequential framing: first record = header, markers equal (readwrite_dumps.f90:87-88)
    (m0,) = struct.unpack_from("<i", raw, 0)
    assert m0 == 8 + 4 + 4 + 8 + 8 + 16 + 4 * ndim + 16 * ndim + 4 + 12
    assert struct.unpack_from("<i", raw, 4 + m0)[0] == m0
    hdr, cols = dumps.read_dump(path)
    assert hdr["t"] == 0.25 and hdr["npart"] == p.npart and hdr["nprint"] == n and hdr["ndim"] == ndim
    assert hdr["ncolumns"] == dumps.ncolumns(ndim, o.imhd, False) == len(dumps.column_names(ndim, o.imhd, False))   # :65-75
    assert hdr["iformat"] == (2 if o.imhd != 0 else 1)
    assert hdr["geom"].startswith("cart") and hdr["ibound"] == [int(o.ibound[d]) for d in range(ndim)]
    # total size = header + ncolumns real records + the itype record
    assert len(raw) == (m0 + 8) + hdr["ncolumns"] * (8 * n + 8) + (4 * n + 8)
    assert np.array_equal(cols["hh"], p.hh[:n]) and np.array_equal(cols["pmass"], p.pmass[:n]) and np.array_equal(cols["itype"], p.itype[:n])
    assert np.array_equal(cols["fx"], p.force[:n, 0]) and np.array_equal(cols["-drhodt/rho"], -p.drhodt[:n] / p.rho[:n])
    if o.imhd != 0:
        assert np.array_equal(cols["By"], p.Bfield[:n, 1]) and np.array_equal(cols["divB"], p.divB[:n])


def test_particles_from_dump_restart_reproduces_the_outputs(tmp_path):
    """write_dump -> read_dump -> primitive2conservative -> derivs gives back the dumped information columns: the route by which a dump
    from a real NDSPMHD build pins the hot path (here exercised on the oracle's own dump, so it checks the plumbing)."""
    o, p = _state(lambda: setups.orszag_tang(ndim=3, nx=10, zfrac=0.5, perturb_amp=0.2, evolved=True))
    path = str(tmp_path / "ot_00000.dat")
    dumps.write_dump(path, 0.0, o, p)
    hdr, cols = dumps.read_dump(path)
    q = dumps.particles_from_dump(hdr, cols, o)
    # the dump holds the CONVERGED h: a restart converges at once and reproduces rho, P, force, div B to round-off of the h iteration
    oracle.derivs(o, q)
    n = p.npart
    for f, tol in (("rho", 1e-3), ("pr", 1e-3), ("divB", 5e-2)):
        a, b = getattr(q, f)[:n], getattr(p, f)[:n]
        assert np.max(np.abs(a - b)) <= tol * np.max(np.abs(b)), f


def test_check_against_dump_tool_with_the_oracle(tmp_path):
    o, p = _state(lambda: setups.hydro_box(ndim=2, nx=16, perturb_amp=0.1))
    path = str(tmp_path / "hydro_00000.dat")
    dumps.write_dump(path, 0.0, o, p)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_against_dump.py"), path, "--oracle", "--set", "iener=2"],
                       capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert "worst" in r.stdout, r.stdout + r.stderr


def test_read_rejects_foreign_files(tmp_path):
    path = str(tmp_path / "junk.dat")
    with open(path, "wb") as f:
        f.write(struct.pack("<i", 48) + b"\0" * 48 + struct.pack("<i", 48))
    with pytest.raises(ValueError):
        dumps.read_dump(path)


def test_onefluid_dust_mhd_dump_is_told_apart_by_its_column_count(tmp_path):
    """iformat = 5 is written for one-fluid dust with AND without the 14 MHD columns (readwrite_dumps.f90:65-86): the reader decides by
    ncolumns, checks the itype record's length, and refuses a column count that fits neither layout."""
    o, p = _state(lambda: setups.dustywave_onefluid(ndim=3, nx=8, mhd=True))
    path = str(tmp_path / "dust_00000.dat")
    n = dumps.write_dump(path, 0.0, o, p)
    hdr, cols = dumps.read_dump(path)
    assert hdr["iformat"] == 5 and hdr["imhd_in_file"] == 1 and hdr["ncolumns"] == dumps.ncolumns(3, 1, True)
    assert np.array_equal(cols["Bz"], p.Bfield[:n, 2]) and np.array_equal(cols["dustfrac"], p.dustfrac[:n])
    assert cols["itype"].dtype == np.int32 and np.array_equal(cols["itype"], p.itype[:n])
    o2, p2 = _state(lambda: setups.dustywave_onefluid(ndim=3, nx=8, mhd=False))
    path2 = str(tmp_path / "dusth_00000.dat")
    dumps.write_dump(path2, 0.0, o2, p2)
    hdr2, cols2 = dumps.read_dump(path2)
    assert hdr2["iformat"] == 5 and hdr2["imhd_in_file"] == 0 and "Bx" not in cols2
    raw = bytearray(open(path, "rb").read())
    struct.pack_into("<i", raw, 4 + 32 + 8, hdr["ncolumns"] + 1)   # ncolumns sits after t, npart, nprint, gamma, hfact, ndim, ndimV
    bad = str(tmp_path / "bad_00000.dat")
    open(bad, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        dumps.read_dump(bad)
