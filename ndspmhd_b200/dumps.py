"""Reader / writer of NDSPMHD's binary dump files: `write_dump` / `read_dump`, src/readwrite_dumps.f90:32-197, :209-440.

The format is Fortran unformatted sequential as written by the reference's build (gfortran, `-fdefault-real-8`, src/Makefile:27):
every record is framed by a 4-byte little-endian length before and after; reals are 8 bytes, integers 4.

    header : t, npart, nprint, gamma, hfact, ndim, ndimV, ncolumns, iformat, ibound(ndim), xmin(ndim), xmax(ndim), len(geom), geom   (:87-88)
    then one record per column, each `nprint` reals, in the order of :104-193, and a last record of `nprint` integers (itype, :194)

Columns (names as in the reference): x(ndim) vel(ndimV) hh dens|rho uu pmass, then
    MHD (iformat 2):   alpha(3) Bfield(ndimV) psi | pr -drhodt/rho divB curlB(ndimV) gradh force(ndimV)
    hydro (iformat 1): alpha(2)                   | pr -drhodt/rho gradh force(ndimV)
    one-fluid dust (iformat 5) appends dustfrac deltav(ndimV) rhogas rhodust.
The part after `|` is "for information only": those are exactly the hot path's outputs, which makes any dump written by a real NDSPMHD
build a set of golden vectors for rho, h, P, drho/dt, div B, curl B, gradh and the force (tools/check_against_dump.py).
Host-side only: nothing here touches the GPU.
"""
from __future__ import annotations

import struct

import numpy as np

NDIMV = 3


def _rec(f, payload: bytes) -> None:
    f.write(struct.pack("<i", len(payload)))
    f.write(payload)
    f.write(struct.pack("<i", len(payload)))


def _read_rec(f) -> bytes:
    head = f.read(4)
    if len(head) < 4:
        raise EOFError("end of dump")
    (n,) = struct.unpack("<i", head)
    payload = f.read(n)
    (m,) = struct.unpack("<i", f.read(4))
    if m != n or len(payload) != n:
        raise ValueError(f"corrupt record: markers {n} / {m}, read {len(payload)}")
    return payload


def column_names(ndim: int, imhd: int, onef_dust: bool = False) -> list[str]:
    """Column order of src/readwrite_dumps.f90:104-193 (cartesian, no self-gravity, no del2v, imhd >= 0)."""
    xyz = "xyz"
    cols = [f"x{xyz[d]}" for d in range(ndim)] + [f"v{xyz[d]}" for d in range(NDIMV)] + ["hh", "rho" if onef_dust else "dens", "uu", "pmass"]
    if imhd != 0:
        cols += ["alpha", "alphau", "alphaB"] + [f"B{xyz[d]}" for d in range(NDIMV)] + ["psi", "pr", "-drhodt/rho", "divB"]
        cols += [f"curlB{xyz[d]}" for d in range(NDIMV)] + ["gradh"] + [f"f{xyz[d]}" for d in range(NDIMV)]
    else:
        cols += ["alpha", "alphau", "pr", "-drhodt/rho", "gradh"] + [f"f{xyz[d]}" for d in range(NDIMV)]
    if onef_dust:
        cols += ["dustfrac"] + [f"deltav{xyz[d]}" for d in range(NDIMV)] + ["rhogas", "rhodust"]
    return cols


def ncolumns(ndim: int, imhd: int, onef_dust: bool = False) -> int:
    """`ncolumns` as the reference computes it (:65-86)."""
    n = ndim + 2 * NDIMV + 4
    n += (8 + 2 * NDIMV) if imhd != 0 else 5
    if onef_dust:
        n += NDIMV + 2 * 1 + 1
    return n


def write_dump(path: str, t: float, opts, p, nprint: int | None = None, geom: str = "cartesian") -> int:
    """Writes the arrays of a Particles container (after a derivs / step + download) exactly as `write_dump` would.  Returns nprint."""
    ndim, npart = p.ndim, p.npart
    nprint = npart if nprint is None else nprint
    imhd, onef = int(opts.imhd), bool(opts.onef_dust)
    iformat = 5 if onef else (2 if imhd != 0 else 1)
    geom12 = geom.ljust(12)[:12].encode()
    r8 = lambda a: np.ascontiguousarray(a, dtype="<f8").tobytes()
    with open(path, "wb") as f:
        hdr = struct.pack("<dii", float(t), npart, nprint) + struct.pack("<dd", float(opts.gamma), float(opts.hfact))
        hdr += struct.pack("<iiii", ndim, NDIMV, ncolumns(ndim, imhd, onef), iformat)
        hdr += struct.pack(f"<{ndim}i", *[int(opts.ibound[d]) for d in range(ndim)])
        hdr += struct.pack(f"<{ndim}d", *[float(opts.xmin[d]) for d in range(ndim)]) + struct.pack(f"<{ndim}d", *[float(opts.xmax[d]) for d in range(ndim)])
        hdr += struct.pack("<i", 12) + geom12
        _rec(f, hdr)
        n = nprint
        for d in range(ndim):
            _rec(f, r8(p.x[:n, d] if ndim > 1 else p.x[:n].reshape(n)))
        for d in range(NDIMV):
            _rec(f, r8(p.vel[:n, d]))
        _rec(f, r8(p.hh[:n]))
        _rec(f, r8(p.rho[:n] if onef else p.dens[:n]))
        _rec(f, r8(p.uu[:n]))
        _rec(f, r8(p.pmass[:n]))
        for d in range(3 if imhd != 0 else 2):
            _rec(f, r8(p.alpha[:n, d]))
        if imhd != 0:
            for d in range(NDIMV):
                _rec(f, r8(p.Bfield[:n, d]))
            _rec(f, r8(p.psi[:n]))
        _rec(f, r8(p.pr[:n]))
        _rec(f, r8(-p.drhodt[:n] / p.rho[:n]))
        if imhd != 0:
            _rec(f, r8(p.divB[:n]))
            for d in range(NDIMV):
                _rec(f, r8(p.curlB[:n, d]))
        _rec(f, r8(p.gradh[:n]))
        for d in range(NDIMV):
            _rec(f, r8(p.force[:n, d]))
        if onef:
            _rec(f, r8(p.dustfrac[:n]))
            for d in range(NDIMV):
                _rec(f, r8(p.deltav[:n, d]))
            _rec(f, r8(p.rhogas[:n]))
            _rec(f, r8(p.rhodust[:n]))
        _rec(f, np.ascontiguousarray(p.itype[:n], dtype="<i4").tobytes())
    return nprint


def read_dump(path: str):
    """Returns (header, columns): header is a dict of the :87-88 fields, columns maps the names of column_names() (+ 'itype') to arrays."""
    with open(path, "rb") as f:
        h = _read_rec(f)
        t, npart, nprint = struct.unpack_from("<dii", h, 0)
        gamma, hfact = struct.unpack_from("<dd", h, 16)
        ndim, ndimv, ncol, iformat = struct.unpack_from("<iiii", h, 32)
        if not (1 <= ndim <= 3) or ndimv != NDIMV:
            raise ValueError(f"not an NDSPMHD double-precision dump: ndim={ndim} ndimV={ndimv} (readwrite_dumps.f90:287-294)")
        off = 48
        ibound = list(struct.unpack_from(f"<{ndim}i", h, off)); off += 4 * ndim
        xmin = list(struct.unpack_from(f"<{ndim}d", h, off)); off += 8 * ndim
        xmax = list(struct.unpack_from(f"<{ndim}d", h, off)); off += 8 * ndim
        (lg,) = struct.unpack_from("<i", h, off); off += 4
        geom = h[off:off + lg].decode(errors="replace").strip()
        header = {"t": t, "npart": npart, "nprint": nprint, "gamma": gamma, "hfact": hfact, "ndim": ndim, "ndimV": ndimv, "ncolumns": ncol,
                  "iformat": iformat, "ibound": ibound, "xmin": xmin, "xmax": xmax, "geom": geom}
        if not geom.startswith("cart"):
            raise ValueError(f"geometry {geom!r}: only cartesian dumps are read")
        onef = iformat == 5
        if onef:
            # the reference writes iformat = 5 for one-fluid dust with or without the MHD columns (:65-86): only the column
            # count tells the two layouts apart
            if ncol == ncolumns(ndim, 1, True):
                imhd = 1
            elif ncol == ncolumns(ndim, 0, True):
                imhd = 0
            else:
                raise ValueError(f"iformat 5 dump with {ncol} columns: neither the MHD ({ncolumns(ndim, 1, True)}) nor the hydro "
                                 f"({ncolumns(ndim, 0, True)}) one-fluid dust layout")
        else:
            imhd = 1 if iformat in (2, 4) else 0
        names = column_names(ndim, imhd, onef)
        if not onef and ncol != len(names):
            raise ValueError(f"iformat {iformat} dump with {ncol} columns, expected {len(names)} (extra columns: igravity, del2v, imhd<0?)")
        cols = {}
        for nm in names:
            rec = _read_rec(f)
            if len(rec) != 8 * nprint:
                raise ValueError(f"column {nm}: {len(rec)} bytes, expected {8 * nprint} (a dump with extra columns: igravity, del2v, imhd<0?)")
            cols[nm] = np.frombuffer(rec, dtype="<f8").copy()
        rec = _read_rec(f)
        if len(rec) != 4 * nprint:
            raise ValueError(f"itype record: {len(rec)} bytes, expected {4 * nprint} (a real-valued column read as itype?)")
        cols["itype"] = np.frombuffer(rec, dtype="<i4").copy()
        header["imhd_in_file"] = imhd
        header["onef_dust_in_file"] = onef
    return header, cols


def particles_from_dump(header: dict, cols: dict, opts, idim: int | None = None):
    """Rebuilds the hot path's inputs from a dump, the way `read_dump` + `primitive2conservative` do for the first-class tuple
    (src/readwrite_dumps.f90:365-440, src/conservative2primitive.f90:481-560): rho = dens, en = uu (iener 2) or the total energy (3),
    Bevol = B (imhd 11) or B/rho (imhd 1).  Only the real rows are kept; ghosts are regenerated by the library."""
    from .abi import Particles
    ndim, n = header["ndim"], header["npart"]
    p = Particles(ndim, n, idim or (n + n // 2 + 4096))
    xyz = "xyz"
    x = np.stack([cols[f"x{xyz[d]}"][:n] for d in range(ndim)], 1)
    p.x[:n] = x if ndim > 1 else x.reshape(p.x[:n].shape)
    p.vel[:n] = np.stack([cols[f"v{xyz[d]}"][:n] for d in range(3)], 1)
    p.hh[:n], p.pmass[:n] = cols["hh"][:n], cols["pmass"][:n]
    rho = cols["rho" if header["onef_dust_in_file"] else "dens"][:n]
    p.rho[:n] = rho
    p.itype[:n] = cols["itype"][:n]
    uu = cols["uu"][:n]
    p.alpha[:n, 0], p.alpha[:n, 1] = cols["alpha"][:n], cols["alphau"][:n]
    B = np.zeros((n, 3))
    if header["imhd_in_file"]:
        p.alpha[:n, 2] = cols["alphaB"][:n]
        B = np.stack([cols[f"B{xyz[d]}"][:n] for d in range(3)], 1)
        p.psi[:n] = cols["psi"][:n]
    if opts.imhd >= 11:
        p.Bevol[:n] = B
    elif opts.imhd > 0:
        p.Bevol[:n] = B / rho[:, None]
    if opts.iener == 3:
        p.en[:n] = 0.5 * np.sum(p.vel[:n] ** 2, 1) + uu + 0.5 * np.sum(B * B, 1) / rho
    else:
        p.en[:n] = uu
    p.ntotal = n
    return p
