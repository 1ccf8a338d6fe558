#!/usr/bin/env python
"""Golden-vector check against a dump written by a real NDSPMHD build.

    python tools/check_against_dump.py DUMPFILE [--imhd 11] [--iener 2] [--idivbzero 2] [--gamma G] [--set name=value ...]

Reads the dump (ndspmhd_b200/dumps.py), rebuilds the conserved inputs from its essential columns, runs one `derivs` through the C-ABI on
cuda:0 and compares the hot path's outputs with the dump's "information only" columns: rho, hh, pr, -drhodt/rho, divB, curlB, gradh,
force.  The run-time options are not stored in a dump (they live in the .in file, src/readwrite_infile.f90): pass the ones that differ
from src/defaults.f90.  A dump is written after the corrector of a step, so its info columns belong to the derivs of the predicted
state: the comparison is exact only for a dump written at t = 0 (`initialise` calls derivs on the dumped state) -- otherwise expect
O(dt) differences and use it as a sanity check.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dump")
    ap.add_argument("--set", action="append", default=[], help="nd_options field, e.g. --set imhd=11 --set iener=2")
    ap.add_argument("--oracle", action="store_true", help="use the CPU oracle instead of the GPU (no CUDA device needed)")
    args = ap.parse_args()
    from ndspmhd_b200 import abi, dumps
    hdr, cols = dumps.read_dump(args.dump)
    o = abi.default_options(hdr["ndim"])
    o.gamma, o.hfact = hdr["gamma"], hdr["hfact"]
    for d in range(hdr["ndim"]):
        o.ibound[d], o.xmin[d], o.xmax[d] = hdr["ibound"][d], hdr["xmin"][d], hdr["xmax"][d]
    o.imhd = 11 if hdr["imhd_in_file"] else 0
    o.device_ghosts = 1
    for kv in args.set:
        k, v = kv.split("=")
        cur = getattr(o, k)
        setattr(o, k, type(cur)(float(v)) if not isinstance(cur, int) else int(v))
    p = dumps.particles_from_dump(hdr, cols, o)
    if args.oracle:
        from oracle import oracle
        oracle.derivs(o, p)
    else:
        from ndspmhd_b200 import lib
        lib.derivs_host(o, p)
    n = hdr["npart"]
    xyz = "xyz"
    got = {"hh": p.hh[:n], "dens": p.dens[:n], "pr": p.pr[:n], "-drhodt/rho": -p.drhodt[:n] / p.rho[:n], "gradh": p.gradh[:n]}
    for d in range(3):
        got[f"f{xyz[d]}"] = p.force[:n, d]
    if hdr["imhd_in_file"]:
        got["divB"] = p.divB[:n]
        for d in range(3):
            got[f"curlB{xyz[d]}"] = p.curlB[:n, d]
    worst = 0.0
    for k, g in got.items():
        if k not in cols:
            continue
        r = cols[k][:n]
        scale = max(float(np.max(np.abs(r))), 1e-300)
        err = float(np.max(np.abs(g - r))) / scale
        worst = max(worst, err)
        print(f"{k:14s} max|diff|/max|dump| = {err:.3e}")
    print(f"worst {worst:.3e}")
    return 0 if worst <= 1e-10 else 1


if __name__ == "__main__":
    sys.exit(main())
