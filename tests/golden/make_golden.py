#!/usr/bin/env python
"""Writes the golden fixtures of tests/golden/: python tests/golden/make_golden.py

What they are -- and are not.  The reference (Fortran 90) cannot be built in this image and ships no expected outputs, so these are
NOT outputs of NDSPMHD.  Each file holds one small seeded case of the hot path: the input particle state, and the state and scalars the
CPU oracle (oracle/nd_oracle.cpp, the restatement of the reference's loops) returns for one `derivs`.  They serve two purposes:
(1) a drift alarm for the oracle itself (tests/test_golden.py re-runs it and must reproduce them), so that what the CUDA path is compared
with cannot change silently between rounds; (2) a parity target for the CUDA path that does not execute the oracle at test time.
A dump from a real NDSPMHD build can be turned into the same kind of target with tools/check_against_dump.py.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from ndspmhd_b200 import setups  # noqa: E402
from oracle import oracle  # noqa: E402

CASES = {
    # BASELINE.json configs[0]: 1-D Brio-Wu tube with fixed ends
    "briowu1d": (lambda: setups.shock1d(nright=40), 1),
    # configs[1]: 2-D Orszag-Tang on the close-packed lattice
    "ot2d_closepacked": (lambda: setups.orszag_tang(ndim=2, nx=32, lattice="cp", perturb_amp=0.2, evolved=True), 1),
    # configs[2]/[4]: 3-D MHD Orszag-Tang slab, the first-class option tuple (want_aux = 0: the FAST kernels, LIGHT density rounds)
    "ot3d_glass_fast": (lambda: setups.orszag_tang(ndim=3, nx=12, zfrac=0.5, perturb_amp=0.25, evolved=True), 0),
    # configs[3]: two-fluid dust + gas
    "dustybox3d": (lambda: setups.dustybox(ndim=3, nx=10), 1),
}


def make_case(name):
    make, aux = CASES[name]
    o, p = make()
    o.device_ghosts = 1
    o.want_aux = aux
    pin = p.copy()
    s, _ = oracle.derivs(o, p)
    nt = int(s["ntotal"])
    out = {"opts": np.frombuffer(bytes(o), dtype=np.uint8), "npart": np.int64(p.npart), "ntotal": np.int64(nt), "ndim": np.int64(p.ndim),
           "aux": np.int64(aux)}
    for k, a in pin.arrays.items():
        out["in_" + k] = a[: pin.npart]
    for k, a in p.arrays.items():
        out["out_" + k] = a[:nt]
    scal = {k: (v.tolist() if isinstance(v, np.ndarray) else (list(v) if isinstance(v, (list, tuple)) else v)) for k, v in s.items()}
    out["scalars"] = np.frombuffer(json.dumps(scal, default=float).encode(), dtype=np.uint8)
    return out


if __name__ == "__main__":
    for name in CASES:
        d = make_case(name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(f"{name}: npart={int(d['npart'])} ntotal={int(d['ntotal'])} -> {os.path.getsize(path) / 1024:.0f} KB")
