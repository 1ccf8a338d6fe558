"""Create / use / destroy contexts repeatedly to flush out teardown bugs: python -X faulthandler tools/stress_close.py MODE [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ndspmhd_b200 import abi, lib, setups
mode = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
o, p0 = setups.orszag_tang(ndim=3, nx=32, zfrac=0.5, perturb_amp=0.2, evolved=True)
o.device_ghosts = 1
o.want_aux = 0
for r in range(reps):
    p = p0.copy()
    hot = lib.Hotpath(o, 3, 0)
    if "h" in mode:
        hot.derivs_host(p, abi.DL_DENSITY | abi.DL_PRIM | abi.DL_RATES)
    if "c" in mode:
        os.environ["NDSPMHD_B200_RATE_CHUNKS"] = "4"
        hot.derivs_host(p, abi.DL_DENSITY | abi.DL_PRIM | abi.DL_RATES)
        os.environ.pop("NDSPMHD_B200_RATE_CHUNKS")
    if "d" in mode:
        p.ntotal = p.npart
        hot.upload(p)
        s = hot.derivs()
    if "s" in mode:
        dt = min(0.25 * s["dtforce"], 0.3 * s["dtcourant"])
        dt, _ = hot.step(dt)
        dt, _ = hot.step(dt)
    if "e" in mode:
        hot.evwrite()
    hot.close()
print("ok", mode, reps)
