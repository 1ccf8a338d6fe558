"""Manual GPU-vs-oracle check (development aid): python tools/gpu_check.py"""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from ndspmhd_b200 import setups, lib, abi
from oracle import oracle
import parity

def run(name, o, p, aux=1):
    o.device_ghosts = 1
    o.want_aux = aux
    po = p.copy(); pg = p.copy()
    t = time.time(); so, ms = oracle.derivs(o, po); tor = time.time() - t
    t = time.time(); sg = lib.derivs_host(o, pg); tg = time.time() - t
    print(f"== {name}: npart={p.npart} ntotal={so['ntotal']}/{sg['ntotal']} oracle {tor:.2f}s gpu(e2e,incl create) {tg:.2f}s its {so['itsdensity']}/{sg['itsdensity']} relink {sg['nrelink']}")
    try:
        errs = parity.assert_parity(pg, po, sg, so, o, aux=bool(aux))
        print("   PARITY OK  max err", max(errs.values()), {k: f"{v:.1e}" for k, v in errs.items()})
    except AssertionError as e:
        print("   PARITY FAIL:", str(e)[:1500])
    except Exception:
        traceback.print_exc()
    return pg, po, sg, so

if __name__ == "__main__":
    print("devices", lib.device_count())
    run("OT3D 16 lattice t=0", *setups.orszag_tang(ndim=3, nx=16, zfrac=0.5, perturb_amp=0.0, evolved=False))
    run("OT3D 16 perturbed evolved", *setups.orszag_tang(ndim=3, nx=16, zfrac=0.5, perturb_amp=0.2, evolved=True))
    run("OT3D 32 cube perturbed evolved imhd=1", *setups.orszag_tang(ndim=3, nx=24, cube=True, perturb_amp=0.3, evolved=True, imhd=1, idivbzero=0))
    run("OT2D 64 cp perturbed", *setups.orszag_tang(ndim=2, nx=64, lattice="cp", perturb_amp=0.2, evolved=True))
    run("hydro 3D 16", *setups.hydro_box(ndim=3, nx=16, perturb_amp=0.2))
    run("hydro 3D 16 noaux", *setups.hydro_box(ndim=3, nx=16, perturb_amp=0.2), aux=0)
    run("shock1D", *setups.shock1d(nright=60))
    run("dustybox 3D 12", *setups.dustybox(ndim=3, nx=12))
    run("dustybox coincident 3D 10", *setups.dustybox(ndim=3, nx=10, coincident=True))
    run("OT3D iso iener=0", *setups.orszag_tang(ndim=3, nx=16, zfrac=0.5, perturb_amp=0.2, evolved=True, iener=0))
