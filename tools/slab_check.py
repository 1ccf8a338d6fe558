"""torchrun --nproc-per-node N tools/slab_check.py [nx] [nccl|callbacks]: slab-decomposed derivs on N GPUs vs the single-GPU result and the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from ndspmhd_b200 import abi, lib, setups, slab
import parity

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    transport = sys.argv[2] if len(sys.argv) > 2 else "nccl"      # "nccl": the library's own NCCL calls; "callbacks": nd_comm over torch.distributed
    ok = True
    for name, kw in [("ot3d_glass", dict(ndim=3, nx=nx, zfrac=0.25, perturb_amp=0.2, evolved=True)),
                     ("ot3d_small_h", dict(ndim=3, nx=nx, zfrac=0.25, perturb_amp=0.3, evolved=True)),
                     ("ot2d", dict(ndim=2, nx=4 * nx, lattice="cp", perturb_amp=0.2, evolved=True))]:
        o, pl, info = setups.orszag_tang(slab=(rank, world), **kw)
        o.device_ghosts = 1; o.want_aux = 0
        if name == "ot3d_small_h":
            pl.hh[: pl.npart] *= 0.6          # forces relinks (h grows past hhmax) and several `density` rounds over all particles
        hot = lib.Hotpath(o, pl.ndim, local)
        if transport == "nccl":
            slab.attach_nccl(hot, rank, world, float(info["edges"][rank]), float(info["edges"][rank + 1]), int(info["nglobal"]))
        else:
            comm = slab.SlabComm(device="cuda")
            slab.attach(hot, comm, float(info["edges"][rank]), float(info["edges"][rank + 1]), int(info["nglobal"]))
        hot.upload(pl)
        s = hot.derivs()
        hot.download(pl)
        nown, nsrc, nt = slab.row_counts(hot)
        # the pipelined host call on the same context, rates in three row chunks (their rows go down while the next chunk runs): must equal
        # the resident path bit for bit -- every target's sums are its own
        o2, pl2, _ = setups.orszag_tang(slab=(rank, world), **kw)
        if name == "ot3d_small_h":
            pl2.hh[: pl2.npart] *= 0.6
        os.environ["NDSPMHD_B200_RATE_CHUNKS"] = "3"
        try:
            s2 = hot.derivs_host(pl2, abi.DL_DENSITY | abi.DL_PRIM | abi.DL_RATES)
        finally:
            os.environ.pop("NDSPMHD_B200_RATE_CHUNKS", None)
        chunk_ok = all(np.array_equal(np.asarray(pl2.arrays[f][: pl.npart]), np.asarray(pl.arrays[f][: pl.npart]))
                       for f in ("rho", "hh", "force", "dBevoldt", "dendt", "dpsidt", "divB", "curlB", "daldt", "drhodt", "dhdt", "numneigh")) \
            and s2["dtcourant"] == s["dtcourant"] and s2["npairs_rates"] == s["npairs_rates"] and s2["rate_chunks"] == 3
        hot.close()
        # single-GPU reference on every rank (same device), compare own rows
        og, pg = setups.orszag_tang(**kw)
        og.device_ghosts = 1; og.want_aux = 0
        if name == "ot3d_small_h":
            pg.hh[: pg.npart] *= 0.6
        sg = lib.derivs_host(og, pg, device=local)
        rows = info["rows"]
        n = pl.npart
        class V:  # view of the single-GPU arrays restricted to my rows
            pass
        ref = V(); ref.npart = n
        for k, v in pg.arrays.items():
            setattr(ref, k, v[rows])
        loc = V(); loc.npart = n
        for k, v in pl.arrays.items():
            setattr(loc, k, v[:n])
        fields = parity.DENSITY_FIELDS + parity.PRIM_FIELDS + [f for f in parity.RATES_FIELDS if f != "graddivv"]
        ns = parity.natural_scales(pg, pg.npart)
        errs = {f: parity.field_error(getattr(loc, f), getattr(ref, f), ns.get(f, 0.0)) for f in fields}
        bad = {k: v for k, v in errs.items() if not v <= parity.RTOL}
        same_nn = np.array_equal(loc.numneigh, ref.numneigh)
        sc_ok = all(abs(s[k] - sg[k]) <= 1e-11 * abs(sg[k]) for k in ("dtcourant", "dtforce", "dtav", "vsigmax", "stressmax", "fhmax", "hhmax")) \
            and s["itsdensity"] == sg["itsdensity"] and s["nneigh_min"] == sg["nneigh_min"] and s["nneigh_max"] == sg["nneigh_max"] \
            and s["ncalctotal"] == sg["ncalctotal"] and s["nrelink"] == sg["nrelink"]
        good = (not bad) and same_nn and sc_ok and chunk_ok
        t = torch.tensor([0.0 if good else 1.0], device="cuda"); dist.all_reduce(t)
        ok = ok and t.item() == 0
        print(f"[rank {rank}] {name}: own {nown} halo {nsrc - nown} ghosts {nt - nsrc} its {s['itsdensity']} relink {s['nrelink']} "
              f"max err {max(errs.values()):.2e} numneigh_equal {same_nn} scalars_ok {sc_ok} chunked_host_call_equal {chunk_ok} -> {'OK' if good else 'FAIL ' + str(bad)}", flush=True)
        if not sc_ok:
            print({k: (s[k], sg[k]) for k in ("dtcourant", "dtforce", "dtav", "vsigmax", "stressmax", "fhmax", "hhmax", "itsdensity", "nneigh_min", "nneigh_max", "ncalctotal", "nrelink")})
    dist.barrier(); dist.destroy_process_group()
    if rank == 0:
        print(f"SLAB CHECK ({transport}, light={os.environ.get('NDSPMHD_B200_SLAB_LIGHT', '0')})", "PASSED" if ok else "FAILED")
    sys.exit(0 if ok else 1)

if __name__ == "__main__":
    main()
