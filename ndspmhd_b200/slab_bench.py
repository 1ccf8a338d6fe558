"""bench.py's N > 1 leg: the same workload, x-slab partitioned over the ranks of one node (strong scaling by default).

Launched as `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N`; one rank per GPU, NCCL.
Timing: barrier + torch.cuda.synchronize() either side of exactly K steps, CUDA events on the library's stream, MAX over
ranks; rank 0 prints the JSON line.  value = global particles / max time.
"""
from __future__ import annotations

import json
import os

import numpy as np


def parity_vs_single(cfg, workload, hot, p, n, s, info, rank, local, torch, dist):
    """The slab-decomposed result against the single-GPU result of the same global particle set, computed in the same run (outside the
    timed region): order-independent integer checksums over ALL ranks -- sum numneigh, ncalctotal, the rates pair count, itsdensity,
    relinks -- must be bit-equal, and rho, h, force, dB/dt, du/dt, div B on rank 0's own rows must agree with the single-GPU run to the
    summation-order tolerance (max-norm error relative to the field's max-norm).  The single-GPU run uses a second context on rank 0's GPU."""
    from . import abi, lib, setups

    hot.download(p, abi.DL_DENSITY | abi.DL_RATES)                     # this rank's own rows of the last timed derivs
    sums = torch.tensor([float(p.arrays["numneigh"][:n].astype(np.int64).sum())], dtype=torch.float64, device="cuda")
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    out = None
    if rank == 0:
        o1, p1 = workload(cfg)
        n1 = p1.npart
        hot1 = lib.Hotpath(o1, 3, local)
        try:
            hot1.upload(p1)
            s1 = hot1.derivs()
            hot1.download(p1, abi.DL_DENSITY | abi.DL_RATES)
        finally:
            hot1.close()
        rows = info["rows"]
        ints = {"sum_numneigh": (int(sums[0].item()), int(p1.arrays["numneigh"][:n1].astype(np.int64).sum())),
                "ncalctotal": (int(s["ncalctotal"]), int(s1["ncalctotal"])), "npairs_rates": (int(s["npairs_rates"]), int(s1["npairs_rates"])),
                "itsdensity": (int(s["itsdensity"]), int(s1["itsdensity"])), "nrelink": (int(s["nrelink"]), int(s1["nrelink"])),
                "nneigh_min": (int(s["nneigh_min"]), int(s1["nneigh_min"])), "nneigh_max": (int(s["nneigh_max"]), int(s1["nneigh_max"]))}
        errs = {}
        for f in ("rho", "hh", "force", "dBevoldt", "dudt", "divB"):
            a, b = np.asarray(p.arrays[f][:n]), np.asarray(p1.arrays[f][rows])
            errs[f] = float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))
        out = {"integers_equal": all(a == b for a, b in ints.values()), "integers_slab_vs_single": {k: list(v) for k, v in ints.items()},
               "numneigh_rows_equal_rank0": bool(np.array_equal(p.arrays["numneigh"][:n], p1.arrays["numneigh"][rows])),
               "max_norm_rel_err_rank0_rows": errs, "tolerance": 1e-12, "ok": None,
               "scalars_rel_err": {k: abs(s[k] - s1[k]) / max(abs(s1[k]), 1e-300) for k in ("dtcourant", "dtforce", "vsigmax")}}
        out["ok"] = bool(out["integers_equal"] and out["numneigh_rows_equal_rank0"] and max(errs.values()) <= 1e-12)
        del p1
    dist.barrier()
    return out


def run(args, cfg, workload, UNIT, config_dict, ClockSampler, measured_peaks):
    import torch
    import torch.distributed as dist

    from . import abi, lib, setups, slab

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"bench.py --gpus {args.gpus} must be launched with torchrun --nproc-per-node {args.gpus} (WORLD_SIZE={world})")
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    weak = args.scaling == "weak"   # the box grows by one x-period of the fields per rank (setups.orszag_tang(weak=True)): fixed work per GPU
    if cfg["kind"] != "ot" or cfg["ndim"] != 3:
        raise SystemExit("bench.py --gpus N: the slab decomposition is benchmarked on the 3-D MHD configs (slab512, cube256, cube128)")
    o, p0, info = workload(cfg, slab=(rank, world), weak=weak)
    n, nglobal = p0.npart, int(info["nglobal"])
    lo, hi = float(info["edges"][rank]), float(info["edges"][rank + 1])
    p = lib.pinned_particles(3, n, p0.idim)
    for k, v in p0.arrays.items():
        p.arrays[k][...] = v
    p.ntotal = n
    del p0
    guess = np.array(p.arrays["hh"], copy=True)
    hot = lib.Hotpath(o, 3, local)
    native = args.transport == "nccl"
    if native:      # the library drives NCCL itself (nd_nccl.cuh); torch.distributed only carries the communicator id and the timing barriers
        slab.attach_nccl(hot, rank, world, lo, hi, nglobal)
    else:           # host-callback transport over torch.distributed (the portable route, what an MPI host would supply)
        comm = slab.SlabComm(device="cuda")
        slab.attach(hot, comm, lo, hi, nglobal)
    stream = torch.cuda.ExternalStream(hot.stream(), device=local)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    mask = abi.DL_DENSITY | abi.DL_PRIM | abi.DL_RATES

    def timed(fn, steps):
        a, b = ev(), ev()
        dist.barrier()
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(steps):
            out = fn()
        b.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    up_names = ["x", "vel", "pmass", "hh", "itype", "ireal", "en", "Bevol", "alpha", "psi", "rho"]
    dn_names = ["hh", "rho", "gradh", "numneigh", "dens", "uu", "pr", "spsound", "Bfield", "drhodt", "dhdt", "force", "dudt", "dendt",
                "dBevoldt", "daldt", "dpsidt", "gradpsi", "divB", "curlB"]
    lean_skip = ["gradh", "numneigh", "dens", "spsound", "gradpsi", "dudt"]   # bench.py LEAN_SKIP: the Fortran shim's contract on ordinary steps
    rowbytes = lambda nm: p.arrays[nm].nbytes // p.idim
    e2e_steps = max(1, min(args.steps, 3))

    def e2e(skip):
        def e2e_step():
            p.arrays["hh"][:n] = guess[:n]       # host-side restore of the guess (the integrator's predictor would do this)
            p.ntotal = n
            return hot.derivs_host(p, mask, skip=skip)
        for _ in range(2):
            e2e_step()
        ms_, s_ = timed(e2e_step, e2e_steps)
        return ms_, s_, sum(rowbytes(nm) for nm in dn_names if nm not in skip) * n

    e2e_full_ms, s, dn_full = e2e([])
    e2e_ms, s, dn_lean = e2e(lean_skip)
    nown, nsrc, nt = slab.row_counts(hot)
    bytes_t = torch.tensor([sum(rowbytes(nm) for nm in up_names) * n, dn_lean, nsrc - nown, nt - nsrc, dn_full],
                           dtype=torch.float64, device="cuda")
    dist.all_reduce(bytes_t, op=dist.ReduceOp.SUM)

    p.arrays["hh"][:n] = guess[:n]
    p.ntotal = n
    hot.upload(p)

    def step():
        hot.rewind()
        return hot.derivs()

    for _ in range(args.warmup):
        step()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = hot.launch_count()
    ar0, sb0 = slab.comm_stats(hot)
    ms, s = timed(step, args.steps)
    launches = hot.launch_count() - l0
    ar1, sb1 = slab.comm_stats(hot)
    halo_bytes = torch.tensor([float(sb1 - sb0) / args.steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(halo_bytes, op=dist.ReduceOp.SUM)
    ck = clocks.stop()
    phases = hot.timings()
    ph = torch.tensor([phases[k] for k in ("link", "density", "c2p_gather", "rates_pair", "rates_final", "rates_pair_kernel")], dtype=torch.float64,
                      device="cuda")
    dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    launches_t = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
    dist.all_reduce(launches_t, op=dist.ReduceOp.SUM)
    # weak scaling: the global set (world x the per-rank set) is not re-run on one GPU; the strong-scaling runs carry the check
    check = parity_vs_single(cfg, workload, hot, p, n, s, info, rank, local, torch, dist) if not (args.no_parity_check or weak) else None
    # ---- whole leapfrog steps on the slab-decomposed resident state: predictor, periodic wrap, row migration between ranks, derivs, corrector
    step_res = None
    try:
        hot.set_row_ids(np.asarray(info["rows"], dtype=np.int64))
        dt_sim = min(0.25 * s["dtforce"], 0.3 * s["dtcourant"])
        for _ in range(3):      # the first steps pay one-off costs (lazy kernel loading, list buffers growing to the partial rounds' sizes)
            dt_sim, _ = hot.step(dt_sim)
        nst = max(1, min(args.steps, 3))
        mo0, mi0, mb0 = hot.migration_stats()
        its_seen = []

        def one_step():
            nonlocal dt_sim
            dt_sim, ss = hot.step(dt_sim)
            its_seen.append(ss["itsdensity"])
            return ss

        st_ms, _ = timed(one_step, nst)
        mo1, mi1, mb1 = hot.migration_stats()
        mig = torch.tensor([float(mo1 - mo0), float(mb1 - mb0), float(slab.row_counts(hot)[0])], dtype=torch.float64, device="cuda")
        dist.all_reduce(mig, op=dist.ReduceOp.SUM)
        step_res = {"api": "ndspmhd_b200_step", "ms_per_step": st_ms, "value": nglobal / (st_ms * 1e-3), "unit": UNIT, "steps": nst, "itsdensity": its_seen[:nst],
                    "rows_migrated_per_step_all_ranks": mig[0].item() / nst, "migration_bytes_per_step_all_ranks": mig[1].item() / nst,
                    "rows_owned_all_ranks": int(mig[2].item()), "pcie_bytes_per_step": 0}
    except Exception as ex:  # never lose the headline line over the extra
        step_res = {"error": str(ex)[:200]}
    if rank == 0:
        peak, peak_src = measured_peaks()
        pair_ms = float(ph[5].item())     # the pair kernel alone (CUDA events on the library's stream), max over ranks
        bytes_rates = 284
        ach = bytes_rates * (nglobal / world) / (pair_ms * 1e-3) / 1e9 if pair_ms > 0 else 0.0
        line = {
            "metric": cfg.get("metric") or ("particle-updates/sec (density+rates), " + cfg["label"].split(",")[0]), "value": nglobal / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args),
            "run": {"npart": nglobal, "npart_per_rank": nglobal // world, "halo_rows_total": int(bytes_t[2].item()),
                    "ghost_rows_total": int(bytes_t[3].item()), "itsdensity": s["itsdensity"], "nrelink": s["nrelink"],
                    "nneigh_min": s["nneigh_min"], "nneigh_max": s["nneigh_max"], "npairs_rates": s["npairs_rates"],
                    "halo_bytes_per_step_all_ranks": float(halo_bytes.item())},
            "e2e": {"value": nglobal / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(bytes_t[0].item()),
                    "d2h_bytes_per_step": int(bytes_t[1].item()), "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "contract": "lean: skips " + ",".join(lean_skip) + "; each rank moves its own rows",
                    "full_contract": {"value": nglobal / (e2e_full_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_full_ms, "d2h_bytes_per_step": int(bytes_t[4].item())}},
            "gpu_launches": int(launches_t.item()),
            "clocks": ck,
            "roofline": {"bound": "hbm", "kernel": "rates_pair_kernel<3,MHD,FAST> (per rank, max over ranks)", "achieved": ach, "peak": peak,
                         "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src, "kernel_ms": pair_ms},
            "phases_ms": dict(zip(["link", "density", "c2p_gather", "rates_pair", "rates_final"], [float(v) for v in ph.tolist()[:5]])),
            "comm": {"allreduces_per_step": (ar1 - ar0) // max(1, args.steps),
                     "transport": "native: ncclAllReduce / grouped ncclSend+ncclRecv / ncclAllGather issued by the library on its stream" if native
                                  else "host callbacks (nd_comm) over torch.distributed NCCL"},
            "parity_vs_single": check,
            "step_resident": step_res,
        }
        print(json.dumps(line), flush=True)
    hot.close()
    lib.free_pinned(p)
    dist.barrier()
    dist.destroy_process_group()
    return 0
