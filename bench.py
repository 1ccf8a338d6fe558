#!/usr/bin/env python
"""bench.py -- particle-updates/s of the NDSPMHD hot path (link + density/h iteration + cons2prim + SPMHD rates).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nx NX] [--scaling strong|weak]

Workload (BASELINE.json metric: "particle-updates/sec (density+rates) at 16M 3D particles"): the reference's own 3D MHD
problem, the Orszag-Tang vortex in a thin periodic slab (src/setup_orszagtang2D_mhd.f90:70-73, default SETUP3D), at
512 x 512 x 64 = 16.8 M particles, as a glass (lattice + 0.2 dp random displacement) with evolved psi/alpha/u fields
so that every term of the rates is live and the h-iteration needs several rounds.  Option tuple: imhd=11,
idivbzero=2, iener=2, iav=2, iavlim=(2,1,0), ikernav=3, cubic spline, periodic ghosts made on the device.

One "step" = one full `derivs` (src/derivs.f90:74-156) from the same unconverged smoothing lengths.
  value : device-resident state, CUDA events on the library's stream, K steps, max over ranks.
  e2e   : the same through the C-ABI with HOST arrays: upload (pinned H2D) + derivs + download (D2H) per step.
  roofline : dominant kernel (rates_pair_kernel) algorithmic bytes / its CUDA-event duration vs measured HBM copy peak.
  cpu_baseline : the CPU oracle (a line-by-line C++ restatement of the reference; the reference itself is Fortran and
                 cannot be built in this image), ONE thread as the reference is serial, on a bounded sample.
--impl reference : the oracle farmed over all host cores as independent serial runs (the reference's only form of
                 parallelism: scripts/doparallel.pl), same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-updates/sec (density+rates), 3D MHD Orszag-Tang, 16.8M particles"
UNIT = "particle-updates/s"
# compulsory read-once/write-once bytes per particle-update, FP64, 3D MHD tuple
BYTES_DENSITY = 68 + 44
BYTES_C2P = 40 + 56
BYTES_RATES = 164 + 120
BYTES_TOTAL = BYTES_DENSITY + BYTES_C2P + BYTES_RATES   # 492


def workload(nx: int, perturb: float = 0.2):
    from ndspmhd_b200 import setups

    o, p = setups.orszag_tang(ndim=3, nx=nx, zfrac=0.125, perturb_amp=perturb, evolved=True, imhd=11, idivbzero=2, iener=2)
    o.device_ghosts = 1
    o.want_aux = 0
    return o, p


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks line of /opt/skills/guides/B200_PROFILING.md, sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_sample(nx: int):
    """One serial oracle `derivs` on the bounded sample; returns (updates/s, npart, seconds, phase ms)."""
    from oracle import oracle

    o, p = workload(nx)
    t = time.perf_counter()
    s, ms = oracle.derivs(o, p)
    dt = time.perf_counter() - t
    return p.npart / dt, p.npart, dt, ms, s["itsdensity"]


# ----------------------------------------------------------------------------------------------------------------------
# --impl reference: the CPU path on all host cores (independent serial runs, one per core)
# ----------------------------------------------------------------------------------------------------------------------
_W = {}


def _ref_init(nx):
    from oracle import oracle  # noqa: F401  (loads the .so once per worker)

    _W["o"], _W["p"] = workload(nx)


def _ref_step(_):
    from oracle import oracle

    p = _W["p"].copy()
    t = time.perf_counter()
    oracle.derivs(_W["o"], p)
    return time.perf_counter() - t, p.npart


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    nx = args.ref_nx or (128 if cores <= 64 else 96)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_ref_init, initargs=(nx,)) as pool:
        times = []
        for step in range(args.warmup + args.steps):
            t = time.perf_counter()
            res = pool.map(_ref_step, range(cores), chunksize=1)
            dt = time.perf_counter() - t
            if step >= args.warmup:
                times.append((dt, sum(r[1] for r in res)))
    total_t = sum(t for t, _ in times)
    total_n = sum(n for _, n in times)
    value = total_n / total_t
    npart = times[0][1] // cores
    sample = (f"{cores} independent serial oracle runs (one per core, the reference's job-farming model, scripts/doparallel.pl), "
              f"each one full derivs on the same workload at {nx}x{nx}x{nx // 8} = {npart} particles")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_t / len(times), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(args, nx_used=nx),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def config_dict(args, nx_used=None, extra=None):
    nx = args.nx
    d = {"workload": f"3D Orszag-Tang MHD vortex, thin periodic slab {nx}x{nx}x{nx // 8} = {nx * nx * (nx // 8)} particles "
                     "(src/setup_orszagtang2D_mhd.f90 in 3D), glass (lattice + 0.2 dp), imhd=11 idivbzero=2 iener=2 iav=2 cubic spline hfact=1.2",
         "step": "one derivs: ghosts + link + density/h iteration to tolh=1e-3 + cons2prim + rates (pair + final)",
         "l2": "inputs (>2 GB of particle state) larger than the 126 MB L2; no flush needed",
         "parallelism": "1 GPU" if args.gpus == 1 else f"x-slabs over {args.gpus} GPUs, NCCL halo exchange"}
    if nx_used and nx_used != nx:
        d["reference_sample_nx"] = nx_used
    if extra:
        d.update(extra)
    return d


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def run_single(args):
    import torch

    from ndspmhd_b200 import abi, lib

    if lib.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; ndspmhd_b200 has no CPU fallback")
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    o, p0 = workload(args.nx)
    n = p0.npart
    # host arrays in page-locked memory (the Fortran module arrays, registered once)
    p = lib.pinned_particles(3, n, p0.idim)
    for k, v in p0.arrays.items():
        p.arrays[k][...] = v
    p.ntotal = n
    del p0
    # derivs overwrites hh with the converged smoothing lengths; every step must start from the same guess
    L = lib.load()
    import ctypes as C
    import numpy as np
    nb = p.arrays["hh"].nbytes
    ptr = L.ndspmhd_b200_host_alloc(nb)
    guess = np.frombuffer((C.c_char * nb).from_address(ptr), dtype=np.float64, count=p.idim)
    guess[...] = p.arrays["hh"]
    p.arrays["hh_guess"] = guess
    p.__dict__["_pinned"].append(ptr)
    hot = lib.Hotpath(o, 3, dev)
    stream = torch.cuda.ExternalStream(hot.stream(), device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    # ---- e2e: host arrays in, host arrays out ----
    mask = abi.DL_DENSITY | abi.DL_PRIM | abi.DL_RATES
    if args.e2e_real_rows:
        mask |= abi.DL_REAL_ROWS            # informational variant: the ghost rows of the output arrays stay on the device
    def e2e_step():
        p.ntotal = n
        return hot.derivs_host(p, mask)     # ndspmhd_b200_derivs_host: upload + derivs + download, copies overlapped with kernels
    for _ in range(max(1, min(args.warmup, 2))):
        s = e2e_step()
    nt = s["ntotal"]
    e2e_steps = max(1, min(args.steps, 3))
    a, b = ev(), ev()
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    b.record(stream)
    torch.cuda.synchronize()
    e2e_ms = a.elapsed_time(b) / e2e_steps
    up_names = ["x", "vel", "pmass", "hh", "itype", "ireal", "en", "Bevol", "alpha", "psi", "rho"]
    dn_names = ["hh", "rho", "gradh", "numneigh", "dens", "uu", "pr", "spsound", "Bfield", "drhodt", "dhdt", "force", "dudt", "dendt",
                "dBevoldt", "daldt", "dpsidt", "gradpsi", "divB", "curlB"]
    rowbytes = lambda nm: p.arrays[nm].nbytes // p.idim
    h2d = sum(rowbytes(nm) for nm in up_names) * n
    d2h = sum(rowbytes(nm) for nm in dn_names) * (n if args.e2e_real_rows else nt)

    # ---- device-resident ----
    p.ntotal = n
    hot.upload(p)
    def step():
        hot.rewind()
        return hot.derivs()
    for _ in range(args.warmup):
        s = step()
    phases = {k: 0.0 for k in ("link", "density", "c2p_gather", "rates_pair", "rates_final", "rates_pair_kernel")}
    clocks = ClockSampler(dev)
    clocks.start()
    l0 = hot.launch_count()
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(args.steps):
        s = step()
        for k, v in hot.timings().items():
            phases[k] += v
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    launches = hot.launch_count() - l0
    ck = clocks.stop()
    for k in phases:
        phases[k] /= args.steps
    value = n / (ms * 1e-3)
    peak, peak_src = measured_peaks()
    pair_ms = phases.pop("rates_pair_kernel")       # CUDA events around the pair kernel's launch alone, on the library's stream
    achieved = BYTES_RATES * n / (pair_ms * 1e-3) / 1e9 if pair_ms > 0 else 0.0
    traffic, ncu_extra = None, {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if int(tj.get("nx", 0)) == args.nx:
                traffic = tj.get("rates_pair_kernel_dram_bytes_per_launch")
                ncu_extra = {k: tj[k] for k in ("fp64_pipe_active_pct", "issue_active_pct", "l1tex_throughput_pct", "ncu_report") if k in tj}
        except Exception:
            pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, extra={"npart": n, "ntotal_with_ghosts": nt, "itsdensity": s["itsdensity"], "nneigh_min": s["nneigh_min"],
                                           "nneigh_max": s["nneigh_max"]}),
        "e2e": {"value": n / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                "steps": e2e_steps},
        "gpu_launches": launches,
        "clocks": ck,
        "roofline": {"bound": "hbm", "kernel": "rates_pair_kernel<3,MHD>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_RATES * n, "kernel_ms": pair_ms, "kernel_share_of_step": pair_ms / ms,
                     "ncu": ncu_extra,
                     "whole_step_GBps": BYTES_TOTAL * n / (ms * 1e-3) / 1e9, "whole_step_frac": BYTES_TOTAL * n / (ms * 1e-3) / 1e9 / peak,
                     "note": "FP64 pairwise gather: the FP64 pipe and L1/shared gather bandwidth bind long before HBM (DESIGN.md); no tensor cores"},
        "phases_ms": phases,
    }
    # ---- whole leapfrog steps on the resident state: ndspmhd_b200_step = predictor + derivs + corrector;
    #      informational: these start from the predicted h of a running simulation, the timed `derivs` above from an unconverged guess
    try:
        dt_sim = min(0.25 * s["dtforce"], 0.3 * s["dtcourant"], 0.9 * s["dtdrag"], 0.25 * s["dtvisc"])
        dt_sim, _ = hot.step(dt_sim)
        nst = max(1, min(args.steps, 3))
        torch.cuda.synchronize()
        a.record(stream)
        its_seen = []
        for _ in range(nst):
            dt_sim, ss = hot.step(dt_sim)
            its_seen.append(ss["itsdensity"])
        b.record(stream)
        torch.cuda.synchronize()
        st_ms = a.elapsed_time(b) / nst
        line["step_resident"] = {"api": "ndspmhd_b200_step", "ms_per_step": st_ms, "value": n / (st_ms * 1e-3), "unit": UNIT, "steps": nst,
                                 "itsdensity": its_seen, "pcie_bytes_per_step": 0}
    except Exception as ex:  # never lose the headline line over the extra
        line["step_resident"] = {"error": str(ex)[:200]}
    if not args.no_cpu:
        v, npc, dt, pms, its = cpu_oracle_sample(args.cpu_nx)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"one serial oracle derivs of the same workload at {args.cpu_nx}x{args.cpu_nx}x{args.cpu_nx // 8} = {npc} particles "
                                          f"({dt:.1f} s, {its} density rounds); the reference is serial (docs/about.rst:12)",
                                "phases_ms": dict(zip(["ghosts", "link", "density", "c2p", "rates"], pms))}
    hot.close()
    lib.free_pinned(p)
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=512, help="particles along x (thin slab nx * nx * nx/8)")
    ap.add_argument("--cpu-nx", type=int, default=224, help="size of the bounded CPU-baseline sample")
    ap.add_argument("--ref-nx", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-real-rows", action="store_true", help="e2e downloads rows [0,npart) only (ND_DL_REAL_ROWS); default: the full contract")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    import __graft_entry__ as g

    g.build()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        from ndspmhd_b200 import slab_bench

        return slab_bench.run(args, METRIC, UNIT, config_dict, ClockSampler, measured_peaks)
    return run_single(args)


if __name__ == "__main__":
    sys.exit(main())
