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

# ----------------------------------------------------------------------------------------------------------------------
# workloads: BASELINE.json configs C2-C5 (C1, the 1000-particle shock tube, is a parity case, not a bench line).  `bytes` = SURVEY.md 8(d)
# compulsory read-once/write-once bytes per particle-update (density, cons2prim, rates), FP64.
# ----------------------------------------------------------------------------------------------------------------------
CONFIGS = {
    # C5 family, thin slab: the reference's own 3-D MHD problem (src/setup_orszagtang2D_mhd.f90:70-73, default SETUP3D) -- the headline
    "slab512": dict(kind="ot", ndim=3, nx=512, cube=False, bytes=(68 + 44, 40 + 56, 164 + 120), metric=METRIC,
                    label="3D Orszag-Tang MHD vortex, thin periodic slab {nx}x{nx}x{nz} = {n} particles (src/setup_orszagtang2D_mhd.f90 in 3D), "
                          "glass (lattice + 0.2 dp), imhd=11 idivbzero=2 iener=2 iav=2 cubic spline hfact=1.2"),
    # C5 / C3: the same fields in a periodic CUBE (twice the halo area per face of the thin slab)
    "cube256": dict(kind="ot", ndim=3, nx=256, cube=True, bytes=(112, 96, 284),
                    label="3D Orszag-Tang MHD vortex, periodic cube {nx}^3 = {n} particles, glass (lattice + 0.2 dp), imhd=11 idivbzero=2 iener=2 iav=2"),
    "cube128": dict(kind="ot", ndim=3, nx=128, cube=True, bytes=(112, 96, 284),
                    label="3D Orszag-Tang MHD vortex, periodic cube {nx}^3 = {n} particles, glass (lattice + 0.2 dp), imhd=11 idivbzero=2 iener=2 iav=2"),
    # C2: 2-D Orszag-Tang on the close-packed lattice, dp = 1/512 -> 512 x 592 (src/setup_orszagtang2D_mhd.f90:63-96)
    "ot2d": dict(kind="ot", ndim=2, nx=512, cube=False, bytes=(60 + 44, 40 + 56, 156 + 120),
                 label="2D Orszag-Tang MHD vortex, close-packed lattice dp=1/{nx} = {n} particles (+0.2 dp), imhd=11 idivbzero=2 iener=2 iav=2"),
    # C4: two-fluid dust + gas, fat periodic box (the thin 1 x 11 x 11 box is ghost-dominated; it is a parity case)
    "dust2m": dict(kind="dust", ndim=3, nx=100, cube=True, bytes=(112, 48, 172),
                   label="3D two-fluid dust+gas box, {nx}^3 gas + {nx}^3 dust = {n} particles, idust=2 idrag_nature=1 K=1, iener=2 iav=2 (src/setup_dustybox.f90 fields, fat box)"),
}


def cfg_of(args):
    c = dict(CONFIGS[args.config])
    if args.nx:
        c["nx"] = args.nx
    return c


def workload(args_or_cfg, nx=None, slab=None, weak=False):
    """(options, particles[, info]) of a config at `nx` particles per unit length (default: the config's own size)."""
    from ndspmhd_b200 import setups

    c = args_or_cfg if isinstance(args_or_cfg, dict) else cfg_of(args_or_cfg)
    nx = nx or c["nx"]
    if c["kind"] == "dust":
        if slab is not None:
            raise SystemExit("bench.py: the dust config is single-GPU")
        o, p = setups.dustybox(ndim=3, nx=nx, perturb_amp=0.05)
        o.device_ghosts, o.want_aux = 1, 0
        return o, p
    kw = dict(ndim=c["ndim"], nx=nx, perturb_amp=0.2, evolved=True, imhd=11, idivbzero=2, iener=2)
    if c["ndim"] == 3:
        kw.update(cube=c["cube"], zfrac=0.125)
    else:
        kw.update(lattice="cp")
    out = setups.orszag_tang(slab=slab, weak=weak, **kw)
    out[0].device_ghosts, out[0].want_aux = 1, 0
    return out


def npart_of(c, nx=None):
    nx = nx or c["nx"]
    if c["kind"] == "dust":
        return 2 * nx ** 3
    if c["ndim"] == 2:
        return None   # the close-packed generator decides (512 x 592 at dp = 1/512)
    return nx ** 3 if c["cube"] else nx * nx * (nx // 8)


def source_sha() -> str:
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_count

    return sass_count.source_sha()


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


def cpu_oracle_runs(c, sizes):
    """Serial oracle `derivs` of the config at several sizes: [(npart, seconds, updates/s, phase ms, its)] -- the reference is serial
    (docs/about.rst:12), so this IS its execution model; the sizes show that updates/s does not depend on N (the fit SURVEY 8d asks for)."""
    from oracle import oracle

    runs = []
    for nx in sizes:
        o, p = workload(c, nx)
        t = time.perf_counter()
        s, ms = oracle.derivs(o, p)
        dt = time.perf_counter() - t
        runs.append({"nx": nx, "npart": int(p.npart), "seconds": dt, "updates_per_s": p.npart / dt, "itsdensity": int(s["itsdensity"]),
                     "phases_ms": dict(zip(["ghosts", "link", "density", "c2p", "rates"], [float(x) for x in ms]))})
        del p
    return runs


def loglog_slope(runs):
    """d log(time) / d log(N): 1.0 = throughput independent of N."""
    import math

    if len(runs) < 2:
        return None
    xs = [math.log(r["npart"]) for r in runs]
    ys = [math.log(r["seconds"]) for r in runs]
    mx, my = sum(xs) / len(xs), sum(ys) / len(ys)
    return sum((x - mx) * (y - my) for x, y in zip(xs, ys)) / max(sum((x - mx) ** 2 for x in xs), 1e-300)


def sample_sizes(c, budget_s: float, rate: float = 1.0e5):
    """Sizes of the config family whose serial oracle runs fit `budget_s` seconds in total (largest first dominates)."""
    cand = {"slab": [256, 224, 192, 160, 128, 96, 64], "cube": [128, 112, 96, 80, 64, 48, 32], "2d": [512, 384, 256, 192, 128],
            "dust": [100, 80, 64, 50, 40, 32]}["dust" if c["kind"] == "dust" else "2d" if c["ndim"] == 2 else "cube" if c["cube"] else "slab"]
    est = lambda nx: (npart_of(c, nx) or int(1.155 * nx * nx)) / rate
    big = next((nx for nx in cand if est(nx) <= 0.7 * budget_s), cand[-1])
    smaller = [nx for nx in cand if nx < big]
    out = [big]
    for nx in (smaller[1:2] + smaller[3:4]):
        if sum(est(k) for k in out) + est(nx) <= budget_s:
            out.append(nx)
    return sorted(out)


# ----------------------------------------------------------------------------------------------------------------------
# --impl reference: the CPU path on all host cores (independent serial runs, one per core)
# ----------------------------------------------------------------------------------------------------------------------
_W = {}


def _ref_init(cfg, nx):
    from oracle import oracle  # noqa: F401  (loads the oracle .so once per worker; the product library is never loaded in this arm)

    _W["o"], _W["p"] = workload(cfg, nx)


def _ref_step(_):
    from oracle import oracle

    p = _W["p"].copy()
    t = time.perf_counter()
    oracle.derivs(_W["o"], p)
    return time.perf_counter() - t, p.npart


def run_reference(args):
    """The reference's CPU implementation of the path on all host cores.  The reference is serial Fortran that cannot be built in this image
    (DESIGN.md), so this times its C++ restatement (oracle/), farmed over the cores as independent serial runs -- the reference's only form of
    parallelism (scripts/doparallel.pl).  Each step is a bounded sample of the workload (the largest size of the same family that keeps the
    whole --steps/--warmup run within a few minutes); single-core runs at smaller sizes show updates/s is flat in N."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp

    c = cfg_of(args)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # time budget: ~200 s for the farmed steps (contention on memory bandwidth costs ~1.4x against a lone core)
    per_step = 200.0 / max(1, args.steps + args.warmup)
    nx = args.ref_nx or sample_sizes(c, per_step / 1.4)[-1]
    fit = cpu_oracle_runs(c, [k for k in sample_sizes(c, 12.0) if k != nx][:2]) if not args.no_fit else []
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_ref_init, initargs=(c, nx)) as pool:
        times = []
        for step in range(args.warmup + args.steps):
            t = time.perf_counter()
            res = pool.map(_ref_step, range(cores), chunksize=1)
            dt = time.perf_counter() - t
            if step >= args.warmup:
                times.append((dt, sum(r[1] for r in res), sum(r[0] for r in res) / len(res)))
    total_t = sum(t[0] for t in times)
    total_n = sum(t[1] for t in times)
    value = total_n / total_t
    npart = times[0][1] // cores
    per_core = npart / (sum(t[2] for t in times) / len(times))
    sample = (f"{cores} independent serial oracle runs (one per core: the reference's job-farming model, scripts/doparallel.pl), each one full derivs of the "
              f"same workload at nx={nx} = {npart} particles per run; {per_core:.3g} updates/s per core inside the farm")
    line = {
        "impl": "reference", "metric": c.get("metric", METRIC_FOR(c)), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_t / len(times), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "single_core_fit": {"runs": [{k: r[k] for k in ("nx", "npart", "seconds", "updates_per_s")} for r in fit],
                                             "farm_per_core_updates_per_s": per_core,
                                             "note": "updates/s of the serial oracle vs N: flat (O(N) cell lists), so the bounded sample stands for the full-size run"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def METRIC_FOR(c):
    return "particle-updates/sec (density+rates), " + c["label"].split(",")[0]


def config_dict(args):
    """The same dict in both arms (the driver compares them): the workload by name and size, nothing run-specific."""
    c = cfg_of(args)
    nx = c["nx"]
    n = npart_of(c) or 512 * 592 * (nx * nx) // (512 * 512)
    return {"workload": args.config + ": " + c["label"].format(nx=nx, nz=nx // 8, n=n),
            "step": "one derivs: ghosts + link + density/h iteration to tolh=1e-3 + cons2prim + rates (pair + final)",
            "l2": "inputs (>2 GB of particle state at the headline size) larger than the 126 MB L2; no flush needed",
            "parallelism": "1 GPU" if args.gpus == 1 else f"x-slabs over {args.gpus} GPUs, NCCL halo exchange"}


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
UP_NAMES = ["x", "vel", "pmass", "hh", "itype", "ireal", "en", "Bevol", "alpha", "psi", "rho"]
DN_FULL = ["hh", "rho", "gradh", "numneigh", "dens", "uu", "pr", "spsound", "Bfield", "drhodt", "dhdt", "force", "dudt", "dendt",
           "dBevoldt", "daldt", "dpsidt", "gradpsi", "divB", "curlB"]
# what neither the leapfrog integrator (src/stepND_leapfrog_mhd.f90:70-216) nor evwrite (src/evwrite_mhd.f90:124-284) reads between two derivs:
# a dump step fetches them with ndspmhd_b200_download.  The Fortran shim's default contract (b200_lean_download).
LEAN_SKIP = ["gradh", "numneigh", "dens", "spsound", "gradpsi", "dudt"]
HYDRO_SKIP = ["Bfield", "dBevoldt", "dpsidt", "gradpsi", "divB", "curlB"]


def fp64_roofline(kernel_label, ntrips, kernel_ms, sm_mhz, num_smsp=592):
    """FP64-pipe floor of the pair kernel: FP64-pipe instructions per pair-loop trip (from the SASS of THIS build, tools/sass_count.py) x the
    warp trips the kernel made (nd_scalars.ntrips_rates) x 2 issue cycles (measured 0.49 warp-instr/clk/SMSP, profiles/r01/fp64_pipe_microbench.txt)
    over 592 SMSPs.  frac = floor / measured kernel time = how busy the binding unit can be at best."""
    try:
        sj = json.load(open(os.path.join(ROOT, "profiles", "sass_fp64.json")))
        if sj.get("source_sha") != source_sha():
            return {"error": "profiles/sass_fp64.json is from another build (run tools/sass_count.py --write)"}
        k = sj["kernels"][kernel_label]
    except Exception as ex:
        return {"error": str(ex)[:120]}
    clk = (sm_mhz or 1965.0) * 1e6
    floor_ms = 1e3 * ntrips * k["fp64_pipe_instructions"] * (1.0 / 0.49) / (num_smsp * clk)
    return {"bound": "fp64_pipe", "kernel": kernel_label, "fp64_pipe_instructions_per_trip": k["fp64_pipe_instructions"], "loop_instructions_per_trip": k["loop_instructions"],
            "warp_trips": int(ntrips), "issue_rate_warp_instr_per_clk_per_smsp": 0.49, "sm_mhz": clk / 1e6,
            "fp64_issue_floor_ms": floor_ms, "kernel_ms": kernel_ms, "frac": floor_ms / kernel_ms if kernel_ms > 0 else None}


def run_single(args):
    import ctypes as C

    import numpy as np
    import torch

    from ndspmhd_b200 import abi, lib

    if lib.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; ndspmhd_b200 has no CPU fallback")
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    c = cfg_of(args)
    o, p0 = workload(c)
    n, ndim = p0.npart, c["ndim"]
    mhd = o.imhd != 0
    # host arrays in page-locked memory (the Fortran module arrays, registered once)
    p = lib.pinned_particles(ndim, n, p0.idim)
    for k, v in p0.arrays.items():
        p.arrays[k][...] = v
    p.ntotal = n
    del p0
    # derivs overwrites hh with the converged smoothing lengths; every step must start from the same guess
    L = lib.load()
    nb = p.arrays["hh"].nbytes
    ptr = L.ndspmhd_b200_host_alloc(nb)
    guess = np.frombuffer((C.c_char * nb).from_address(ptr), dtype=np.float64, count=p.idim)
    guess[...] = p.arrays["hh"]
    p.arrays["hh_guess"] = guess
    p.__dict__["_pinned"].append(ptr)
    hot = lib.Hotpath(o, ndim, dev)
    stream = torch.cuda.ExternalStream(hot.stream(), device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    a, b = ev(), ev()
    rowbytes = lambda nm: p.arrays[nm].nbytes // p.idim

    # ---- e2e: host arrays in, host arrays out, through ndspmhd_b200_derivs_host (upload + derivs + download, copies overlapped with kernels) ----
    def e2e(skip, real_rows):
        mask = abi.DL_DENSITY | abi.DL_PRIM | abi.DL_RATES | (abi.DL_REAL_ROWS if real_rows else 0)
        def one():
            p.ntotal = n
            return hot.derivs_host(p, mask, skip=skip)
        for _ in range(max(1, min(args.warmup, 2))):
            s_ = one()
        k = max(1, min(args.steps, 3))
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(k):
            one()
        b.record(stream)
        torch.cuda.synchronize()
        ms_ = a.elapsed_time(b) / k
        names = [nm for nm in DN_FULL if nm not in skip and (mhd or nm not in HYDRO_SKIP)]
        up = [nm for nm in UP_NAMES if mhd or nm not in ("Bevol", "psi")]
        return {"value": n / (ms_ * 1e-3), "unit": UNIT, "h2d_bytes_per_step": sum(rowbytes(nm) for nm in up) * n,
                "d2h_bytes_per_step": sum(rowbytes(nm) for nm in names) * (n if real_rows else s_["ntotal"]), "ms_per_step": ms_, "steps": k}, s_
    e2e_full, s = e2e([], False)
    e2e_lean, s = e2e(LEAN_SKIP, True)
    nt = s["ntotal"]

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
    b_dens, b_c2p, b_rates = c["bytes"]
    b_total = b_dens + b_c2p + b_rates
    achieved = b_rates * n / (pair_ms * 1e-3) / 1e9 if pair_ms > 0 else 0.0
    kern = ("rates_pair_kernel<%d,%s>" % (ndim, "hydro,DRAG" if c["kind"] == "dust" else "MHD,FAST=2"))
    traffic, ncu_extra = None, {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("config") == args.config and int(tj.get("nx", 0)) == c["nx"]:
                if tj.get("source_sha") == source_sha():
                    traffic = tj.get("rates_pair_kernel_dram_bytes_per_launch")
                    ncu_extra = {k: tj[k] for k in ("fp64_pipe_active_pct", "issue_active_pct", "l1tex_throughput_pct", "gpu_time_ms_under_ncu", "ncu_report") if k in tj}
                else:
                    ncu_extra = {"stale": "profiles/traffic.json was captured from another build of the kernels (source_sha differs): not reported"}
        except Exception:
            pass
    line = {
        "metric": c.get("metric", METRIC_FOR(c)), "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args),
        "run": {"npart": n, "ntotal_with_ghosts": nt, "itsdensity": s["itsdensity"], "nrelink": s["nrelink"], "nneigh_min": s["nneigh_min"], "nneigh_max": s["nneigh_max"],
                "npairs_rates": s["npairs_rates"], "lmax": s["lmax"], "source_sha": source_sha()},
        # headline e2e: the contract the Fortran shim uses on ordinary steps (rows [0,npart); arrays no host code reads between two derivs stay
        # on the device); full_contract: every output array of the reference's density/cons2prim/get_rates, ghost rows included
        "e2e": dict(e2e_lean, contract="lean: skips " + ",".join(LEAN_SKIP) + " + ghost rows (ND_DL_REAL_ROWS)", full_contract=e2e_full),
        "gpu_launches": launches,
        "clocks": ck,
        "roofline": {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": b_rates * n, "kernel_ms": pair_ms, "kernel_share_of_step": pair_ms / ms,
                     "ncu": ncu_extra,
                     "whole_step_GBps": b_total * n / (ms * 1e-3) / 1e9, "whole_step_frac": b_total * n / (ms * 1e-3) / 1e9 / peak,
                     "fp64": fp64_roofline(kern, s["ntrips_rates"], pair_ms, ck.get("sm_mhz")),
                     "note": "FP64 pairwise gather: the FP64 pipe and L1 gather bandwidth bind long before HBM (DESIGN.md); no tensor cores"},
        "phases_ms": phases,
    }
    # ---- whole leapfrog steps on the resident state: ndspmhd_b200_step = predictor + derivs + corrector;
    #      informational: these start from the predicted h of a running simulation, the timed `derivs` above from an unconverged guess
    try:
        dt_sim = min(0.25 * s["dtforce"], 0.3 * s["dtcourant"], 0.9 * s["dtdrag"], 0.25 * s["dtvisc"])
        for _ in range(3):      # the first steps pay one-off costs (lazy kernel loading, list buffers growing to the partial rounds' sizes)
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
    hot.close()
    lib.free_pinned(p)
    if not args.no_cpu:
        runs = cpu_oracle_runs(c, [args.cpu_nx] if args.cpu_nx else sample_sizes(c, 24.0))
        big = runs[-1]
        line["cpu_baseline"] = {"value": big["updates_per_s"], "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"one serial oracle derivs of the same workload at nx={big['nx']} = {big['npart']} particles ({big['seconds']:.1f} s, "
                                          f"{big['itsdensity']} density rounds); the reference is serial (docs/about.rst:12)",
                                "phases_ms": big["phases_ms"],
                                "fit": {"runs": [{k: r[k] for k in ("nx", "npart", "seconds", "updates_per_s")} for r in runs], "loglog_slope_time_vs_n": loglog_slope(runs),
                                        "note": "slope 1.0 = updates/s independent of N: the bounded sample stands for the full-size run"}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="slab512", choices=sorted(CONFIGS), help="workload (BASELINE.json configs); default = the headline 16.8 M thin slab")
    ap.add_argument("--nx", type=int, default=0, help="override the config's particles per unit length")
    ap.add_argument("--cpu-nx", type=int, default=0, help="size of the bounded CPU-baseline sample (default: three sizes within ~25 s)")
    ap.add_argument("--ref-nx", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fit", action="store_true", help="--impl reference: skip the single-core runs at smaller sizes")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-parity-check", action="store_true", help="N > 1: skip the single-GPU re-run that fills parity_vs_single")
    ap.add_argument("--transport", default="nccl", choices=["nccl", "callbacks"], help="N > 1: native NCCL in the library (default) or the nd_comm host callbacks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    import __graft_entry__ as g

    if args.impl == "reference":
        g.build(load=False)   # compiles; the product library is not loaded into this process
        return run_reference(args)
    g.build()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        from ndspmhd_b200 import slab_bench

        return slab_bench.run(args, cfg_of(args), workload, UNIT, config_dict, ClockSampler, measured_peaks)
    return run_single(args)


if __name__ == "__main__":
    sys.exit(main())
