/*
 * ndspmhd_b200.h -- C-ABI of the B200-native NDSPMHD hot path.
 *
 * The reference (danieljprice/ndspmhd, Fortran 90) has no FFI: its per-step hot
 * path is three argument-less external subroutines that talk through module
 * globals, called from `derivs` (src/derivs.f90:74-156):
 *
 *     call set_linklist        src/derivs.f90:82    -> src/linkND.f90:45
 *     call iterate_density     src/derivs.f90:92    -> src/iterate_density.f90:43
 *     call conservative2primitive  derivs.f90:98    -> src/conservative2primitive.f90:42
 *     call get_rates           src/derivs.f90:156   -> src/ratesND_mhd.f90:29
 *
 * This header declares what an ISO_C_BINDING shim behind those call sites binds
 * (see INTEGRATION.md and ndspmhd_b200/fortran/ for the shim sources).
 *
 * Conventions
 *  - every real is IEEE double (the reference is built with -fdefault-real-8,
 *    src/Makefile:27); every integer is 32-bit; logical options are int 0/1.
 *  - particle arrays are passed in the reference's NATIVE layout: column-major,
 *    1-based on the Fortran side, i.e. x(ndim,idim) is an array of `idim`
 *    records of `ndim` doubles; vel/Bfield/Bevol/alpha/force/... are (3,idim).
 *    `idim` is the allocated length (src/allocateND.f90:313), >= ntotal.
 *  - rows [0,npart) are real particles, rows [npart,ntotal) are ghosts made by
 *    set_ghost_particles (src/ghostND_mhd.f90:33); `ireal` holds the 1-based
 *    parent index exactly as the Fortran array does (src/ghostND_mhd.f90:423).
 *  - all functions return 0 on success or an ND_ERR_* code; the message is
 *    available from ndspmhd_b200_last_error().
 *  - no pointer is retained across calls (Fortran re-allocates its arrays when
 *    ghosts overflow, src/ghostND_mhd.f90:383-386).
 */
#ifndef NDSPMHD_B200_H
#define NDSPMHD_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (reference behaviour: print + `call quit`, ndspmhd.f90:369) ---- */
enum {
  ND_OK = 0,
  ND_ERR_INVALID_ARG = 1,
  ND_ERR_UNSUPPORTED_OPTION = 2, /* option tuple outside the compiled set: shim falls back to the CPU routine */
  ND_ERR_H_NONPOSITIVE = 3,      /* iterate_density.f90:99-102, ratesND_mhd.f90:384-387 */
  ND_ERR_RHO_NONPOSITIVE = 4,    /* iterate_density.f90:168-170 */
  ND_ERR_DENSITY_NOT_CONVERGED = 5, /* iterate_density.f90:349-351 */
  ND_ERR_VSIG_DET = 6,           /* ratesND_mhd.f90:1422-1429 */
  ND_ERR_CUDA = 7,
  ND_ERR_NO_DEVICE = 8,
  ND_ERR_LINK = 9,               /* linkND.f90:73-76, :94-99, :122-125 */
  ND_ERR_NEIGHBOUR_OVERFLOW = 10,
  ND_ERR_STATE = 11,             /* calls out of order (e.g. get_rates before upload) */
  ND_ERR_COMM = 12,              /* a transport callback failed, or another rank reported an error */
  ND_NEED_RELINK = 100           /* not an error: host-ghost mode, h grew past hhmax (iterate_density.f90:122-126, :260-262);
                                    caller must re-run set_ghost_particles, upload again and call iterate_density(resume=1) */
};

/* particle types, src/variablesND.f90:165-171 */
enum {
  ND_ITYPE_GAS = 0, ND_ITYPE_BND = 1, ND_ITYPE_DUST = 2, ND_ITYPE_GAS1 = 3,
  ND_ITYPE_GAS2 = 4, ND_ITYPE_BND2 = 11, ND_ITYPE_BNDDUST = 12
};

/*
 * Run-time options consumed by the path.  Names and meanings are those of the
 * reference modules `options`, `artvi`, `eos`, `setup_params`, `bound`, `part`
 * (src/variablesND.f90:33-37, :44-50, :135-155, :245; src/eos.f90:36); defaults
 * are src/defaults.f90:47-118.  ndspmhd_b200_default_options() fills them.
 */
typedef struct nd_options {
  /* module options */
  int iener, icty, iav, ikernav, ihvar, iprterm;
  int imhd, imagforce, idivbzero, iresist;
  int idust, idrag_nature;
  int ixsph, igravity, iexternal_force;
  int ikernel, ikernelalt;
  int maxdensits;
  int iavlim[3];
  int ibound[3];              /* per dimension: 0 none, 1 fixed, 2 reflecting ghosts, 3 periodic ghosts */
  int usenumdens, ibiascorrection, onef_dust, use_smoothed_rhodust;
  int islope_limiter, iuse_exact_derivs, iambipolar, ivisc, iquantum, ind_timesteps;
  int nsubsteps_divB;         /* uninitialised under leapfrog in the reference; treated as 0 (ratesND_mhd.f90:727) */
  /* library-side switches (no reference counterpart) */
  int device_ghosts;          /* 1: library makes the ghost rows itself (restating ghostND_mhd.f90) from rows [0,npart) */
  int want_aux;               /* 1: also produce rhoalt, gradhn, gradsoft, gradgradh (dead for the first-class tuple) */
  int idustevol;              /* options:idustevol (src/defaults.f90:98); only 0 (dust fraction itself) is supported with idust=1 */
  int reserved_i[5];
  /* reals */
  double hfact, psep, tolh;   /* setup_params, options */
  double gamma, polyk;        /* eos */
  double alphamin, alphaumin, alphaBmin, beta, avdecayconst, avfact; /* artvi */
  double psidecayfact, etamhd, Kdrag, damp, pext;
  double xmin[3], xmax[3];    /* bound */
  double Bconst[3];           /* part */
  double hhmax;               /* bound:hhmax as set by set_ghost_particles (host-ghost mode); ignored with device_ghosts */
  double reserved_d[8];
} nd_options;

/*
 * Pointers to the host (Fortran module) arrays.  NULL is allowed for arrays the
 * active option tuple does not touch.  "in" = read by upload, "out" = written
 * by download.  Shapes use `idim` as the trailing extent.
 */
typedef struct nd_arrays {
  /* --- in --- */
  const double *x;       /* (ndim,idim)  part:x       */
  const double *vel;     /* (3,idim)     part:vel     */
  const double *pmass;   /* (idim)       part:pmass   */
  const double *hh_in;   /* (idim)       part:hh (guess on entry) */
  const int    *itype;   /* (idim)       part:itype   */
  const int    *ireal;   /* (idim)       bound:ireal, 1-based, 0 = unset */
  const double *en;      /* (idim)       part:en      */
  const double *Bevol;   /* (3,idim)     part:Bevol   */
  const double *alpha;   /* (3,idim)     part:alpha   */
  const double *psi;     /* (idim)       part:psi     */
  const double *rho_in;  /* (idim)       part:rho on entry (kept on fixed particles with ireal=0) */
  /* --- out: density phase (iterate_density) --- */
  double *hh;            /* (idim) */
  double *rho;           /* (idim) */
  double *gradh;         /* (idim) hterms:gradh */
  double *drhodt;        /* (idim) rates:drhodt */
  double *dhdt;          /* (idim) rates:dhdt   */
  int    *numneigh;      /* (idim) linklist:numneigh */
  double *rhoalt;        /* (idim) want_aux */
  double *gradhn;        /* (idim) want_aux */
  double *gradsoft;      /* (idim) want_aux */
  double *gradgradh;     /* (idim) want_aux */
  /* --- out: conservative2primitive --- */
  double *dens;          /* (idim) */
  double *uu;            /* (idim) */
  double *pr;            /* (idim) */
  double *spsound;       /* (idim) */
  double *Bfield;        /* (3,idim) */
  /* --- out: get_rates --- */
  double *force;         /* (3,idim) */
  double *dudt;          /* (idim) */
  double *dendt;         /* (idim) */
  double *dBevoldt;      /* (3,idim) */
  double *daldt;         /* (3,idim) */
  double *dpsidt;        /* (idim) */
  double *gradpsi;       /* (3,idim) */
  double *divB;          /* (idim) */
  double *curlB;         /* (3,idim) */
  double *graddivv;      /* (3,idim) */
  double *del2u;         /* (idim)  local array in the reference (ratesND_mhd.f90:168); exposed for parity tests */
  /* --- out: ghost rows when device_ghosts=1 (rows [npart,ntotal)) --- */
  double *x_out;         /* (ndim,idim) */
  double *vel_out;       /* (3,idim) */
  int    *ireal_out;     /* (idim) */
  int    *itype_out;     /* (idim) */
  /* --- one-fluid dust, idust=1 (src/allocateND.f90:386-392); NULL otherwise --- */
  const double *dustevol;    /* (idim)   in:  part:dustevol(1,:), the evolved dust variable (idustevol=0: the dust fraction) */
  const double *dustfrac_in; /* (idim)   in:  part:dustfrac(1,:) on entry -- `density` reads the value left by the previous
                                              conservative2primitive (src/density_sums.f90:280-282), not the one made from dustevol */
  const double *deltav;      /* (3,idim) in:  part:deltav */
  double *dustfrac;          /* (idim)   out: part:dustfrac after conservative2primitive (src/conservative2primitive.f90:76-113) */
  double *rhogas;            /* (idim)   out: part:rhogas  (density sums, src/density_sums.f90:278-292) */
  double *rhodust;           /* (idim)   out: part:rhodust(1,:) */
  double *ddustevoldt;       /* (idim)   out: rates:ddustevoldt(1,:) */
  double *ddeltavdt;         /* (3,idim) out: rates:ddeltavdt */
  /* --- iavlim(3) = 2: conservative2primitive rewrites part:alpha(3,:) (Tricco & Price resistivity switch, conservative2primitive.f90:299-311);
         normally the same memory as `alpha`.  Comes down with ND_DL_PRIM; NULL = not wanted --- */
  double *alpha_out;         /* (3,idim) */
} nd_arrays;

/* download masks */
enum {
  ND_DL_DENSITY = 1u,   /* hh rho gradh drhodt dhdt numneigh (+aux) (+rhogas rhodust with idust=1) */
  ND_DL_PRIM    = 2u,   /* dens uu pr spsound Bfield (+dustfrac) */
  ND_DL_RATES   = 4u,   /* force dudt dendt dBevoldt daldt dpsidt gradpsi divB curlB graddivv del2u (+ddustevoldt ddeltavdt) */
  ND_DL_GHOSTS  = 8u,   /* x_out vel_out ireal_out itype_out rows [npart,ntotal) */
  ND_DL_ALL     = 15u,
  /* modifier: the groups above come down for rows [0,npart) only.  The ghost rows of those arrays are copies of their parents
     (src/iterate_density.f90:330-344, src/conservative2primitive.f90:441-467) or zeros (src/ratesND_mhd.f90:949-965) and are rebuilt by
     the next set_ghost_particles; leave the bit clear when the caller reads them (e.g. a dump of ghost rows). */
  ND_DL_REAL_ROWS = 16u
};

/* scalars returned by the path: module timestep (src/variablesND.f90:255-273), hterms:itsdensity, bound:hhmax */
typedef struct nd_scalars {
  double dtcourant, dtforce, dtav, dtdrag, dtvisc, vsig2max, vsigmax;
  double stressmax, ts_min, h_on_csts_max, fhmax;
  double hhmax, dxcell;
  double fmean[3];        /* sum m*force, the reference's momentum-conservation diagnostic (ratesND_mhd.f90:678) */
  int itsdensity, nneigh_min, nneigh_max, nclumped;
  int ntotal, ncells, ncellsx[3], nrelink;
  long long ncalctotal;   /* total particle-density evaluations over all rounds (iterate_density.f90:154) */
  /* diagnostics of the list machinery (no reference counterpart): column capacity of the neighbour lists, how many list builds had
     to be repeated with a larger capacity since create, and in how many row chunks the last get_rates ran (derivs_host pipelines
     the download of one chunk with the pair kernel of the next) */
  int lmax, list_overflows, rate_chunks;
  int reserved_i[5];
  /* ordered pairs (i real, j any row) the last get_rates evaluated = sum of the rates list lengths: an order-independent integer
     checksum of the neighbour finding (summed over ranks with slabs; bench.py --gpus N compares it with the single-GPU run) */
  long long npairs_rates;
  /* warp trips of the rates pair loop on THIS context (sum over 32-target list blocks of the longest list): bench.py's FP64-pipe floor */
  long long ntrips_rates;
} nd_scalars;

typedef struct nd_ctx nd_ctx;

/* fills `o` with src/defaults.f90:47-118 (+ avfact from initialiseND_mhd.f90:168-172 for the default gamma) */
int ndspmhd_b200_default_options(nd_options *o);

/* library / device probing; no compute */
int ndspmhd_b200_version(void);
int ndspmhd_b200_device_count(void);

/* one context per host thread and GPU.  `device` = CUDA ordinal.  Builds the kernel tables
 * (setkernels + setkerndrag, src/initialiseND_mhd.f90:179-216, src/kernelND.f90:127-4289). */
int ndspmhd_b200_create(const nd_options *o, int ndim, int device, nd_ctx **out);
int ndspmhd_b200_set_options(nd_ctx *c, const nd_options *o);
int ndspmhd_b200_destroy(nd_ctx *c);
const char *ndspmhd_b200_last_error(const nd_ctx *c);

/* copy of the kernel tables the device uses: w, grw, grgrw, wdrag each (0:4000); returns radkern2, dq2table */
int ndspmhd_b200_get_kernel_tables(const nd_ctx *c, double *wij, double *grwij, double *grgrwij,
                                   double *wijdrag, double *radkern2, double *dq2table);

/* host -> device.  With device_ghosts=0 rows [0,ntotal) are taken as given; with device_ghosts=1 only
 * rows [0,npart) are read and `ntotal` is ignored. */
int ndspmhd_b200_upload(nd_ctx *c, const nd_arrays *a, int npart, int ntotal, int idim);

/* host-ghost mode only, after ND_NEED_RELINK: the caller has downloaded hh, re-run set_ghost_particles
 * (src/iterate_density.f90:123) and hands over the new ghost rows [npart,ntotal) of x, vel, itype, ireal and the new
 * bound:hhmax; then calls ndspmhd_b200_iterate_density(c, 1, s). */
int ndspmhd_b200_update_ghosts(nd_ctx *c, const nd_arrays *a, int ntotal, int idim, double hhmax);

/* replaces `call set_linklist` (src/linkND.f90:45): (ghosts if device_ghosts) + cell grid + cell-sorted SoA */
int ndspmhd_b200_link(nd_ctx *c);

/* replaces `call iterate_density` (src/iterate_density.f90:43).  resume=1 continues after ND_NEED_RELINK. */
int ndspmhd_b200_iterate_density(nd_ctx *c, int resume, nd_scalars *s);

/* replaces `call conservative2primitive` (element-wise branches, src/conservative2primitive.f90:42) + eos.f90:40 */
int ndspmhd_b200_cons2prim(nd_ctx *c);

/* replaces `call get_rates` (src/ratesND_mhd.f90:29) */
int ndspmhd_b200_get_rates(nd_ctx *c, nd_scalars *s);

/* link + iterate_density + cons2prim + get_rates on the resident state = one `derivs` (src/derivs.f90:74-156) */
int ndspmhd_b200_derivs(nd_ctx *c, nd_scalars *s);

/* upload + derivs + download of the arrays selected by `mask` in one call, copies overlapped with the kernels on separate
 * streams.  This is the call a host that keeps its state in host memory (the reference's integrators) should make. */
int ndspmhd_b200_derivs_host(nd_ctx *c, nd_arrays *a, int npart, int ntotal, int idim, unsigned mask, nd_scalars *s);

/* device -> host, rows [0,ntotal) of the arrays selected by `mask` (NULL pointers skipped) */
int ndspmhd_b200_download(nd_ctx *c, nd_arrays *a, unsigned mask, int idim);

/*
 * Multi-GPU: 1-D slab decomposition along x, one context (one rank) per GPU.  The reference is serial; this is the
 * device-side equivalent of its periodic ghosts (src/ghostND_mhd.f90:166-346) applied at slab faces: every rank keeps
 * the particles with slab_lo <= x(1) < slab_hi, and the library appends HALO rows -- copies of the neighbours'
 * particles within radkern*hhmax of the face, shifted with the reference's ghost arithmetic across the periodic wrap --
 * before it makes the y/z ghosts locally.  The library owns selection, packing, row layout and the order of
 * operations; the host supplies the transport as callbacks (torch.distributed/NCCL in ndspmhd_b200/slab.py; MPI from a
 * Fortran host).  Per derivs: one all-reduce of hhmax, one halo exchange of the inputs (x, vel, pmass, hh, en, Bevol,
 * alpha, psi, rho, itype), one small all-reduce per density round (unconverged count, relink flag), one halo exchange of
 * (hh, rho, gradh) after the iteration, all-reduces of stressmax, vsigmax and the timestep scalars.
 * Requires device_ghosts = 1 and ibound(1) in {0, 1, 3}.  With nranks = 1 the callbacks are never used.
 */
typedef struct nd_comm {
  void *user;                   /* passed back to every callback */
  int rank, nranks;
  double slab_lo, slab_hi;      /* this rank's x interval */
  long long nglobal;            /* real particles over all ranks (decides `density` vs `density_partial`, iterate_density.f90:131) */
  /* in-place all-reduce of n HOST doubles; op 0 = max, 1 = min, 2 = sum */
  int (*allreduce)(void *user, double *v, int n, int op);
  /* byte counts handshake with the two x-neighbours: side 0 = rank-1, side 1 = rank+1 (periodic ring) */
  int (*sendrecv_counts)(void *user, const long long sendbytes[2], long long recvbytes[2]);
  /* payload: DEVICE buffers; sendbuf[s] -> neighbour s, recvbuf[s] <- neighbour s; ordered after/before work on `stream` */
  int (*sendrecv)(void *user, void *const sendbuf[2], const long long sendbytes[2], void *const recvbuf[2],
                  const long long recvbytes[2], void *stream);
} nd_comm;

/* attach (or detach with NULL) the transport; call before upload.  upload then takes only this rank's own rows. */
int ndspmhd_b200_set_comm(nd_ctx *c, const nd_comm *comm);
/*
 * The native transport: the library calls NCCL itself, on its own stream (all-reduces of device scalars, grouped ncclSend/ncclRecv of the
 * halo buffers, all-gather of the halo byte counts) -- no host callback on the path.  libnccl.so.2 is opened with dlopen at the first call
 * (the copy the host process already holds, e.g. torch's, or NDSPMHD_B200_NCCL_LIB).  Rank 0 makes the 128-byte id, the host program
 * broadcasts it by its own means (torch.distributed, MPI_Bcast), every rank then calls set_comm_nccl (collective: ncclCommInitRank).
 */
int ndspmhd_b200_nccl_unique_id(unsigned char id[128]);
int ndspmhd_b200_set_comm_nccl(nd_ctx *c, const unsigned char id[128], int rank, int nranks, double slab_lo, double slab_hi, long long nglobal);
/* counters since create: all-reduces issued (either transport), halo payload bytes this rank sent */
int ndspmhd_b200_comm_stats(const nd_ctx *c, long long *n_allreduce, long long *halo_bytes_sent);
/*
 * Row ids.  upload gives every row its row number; a slab-decomposed run sets global ids after upload.  ndspmhd_b200_step on a
 * slab-decomposed context re-owns the rows whose x left [slab_lo, slab_hi) after the predictor and the periodic wrap
 * (src/stepND_leapfrog_mhd.f90:145, src/boundaryND.f90:65-93): they travel to the adjacent rank with their evolved state and the
 * integrator's `*in` copies, holes are filled from the tail of the own rows and arrivals are appended -- so row numbers change and the
 * ids are how a caller follows a particle.  Moving farther than the adjacent slab in one step is ND_ERR_INVALID_ARG, and so is leaving
 * [xmin, xmax] through an open x boundary (ibound(1) = 0: the end slabs are not open-ended); fixed-particle boundaries are refused on
 * slab contexts (fixed rows are tied to their partners' row numbers).
 */
int ndspmhd_b200_set_row_ids(nd_ctx *c, const long long *ids, int n);
int ndspmhd_b200_get_row_ids(nd_ctx *c, long long *ids, int cap);
/* counters since create: rows that left / arrived, payload bytes this rank sent for them */
int ndspmhd_b200_migration_stats(const nd_ctx *c, long long *rows_out, long long *rows_in, long long *bytes_sent);
/* rows after the last link: own rows [0,nown), halo rows [nown,nsrc), ghosts [nsrc,ntotal) */
int ndspmhd_b200_row_counts(const nd_ctx *c, int *nown, int *nsrc, int *ntotal);

/*
 * Leapfrog integrator on the RESIDENT state: one call = `call step`
 * (src/stepND_leapfrog_mhd.f90:39-300) -- predictor (:108-163), `call derivs` (:167), corrector (:171-209),
 * `call boundary` (:216 -> src/boundaryND.f90:65-93, particles crossing a periodic domain), new timestep (:239-253).
 * Between dumps no particle array crosses PCIe.  Needs a prior upload + derivs (the reference enters `step` with the
 * rates of the previous call) and, with ghost boundaries, device_ghosts = 1.
 * nd_step_opts mirrors module timestep (C_cour, C_force, dtfixed: src/variablesND.f90:255-273); `damp` is nd_options.damp.
 * *dt_inout: in = dt of this step, out = dt of the next: min(C_force*dtforce, C_cour*dtcourant, 0.9*dtdrag, C_force*dtvisc).
 */
typedef struct nd_step_opts {
  double C_cour, C_force;
  int dtfixed;
  int reserved;
} nd_step_opts;
int ndspmhd_b200_step(nd_ctx *c, const nd_step_opts *so, double *dt_inout, nd_scalars *s);

/* the evolved state after steps: host arrays for rows [0,npart), Fortran layout; NULL pointers are skipped */
typedef struct nd_state_out {
  double *x, *vel, *hh, *en, *Bevol, *alpha, *psi, *rho;
  double *dustevol, *deltav;   /* idust = 1 */
} nd_state_out;
int ndspmhd_b200_download_state(nd_ctx *c, const nd_state_out *st, int idim);

/*
 * Per-step diagnostics on the resident state: the sums of `evwrite`, src/evwrite_mhd.f90:27-320, called
 * every step by the reference (src/evolve.f90:62,165) -- as device reductions (fixed-order, run-to-run deterministic) so
 * that a simulation stepping on the device does not have to download the particles to write its .ev line.
 * Not produced: fdotBav/fdotBmax/force_err_av/force_err_max (they need fmagarray:fmag, which the hot path does not keep)
 * and epot (external forces / self-gravity are outside the supported tuple: 0).
 */
typedef struct nd_evwrite {
  double ekin, etherm, emag, epot, etot, momtot, angtot, rhomax, rhomean, rhomin;   /* :296-301 columns 2-10 */
  double emagp, crosshel, betamhdmin, betamhdav, divBav, divBmax, divBtot;           /* MHD columns */
  double omegamhdav, omegamhdmax, fracdivBok, fluxtotmag;
  double ekiny, dmomtot, totmassgas, totmassdust;                                    /* hydro / one-fluid dust columns */
  double mom[3], dmom[3], ang[3], fluxtot[3];
  double reserved[8];
} nd_evwrite;
int ndspmhd_b200_evwrite(nd_ctx *c, nd_evwrite *ev);

/*
 * `get_curl` (src/get_curl.f90:64-287) as an operator on the resident state, on the same cell grid / neighbour-list / pair
 * engine as the rates: curl of Bvec (host, (3,idim), rows [0,npart) read; ghost rows take their parent's value) into curlB (host, (3,idim),
 * rows [0,npart) written) and, for icurltype = 1, optionally grad Bvec into gradB ((3,3,idim), gradB(l,k,i) = d Bvec_k / d x_l).
 * icurltype: 1 differenced, mass-weighted with grad-h (default); 2 symmetric; 3 constant weights; 4 m_j/rho_j^2 weights.
 * Needs a prior link + iterate_density (it uses the converged rho, h, gradh).  The reference's call site on the path -- the Tricco & Price
 * resistivity switch inside conservative2primitive, iavlim(3) = 2 -- is run by ndspmhd_b200_cons2prim / _derivs themselves.
 */
int ndspmhd_b200_get_curl(nd_ctx *c, int icurltype, const double *Bvec, double *curlB, double *gradB, int idim);

/* page-locked host memory for the caller's particle arrays (makes upload/download run at PCIe speed) */
void *ndspmhd_b200_host_alloc(size_t bytes);
void ndspmhd_b200_host_free(void *p);

/* bench/diagnostic hooks: per-phase device time of the last derivs call (ms): link, density, c2p, rates pair, rates final */
int ndspmhd_b200_last_timings(const nd_ctx *c, double ms[8]);
/* puts hh back to the guess handed over by the last upload (a D2D copy), so that a repeated derivs() on the resident
 * state repeats the whole smoothing-length iteration instead of starting from the converged answer */
int ndspmhd_b200_rewind(nd_ctx *c);
/* device self-test of the branch-free FP64 sqrt / 1/sqrt the pair kernels use (host arrays in and out) */
int ndspmhd_b200_selftest_math(nd_ctx *c, const double *in, double *out_sqrt, double *out_rsqrt, int n);
/* number of kernels this library launched since create (bench.py's gpu_launches) */
long long ndspmhd_b200_launch_count(const nd_ctx *c);
/* the CUDA stream the context launches on (cudaStream_t as void*), for event timing in bench.py */
void *ndspmhd_b200_stream(const nd_ctx *c);
/* neighbour pair list of the rates pass for parity tests: fills up to `cap` (i,j) pairs (1-based, i real) in
 * unspecified order, returns the total count in *npairs */
int ndspmhd_b200_rates_pairs(nd_ctx *c, int *pair_i, int *pair_j, long long cap, long long *npairs);

#ifdef __cplusplus
}
#endif
#endif /* NDSPMHD_B200_H */
