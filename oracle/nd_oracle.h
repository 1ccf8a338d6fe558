/*
 * nd_oracle.h -- CPU oracle for the NDSPMHD hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (ndspmhd_b200/) never links, imports or calls it.
 *
 * PARITY UNPINNED BY THE REFERENCE: danieljprice/ndspmhd ships no tests, golden vectors or example
 * outputs, and neither this container nor the GPU box has a Fortran compiler, so the reference
 * cannot be run to produce fixtures.  The oracle is a line-by-line restatement of the cited Fortran
 * (serial, same loop order, doubles, no FMA contraction, no re-association) and is additionally
 * pinned by algorithm-independent invariants in tests/ (O(N^2) neighbour sets, kernel-table known
 * answers, momentum/energy conservation, lattice symmetry).
 */
#ifndef ND_ORACLE_H
#define ND_ORACLE_H
#include "../include/ndspmhd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* all arrays mutable, Fortran layout, capacity idim rows; every pointer must be non-NULL */
typedef struct ndo_arrays {
  double *x, *vel, *pmass, *hh, *en, *Bevol, *alpha, *psi, *rho;
  int *itype, *ireal;
  double *gradh, *gradhn, *gradsoft, *gradgradh, *rhoalt, *drhodt, *dhdt;
  int *numneigh;
  double *dens, *uu, *pr, *spsound, *Bfield, *sqrtg;
  double *force, *dudt, *dendt, *dBevoldt, *daldt, *dpsidt, *gradpsi, *fmag, *divB, *curlB,
         *graddivv, *del2u, *xsphterm;
  /* one-fluid dust (idust=1): dustfrac is in/out like the module array (read by `density`, rewritten by c2p) */
  double *dustevol, *dustfrac, *deltav, *rhogas, *rhodust, *ddustevoldt, *ddeltavdt;
} ndo_arrays;

enum { NDO_GHOSTS = 1, NDO_LINK = 2, NDO_DENSITY = 4, NDO_C2P = 8, NDO_RATES = 16, NDO_ALL = 31 };

/* one `derivs` (src/derivs.f90:74-156) restricted to the phases in `phases`.
 * *ntotal is in/out (set_ghost_particles rewrites it).  ms[5] = wall ms of ghosts, link, density, c2p, rates. */
int ndo_derivs(const nd_options *o, int ndim, ndo_arrays *a, int npart, int *ntotal, int idim,
               int phases, nd_scalars *s, double *ms);

/* one leapfrog `step` (src/stepND_leapfrog_mhd.f90:39-300) incl. its `call derivs` and the periodic wrap of `boundary`
 * (src/boundaryND.f90:65-93): on entry the rates arrays hold the previous derivs; *dt_inout in = this step's dt, out = the next
 * (:253) unless dtfixed.  `s` (required) receives the scalars of the inner derivs. */
int ndo_step(const nd_options *o, int ndim, ndo_arrays *a, int npart, int *ntotal, int idim, double *dt_inout, double C_cour, double C_force,
             int dtfixed, nd_scalars *s);

/* the sums of `evwrite` (src/evwrite_mhd.f90:124-284) over the host arrays, serial in the reference's order */
int ndo_evwrite(const nd_options *o, int ndim, const ndo_arrays *a, int npart, nd_evwrite *ev);

/* kernel tables as built by setkernels/setkerndrag (src/kernelND.f90:127-4289) */
int ndo_kernel_tables(int ikernel, int ikerneldrag, int ndim, double *wij, double *grwij, double *grgrwij,
                      double *wijdrag, double *radkern2, double *dq2table);
/* src/kernelND.f90:4426 and :4599 for one q2 (w, grw, grgrw) */
int ndo_interpolate(int ikernel, int ndim, double q2, double *w, double *grw, double *grgrw);

/* src/random.f90:61 ran1; state is process-global like the Fortran `save` */
double ndo_ran1(int *iseed);

/* O(N^2) pair finder after src/check_neighbourlist.f90:149-173: all (i<j) with q2i<radkern2 or q2j<radkern2,
 * i or j real.  Returns count; fills up to cap pairs (1-based). */
long long ndo_bruteforce_pairs(int ndim, const double *x, const double *hh, int npart, int ntotal,
                               double radkern2, int *pi, int *pj, long long cap);
/* pairs visited by the reference's rates loop (half stencil, src/ratesND_mhd.f90:304-467); needs a prior link */
long long ndo_linklist_pairs(const nd_options *o, int ndim, ndo_arrays *a, int npart, int ntotal, int idim,
                             int *pi, int *pj, long long cap);

/* `get_curl` (src/get_curl.f90:64-287) on the arrays as they stand after a derivs (rho, hh, gradh; rows npart..ntotal-1
 * are that derivs' ghosts): icurltype 1..4; Bvec and curlB are (3,idim), gradB (3,3,idim) or NULL (filled for icurltype 1 only). */
int ndo_get_curl(const nd_options *o, int ndim, ndo_arrays *a, int npart, int ntotal, int idim, int icurltype, const double *Bvec,
                 double *curlB, double *gradB);

const char *ndo_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
