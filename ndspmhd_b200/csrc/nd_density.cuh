// Density / smoothing-length iteration kernels.
//
// Replaces the reference's `density` (src/density_sums.f90:38-379), `density_partial` (:396-658) and the
// per-particle Newton-Raphson of `iterate_density` (src/iterate_density.f90:163-278) with one gather kernel per
// round: round 1 over all real particles (FIRST), later rounds over the compacted list of unconverged particles.
// A target's sums only involve W(h_target), so the one-sided gather reproduces every pair term of the reference's
// symmetric loop; convergence is decided per particle on the device.
#pragma once
#include "nd_device.cuh"

namespace ndk {

struct DensityArgs {
  // original-order arrays (row = Fortran index - 1)
  double *hh;            // in/out current h
  const double *hhin;    // h on entry to iterate_density (convergence test, :242)
  double *rho, *gradh, *drhodt, *dhdt;
  int *numneigh;
  double *rhoalt, *gradhn, *gradsoft, *gradgradh;   // AUX
  const int *list;       // sorted slots to process (NULL in the FIRST round: slot = global thread id)
  int nlist;
  int *redo;             // [ntotal] by sorted slot: 1 = not converged, recompute next round
  int *flags;            // [0] relink requested, [1] error code, [2] rho<=1e-6 count
  int itsdensity, itsdensitymax;
  double hfact, psep, tolh, hhmax;
};

constexpr int DENS_BLOCK = 128;
constexpr int DENS_CAP = 96;

template <int NDIM, bool FIRST, bool AUX>
__global__ void __launch_bounds__(DENS_BLOCK) density_round_kernel(Grid G, DensityArgs A) {
  extern __shared__ unsigned nlist_smem[];
  const int gid = blockIdx.x * DENS_BLOCK + threadIdx.x;
  int s = -1;
  if (FIRST) { if (gid < G.ntotal) s = gid; }
  else if (gid < A.nlist) s = A.list[gid];
  int orig = -1, ti = 0;
  bool active = false;
  if (s >= 0) {
    orig = G.perm[s];
    ti = G.typ[s];
    active = orig < G.npart;                       // ghosts are sources only
    if (!FIRST && ti == T_BND) active = false;     // density_sums.f90:510
  }
  double xi = 0, yi = 0, zi = 0, vxi = 0, vyi = 0, vzi = 0, mi = 0, hi = 1;
  int celli = 0;
  if (active) {
    double4 p = ld4(G.posh + s), v = ld4(G.vm + s);
    xi = p.x; yi = p.y; zi = p.z;
    vxi = v.x; vyi = v.y; vzi = v.z; mi = v.w;
    hi = A.hh[orig];                               // current h (== p.w in the first round)
    celli = G.cellOf[s];
  }
  const double hi1 = 1.0 / hi;                     // h1(i) = 1./hh(i), density_sums.f90:130
  const double hi21 = __dmul_rn(hi1, hi1);
  const double hfacwabi = powndim<NDIM>(hi1);
  int nneigh = 0;
  double rho = 0, gradh = 0, drhodt = 0, densn = 0, gradhn = 0, gradgradh = 0;

  // ---- phase 1: inclusion test (bit-exact arithmetic) ----
  auto cull = [&](int k) -> bool {
    const double4 pj = ld4(G.posh + k);
    const int tj = __ldg(G.typ + k);
    const double rij2 = dist2_exact(xi - pj.x, yi - pj.y, zi - pj.z);
    if (FIRST) {
      if (!types_interact(ti, tj)) return false;   // density_sums.f90:169-174
      // The reference visits the pair once; "i" is the particle met first: lower cell index, or the later-inserted
      // (higher index) particle of the same chain.  q2 of "i" is rij2*hi21, q2 of "j" is (rij2*hj1)*hj1 (:182-183).
      const int cellj = __ldg(G.cellOf + k), origj = __ldg(G.perm + k);
      const bool iam_i = (celli < cellj) || (celli == cellj && orig >= origj);
      const double hj1 = 1.0 / pj.w;
      double q2me, q2ot;
      if (iam_i) { q2me = __dmul_rn(rij2, hi21); q2ot = __dmul_rn(__dmul_rn(rij2, hj1), hj1); }
      else { q2me = __dmul_rn(__dmul_rn(rij2, hi1), hi1); q2ot = __dmul_rn(rij2, __dmul_rn(hj1, hj1)); }
      // :189-190 with the target real: q2i<radkern2 .or. q2j<radkern2
      const bool mine = q2me < G.radkern2;
      if (mine || q2ot < G.radkern2) nneigh++;     // :196-197
      return mine;                                  // terms with q2me >= radkern2 are exact zeros (table end = 0)
    } else {
      if (tj != ti && tj != T_BND) return false;   // density_sums.f90:517
      const double q2i = __dmul_rn(rij2, hi21);
      if (q2i < G.radkern2) { nneigh++; return true; }   // :528-532
      return false;
    }
  };

  // ---- phase 2: pair sums ----
  auto body = [&](int k) {
    const double4 pj = ld4(G.posh + k);
    const double4 vj = ld4(G.vm + k);
    const double dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
    const double rij2 = dist2_exact(dx, dy, dz);
    double q2i;
    if (FIRST) {
      const int cellj = __ldg(G.cellOf + k), origj = __ldg(G.perm + k);
      const bool iam_i = (celli < cellj) || (celli == cellj && orig >= origj);
      q2i = iam_i ? __dmul_rn(rij2, hi21) : __dmul_rn(__dmul_rn(rij2, hi1), hi1);
    } else q2i = __dmul_rn(rij2, hi21);
    const double rij = sqrt(rij2);
    const bool self = (k == s);
    const double pmassj = vj.w;
    double wabi, grkerni, grgrkerni = 0.;
    if (AUX) interp_wggg(G, q2i, wabi, grkerni, grgrkerni);
    else interp_wg(G, q2i, wabi, grkerni);
    wabi = wabi * hfacwabi;                        // :237-241 / :549-553
    grkerni = grkerni * hfacwabi * hi1;
    const double dwdhi = -rij * grkerni * hi1 - NDIM * wabi * hi1;   // :260
    // self pair: the symmetric loop adds weight 1/2 twice (:204-208, :274, :286); the gather adds it once in full
    const bool bnd_first = FIRST && ti == T_BND;   // :273, :321 -- fixed particles keep rho, gradh
    if (!bnd_first) {
      rho += pmassj * wabi;
      gradh += pmassj * dwdhi;
      if (AUX) {
        densn += wabi;                             // wabalt == wab: ikernelalt = ikernel
        gradhn += dwdhi;                           // :323 / :595
        grgrkerni = grgrkerni * hfacwabi * hi1 * hi1;
        const double dwdhdhi = NDIM * (NDIM + 1) * wabi * (hi1 * hi1) + 2. * (NDIM + 1) * rij * (hi1 * hi1) * grkerni +
                               (rij * rij) * (hi1 * hi1) * grgrkerni;        // :265
        gradgradh += pmassj * dwdhdhi;
      }
    }
    if (!self) {                                   // :297-303
      const double rinv = 1.0 / (rij + 2.220446049250313e-16);   // dr = dx/(rij + epsilon(rij)), :199
      const double dvdotr = ((vxi - vj.x) * (dx * rinv) + (vyi - vj.y) * (dy * rinv)) + (vzi - vj.z) * (dz * rinv);
      drhodt += pmassj * dvdotr * grkerni;
    }
  };

  neighbour_walk<NDIM, DENS_CAP, DENS_BLOCK>(G, active, celli, nlist_smem, cull, body);

  if (s < 0 || orig >= G.npart) return;
  if (!active) {                                   // fixed particle skipped by density_partial: stays as it was
    A.redo[s] = 0;
    return;
  }
  // ---- Newton-Raphson update, src/iterate_density.f90:163-278 ----
  A.numneigh[orig] = nneigh;
  int redo = 0;
  if (ti != T_BND && ti != T_BNDDUST) {
    if (rho <= 1.e-6) {
      if (rho <= 0.) { atomicCAS(&A.flags[1], 0, 4 /*ND_ERR_RHO_NONPOSITIVE*/); }
      else atomicAdd(&A.flags[2], 1);
    }
    const double rhoi = mi / powndim<NDIM>(hi / A.hfact);            // :192 (h_min = rhomin = 0)
    const double dhdrhoi = -hi / (NDIM * rho);                        // :193
    const double dwdhsumi = gradh;
    double omegai = 1. - dhdrhoi * gradh;                             // :196
    if (omegai < 1.e-5) { if (fabs(omegai) == 0.) omegai = 1.; }
    const double gradh_out = 1. / omegai;                             // :201
    const double func = rhoi - rho;
    const double dfdh = omegai / dhdrhoi;
    double hnew = hi - func / dfdh;                                   // :212
    if (hnew > 1.2 * hi) hnew = 1.2 * hi;
    else if (hnew < 0.8 * hi) hnew = 0.8 * hi;
    if (hnew <= 0. || gradh_out <= 2.2250738585072014e-308) hnew = A.hfact * pow(mi / rho, 1.0 / NDIM);   // :225-227
    else if (A.itsdensity > 100) hnew = A.hfact * pow(mi / rho, 1.0 / NDIM);                               // :228-229
    bool relink = false;
    if (nneigh <= 1) { hnew = hi + A.psep; relink = true; }           // :231-237
    const bool converged = (fabs((hnew - hi) / A.hhin[orig]) < A.tolh && omegai > 0.) || A.itsdensitymax == 0;   // :242
    A.rho[orig] = rho;
    A.gradh[orig] = gradh_out;
    if (AUX) {
      A.rhoalt[orig] = densn;
      A.gradhn[orig] = gradhn;                                        // raw sum (usenumdens = .false.)
      const double d2hdrho2i = hi * (NDIM + 1) / ((rho * NDIM) * (rho * NDIM));   // :206
      A.gradgradh[orig] = rho * (d2hdrho2i * dwdhsumi + (dhdrhoi * dhdrhoi) * gradgradh);
      A.gradsoft[orig] = 0.;                                          // :204 with igravity = 0
    }
    if (!converged) {
      redo = 1;
      if (A.itsdensity <= A.itsdensitymax) A.hh[orig] = hnew;         // :253-255
      if (hnew > A.hhmax) relink = true;                              // :260-262
      A.drhodt[orig] = drhodt;
    } else {
      const double d = drhodt * gradh_out;                            // :272-273
      A.drhodt[orig] = d;
      A.dhdt[orig] = dhdrhoi * d;
    }
    if (relink) A.flags[0] = 1;
  } else {
    A.drhodt[orig] = drhodt;                                          // density_sums.f90:300 runs for fixed particles too
  }
  A.redo[s] = redo;
}

}  // namespace ndk
