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
  const int *list;       // sorted slots to process (NULL in the FIRST round: slot = s0 + thread id)
  int nlist, s0;
  int *redo;             // [ntotal] by sorted slot: 1 = not converged, recompute next round
  int *flags;            // [0] relink requested, [1] error code, [2] rho<=1e-6 count
  int itsdensity, itsdensitymax;
  double hfact, psep, tolh, hhmax;
  // one-fluid dust (AUX instantiation only; sdf NULL otherwise): rhogas, rhodust sums of src/density_sums.f90:278-292, :569-573
  const double *sdf; double *rhogas, *rhodust;
  int *sched;            // work counter of the persistent warps (zeroed before each launch)
};

#ifndef ND_DENS_LIGHT
#define ND_DENS_LIGHT 1   // 1: in a fused derivs of the fast option tuple the density rounds run LIGHT (no drho/dt sum, one gather a pair) and the rates pair kernel makes drho/dt
#endif
// The {W, slope} and {grad W, slope} rows live in shared memory (two TMA bulk copies per persistent block, one block per SM).
#ifndef ND_DENS_BLOCK
#define ND_DENS_BLOCK 512
#endif
#ifndef ND_DENS_BLOCK_LIGHT
#define ND_DENS_BLOCK_LIGHT 1024   // LIGHT instantiations: 64 registers without spills, 32 warps on an SM instead of 16
#endif
constexpr int DENS_BLOCK = ND_DENS_BLOCK;
constexpr int DENS_BLOCK_LIGHT = ND_DENS_BLOCK_LIGHT;
constexpr int DENS_TAB_BYTES = (IKERN + 1) * 16;                         // one {value, slope} table, 64016 B
constexpr int DENS_TAB_STRIDE = ((DENS_TAB_BYTES + 127) / 128) * 128;
constexpr int DENS_SMEM_BYTES = 128 + 2 * DENS_TAB_STRIDE;

// LIGHT: the round makes rho, gradh and the Newton-Raphson update only.  drho/dt is left to the rates pair kernel of the same derivs
// (fast option tuple), which visits the same pairs with the same grad W anyway: a neighbour then costs ONE 32-byte gather
// ({x,y,z,m}) instead of two (position + velocity records), and the L1 gather path is what bounds this kernel.
template <int NDIM, bool FIRST, bool AUX, bool LIGHT>
__device__ __forceinline__ void density_target(const Grid &G, const DensityArgs &A, const NbrLists &L, int t, const double2 *tabw, const double2 *tabg) {
  const int s = FIRST ? A.s0 + t : A.list[t];
  const int orig = G.perm[s];
  const int ti = G.typ[s];
  bool active = orig < G.nown;                     // ghosts and halo rows are sources only
  if (!FIRST && ti == T_BND) active = false;       // density_sums.f90:510
  double xi = 0, yi = 0, zi = 0, vxi = 0, vyi = 0, vzi = 0, mi = 0, hi = 1;
  int cs0 = 0, cs1 = 0, cnt = 0;
  if (active) {
    if (LIGHT) {                                   // LIGHT rounds read no velocity at all (derivs_host may still be uploading them)
      const double4 p = ld4(G.posm + s);
      xi = p.x; yi = p.y; zi = p.z; mi = p.w;
    } else {
      const double4 p = ld4(G.posh + s), v = ld4(G.vm + s);
      xi = p.x; yi = p.y; zi = p.z;
      vxi = v.x; vyi = v.y; vzi = v.z; mi = v.w;
    }
    hi = A.hh[orig];                               // current h (1/h == p.w in the first round)
    const int celli = G.cellOf[s];
    cs0 = cell_begin(G, celli); cs1 = cell_begin(G, celli + 1);
    cnt = L.cnt[t];
  }
  const double hi1 = 1.0 / hi;                     // h1(i) = 1./hh(i), density_sums.f90:130
  const double hi21 = __dmul_rn(hi1, hi1);
  const double hfacwabi = powndim<NDIM>(hi1);
  double rho = 0, gradh = 0, drhodt = 0, densn = 0, gradhn = 0, gradgradh = 0, rhogas = 0, rhodust = 0;
  const bool onef = AUX && A.sdf != nullptr;

  // ---- pair sums over the neighbour list (built by build_lists_kernel with the reference's inclusion test) ----
  auto body = [&](int k, const double4 &pj, const double4 &vj) {
    const double dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
    const double rij2 = dist2_exact(dx, dy, dz);
    double q2i;
    // the reference rounds q2 of the first-met particle of a pair as rij2*hi21 and of the other as (rij2*hi1)*hi1 (:182-183);
    // inside a cell "first met" is decided here by slot order, (fine bin, index), where the reference's chain order is the
    // index alone: a last-bit difference of q2 for some same-cell pairs, far inside the 1e-12 tolerance (the neighbour SET is
    // decided exactly by build_lists_kernel, which compares the original rows)
    if (FIRST) q2i = ((k >= cs1) || (k >= cs0 && k <= s)) ? __dmul_rn(rij2, hi21) : __dmul_rn(__dmul_rn(rij2, hi1), hi1);
    else q2i = __dmul_rn(rij2, hi21);
    // rij = sqrt(rij2) and dr = dx/(rij + epsilon(rij)) (:199) without a divide: 1/(r+e) = (1/r)(1 - e/r) to O((e/r)^2) ~ 1e-26
    const double rinv = rsqrt_nr(rij2);          // 0 for the self pair
    const double rij = rij2 * rinv;
    const double rinve = rinv - 2.220446049250313e-16 * rinv * rinv;
    const double pmassj = LIGHT ? pj.w : vj.w;
    double wabi, grkerni, grgrkerni = 0.;
    {
      const int idx = tab_index(q2i, G.ddq2table);
      const double2 rw = tabw[idx], rg = tabg[idx];
      const double dxx = q2i - __dmul_rn((double)idx, G.dq2table);
      wabi = rw.x + rw.y * dxx;
      grkerni = rg.x + rg.y * dxx;
      if (AUX) { const double2 r2 = __ldg(reinterpret_cast<const double2 *>(G.tab2 + idx)); grgrkerni = r2.x + r2.y * dxx; }
    }
    wabi = wabi * hfacwabi;                        // :237-241 / :549-553
    grkerni = grkerni * hfacwabi * hi1;
    const double dwdhi = -rij * grkerni * hi1 - NDIM * wabi * hi1;   // :260
    // self pair: the symmetric loop adds weight 1/2 twice (:204-208, :274, :286); the gather adds it once in full
    // (fixed particles keep rho and gradh in the first round, :273, :321: their sums are simply not used after the loop)
    {
      rho += pmassj * wabi;
      gradh += pmassj * dwdhi;
      if (AUX) {
        densn += wabi;                             // wabalt == wab: ikernelalt = ikernel
        gradhn += dwdhi;                           // :323 / :595
        grgrkerni = grgrkerni * hfacwabi * hi1 * hi1;
        const double dwdhdhi = NDIM * (NDIM + 1) * wabi * (hi1 * hi1) + 2. * (NDIM + 1) * rij * (hi1 * hi1) * grkerni +
                               (rij * rij) * (hi1 * hi1) * grgrkerni;        // :265
        gradgradh += pmassj * dwdhdhi;
        if (onef) {
          const double dfj = __ldg(A.sdf + k);
          rhodust += pmassj * dfj * wabi;
          rhogas += pmassj * (1. - dfj) * wabi;
        }
      }
    }
    if (!LIGHT) {                                  // :297-303 (j /= i: the self pair has dx = dv = 0 and adds an exact zero)
      const double dvdotr = ((vxi - vj.x) * (dx * rinve) + (vyi - vj.y) * (dy * rinve)) + (vzi - vj.z) * (dz * rinve);
      drhodt += pmassj * dvdotr * grkerni;
    }
  };
  if (cnt > 0) {
    const unsigned *col = L.nbr + ((size_t)(t >> 5) * L.lmax) * 32 + (t & 31);
    // register pipeline: the next neighbour's records load while this pair is evaluated (list read in batches, walk_list)
    if (LIGHT) {
      double4 pn = ld4(G.posm + (int)col[0]);
      walk_list(col, cnt, [&](int n, int k, int k1, int k2) {
        const double4 pc = pn;
        pn = ld4(G.posm + k1);
        body(k, pc, pc);
      });
    } else {
      double4 pn = ld4(G.posh + (int)col[0]), vn = ld4(G.vm + (int)col[0]);
      walk_list(col, cnt, [&](int n, int k, int k1, int k2) {
        const double4 pc = pn, vc = vn;
        pn = ld4(G.posh + k1); vn = ld4(G.vm + k1);
        body(k, pc, vc);
      });
    }
  }
  const int nneigh = active ? A.numneigh[orig] : 0;   // counted by build_lists_kernel (:196-197 / :532)

  if (orig >= G.nown) return;
  if (!active) {                                   // fixed particle skipped by density_partial: stays as it was
    A.redo[s] = 0;
    return;
  }
  // ---- Newton-Raphson update, src/iterate_density.f90:163-278 ----
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
      if (onef) { A.rhogas[orig] = rhogas; A.rhodust[orig] = rhodust; }
    }
    if (!converged) {
      redo = 1;
      if (A.itsdensity <= A.itsdensitymax) A.hh[orig] = hnew;         // :253-255
      if (hnew > A.hhmax) relink = true;                              // :260-262
      if (!LIGHT) A.drhodt[orig] = drhodt;
    } else if (!LIGHT) {
      const double d = drhodt * gradh_out;                            // :272-273
      A.drhodt[orig] = d;
      A.dhdt[orig] = dhdrhoi * d;
    }
    if (relink) A.flags[0] = 1;
  } else if (!LIGHT) {
    A.drhodt[orig] = drhodt;                                          // density_sums.f90:300 runs for fixed particles too
  }
  A.redo[s] = redo;
}

template <int NDIM, bool FIRST, bool AUX, bool LIGHT>
__global__ void __launch_bounds__(LIGHT ? DENS_BLOCK_LIGHT : DENS_BLOCK, 1) density_round_kernel(Grid G, DensityArgs A, NbrLists L) {
  // One persistent block per SM: the two interpolation tables (128 KB) arrive by TMA bulk copies, then every warp draws
  // 32-target units from a global counter until the work is gone.
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw);
  const double2 *tabw = reinterpret_cast<const double2 *>(smem_raw + 128);
  const double2 *tabg = reinterpret_cast<const double2 *>(smem_raw + 128 + DENS_TAB_STRIDE);
  if (threadIdx.x == 0) {
    mbar_init(mbar, 1);
    fence_mbar_init();
    mbar_expect_tx(mbar, 2 * DENS_TAB_BYTES);
    bulk_g2s(smem_raw + 128, G.tabw, DENS_TAB_BYTES, mbar);
    bulk_g2s(smem_raw + 128 + DENS_TAB_STRIDE, G.tabg, DENS_TAB_BYTES, mbar);
  }
  __syncthreads();
  mbar_wait(mbar, 0);
  const int nunits = (A.nlist + 31) >> 5;
#pragma unroll 1
  for (;;) {
    int unit = 0;
    if ((threadIdx.x & 31) == 0) unit = atomicAdd(A.sched, 1);
    unit = __shfl_sync(FULL, unit, 0);
    if (unit >= nunits) break;
    const int t = unit * 32 + (threadIdx.x & 31);
    if (t < A.nlist) density_target<NDIM, FIRST, AUX, LIGHT>(G, A, L, t, tabw, tabg);
  }
}

}  // namespace ndk
