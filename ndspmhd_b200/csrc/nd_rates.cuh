// SPMHD rates: pair kernel + per-particle finalisation.
//
// Replaces the pair loop of `get_rates` (src/ratesND_mhd.f90:304-467) with its contained procedures `rates_core`
// (:1175-1690), `artificial_dissipation` (:1700-1894), `artificial_dissipation_phantom` (:1903-1959), `mhd_terms`
// (:2377-2720, default tensor force) and `drag_forces` (:1074-1169), and the finalisation loop (:532-965).
//
// The reference visits each pair once and updates both particles; here every real particle gathers its own side.
// For the supported option tuple every pair quantity is exactly symmetric or antisymmetric under the role swap
// (SURVEY.md 7a), so writing each term "as particle i" reproduces the reference's contribution to either side;
// only the order of accumulation differs.
#pragma once
#include "nd_device.cuh"

namespace ndk {

struct RatesOpts {
  int iener, iav, imhd, idivbzero, iresist, idust, idrag_nature, ikernav;
  int iavlim0, iavlim1, iavlim2, nsubsteps_divB;
  double beta, pext, etamhd, Kdrag, stressmax, gamma;
  double alphamin, alphaumin, alphaBmin, avdecayconst, avfact, psidecayfact;
  double Bconst[3];
};

struct RatesIn {
  const double4 *bpsi;     // {Bx,By,Bz,psi}
  const double4 *thermo;   // {1/rho, max(pr - pext, 0), spsound, uu}   (1/rho correctly rounded: rho1i = 1./rhoi, ratesND_mhd.f90:325)
  const double4 *gal;      // {gradh, alpha, alphau, alphaB}
  const double *srho;      // rho by sorted slot (drag and phantom-AV branches only)
  const double4 *dusta;    // one-fluid dust: {dustfrac, deltav xyz}
  const double2 *dustb;    // one-fluid dust: {rhogas, rhodust} (smoothed sums or rho*(1-eps), rho*eps: ratesND_mhd.f90:346-352)
};

struct RatesSums {         // per sorted slot, written by the pair kernel, read by the final kernel
  double4 *F;              // {force xyz, dudt}
  double4 *dB;             // {dBevoldt xyz, divB}
  double4 *C;              // {curlB xyz, del2u}
  double4 *P;              // {gradpsi xyz, total-energy dissipation pair sum (iener=3)}
  double4 *V;              // {graddivv xyz, -}
  double4 *D;              // one-fluid dust: {ddeltavdt xyz, ddustevoldt}
};

// global reductions (order-preserving u64 keys, see dkey)
struct RatesRed {
  unsigned long long *dtcourant_min, *vsigmax_max, *dtav_min, *ts_min, *h_on_csts_max, *fhmax_max, *dtforce_min;
  double *fmean;           // [3]
  int *nclumped, *err;
  int *sched;              // work counter of the persistent warps (zeroed before each launch)
};

// max(a, b) for finite a of either sign and b >= +0 as a signed 64-bit integer compare: IEEE doubles of equal sign order like their bit
// patterns and a negative a is a negative integer (DSETP/DMNMX issue on the half-rate FP64 pipe, the integer compare does not)
__device__ __forceinline__ double fmax_nonneg(double a, double b) { return __double_as_longlong(a) > __double_as_longlong(b) ? a : b; }
#define ND_FMAX(a, b) fmax_nonneg(a, b)
// One 256-thread block an SM (2 warps a scheduler either way: 240 registers): one copy of the 64 KB table instead of two leaves the L1
// 32 KB more (hit rate 43 -> 65 %): 26.5 against 27.3 ms at 16.8 M particles.
#ifndef ND_RATES_MINB
#define ND_RATES_MINB 1
#endif
#ifndef ND_RATES_BLOCK
#define ND_RATES_BLOCK 256
#endif
constexpr int RATES_BLOCK = ND_RATES_BLOCK;
constexpr int RATES_TABG_BYTES = (IKERN + 1) * 16;                       // {grad W, slope} rows, 64016 B
constexpr int RATES_SMEM_BYTES = ((16 + RATES_TABG_BYTES + 127) / 128) * 128;

__device__ __forceinline__ double get_tstop(int idrag_nature, double rhogas, double rhodust, double Kdrag) {
  // src/dust.f90:77-102
  const double rho = rhogas + rhodust;
  if (idrag_nature == 1) return rhodust * rhogas / (Kdrag * rho);
  if (idrag_nature == 2 || idrag_nature == 4) return Kdrag;
  return 1.7976931348623157e308;
}

// FAST (1: iener=0, 2: iener=2) = the first-class option tuple compiled without run-time option tests: iav=2, ikernav=3, iresist=0,
// iavlim(1) /= 3, iavlim(3) /= 2, pext folded into thermo.  Everything else runs the generic instantiation.
// ONEF = one-fluid dust (idust=1): dust_derivs (:2726-2807) and artificial_dissipation_dust (:1969-2148) on the generic path.
template <int NDIM, bool MHD, bool DRAG, int FAST, bool ONEF>
__global__ void __launch_bounds__(RATES_BLOCK, ND_RATES_MINB) rates_pair_kernel(Grid G, RatesIn I, RatesOpts O, RatesSums S, RatesRed R, NbrLists L,
                                                                                 int s0, int ntargets, const int *targets) {
  // dynamic shared memory: [0,16) mbarrier, then the {grad W, slope} rows of the kernel table (64 KB; every pair does two
  // random lookups, which as global loads cost ~30 L1 wavefronts each and were the largest long-scoreboard stall).
  // The table arrives by one TMA bulk copy per block; blocks are persistent (grid = resident blocks) so it is loaded once
  // per SM slot, not once per 128 targets.
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw);
  const double2 *tabs = reinterpret_cast<const double2 *>(smem_raw + 16);
  if (threadIdx.x == 0) {
    mbar_init(mbar, 1);
    fence_mbar_init();
    mbar_expect_tx(mbar, RATES_TABG_BYTES);
    bulk_g2s(smem_raw + 16, G.tabg, RATES_TABG_BYTES, mbar);
  }
  __syncthreads();
  mbar_wait(mbar, 0);
  const int iav = FAST ? 2 : O.iav, iener = FAST == 2 ? 2 : (FAST == 1 ? 0 : O.iener), iresist = FAST ? 0 : O.iresist;
  // FAST excludes iavlim(1) = 3 and iavlim(3) = 2 at compile time; the remaining run-time options of the fast tuple enter the pair
  // body as 0/1 multipliers instead of (uniform) branches, which would split the scheduling block
  const int iavlim0 = FAST ? 0 : O.iavlim0, iavlim1 = O.iavlim1, iavlim2 = FAST ? 0 : O.iavlim2;
  const double sel_del2u = (O.iavlim1 > 0) ? 1. : 0., sel_gradpsi = (O.idivbzero >= 2) ? 1. : 0.;
  const double zero = 1.e-10;
  const double eps = 2.220446049250313e-16;
  // block-wide reductions are carried across the chunks of a persistent block and flushed once
  double dtc_den = 0., dtav_den = 0., vsigmax = 0., ts_min = 1.7976931348623157e308, h_on_csts_max = 0.;
  int nclumped = 0;
  // Work unit = the 32 targets of one list column block.  Persistent warps draw units from a global counter (a static
  // round-robin of 128-target chunks over blocks measured 8.9 ms against 6.5 ms for one block per chunk: unit costs differ
  // widely -- ghost-only units are free -- and a block waits for its slowest warp); the non-persistent launch maps
  // unit = global warp index.
  const int nunits = (ntargets + 31) >> 5;
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int trip = 0;; trip++) {
    int unit = 0;
    if (lane == 0) unit = atomicAdd(R.sched, 1);
    unit = __shfl_sync(FULL, unit, 0);
    if (unit >= nunits) break;
  const int tix = unit * 32 + lane;   // target index within this launch
  int s = s0 + tix;                                  // target slot: a contiguous range, or the entries of a target list
  int orig = -1, ti = 0, cnt = 0;
  bool active = false;
  if (tix < ntargets) {
    if (targets) s = targets[tix];
    orig = G.perm[s];
    active = orig < G.nown;       // rates are gathered for the caller's own rows (fixed particles included: they feed the dt minima)
  }
  double xi = 0, yi = 0, zi = 0, hi1 = 1, vxi = 0, vyi = 0, vzi = 0, pmassi = 0;
  double Bxi = 0, Byi = 0, Bzi = 0, psii = 0, rho1i = 1, rhoi = 1, pri = 0, spsoundi = 0, uui = 0, gradhi = 0, alphai = 0, alphaui = 0, alphaBi = 0;
  if (active) {
    double4 p = ld4(G.posh + s), v = ld4(G.vm + s), t = ld4(I.thermo + s), g = ld4(I.gal + s);
    xi = p.x; yi = p.y; zi = p.z; hi1 = p.w;                                   // hi1 = 1./hi, :215, :389
    vxi = v.x; vyi = v.y; vzi = v.z; pmassi = v.w;
    rho1i = t.x; pri = t.y; spsoundi = t.z; uui = t.w;                          // :325-328
    gradhi = g.x; alphai = g.y; alphaui = g.z; alphaBi = g.w;
    if (MHD) { double4 b = ld4(I.bpsi + s); Bxi = b.x; Byi = b.y; Bzi = b.z; psii = b.w; }
    ti = G.typ[s];
    rhoi = I.srho[s];
    cnt = L.cnt[tix];
  }
  double dustfraci = 0., rhogasi = rhoi, rhodusti = 0., dvix = 0., dviy = 0., dviz = 0., deltav2i = 0., rhogrhodonrhoi = 0.;
  if (ONEF && active) {                                                        // :344-356
    const double4 da = ld4(I.dusta + s);
    const double2 db = __ldg(I.dustb + s);
    dustfraci = da.x; dvix = da.y; dviy = da.z; dviz = da.w;
    rhogasi = db.x; rhodusti = db.y;
    deltav2i = (dvix * dvix + dviy * dviy) + dviz * dviz;
    rhogrhodonrhoi = rhogasi * rhodusti * rho1i;
  }
  double fgx = 0, fgy = 0, fgz = 0, ddvx = 0, ddvy = 0, ddvz = 0, ddust = 0;      // gas-only force sum, ddeltavdt, ddustevoldt
  const double rho21i = rho1i * rho1i;                                         // :326
  const double Prho2i = pri * rho21i;                                          // :332
  const double hi21 = __dmul_rn(hi1, hi1);
  const double hfacwabi = powndim<NDIM>(hi1), hfacgrkerni = hfacwabi * hi1;    // :390-391
  double Brhoxi = 0, Brhoyi = 0, Brhozi = 0, Brho2i = 0, valfven2i = 0, vsig2i = 0;
  if (MHD) {                                                                   // :362-373
    Brhoxi = Bxi * rho1i; Brhoyi = Byi * rho1i; Brhozi = Bzi * rho1i;
    const double B2i = (Bxi * Bxi + Byi * Byi) + Bzi * Bzi;
    Brho2i = B2i * rho21i;
    valfven2i = B2i * rho1i;
    vsig2i = spsoundi * spsoundi + valfven2i;
  }
  // accumulators
  double fx = 0, fy = 0, fz = 0, dudt = 0, dBx = 0, dBy = 0, dBz = 0, divB = 0, cBx = 0, cBy = 0, cBz = 0, del2u = 0;
  double gpx = 0, gpy = 0, gpz = 0, gvx = 0, gvy = 0, gvz = 0, endiss = 0;
  double drho = 0;   // FAST: drho/dt of the density sums (density_sums.f90:297-303 with dr = dx/(rij + epsilon)), times gradh through grkerni
  // dtcourant = min over pairs of min(hi,hj)/vsigdtc = 1/max(max(1/hi,1/hj)*vsigdtc): track the denominator, divide once

  bool any_coincident = false, vsig_det_bad = false;
  // ---- pair terms over the neighbour list (build_lists_kernel<LIST_RATES> applied src/ratesND_mhd.f90:401-415) ----
  // KIND (compile time): 0 = the type rule decides per pair; 1 = a pair of the list's front part (types interact: rates_core);
  // 2 = a pair of its back part (gas-dust: drag_forces).  The builder splits the lists of drag runs by that rule, so a warp's lanes no
  // longer take both branches of every trip (gas and dust rows alternate in a cell).
  auto body = [&](auto kind_c, int k, const double4 &pj, const double4 &vj, const double4 &tj4, const double4 &gj, const double4 &bj) {
    constexpr int KIND = decltype(kind_c)::value;
    const int tj = (DRAG && KIND != 1) ? __ldg(G.typ + k) : 0;   // only the drag branch needs the neighbour's type
    const double dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
    const double rij2 = dist2_exact(dx, dy, dz);
    const double hj1 = pj.w, hj21 = __dmul_rn(hj1, hj1);
    const double q2i = __dmul_rn(rij2, hi21), q2j = __dmul_rn(rij2, hj21);
    // rij = sqrt(rij2), dr = dx/rij (:416, :429); coincident particles (rij <= epsilon, :417-427) get dr = 0 through rinv = 0
    const double rinv = rij2 > 4.930380657631324e-32 ? rsqrt_nr(rij2) : 0.;
    const double rij = rij2 * rinv;
    const double drx = dx * rinv, dry = dy * rinv, drz = dz * rinv;
    // coincident pairs (:417-427) are only flagged here -- no branch in the pair body, which would end ptxas's scheduling
    // block; their bookkeeping is redone from the list after the loop (rare)
    any_coincident |= (rinv == 0.);
    const double pmassj = vj.w;
    const double dvx = vxi - vj.x, dvy = vyi - vj.y, dvz = vzi - vj.z;
    const double h1max = ND_FMAX(hi1, hj1);
    if (!DRAG || KIND == 1 || (KIND == 0 && types_interact(ti, tj))) {
      // =============================== rates_core ===============================
      // kernel gradient table rows for q2i, q2j: the loads are issued here, the interpolation (their first use) comes after the
      // signal-velocity block so that ~250 independent FP64 instructions cover the lookup latency
      const int idxi = tab_index(q2i, G.ddq2table), idxj = tab_index(q2j, G.ddq2table);
      const double2 rowi = tabs[idxi], rowj = tabs[idxj];
      const double dvdotr = (dvx * drx + dvy * dry) + dvz * drz;   // :1250
      const double rho1j = tj4.x, rho21j = rho1j * rho1j;          // :1256-1258
      const double rhoav1 = 0.5 * (rho1i + rho1j);                 // :1261
      double dustfracj = 0., rhogasj = 0., rhodustj = 0., dvjx = 0., dvjy = 0., dvjz = 0., deltav2j = 0., rhogrhodonrhoj = 0.;
      double projdvgas = dvdotr, projdeltavi = 0., projdeltavj = 0.;
      if (ONEF) {                                                  // :1262-1280
        const double4 da = ld4(I.dusta + k);
        const double2 db = __ldg(I.dustb + k);
        dustfracj = da.x; dvjx = da.y; dvjy = da.z; dvjz = da.w;
        rhogasj = db.x; rhodustj = db.y;
        rhogrhodonrhoj = rhogasj * rhodustj * rho1j;
        deltav2j = (dvjx * dvjx + dvjy * dvjy) + dvjz * dvjz;
        const double gx = (vxi - dustfraci * dvix) - (vj.x - dustfracj * dvjx), gy = (vyi - dustfraci * dviy) - (vj.y - dustfracj * dvjy),
                     gz = (vzi - dustfraci * dviz) - (vj.z - dustfracj * dvjz);
        projdvgas = (gx * drx + gy * dry) + gz * drz;
        projdeltavi = (dvix * drx + dviy * dry) + dviz * drz;
        projdeltavj = (dvjx * drx + dvjy * dry) + dvjz * drz;
      }
      const double prj = tj4.y;                                    // :1285 (pext already subtracted)
      const double Prho2j = prj * rho21j;
      const double spsoundj = tj4.z, uuj = tj4.w;
      double Bxj = 0, Byj = 0, Bzj = 0, psij = 0, dBxx = 0, dByy = 0, dBzz = 0;
      double projBi = 0, projBj = 0, projdB = 0, projBrhoi = 0, projBrhoj = 0, Brho2j = 0, valfven2j = 0;
      double Brhoxj = 0, Brhoyj = 0, Brhozj = 0;
      if (MHD) {                                                // :1295-1313
        Bxj = bj.x; Byj = bj.y; Bzj = bj.z; psij = bj.w;
        Brhoxj = Bxj * rho1j; Brhoyj = Byj * rho1j; Brhozj = Bzj * rho1j;
        dBxx = Bxi - Bxj; dByy = Byi - Byj; dBzz = Bzi - Bzj;
        projBi = (Bxi * drx + Byi * dry) + Bzi * drz;
        projBj = (Bxj * drx + Byj * dry) + Bzj * drz;
        projdB = (dBxx * drx + dByy * dry) + dBzz * drz;
        projBrhoi = projBi * rho1i;                             // dot_product(Brhoi,dr)
        projBrhoj = projBj * rho1j;
        const double B2j = (Bxj * Bxj + Byj * Byj) + Bzj * Bzj;
        valfven2j = B2j * rho1j;
        Brho2j = B2j * rho21j;
      }
      // ---- signal velocities :1417-1465 ----  (the independent square roots advance in lockstep, see sqrt_n)
      double vsigi, vsigj, vsigB, vsigu;
      const double pdiff = fabs(pri - prj) * rhoav1;                               // :1459 (pequil = 0)
      if (MHD) {
        const double vsig2j = spsoundj * spsoundj + valfven2j;
        const double vsigproji = vsig2i * vsig2i - 4. * ((spsoundi * projBi) * (spsoundi * projBi)) * rho1i;
        const double vsigprojj = vsig2j * vsig2j - 4. * ((spsoundj * projBj) * (spsoundj * projBj)) * rho1j;
        vsig_det_bad |= (vsigproji < 0. || vsigprojj < 0.);   // ND_ERR_VSIG_DET, raised after the loop
        const double a4[4] = {vsigproji, vsigprojj, (dvx * dvx + dvy * dvy) + dvz * dvz /* norm2(dvel), :1433 */, pdiff};
        double r4[4];
        sqrt_n<4>(a4, r4);
        const double a2[2] = {0.5 * (vsig2i + r4[0]), 0.5 * (vsig2j + r4[1])};
        double r2[2];
        sqrt_n<2>(a2, r2);
        vsigi = r2[0]; vsigj = r2[1]; vsigu = r4[3];
        if (iavlim2 != 2) vsigB = r4[2];
        else vsigB = 0.5 * (vsigi + vsigj) + fabs(dvdotr);
      } else {
        vsigi = spsoundi; vsigj = spsoundj; vsigB = 0.;
        vsigu = sqrt_nr(pdiff);
      }
      double vsig = 0.5 * (ND_FMAX(vsigi + vsigj - O.beta * dvdotr, 0.0));          // :1452
      double vsigdtc = ND_FMAX(vsig, ND_FMAX(0.5 * (vsigi + vsigj + O.beta * fabs(dvdotr)), vsigB));   // :1465
      if (ONEF) vsigdtc = vsigdtc + sqrt(deltav2i + deltav2j);                                   // :1466-1468
      {                                                                         // :1472-1481 as selects
        const bool dust = (ti == T_DUST);
        vsig = dust ? 0. : vsig; vsigu = dust ? 0. : vsigu;
        vsigmax = ND_FMAX(vsigmax, dust ? 0. : vsigdtc);
        dtc_den = ND_FMAX(dtc_den, (!dust && vsigdtc > zero) ? h1max * vsigdtc : 0.);
      }
      // ---- kernel gradients :1208-1241 (w = w[index] + dwdx*(q2 - index*dq2table), src/kernelND.f90:4443-4455) ----
      double grkerni = rowi.x + rowi.y * (q2i - __dmul_rn((double)idxi, G.dq2table));
      double grkernj = rowj.x + rowj.y * (q2j - __dmul_rn((double)idxj, G.dq2table));
      // :1215-1237 with ikernav = 3 (the only supported value, check_options): h^-(ndim+1) * gradh as one factor per side -- the
      // target's is loop-invariant, the neighbour's reuses hj21 = (1/h_j)^2
      const double hgj = (NDIM == 3 ? hj21 * hj21 : NDIM == 2 ? hj21 * hj1 : hj21) * gj.x;
      grkerni = grkerni * (hfacgrkerni * gradhi);
      grkernj = grkernj * hgj;
      const double grkern = 0.5 * (grkerni + grkernj);
      double fix = 0, fiy = 0, fiz = 0;   // forcei contribution of this pair
      // FUSEDR (every instantiation but one-fluid dust, which needs forcei on its own): the scalar coefficients of dr in the force
      // (AV, pressure, isotropic magnetic part) are summed before they meet dr -- forcei = pmassj * ((fix, fiy, fiz) - cdr * dr)
      constexpr bool FUSEDR = !ONEF;
      double cdr = 0.;
      double vsigav = 0.;
      if (ONEF && iav > 0) {
        // =============================== artificial_dissipation_dust (iav = 1, 2, 3) ===============================
        const double alphaav = 0.5 * (alphai + gj.y), alphau = 0.5 * (alphaui + gj.z), alphaB = 0.5 * (alphaBi + gj.w);
        vsigav = ND_FMAX(alphaav, alphau) * vsig;                   // :1984
        const double dustfracav = 0.5 * (dustfraci + dustfracj);
        const double projdvgasav = dvdotr - dustfracav * (projdeltavi - projdeltavj);
        const double ddx_ = dvix - dvjx, ddy_ = dviy - dvjy, ddz_ = dviz - dvjz;
        const double projddeltav = (ddx_ * drx + ddy_ * dry) + ddz_ * drz;
        const double dpmomdotr = (iav == 3) ? projdvgasav : (iav == 2) ? projdvgas : dvdotr;   // :1993-1999
        const double term = vsig * rhoav1 * grkern;
        double termv = term;
        const bool allpairs = (iav == 1);
        if (iav == 2 || iav == 3) termv = termv * (1. - dustfracav);
        const double termu = vsigu * rhoav1 * grkern * (1. - dustfracav);
        const bool on = (projdvgas < 0.) || allpairs;
        if (on) {                                                // :2023-2038
          const double visc = alphaav * termv * dpmomdotr;
          const double c = pmassj * visc;
          if (iav == 1 || iav == 3) { fx += c * drx; fy += c * dry; fz += c * drz; }   // fextra: acts on the whole fluid
          else { fix += c * drx; fiy += c * dry; fiz += c * drz; }
        }
        if (iener > 0) {                                         // :2049-2146
          double vissv = 0., vissdv = 0., termdv = 0.;
          if (on) vissv = (iav == 3) ? -alphaav * 0.5 * (projdvgasav * projdvgasav) : (iav == 2) ? -alphaav * 0.5 * (projdvgas * projdvgas) : -alphaav * 0.5 * (dvdotr * dvdotr);
          const double vissu = (iener == 1) ? 0. : alphau * (uui - uuj);
          const double faci = 1. / (dustfraci * (1. - dustfraci));   // :2078-2079
          if (iav == 1 || iav == 3) {
            if (iav == 1) { vissdv = -0.5 * ((ddx_ * ddx_ + ddy_ * ddy_) + ddz_ * ddz_); termdv = alphaav * vsig * rhoav1 * dustfracav * grkern; }
            else { vissdv = 0.; termdv = alphaav * vsig * rhoav1 * dustfracav * grkern * (1. - dustfracav); }
            if (iav == 3) { const double c = faci * pmassj * termdv * (-projdvgasav); ddvx += c * drx; ddvy += c * dry; ddvz += c * drz; }
            else { const double c = faci * pmassj * termdv; ddvx += c * ddx_; ddvy += c * ddy_; ddvz += c * ddz_; }
          } else if (iav == 2 && projddeltav < 0.) {             // :2111-2125
            const double vsigdv = 0.5 * (spsoundi + spsoundj);
            termdv = alphaav * vsigdv * rhoav1 * grkern * dustfracav * (1. - dustfracav);
            const double c = faci * pmassj * termdv * projddeltav;
            ddvx += c * drx; ddvy += c * dry; ddvz += c * drz;
            vissdv = -0.5 * (projddeltav * projddeltav);
          }
          dudt += (rhoi / rhogasi) * pmassj * (termv * vissv + termu * vissu + termdv * vissdv);   // :2133-2138
          const double vsigeps = 0.5 * (spsoundi + spsoundj);    // :2140-2145
          ddust += pmassj * (alphaB * rhoav1 * vsigeps * (dustfraci - dustfracj) * grkern);
        }
      } else if (iav > 0 && iav != 3) {
        // =============================== artificial_dissipation ===============================
        const double alphaav = 0.5 * (alphai + gj.y), alphau = 0.5 * (alphaui + gj.z), alphaB = 0.5 * (alphaBi + gj.w);   // :1712-1714
        vsigav = ND_FMAX(alphaav, ND_FMAX(alphau, alphaB)) * vsig;
        const double rg = rhoav1 * grkern;
        const double term = vsig * rg;                           // :1723
        const double termu = vsigu * rg;                         // :1727
        const double termB = vsigB * rg;                         // :1732
        const double approaching = dvdotr < 0. ? 1. : 0.;        // :1745-1748 as a select: no divergent branch in the pair body
        {
          const double visc = alphaav * term * (-dvdotr);
          if (FUSEDR) cdr += visc * approaching;
          else {
          const double c = pmassj * visc * approaching;
          fix -= c * drx; fiy -= c * dry; fiz -= c * drz;
          }
        }
        if (MHD) {                                               // :1762-1774
          double bvx, bvy, bvz;
          if (iav >= 2) { bvx = dBxx * rhoav1; bvy = dByy * rhoav1; bvz = dBzz * rhoav1; }
          else { bvx = (dBxx - drx * projdB) * rhoav1; bvy = (dByy - dry * projdB) * rhoav1; bvz = (dBzz - drz * projdB) * rhoav1; }
          const double c = rhoi * pmassj, ab = alphaB * termB;   // dBevoldti + rhoi*pmassj*dBdtvisc, :1773
          dBx += c * (ab * bvx); dBy += c * (ab * bvy); dBz += c * (ab * bvz);
        }
        if (iener == 3) {                                        // :1792-1830 total energy: pair part of dendt
          double qdiff = 0.;
          const double projvi = (vxi * drx + vyi * dry) + vzi * drz, projvj = (vj.x * drx + vj.y * dry) + vj.z * drz;
          qdiff += approaching * (term * alphaav * 0.5 * (projvi * projvi - projvj * projvj));
          qdiff += alphau * termu * (uui - uuj);
          if (MHD) {
            double B2i_, B2j_;
            if (iav >= 2) { B2i_ = (Bxi * Bxi + Byi * Byi) + Bzi * Bzi; B2j_ = (Bxj * Bxj + Byj * Byj) + Bzj * Bzj; }
            else { B2i_ = ((Bxi * Bxi + Byi * Byi) + Bzi * Bzi) - projBi * projBi; B2j_ = ((Bxj * Bxj + Byj * Byj) + Bzj * Bzj) - projBj * projBj; }
            qdiff += alphaB * termB * 0.5 * (B2i_ - B2j_) * rhoav1;
          }
          endiss += pmassj * qdiff;                           // :1829
        } else if (iener > 0) {                                  // :1835-1875 thermal energy
          double vissB = 0.;
          const double tv = ((vxi * drx + vyi * dry) + vzi * drz) - ((vj.x * drx + vj.y * dry) + vj.z * drz);
          const double vissv = -alphaav * 0.5 * (tv * tv) * approaching;
          const double vissu = alphau * (uui - uuj);
          if (MHD) {
            const double dB2 = (dBxx * dBxx + dByy * dByy) + dBzz * dBzz;
            if (iav >= 2) vissB = -alphaB * 0.5 * dB2 * rhoav1;
            else vissB = -alphaB * 0.5 * (dB2 - projdB * projdB) * rhoav1;
          }
          dudt += pmassj * (term * vissv + termu * vissu + termB * vissB);
        }
      } else if (iav == 3) {
        // =============================== artificial_dissipation_phantom ===============================
        const double rhoj = __ldg(I.srho + k);
        double dudti = 0.;
        if (dvdotr < 0.) {
          const double vsi = ND_FMAX(alphai * spsoundi - O.beta * dvdotr, 0.);
          const double vsj = ND_FMAX(gj.y * spsoundj - O.beta * dvdotr, 0.);
          const double qi = -0.5 * rhoi * vsi * dvdotr, qj = -0.5 * rhoj * vsj * dvdotr;
          const double visc = (qi * rho21i * grkerni + qj * rho21j * grkernj);
          if (FUSEDR) cdr += visc;
          else {
          const double c = pmassj * visc;
          fix -= c * drx; fiy -= c * dry; fiz -= c * drz;
          }
          dudti = qi * rho21i * pmassj * dvdotr * grkerni;
        }
        const double du = uui - uuj;
        const double cfaci = 0.5 * alphaui * rhoi * vsigu * du, cfacj = 0.5 * gj.z * rhoj * vsigu * du;
        const double diffu = cfaci * grkerni * (rho1i * rho1i) + cfacj * grkernj * (rho1j * rho1j);
        dudt += dudti + pmassj * diffu;
      }
      dtav_den = ND_FMAX(dtav_den, vsigav > zero ? h1max * vsigav : 0.);              // :1500
      if (FAST && ND_DENS_LIGHT) drho += pmassj * (dvdotr * (1. - eps * rinv)) * grkerni;   // sum m_j (dv.dr) grad W_i: the density loop's drhodt (LIGHT rounds skip it)
      {                                                          // pressure, :1538-1567 (phi = 1, sqrtg = 1)
        const double prterm = Prho2i * grkerni + Prho2j * grkernj;
        if (FUSEDR) cdr += prterm;
        else {
        const double c = pmassj * prterm;
        fix -= c * drx; fiy -= c * dry; fiz -= c * drz;
        }
      }
      if (MHD) {
        // =============================== mhd_terms ===============================
        const double fiso = 0.5 * (Brho2i * grkerni + Brho2j * grkernj);        // :2526
        const double sm = O.stressmax;
        // faniso = (Brho_i (Brho_i.dr) - stressmax dr/rho_i^2) grkern_i + (same for j), :2534-2537; the force is faniso - fiso dr
        const double ai = projBrhoi * grkerni, aj = projBrhoj * grkernj;
        const double sdr = sm * (rho21i * grkerni + rho21j * grkernj) + fiso;
        if (FUSEDR) { cdr += sdr; fix = Brhoxi * ai + Brhoxj * aj; fiy = Brhoyi * ai + Brhoyj * aj; fiz = Brhozi * ai + Brhozj * aj; }
        else {
        fix += pmassj * ((Brhoxi * ai + Brhoxj * aj) - sdr * drx);               // :2541, :2629
        fiy += pmassj * ((Brhoyi * ai + Brhoyj * aj) - sdr * dry);
        fiz += pmassj * ((Brhozi * ai + Brhozj * aj) - sdr * drz);
        }
        const double mg = pmassj * grkern;
        divB -= mg * projdB;                                     // :2552
        cBx += (dByy * drz - dBzz * dry) * mg;                   // :2601-2602 curlB += pmassj*(dB x dr)*grkern
        cBy += (dBzz * drx - dBxx * drz) * mg;
        cBz += (dBxx * dry - dByy * drx) * mg;
        const double ci = pmassj * projBrhoi * grkerni;          // :2664-2665 induction (imhd = 1, 11)
        dBx -= dvx * ci; dBy -= dvy * ci; dBz -= dvz * ci;
        if (iresist == 1) {                                      // :2685-2709
          const double etaij = O.etamhd;                         // 0.5*(etai + etaj) with constant eta
          const double f = -2. * etaij / (rij + eps);
          const double c = rhoi * pmassj * 0.5 * ((rho1i * rho1i) * grkerni + (rho1j * rho1j) * grkernj);
          dBx -= c * (f * dBxx); dBy -= c * (f * dByy); dBz -= c * (f * dBzz);
          if (iener > 0) dudt += pmassj * (-etaij * rho1i * rho1j * ((dBxx * dBxx + dByy * dByy) + dBzz * dBzz) * grkern / rij);
        }
        if (FAST || O.idivbzero >= 2) {                          // :2712-2716
          const double gradpsiterm = psii * rho21i * grkerni + psij * rho21j * grkernj;
          const double c = (FAST ? sel_gradpsi : 1.) * (pmassj * gradpsiterm);
          gpx -= c * drx; gpy -= c * dry; gpz -= c * drz;
        }
      }
      if (ONEF) {
        // =============================== dust_derivs (idustevol = 0) ===============================
        ddust -= pmassj * (rhogrhodonrhoi * projdeltavi * rho21i * grkerni + rhogrhodonrhoj * projdeltavj * rho21j * grkernj);   // :2753-2758
        const double dterm = 0.5 * ((rhogasi - rhodusti) * rho1i * deltav2i - (rhogasj - rhodustj) * rho1j * deltav2j);          // :2767-2769
        const double c = rho1i * pmassj * grkerni;               // :2776 (the -rho/rhogas*forcei term follows the pair loop, :460)
        ddvx += c * (dvx * projdeltavi + dterm * drx); ddvy += c * (dvy * projdeltavi + dterm * dry); ddvz += c * (dvz * projdeltavi + dterm * drz);
        const double pi_ = rhogrhodonrhoi * projdeltavi * rho21i * grkerni, pj_ = rhogrhodonrhoj * projdeltavj * rho21j * grkernj;   // :2792-2795
        fx -= pmassj * (pi_ * dvix + pj_ * dvjx); fy -= pmassj * (pi_ * dviy + pj_ * dvjy); fz -= pmassj * (pi_ * dviz + pj_ * dvjz);
        if (iener > 0) dudt += pmassj * (pri * rho1i / rhogasi * projdvgas - rhodusti * rho21i * (uui - uuj) * projdeltavi) * grkerni;   // :2801-2803
        fgx += fix; fgy += fiy; fgz += fiz;
      }
      if (FUSEDR) {
        const double cm = pmassj * cdr;
        if (MHD) { fx = fma(pmassj, fix, fx); fy = fma(pmassj, fiy, fy); fz = fma(pmassj, fiz, fz); }
        fx = fma(-cm, drx, fx); fy = fma(-cm, dry, fy); fz = fma(-cm, drz, fz);
      } else {
      fx += fix; fy += fiy; fz += fiz;
      }
      if (iav > 0) {                                             // :1639-1656 switch sources
        if (FAST) del2u += sel_del2u * (pmassj * rho1j * ((uui - uuj) * rinv) * grkerni);
        else if (iavlim1 > 0) del2u += pmassj * rho1j * ((uui - uuj) * rinv) * grkerni;
        if (iavlim0 == 3) {
          const double c = pmassj * rho1j * rinv * dvdotr * grkerni;
          gvx += c * drx; gvy += c * dry; gvz += c * drz;
        } else if (!FAST) {                                      // graddivv is only read back by the iavlim(1)=3 switch
          const double c = pmassj * grkerni;
          gvx += c * (dvx - dvdotr); gvy += c * (dvy - dvdotr); gvz += c * (dvz - dvdotr);
        }
      }
    } else if (DRAG && KIND != 1) {
      // =============================== drag_forces ===============================
      const double rhoj = __ldg(I.srho + k);
      const double dv2 = (dvx * dvx + dvy * dvy) + dvz * dvz;
      double ddx = drx, ddy = dry, ddz = drz;
      bool skip = false;
      if (rij <= eps) {                                          // :1091-1098
        if (dv2 <= eps) skip = true;
        else { const double v1 = 1. / sqrt(dv2); ddx = dvx * v1; ddy = dvy * v1; ddz = dvz * v1; }
      }
      const bool igas = (ti == T_GAS || ti == T_BND), jgas = (tj == T_GAS || tj == T_BND);
      if (!skip && (igas || jgas)) {                             // :1133-1144: kernel and sound speed of the gas particle
        const double projv = (dvx * ddx + dvy * ddy) + dvz * ddz;
        double wab, spsoundgas, ts, hgas1;
        if (igas) { wab = interp_drag(G, q2i) * hfacwabi; spsoundgas = spsoundi; ts = get_tstop(O.idrag_nature, rhoi, rhoj, O.Kdrag); hgas1 = hi1; }
        else { wab = interp_drag(G, q2j) * powndim<NDIM>(hj1); spsoundgas = tj4.z; ts = get_tstop(O.idrag_nature, rhoj, rhoi, O.Kdrag); hgas1 = hj1; }
        h_on_csts_max = fmax(h_on_csts_max, 1. / (hgas1 * (spsoundgas * ts)));
        ts_min = fmin(ts_min, ts);
        const double dragterm = NDIM * wab / ((rhoi + rhoj) * ts) * projv;   // :1156 (projvstar = projv)
        const double c = dragterm * pmassj;
        fx -= c * ddx; fy -= c * ddy; fz -= c * ddz;
        if (ti == T_GAS) dudt += pmassj * (dragterm * projv);   // :1162-1164
      }
    }
  };

  // drag runs: the builder packs two counts -- low half = front part (pairs whose types interact), high half = back part (gas-dust pairs,
  // stored from the end of the column downwards)
  int cnt_drag = 0;
  if (DRAG && L.split) { cnt_drag = cnt >> 16; cnt &= 0xffff; }
  {
    const unsigned *col0 = L.nbr + ((size_t)(tix >> 5) * L.lmax) * 32 + (tix & 31);
    const double4 zero4 = make_double4(0., 0., 0., 0.);
    // walks `count` entries of the column from `col` with stride `step` lines (+1 front part, -1 back part)
    auto run = [&](auto kind_c, const unsigned *col, int count, long long step) {
      if (count <= 0) return;
      if (ONEF) {   // the one-fluid dust instantiations have no registers to spare: direct loads
        walk_list(col, count, [&](int n, int k, int k1, int k2) { body(kind_c, k, ld4(G.posh + k), ld4(G.vm + k), ld4(I.thermo + k), ld4(I.gal + k), MHD ? ld4(I.bpsi + k) : zero4); });
      } else {
        // Register software pipeline: all five records of the next neighbour are in flight while this pair is evaluated.  The list
        // column is read two entries ahead with plain rotation (k <- k1 <- k2 <- load): one coalesced load per pair and no branch in
        // the loop besides its own.
        const int last = count - 1;
        int k = (int)__ldcs(col), k1 = (int)__ldcs(col + (long long)min(1, last) * 32 * step);
        double4 pn = ld4(G.posh + k), vn = ld4(G.vm + k), tn = ld4(I.thermo + k), gn = ld4(I.gal + k), bn = MHD ? ld4(I.bpsi + k) : zero4;
#pragma unroll 1
        for (int n = 0; n < count; n++) {
          const int k2 = (int)__ldcs(col + (long long)min(n + 2, last) * 32 * step);
          const double4 pc = pn, vc = vn, tc = tn, gc = gn, bc = bn;
          pn = ld4(G.posh + k1); vn = ld4(G.vm + k1); tn = ld4(I.thermo + k1); gn = ld4(I.gal + k1); bn = MHD ? ld4(I.bpsi + k1) : zero4;
          body(kind_c, k, pc, vc, tc, gc, bc);
          k = k1; k1 = k2;
        }
      }
    };
    if (DRAG && L.split) {
      run(std::integral_constant<int, 1>{}, col0, cnt, 1);
      run(std::integral_constant<int, 2>{}, col0 + (size_t)(L.lmax - 1) * 32, cnt_drag, -1);
    } else {
      run(std::integral_constant<int, 0>{}, col0, cnt, 1);
    }
  }

  if (vsig_det_bad) atomicCAS(R.err, 0, 6 /*ND_ERR_VSIG_DET*/);
  if (any_coincident) {                                          // bookkeeping of :417-427 for the flagged targets
    const unsigned *col = L.nbr + ((size_t)(tix >> 5) * L.lmax) * 32 + (tix & 31);
    for (int n = 0; n < cnt; n++) {
      const int k = (int)col[(size_t)n * 32];
      const double4 pj = ld4(G.posh + k);
      const double rij2 = dist2_exact(xi - pj.x, yi - pj.y, zi - pj.z);
      if (!(rij2 > 4.930380657631324e-32) && __ldg(G.typ + k) == ti) {
        const int origj = G.perm[k];
        if (origj >= G.nown || orig > origj) nclumped++;
        if (rij2 == 0. && ti != 2) atomicCAS(R.err, 0, 1 /*ND_ERR_INVALID_ARG: dx = 0*/);
      }
    }
  }
  if (active) {
    st4(S.F + s, make_double4(fx, fy, fz, dudt));
    st4(S.dB + s, make_double4(dBx, dBy, dBz, divB));
    st4(S.C + s, make_double4(cBx, cBy, cBz, del2u));
    st4(S.P + s, make_double4(gpx, gpy, gpz, endiss));
    st4(S.V + s, make_double4(gvx, gvy, gvz, drho));
    if (ONEF) {                                                  // ddeltavdt(:,i) - rhoi/rhogasi*forcei(:), :460
      const double c = rhoi / rhogasi;
      st4(S.D + s, make_double4(ddvx - c * fgx, ddvy - c * fgy, ddvz - c * fgz, ddust));
    }
  }
  }   // persistent chunk loop
  // block-free warp reductions into global min/max keys
  double dtcourant = dtc_den > 0. ? fmin(1.e6, 1. / dtc_den) : 1.e6;            // initial value 1.e6, :251
  double dtav = dtav_den > 0. ? 1. / dtav_den : 1.7976931348623157e308;
  dtcourant = warp_min(dtcourant); vsigmax = warp_max(vsigmax); dtav = warp_min(dtav);
  int ncl = nclumped;
  for (int o = 16; o; o >>= 1) ncl += __shfl_xor_sync(FULL, ncl, o);
  if (DRAG) { ts_min = warp_min(ts_min); h_on_csts_max = warp_max(h_on_csts_max); }
  if ((threadIdx.x & 31) == 0) {
    atomic_min_d(R.dtcourant_min, dtcourant);
    atomic_max_d(R.vsigmax_max, vsigmax);
    atomic_min_d(R.dtav_min, dtav);
    if (DRAG) { atomic_min_d(R.ts_min, ts_min); atomic_max_d(R.h_on_csts_max, h_on_csts_max); }
    if (ncl) atomicAdd(R.nclumped, ncl);
  }
}

}  // namespace ndk
