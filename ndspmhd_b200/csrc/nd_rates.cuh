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
  const double4 *thermo;   // {rho, pr, spsound, uu}
  const double4 *gal;      // {gradh, alpha, alphau, alphaB}
};

struct RatesSums {         // per sorted slot, written by the pair kernel, read by the final kernel
  double4 *F;              // {force xyz, dudt}
  double4 *dB;             // {dBevoldt xyz, divB}
  double4 *C;              // {curlB xyz, del2u}
  double4 *P;              // {gradpsi xyz, total-energy dissipation pair sum (iener=3)}
  double4 *V;              // {graddivv xyz, -}
};

// global reductions (order-preserving u64 keys, see dkey)
struct RatesRed {
  unsigned long long *dtcourant_min, *vsigmax_max, *dtav_min, *ts_min, *h_on_csts_max, *fhmax_max, *dtforce_min;
  double *fmean;           // [3]
  int *nclumped, *err;
};

constexpr int RATES_BLOCK = 128;
constexpr int RATES_CAP = 96;

__device__ __forceinline__ double get_tstop(int idrag_nature, double rhogas, double rhodust, double Kdrag) {
  // src/dust.f90:77-102
  const double rho = rhogas + rhodust;
  if (idrag_nature == 1) return rhodust * rhogas / (Kdrag * rho);
  if (idrag_nature == 2 || idrag_nature == 4) return Kdrag;
  return 1.7976931348623157e308;
}

template <int NDIM, bool MHD, bool DRAG>
__global__ void __launch_bounds__(RATES_BLOCK, 2) rates_pair_kernel(Grid G, RatesIn I, RatesOpts O, RatesSums S, RatesRed R, int *pair_out_i,
                                                                      int *pair_out_j, unsigned long long *pair_count, long long pair_cap) {
  extern __shared__ unsigned nlist_smem[];
  const int s = blockIdx.x * RATES_BLOCK + threadIdx.x;
  int orig = -1, ti = 0, celli = 0;
  bool active = false;
  if (s < G.ntotal) {
    orig = G.perm[s];
    active = orig < G.npart;      // rates are gathered for rows 1..npart (fixed particles included: they feed the dt minima)
  }
  double xi = 0, yi = 0, zi = 0, hi = 1, vxi = 0, vyi = 0, vzi = 0, pmassi = 0;
  double Bxi = 0, Byi = 0, Bzi = 0, psii = 0, rhoi = 1, pri = 0, spsoundi = 0, uui = 0, gradhi = 0, alphai = 0, alphaui = 0, alphaBi = 0;
  if (active) {
    double4 p = ld4(G.posh + s), v = ld4(G.vm + s), t = ld4(I.thermo + s), g = ld4(I.gal + s);
    xi = p.x; yi = p.y; zi = p.z; hi = p.w;
    vxi = v.x; vyi = v.y; vzi = v.z; pmassi = v.w;
    rhoi = t.x; pri = fmax(t.y - O.pext, 0.); spsoundi = t.z; uui = t.w;       // :328
    gradhi = g.x; alphai = g.y; alphaui = g.z; alphaBi = g.w;
    if (MHD) { double4 b = ld4(I.bpsi + s); Bxi = b.x; Byi = b.y; Bzi = b.z; psii = b.w; }
    ti = G.typ[s];
    celli = G.cellOf[s];
    if (hi <= 0.) atomicCAS(R.err, 0, 3 /*ND_ERR_H_NONPOSITIVE*/);             // :384-387
  }
  const double rho1i = 1. / rhoi, rho21i = rho1i * rho1i;                     // :325-326
  const double Prho2i = pri * rho21i;                                          // :332
  const double hi1 = 1. / hi, hi21 = __dmul_rn(hi1, hi1);                      // :215, :389
  const double hfacwabi = powndim<NDIM>(hi1), hfacgrkerni = hfacwabi * hi1;    // :390-391
  double Brhoxi = 0, Brhoyi = 0, Brhozi = 0, Brho2i = 0, valfven2i = 0;
  if (MHD) {                                                                   // :362-373
    Brhoxi = Bxi * rho1i; Brhoyi = Byi * rho1i; Brhozi = Bzi * rho1i;
    const double B2i = (Bxi * Bxi + Byi * Byi) + Bzi * Bzi;
    Brho2i = B2i * rho21i;
    valfven2i = B2i * rho1i;
  }
  // accumulators
  double fx = 0, fy = 0, fz = 0, dudt = 0, dBx = 0, dBy = 0, dBz = 0, divB = 0, cBx = 0, cBy = 0, cBz = 0, del2u = 0;
  double gpx = 0, gpy = 0, gpz = 0, gvx = 0, gvy = 0, gvz = 0, endiss = 0;
  double dtcourant = 1.e6, vsigmax = 0., dtav = 1.7976931348623157e308, ts_min = 1.7976931348623157e308, h_on_csts_max = 0.;
  int nclumped = 0;
  const double zero = 1.e-10;

  // ---- phase 1: inclusion test, src/ratesND_mhd.f90:401-415 (bit-exact arithmetic) ----
  auto cull = [&](int k) -> bool {
    if (k == s) return false;                                   // j /= i (both-ghost pairs cannot occur: the target is real)
    const double4 pj = ld4(G.posh + k);
    const double rij2 = dist2_exact(xi - pj.x, yi - pj.y, zi - pj.z);
    const double hj1 = 1. / pj.w;
    const double q2i = __dmul_rn(rij2, hi21), q2j = __dmul_rn(rij2, __dmul_rn(hj1, hj1));
    if (!((q2i < G.radkern2) || (q2j < G.radkern2))) return false;
    if (pair_out_i) {                                           // parity-test hook: record the accepted pair
      unsigned long long n = atomicAdd(pair_count, 1ull);
      if ((long long)n < pair_cap) { pair_out_i[n] = orig + 1; pair_out_j[n] = G.perm[k] + 1; }
    }
    const int tj = __ldg(G.typ + k);
    return types_interact(ti, tj) || (DRAG && O.idrag_nature > 0);   // :436-446
  };

  // ---- phase 2: pair terms ----
  auto body = [&](int k) {
    const double4 pj = ld4(G.posh + k);
    const double4 vj = ld4(G.vm + k);
    const double4 tj4 = ld4(I.thermo + k);
    const int tj = __ldg(G.typ + k);
    const double dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
    const double rij2 = dist2_exact(dx, dy, dz);
    const double hj = pj.w, hj1 = 1. / hj, hj21 = __dmul_rn(hj1, hj1);
    const double q2i = __dmul_rn(rij2, hi21), q2j = __dmul_rn(rij2, hj21);
    const double rij = sqrt(rij2);
    double drx, dry, drz;
    const double eps = 2.220446049250313e-16;
    if (rij <= eps) {                                           // :417-427 coincident particles
      drx = dry = drz = 0.;
      if (tj == ti) {
        const int origj = G.perm[k];
        if (origj >= G.npart || orig > origj) nclumped++;
        if (rij < 2.2250738585072014e-308 && ti != 2) atomicCAS(R.err, 0, 1 /*ND_ERR_INVALID_ARG: dx = 0*/);
      }
    } else {
      const double r1 = 1. / rij;                               // dr = dx/rij, :429
      drx = dx * r1; dry = dy * r1; drz = dz * r1;
    }
    const double pmassj = vj.w;
    const double dvx = vxi - vj.x, dvy = vyi - vj.y, dvz = vzi - vj.z;
    const double rhoj = tj4.x;
    if (types_interact(ti, tj)) {
      // =============================== rates_core ===============================
      const double4 gj = ld4(I.gal + k);
      double wabi, grkerni, wabj, grkernj;
      interp_wg(G, q2i, wabi, grkerni);                         // :1208-1211
      grkerni = grkerni * hfacgrkerni;
      const double hfacwabj = powndim<NDIM>(hj1), hfacgrkernj = hfacwabj * hj1;   // :1215-1216
      interp_wg(G, q2j, wabj, grkernj);                         // :1217-1220
      grkernj = grkernj * hfacgrkernj;
      double grkern;
      if (O.ikernav == 3) {                                     // :1227-1237
        grkerni = grkerni * gradhi;
        grkernj = grkernj * gj.x;
        grkern = 0.5 * (grkerni + grkernj);
      } else {                                                  // :1239-1241
        grkern = 0.5 * (grkerni + grkernj);
        grkerni = grkern; grkernj = grkern;
      }
      const double dvdotr = (dvx * drx + dvy * dry) + dvz * drz;   // :1250
      const double rho1j = 1. / rhoj, rho21j = rho1j * rho1j;      // :1256-1258
      const double rhoav1 = 0.5 * (rho1i + rho1j);                 // :1261
      const double prj = fmax(tj4.y - O.pext, 0.);                 // :1285
      const double Prho2j = prj * rho21j;
      const double spsoundj = tj4.z, uuj = tj4.w;
      double Bxj = 0, Byj = 0, Bzj = 0, psij = 0, dBxx = 0, dByy = 0, dBzz = 0;
      double projBi = 0, projBj = 0, projdB = 0, projBrhoi = 0, projBrhoj = 0, Brho2j = 0, valfven2j = 0;
      double Brhoxj = 0, Brhoyj = 0, Brhozj = 0;
      if (MHD) {                                                // :1295-1313
        const double4 bj = ld4(I.bpsi + k);
        Bxj = bj.x; Byj = bj.y; Bzj = bj.z; psij = bj.w;
        Brhoxj = Bxj * rho1j; Brhoyj = Byj * rho1j; Brhozj = Bzj * rho1j;
        dBxx = Bxi - Bxj; dByy = Byi - Byj; dBzz = Bzi - Bzj;
        projBi = (Bxi * drx + Byi * dry) + Bzi * drz;
        projBj = (Bxj * drx + Byj * dry) + Bzj * drz;
        projdB = (dBxx * drx + dByy * dry) + dBzz * drz;
        projBrhoi = (Brhoxi * drx + Brhoyi * dry) + Brhozi * drz;
        projBrhoj = (Brhoxj * drx + Brhoyj * dry) + Brhozj * drz;
        const double B2j = (Bxj * Bxj + Byj * Byj) + Bzj * Bzj;
        valfven2j = B2j * rho1j;
        Brho2j = B2j * rho21j;
      }
      // ---- signal velocities :1417-1465 ----
      double vsigi, vsigj, vsigB;
      if (MHD) {
        const double vsig2i = spsoundi * spsoundi + valfven2i;
        const double vsig2j = spsoundj * spsoundj + valfven2j;
        const double vsigproji = vsig2i * vsig2i - 4. * ((spsoundi * projBi) * (spsoundi * projBi)) * rho1i;
        const double vsigprojj = vsig2j * vsig2j - 4. * ((spsoundj * projBj) * (spsoundj * projBj)) * rho1j;
        if (vsigproji < 0. || vsigprojj < 0.) atomicCAS(R.err, 0, 6 /*ND_ERR_VSIG_DET*/);
        vsigi = sqrt(0.5 * (vsig2i + sqrt(vsigproji)));
        vsigj = sqrt(0.5 * (vsig2j + sqrt(vsigprojj)));
        if (O.iavlim2 != 2) vsigB = sqrt((dvx * dvx + dvy * dvy) + dvz * dvz);   // norm2(dvel), :1433
        else vsigB = 0.5 * (vsigi + vsigj) + fabs(dvdotr);
      } else {
        vsigi = spsoundi; vsigj = spsoundj; vsigB = 0.;
      }
      double vsig = 0.5 * (fmax(vsigi + vsigj - O.beta * dvdotr, 0.0));          // :1452
      double vsigu = sqrt(fabs(pri - prj) * rhoav1);                            // :1459 (pequil = 0)
      const double vsigdtc = fmax(vsig, fmax(0.5 * (vsigi + vsigj + O.beta * fabs(dvdotr)), vsigB));   // :1465
      if (ti == T_DUST) { vsig = 0.; vsigu = 0.; }                              // :1472-1474
      else {
        const double dvsigdtc = 1. / vsigdtc;                                   // :1476-1481
        vsigmax = fmax(vsigmax, vsigdtc);
        if (vsigdtc > zero) dtcourant = fmin(dtcourant, fmin(hi * dvsigdtc, hj * dvsigdtc));
      }
      double fix = 0, fiy = 0, fiz = 0;   // forcei contribution of this pair
      double vsigav = 0.;
      if (O.iav > 0 && O.iav != 3) {
        // =============================== artificial_dissipation ===============================
        const double alphaav = 0.5 * (alphai + gj.y), alphau = 0.5 * (alphaui + gj.z), alphaB = 0.5 * (alphaBi + gj.w);   // :1712-1714
        vsigav = fmax(alphaav, fmax(alphau, alphaB)) * vsig;
        const double term = vsig * rhoav1 * grkern;              // :1723
        const double termu = vsigu * rhoav1 * grkern;            // :1727
        const double termB = vsigB * rhoav1 * grkern;            // :1732
        if (dvdotr < 0) {                                        // :1745-1748
          const double visc = alphaav * term * (-dvdotr);
          const double c = pmassj * visc;
          fix -= c * drx; fiy -= c * dry; fiz -= c * drz;
        }
        if (MHD) {                                               // :1762-1774
          double bvx, bvy, bvz;
          if (O.iav >= 2) { bvx = dBxx * rhoav1; bvy = dByy * rhoav1; bvz = dBzz * rhoav1; }
          else { bvx = (dBxx - drx * projdB) * rhoav1; bvy = (dByy - dry * projdB) * rhoav1; bvz = (dBzz - drz * projdB) * rhoav1; }
          const double c = rhoi * pmassj, ab = alphaB * termB;
          dBx += c * (ab * bvx); dBy += c * (ab * bvy); dBz += c * (ab * bvz);
        }
        if (O.iener == 3) {                                      // :1792-1830 total energy: pair part of dendt
          double qdiff = 0.;
          const double projvi = (vxi * drx + vyi * dry) + vzi * drz, projvj = (vj.x * drx + vj.y * dry) + vj.z * drz;
          if (dvdotr < 0) qdiff += term * alphaav * 0.5 * (projvi * projvi - projvj * projvj);
          qdiff += alphau * termu * (uui - uuj);
          if (MHD) {
            double B2i_, B2j_;
            if (O.iav >= 2) { B2i_ = (Bxi * Bxi + Byi * Byi) + Bzi * Bzi; B2j_ = (Bxj * Bxj + Byj * Byj) + Bzj * Bzj; }
            else { B2i_ = ((Bxi * Bxi + Byi * Byi) + Bzi * Bzi) - projBi * projBi; B2j_ = ((Bxj * Bxj + Byj * Byj) + Bzj * Bzj) - projBj * projBj; }
            qdiff += alphaB * termB * 0.5 * (B2i_ - B2j_) * rhoav1;
          }
          endiss += pmassj * qdiff;                           // :1829
        } else if (O.iener > 0) {                                // :1835-1875 thermal energy
          double vissv = 0., vissB = 0.;
          if (dvdotr < 0) { const double t = ((vxi * drx + vyi * dry) + vzi * drz) - ((vj.x * drx + vj.y * dry) + vj.z * drz); vissv = -alphaav * 0.5 * (t * t); }
          const double vissu = alphau * (uui - uuj);
          if (MHD) {
            const double dB2 = (dBxx * dBxx + dByy * dByy) + dBzz * dBzz;
            if (O.iav >= 2) vissB = -alphaB * 0.5 * dB2 * rhoav1;
            else vissB = -alphaB * 0.5 * (dB2 - projdB * projdB) * rhoav1;
          }
          dudt += pmassj * (term * vissv + termu * vissu + termB * vissB);
        }
      } else if (O.iav == 3) {
        // =============================== artificial_dissipation_phantom ===============================
        double dudti = 0.;
        if (dvdotr < 0.) {
          const double vsi = fmax(alphai * spsoundi - O.beta * dvdotr, 0.);
          const double vsj = fmax(gj.y * spsoundj - O.beta * dvdotr, 0.);
          const double qi = -0.5 * rhoi * vsi * dvdotr, qj = -0.5 * rhoj * vsj * dvdotr;
          const double visc = (qi * rho21i * grkerni + qj * rho21j * grkernj);
          const double c = pmassj * visc;
          fix -= c * drx; fiy -= c * dry; fiz -= c * drz;
          dudti = qi * rho21i * pmassj * dvdotr * grkerni;
        }
        const double du = uui - uuj;
        const double cfaci = 0.5 * alphaui * rhoi * vsigu * du, cfacj = 0.5 * gj.z * rhoj * vsigu * du;
        const double diffu = cfaci * grkerni * (rho1i * rho1i) + cfacj * grkernj * (rho1j * rho1j);
        dudt += dudti + pmassj * diffu;
      }
      if (vsigav > zero) dtav = fmin(dtav, fmin(hi / vsigav, hj / vsigav));     // :1500
      {                                                          // pressure, :1538-1567 (phi = 1, sqrtg = 1)
        const double prterm = Prho2i * grkerni + Prho2j * grkernj;
        const double c = pmassj * prterm;
        fix -= c * drx; fiy -= c * dry; fiz -= c * drz;
      }
      if (MHD) {
        // =============================== mhd_terms ===============================
        const double fiso = 0.5 * (Brho2i * grkerni + Brho2j * grkernj);        // :2526
        const double sm = O.stressmax;
        const double fax = (Brhoxi * projBrhoi - sm * drx * rho21i) * grkerni + (Brhoxj * projBrhoj - sm * drx * rho21j) * grkernj;   // :2534-2537
        const double fay = (Brhoyi * projBrhoi - sm * dry * rho21i) * grkerni + (Brhoyj * projBrhoj - sm * dry * rho21j) * grkernj;
        const double faz = (Brhozi * projBrhoi - sm * drz * rho21i) * grkerni + (Brhozj * projBrhoj - sm * drz * rho21j) * grkernj;
        fix += pmassj * (fax - fiso * drx);                      // :2541, :2629
        fiy += pmassj * (fay - fiso * dry);
        fiz += pmassj * (faz - fiso * drz);
        divB -= pmassj * projdB * grkern;                        // :2552
        const double mg = pmassj * grkern;                       // :2601-2602 curlB += pmassj*(dB x dr)*grkern
        cBx += (dByy * drz - dBzz * dry) * mg;
        cBy += (dBzz * drx - dBxx * drz) * mg;
        cBz += (dBxx * dry - dByy * drx) * mg;
        const double ci = pmassj * projBrhoi * grkerni;          // :2664-2665 induction (imhd = 1, 11)
        dBx -= dvx * ci; dBy -= dvy * ci; dBz -= dvz * ci;
        if (O.iresist == 1) {                                    // :2685-2709
          const double etaij = O.etamhd;                         // 0.5*(etai + etaj) with constant eta
          const double f = -2. * etaij / (rij + eps);
          const double c = rhoi * pmassj * 0.5 * ((rho1i * rho1i) * grkerni + (rho1j * rho1j) * grkernj);
          dBx -= c * (f * dBxx); dBy -= c * (f * dByy); dBz -= c * (f * dBzz);
          if (O.iener > 0) dudt += pmassj * (-etaij * rho1i * rho1j * ((dBxx * dBxx + dByy * dByy) + dBzz * dBzz) * grkern / rij);
        }
        if (O.idivbzero >= 2) {                                  // :2712-2716
          const double gradpsiterm = psii * rho21i * grkerni + psij * rho21j * grkernj;
          const double c = pmassj * gradpsiterm;
          gpx -= c * drx; gpy -= c * dry; gpz -= c * drz;
        }
      }
      fx += fix; fy += fiy; fz += fiz;
      if (O.iav > 0) {                                           // :1639-1656 switch sources
        if (O.iavlim1 > 0) del2u += pmassj * rho1j * ((uui - uuj) / rij) * grkerni;
        if (O.iavlim0 == 3) {
          const double c = pmassj * rho1j / rij * dvdotr * grkerni;
          gvx += c * drx; gvy += c * dry; gvz += c * drz;
        } else {
          const double c = pmassj * grkerni;
          gvx += c * (dvx - dvdotr); gvy += c * (dvy - dvdotr); gvz += c * (dvz - dvdotr);
        }
      }
    } else if (DRAG) {
      // =============================== drag_forces ===============================
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
        double wab, spsoundgas, ts, hgas;
        if (igas) { wab = interp_drag(G, q2i) * hfacwabi; spsoundgas = spsoundi; ts = get_tstop(O.idrag_nature, rhoi, rhoj, O.Kdrag); hgas = hi; }
        else { wab = interp_drag(G, q2j) * powndim<NDIM>(hj1); spsoundgas = tj4.z; ts = get_tstop(O.idrag_nature, rhoj, rhoi, O.Kdrag); hgas = hj; }
        h_on_csts_max = fmax(h_on_csts_max, hgas / (spsoundgas * ts));
        ts_min = fmin(ts_min, ts);
        const double dragterm = NDIM * wab / ((rhoi + rhoj) * ts) * projv;   // :1156 (projvstar = projv)
        const double c = dragterm * pmassj;
        fx -= c * ddx; fy -= c * ddy; fz -= c * ddz;
        if (ti == T_GAS) dudt += pmassj * (dragterm * projv);   // :1162-1164
      }
    }
  };

  neighbour_walk<NDIM, RATES_CAP, RATES_BLOCK>(G, active, celli, nlist_smem, cull, body);

  if (active) {
    S.F[s] = make_double4(fx, fy, fz, dudt);
    S.dB[s] = make_double4(dBx, dBy, dBz, divB);
    S.C[s] = make_double4(cBx, cBy, cBz, del2u);
    S.P[s] = make_double4(gpx, gpy, gpz, endiss);
    S.V[s] = make_double4(gvx, gvy, gvz, 0.);
  }
  // block-free warp reductions into global min/max keys
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
