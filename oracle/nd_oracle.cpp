/*
 * nd_oracle.cpp -- CPU restatement of the NDSPMHD hot path.  TEST INFRASTRUCTURE ONLY (see nd_oracle.h).
 * PARITY UNPINNED by reference fixtures (none exist; no Fortran compiler here) -- pinned by invariants in tests/.
 *
 * Every routine cites the reference file:line it follows (paths relative to /root/reference/).
 * Style: serial, same loop order as the Fortran, 1-based particle/cell indices kept through accessor
 * macros so the code reads side by side with the source.  Build: g++ -O2 -ffp-contract=off -fno-fast-math
 * (mirrors src/Makefile:25-27: -O3 without -ffast-math, default real = 8 bytes, no FMA on baseline x86-64).
 */
#include "nd_oracle.h"
#include <cmath>
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <chrono>
#include <vector>
#include <algorithm>
#include <string>

namespace {

const int ikern = 4000;                    // src/kernelND.f90:39
const double pi_k = 3.141592653589;         // src/kernelND.f90:41 (truncated on purpose)
std::string g_err;

inline double powi(double x, int n) {      // gfortran integer powers: repeated multiplication (GCC powi_table)
  switch (n) {
    case 0: return 1.0;
    case 1: return x;
    case 2: return x * x;
    case 3: return (x * x) * x;
    case 4: { double x2 = x * x; return x2 * x2; }
    case 5: { double x2 = x * x; return (x2 * x) * x2; }
    default: { double r = 1.0; for (int i = 0; i < n; i++) r *= x; return r; }
  }
}

struct Kern {
  double wij[ikern + 1], grwij[ikern + 1], grgrwij[ikern + 1];
  double wijalt[ikern + 1], grwijalt[ikern + 1], grgrwijalt[ikern + 1];
  double wijdrag[ikern + 1], grwijdrag[ikern + 1], grgrwijdrag[ikern + 1];
  double radkern, radkern2, dq2table, ddq2table;
};

// src/kernelND.f90:127-4289 setkerntable (cases 0, 2, 3, 41, 42)
int setkerntable(Kern &K, int ikernel, int ndim, double *wkern, double *grwkern, double *grgrwkern) {
  double cnormk = 0.0;
  for (int i = 0; i <= ikern; i++) { wkern[i] = 0.; grwkern[i] = 0.; grgrwkern[i] = 0.; }
  switch (ikernel) {
    case 2: {  // M5 quartic, :152-203
      K.radkern = std::max(K.radkern, 2.5);
      K.radkern2 = K.radkern * K.radkern;
      K.dq2table = K.radkern2 / double(ikern);
      if (ndim == 1) cnormk = 1. / 24.;
      else if (ndim == 2) cnormk = 96. / (1199. * pi_k);
      else cnormk = 0.05 / pi_k;
      for (int i = 0; i <= ikern; i++) {
        double q2 = i * K.dq2table, q = std::sqrt(q2);
        if (q < 0.5) {
          wkern[i] = 6.0 * q2 * q2 - 15.0 * q2 + 14.375;
          grwkern[i] = q * (24.0 * q2 - 30.0);
          grgrwkern[i] = 72.0 * q2 - 30.0;
        } else if (q < 1.5) {
          wkern[i] = -5.0 * powi(-q + 1.5, 4) + powi(-q + 2.5, 4);
          grwkern[i] = -16.0 * q2 * q + 60.0 * q2 - 60.0 * q + 5.0;
          grgrwkern[i] = -48.0 * q2 + 120.0 * q - 60.0;
        } else if (q < 2.5) {
          wkern[i] = powi(-q + 2.5, 4);
          grwkern[i] = -4.0 * powi(-q + 2.5, 3);
          grgrwkern[i] = 12.0 * q2 - 60.0 * q + 75.0;
        }
      }
    } break;
    case 3: {  // M6 quintic, :205-269
      K.radkern = std::max(K.radkern, 3.0);
      K.radkern2 = K.radkern * K.radkern;
      K.dq2table = K.radkern2 / double(ikern);
      if (ndim == 1) cnormk = 1. / 120.;
      else if (ndim == 2) cnormk = 7. / (478 * pi_k);
      else cnormk = 1. / (120. * pi_k);
      for (int i = 0; i <= ikern; i++) {
        double q2 = i * K.dq2table, q4 = q2 * q2, q = std::sqrt(q2);
        double term1 = -5. * std::pow(3. - q, 4.);  // real exponent in the source (:230)
        if (q < 1.0) {
          wkern[i] = 66. - 60. * q2 + 30. * q4 - 10. * q4 * q;
          grwkern[i] = term1 + 30 * powi(2. - q, 4) - 75. * powi(1. - q, 4);
          grgrwkern[i] = 20. * powi(3. - q, 3) - 120. * powi(2. - q, 3) + 300. * powi(1. - q, 3);
        } else if (q >= 1.0 && q < 2.0) {
          wkern[i] = powi(3. - q, 5) - 6. * powi(2. - q, 5);
          grwkern[i] = term1 + 30 * powi(2. - q, 4);
          grgrwkern[i] = 20. * powi(3. - q, 3) - 120. * powi(2. - q, 3);
        } else if (q >= 2.0 && q <= 3.0) {
          wkern[i] = powi(3. - q, 5);
          grwkern[i] = term1;
          grgrwkern[i] = 20. * powi(3. - q, 3);
        }
      }
    } break;
    case 41: {  // double hump M4, :1659-1697
      K.radkern = std::max(K.radkern, 2.0);
      K.radkern2 = K.radkern * K.radkern;
      K.dq2table = K.radkern * K.radkern / double(ikern);
      if (ndim == 1) cnormk = 2.0;
      else if (ndim == 2) cnormk = 70. / (31. * pi_k);
      else cnormk = 10. / (9. * pi_k);
      for (int i = 0; i <= ikern; i++) {
        double q2 = i * K.dq2table, q = std::sqrt(q2), q4 = q2 * q2;
        if (q < 1.0) {
          wkern[i] = q2 - 1.5 * q4 + 0.75 * q4 * q;
          grwkern[i] = 2. * q - 6. * q2 * q + 3.75 * q4;
          grgrwkern[i] = 2. - 18. * q2 + 15. * q2 * q;
        } else if (q >= 1.0 && q <= 2.0) {
          wkern[i] = 0.25 * q2 * powi(2. - q, 3);
          grwkern[i] = 0.5 * q * powi(2. - q, 3) - 0.75 * q2 * powi(2. - q, 2);
          grgrwkern[i] = 0.5 * powi(2. - q, 3) - 3. * q * powi(2. - q, 2) + 1.5 * q2 * (2. - q);
        }
      }
    } break;
    case 42: {  // double hump M5, :1698-1757
      K.radkern = 2.5;
      K.radkern2 = K.radkern * K.radkern;
      K.dq2table = K.radkern2 / double(ikern);
      if (ndim == 1) cnormk = 1. / 10.;
      else if (ndim == 2) cnormk = 3584. / (35783. * pi_k);
      else cnormk = 1. / (23. * pi_k);
      for (int i = 0; i <= ikern; i++) {
        double q2 = i * K.dq2table, q4 = q2 * q2, q = std::sqrt(q2);
        if (q < 0.5) {
          wkern[i] = q2 * (6.0 * q4 - 15.0 * q2 + 14.375);
          grwkern[i] = q * (36.0 * q4 - 60.0 * q2 + 28.75);
          grgrwkern[i] = 180.0 * q4 - 180.0 * q2 + 28.75;
        } else if (q < 1.5) {
          wkern[i] = q2 * (powi(q - 2.5, 4) - 5.0 * powi(q - 1.5, 4));
          grwkern[i] = q * (-24.0 * q4 + 100.0 * q2 * q - 120.0 * q2 + 15.0 * q + 27.5);
          grgrwkern[i] = -120.0 * q4 + 400.0 * q2 * q - 360.0 * q2 + 30.0 * q + 27.5;
        } else if (q < 2.5) {
          wkern[i] = q2 * powi(q - 2.5, 4);
          grwkern[i] = q * powi(q - 2.5, 3) * (6.0 * q - 5.0);
          grgrwkern[i] = 30.0 * q4 - 200.0 * q2 * q + 450.0 * q2 - 375.0 * q + 78.125;
        }
      }
    } break;
    case 0: {  // M4 cubic spline, :4143-4193
      K.radkern = std::max(K.radkern, 2.0);
      K.radkern2 = K.radkern * K.radkern;
      K.dq2table = K.radkern * K.radkern / double(ikern);
      if (ndim == 1) cnormk = 0.66666666666;
      else if (ndim == 2) cnormk = 10. / (7. * pi_k);
      else cnormk = 1. / pi_k;
      for (int i = 0; i <= ikern; i++) {
        double q2 = i * K.dq2table, q = std::sqrt(q2);
        if (q < 1.0) {
          wkern[i] = 1. - 1.5 * q2 + 0.75 * q * q2;
          grwkern[i] = -3. * q + 2.25 * q2;
          grgrwkern[i] = -3. + 4.5 * q;
        } else if (q >= 1.0 && q <= 2.0) {
          wkern[i] = 0.25 * powi(2. - q, 3);
          grwkern[i] = -0.75 * powi(2. - q, 2);
          grgrwkern[i] = 1.5 * (2. - q);
        }
      }
    } break;
    default:
      return ND_ERR_UNSUPPORTED_OPTION;
  }
  // :4279-4289 normalise, ddq2table
  for (int i = 0; i <= ikern; i++) { wkern[i] = cnormk * wkern[i]; grwkern[i] = cnormk * grwkern[i]; grgrwkern[i] = cnormk * grgrwkern[i]; }
  K.ddq2table = 1. / K.dq2table;
  return 0;
}

// src/initialiseND_mhd.f90:179-216
int setkernels(Kern &K, int ikernel, int ikernelalt, int idust, int ndim) {
  K.radkern = 2.;  // kernelND.f90:111
  int e = setkerntable(K, ikernel, ndim, K.wij, K.grwij, K.grgrwij);
  if (e) return e;
  e = setkerntable(K, ikernelalt, ndim, K.wijalt, K.grwijalt, K.grgrwijalt);
  if (e) return e;
  double radkernold = K.radkern;
  for (int i = 0; i <= ikern; i++) { K.wijdrag[i] = 0; K.grwijdrag[i] = 0; K.grgrwijdrag[i] = 0; }
  if (idust != 0) {
    int kd = (ikernel == 0) ? 41 : (ikernel == 2) ? 42 : -1;
    if (kd < 0) return ND_ERR_UNSUPPORTED_OPTION;
    e = setkerntable(K, kd, ndim, K.wijdrag, K.grwijdrag, K.grgrwijdrag);
    if (e) return e;
    if (K.radkern != radkernold) return ND_ERR_UNSUPPORTED_OPTION;
  }
  return 0;
}

inline void kindex(const Kern &K, double q2, int &index, int &index1) {
  // index = int(q2*ddq2table); clamp (kernelND.f90:4435-4438).  Guard the cast (Fortran int() of a huge real
  // is processor-dependent; gfortran yields a negative number that the clamp maps to ikern).
  double t = q2 * K.ddq2table;
  if (!(t < 2147483000.0) || t < 0) { index = ikern; index1 = ikern; return; }
  index = (int)t;
  index1 = index + 1;
  if (index > ikern || index < 0) index = ikern;
  if (index1 > ikern || index1 < 0) index1 = ikern;
}

// src/kernelND.f90:4367 interpolate_kernel
inline void interpolate_kernel(const Kern &K, double q2, double &w, double &gradw) {
  int index, index1; kindex(K, q2, index, index1);
  double dxx = q2 - index * K.dq2table;
  double dwdx = (K.wij[index1] - K.wij[index]) * K.ddq2table;
  w = (K.wij[index] + dwdx * dxx);
  double dgrwdx = (K.grwij[index1] - K.grwij[index]) * K.ddq2table;
  gradw = (K.grwij[index] + dgrwdx * dxx);
}
// src/kernelND.f90:4398 interpolate_kerneldrag
inline void interpolate_kerneldrag(const Kern &K, double q2, double &w) {
  int index, index1; kindex(K, q2, index, index1);
  double dxx = q2 - index * K.dq2table;
  double dwdx = (K.wijdrag[index1] - K.wijdrag[index]) * K.ddq2table;
  w = (K.wijdrag[index] + dwdx * dxx);
}
// src/kernelND.f90:4426 interpolate_kernels
inline void interpolate_kernels(const Kern &K, double q2, double &w, double &gradw, double &gradwalt, double &gradgradwalt) {
  int index, index1; kindex(K, q2, index, index1);
  double dxx = q2 - index * K.dq2table;
  w = K.wij[index];
  double dwdx = (K.wij[index1] - w) * K.ddq2table;
  w = w + dwdx * dxx;
  gradw = K.grwij[index];
  double dgrwdx = (K.grwij[index1] - gradw) * K.ddq2table;
  gradw = gradw + dgrwdx * dxx;
  gradwalt = K.grwijalt[index];
  double dgrwaltdx = (K.grwijalt[index1] - gradwalt) * K.ddq2table;
  gradwalt = gradwalt + dgrwaltdx * dxx;
  gradgradwalt = K.grgrwijalt[index];
  double dgrgrwaltdx = (K.grgrwijalt[index1] - gradgradwalt) * K.ddq2table;
  gradgradwalt = gradgradwalt + dgrgrwaltdx * dxx;
}
// src/kernelND.f90:4599 interpolate_kernels_dens
inline void interpolate_kernels_dens(const Kern &K, double q2, double &w, double &gradw, double &gradgradw, double &walt, double &gradwalt) {
  int index, index1; kindex(K, q2, index, index1);
  double dxx = q2 - index * K.dq2table;
  double dwdx = (K.wij[index1] - K.wij[index]) * K.ddq2table;
  w = (K.wij[index] + dwdx * dxx);
  double dgrwdx = (K.grwij[index1] - K.grwij[index]) * K.ddq2table;
  gradw = (K.grwij[index] + dgrwdx * dxx);
  double dgrgrwdx = (K.grgrwij[index1] - K.grgrwij[index]) * K.ddq2table;
  gradgradw = (K.grgrwij[index] + dgrgrwdx * dxx);
  walt = K.wijalt[index];
  double dwaltdx = (K.wijalt[index1] - walt) * K.ddq2table;
  walt = walt + dwaltdx * dxx;
  gradwalt = K.grwijalt[index];
  double dgrwaltdx = (K.grwijalt[index1] - gradwalt) * K.ddq2table;
  gradwalt = gradwalt + dgrwaltdx * dxx;
}

// ------------------------------------------------------------------------------------------------
struct St {
  const nd_options *o;
  int ndim, npart, ntotal, idim;
  ndo_arrays a;
  Kern K;
  // module linklist (src/variablesND.f90:111-118)
  std::vector<int> ll, ifirstincell, iamincell;
  int ncellsx[3], ncells, ncellsloop;
  double dxcell;
  double hhmax;   // module bound
  // module hterms / timestep
  int itsdensity;
  long long ncalctotal;
  double dtcourant, dtforce, dtav, dtdrag, dtvisc, vsig2max, vsigmax_out, stressmax_out, ts_min_out, h_on_csts_max_out, fhmax_out;
  double fmean[3];
  int nclumped;
  std::vector<double> gradB;   // module derivB: gradB(3,3,idim), filled by get_curl when iavlim(3) = 2
  long long npairs_rates;   // ordered pairs evaluated by get_rates, counted on the side(s) that are real particles (test checksum, no reference counterpart)
  int err;
};

// 1-based accessors in the reference's layout
#define X_(k, i) S.a.x[(size_t)((i)-1) * S.ndim + ((k)-1)]
#define V3(arr, k, i) S.a.arr[(size_t)((i)-1) * 3 + ((k)-1)]
#define A1(arr, i) S.a.arr[(size_t)((i)-1)]

int fail(St &S, int code, const char *msg) { g_err = msg; S.err = code; return code; }

// src/copy_particle.f90:28-115
void copy_particle(St &S, int i, int j) {
  const nd_options &o = *S.o;
  A1(pmass, i) = A1(pmass, j);
  A1(rho, i) = A1(rho, j);
  A1(rhoalt, i) = A1(rhoalt, j);
  A1(hh, i) = A1(hh, j);
  A1(uu, i) = A1(uu, j);
  A1(en, i) = A1(en, j);
  if (!(o.imhd < 0 && A1(itype, i) > 0)) for (int k = 1; k <= 3; k++) V3(Bevol, k, i) = V3(Bevol, k, j);
  for (int k = 1; k <= 3; k++) V3(Bfield, k, i) = V3(Bfield, k, j);
  for (int k = 1; k <= 3; k++) V3(alpha, k, i) = V3(alpha, k, j);
  A1(psi, i) = A1(psi, j);
  A1(gradh, i) = A1(gradh, j);
  A1(gradhn, i) = A1(gradhn, j);
  A1(gradsoft, i) = A1(gradsoft, j);
  A1(gradgradh, i) = A1(gradgradh, j);
  A1(sqrtg, i) = A1(sqrtg, j);
  A1(spsound, i) = A1(spsound, j);
  A1(pr, i) = A1(pr, j);
  A1(dens, i) = A1(dens, j);
  int ti = A1(itype, i);
  if (ti != ND_ITYPE_BND && ti != ND_ITYPE_BNDDUST && ti != ND_ITYPE_BND2 && A1(itype, j) == ND_ITYPE_GAS) A1(itype, i) = A1(itype, j);
  for (int k = 1; k <= 3; k++) V3(force, k, i) = V3(force, k, j);
  A1(drhodt, i) = A1(drhodt, j);
  A1(dudt, i) = A1(dudt, j);
  A1(dendt, i) = A1(dendt, j);
  for (int k = 1; k <= 3; k++) V3(dBevoldt, k, i) = V3(dBevoldt, k, j);
  for (int k = 1; k <= 3; k++) V3(gradpsi, k, i) = V3(gradpsi, k, j);
  A1(dhdt, i) = A1(dhdt, j);
  for (int k = 1; k <= 3; k++) V3(daldt, k, i) = V3(daldt, k, j);
  A1(dpsidt, i) = A1(dpsidt, j);
  for (int k = 1; k <= 3; k++) V3(xsphterm, k, i) = V3(xsphterm, k, j);
  for (int k = 1; k <= 3; k++) V3(fmag, k, i) = V3(fmag, k, j);
  A1(divB, i) = A1(divB, j);
  for (int k = 1; k <= 3; k++) V3(curlB, k, i) = V3(curlB, k, j);
  for (int k = 1; k <= 3; k++) V3(graddivv, k, i) = V3(graddivv, k, j);
  if (o.onef_dust) {                                            // :84-92
    A1(dustfrac, i) = A1(dustfrac, j); A1(dustevol, i) = A1(dustevol, j); A1(ddustevoldt, i) = A1(ddustevoldt, j);
    for (int k = 1; k <= 3; k++) V3(deltav, k, i) = V3(deltav, k, j);
    for (int k = 1; k <= 3; k++) V3(ddeltavdt, k, i) = V3(ddeltavdt, k, j);
    A1(rhodust, i) = A1(rhodust, j); A1(rhogas, i) = A1(rhogas, j);
  }
}

// src/ghostND_mhd.f90:363-431 makeghost
int makeghost(St &S, int jpart, const double *xghost, const double *vghost) {
  int ipart = S.ntotal + 1;
  if (ipart > S.idim) return fail(S, ND_ERR_INVALID_ARG, "ghost: ntotal > array size (idim too small)");
  S.ntotal = ipart;
  for (int k = 1; k <= S.ndim; k++) X_(k, ipart) = xghost[k - 1];
  copy_particle(S, ipart, jpart);
  for (int k = 1; k <= 3; k++) V3(vel, k, ipart) = vghost[k - 1];
  A1(ireal, ipart) = jpart;
  return 0;
}

// src/ghostND_mhd.f90:33-358 set_ghost_particles (periodic ibound=3 and reflecting ibound=2; no shearing box)
int set_ghost_particles(St &S) {
  const nd_options &o = *S.o;
  const int ndim = S.ndim;
  S.ntotal = S.npart;
  double hhmax = A1(hh, 1);
  for (int i = 2; i <= S.npart; i++) hhmax = std::max(hhmax, A1(hh, i));
  S.hhmax = hhmax;                                              // :79
  double dxbound[3], xpart[3], vpart[3], xnew[3][2];
  bool imakeghost[3][2];
  for (int d = 0; d < 3; d++) dxbound[d] = S.K.radkern * hhmax; // :80
  for (int jpart = 1; jpart <= S.npart; jpart++) {              // :166
    for (int d = 0; d < ndim; d++)
      if (o.ibound[d] == 2 || o.ibound[d] == 4 || o.ibound[d] == 6) dxbound[d] = S.K.radkern * A1(hh, jpart);  // :173
    for (int idimen = 1; idimen <= ndim; idimen++) {            // :176
      if (o.ibound[idimen - 1] <= 1) {
        imakeghost[idimen - 1][0] = imakeghost[idimen - 1][1] = false;
        continue;
      }
      for (int imaxmin = 1; imaxmin <= 2; imaxmin++) {          // :184
        for (int k = 1; k <= ndim; k++) xpart[k - 1] = X_(k, jpart);
        for (int k = 1; k <= 3; k++) vpart[k - 1] = V3(vel, k, jpart);
        double xbound, xperbound, dx;
        if (imaxmin == 1) { xbound = o.xmax[idimen - 1]; xperbound = o.xmin[idimen - 1]; dx = o.xmax[idimen - 1] - X_(idimen, jpart); }
        else { xbound = o.xmin[idimen - 1]; xperbound = o.xmax[idimen - 1]; dx = X_(idimen, jpart) - o.xmin[idimen - 1]; }
        imakeghost[idimen - 1][imaxmin - 1] = ((dx < dxbound[idimen - 1]) && (dx > 0));   // :205
        if (!imakeghost[idimen - 1][imaxmin - 1]) continue;
        double dxshift = X_(idimen, jpart) - xbound;            // :212
        int ib = o.ibound[idimen - 1];
        if (ib == 3) xnew[idimen - 1][imaxmin - 1] = xperbound + dxshift;                 // :225
        else if (ib == 2 || ib == 4 || ib == 6) { xnew[idimen - 1][imaxmin - 1] = xbound - dxshift; vpart[idimen - 1] = -V3(vel, idimen, jpart); }
        xpart[idimen - 1] = xnew[idimen - 1][imaxmin - 1];      // :233
        if (int e = makeghost(S, jpart, xpart, vpart)) return e; // :248
        if (idimen > 1) {                                       // :256 edges
          for (int idimenprev = 1; idimenprev <= idimen - 1; idimenprev++) {
            for (int imaxminprev = 1; imaxminprev <= 2; imaxminprev++) {
              if (!imakeghost[idimenprev - 1][imaxminprev - 1]) continue;
              xpart[idimenprev - 1] = xnew[idimenprev - 1][imaxminprev - 1];   // :266
              int ibp = o.ibound[idimenprev - 1];
              if (ibp == 2 || ibp == 4 || ibp == 6) for (int k = 1; k <= 3; k++) vpart[k - 1] = V3(vel, k, jpart);  // :283-284 (sic)
              if (int e = makeghost(S, jpart, xpart, vpart)) return e;        // :287
              if (idimenprev >= 2) {                            // :293 corners
                int idpp = idimenprev - 1;
                for (int imaxminpp = 1; imaxminpp <= 2; imaxminpp++) {
                  if (!imakeghost[idpp - 1][imaxminpp - 1]) continue;
                  xpart[idpp - 1] = xnew[idpp - 1][imaxminpp - 1];            // :300
                  int ibpp = o.ibound[idpp - 1];
                  if (ibpp == 2 || ibpp == 4 || ibpp == 6) for (int k = 1; k <= 3; k++) vpart[k - 1] = -V3(vel, k, jpart); // :307-309 (sic)
                  if (int e = makeghost(S, jpart, xpart, vpart)) return e;    // :316
                  xpart[idpp - 1] = X_(idpp, jpart);            // :318
                }
              }
              xpart[idimenprev - 1] = X_(idimenprev, jpart);    // :327
            }
          }
        }
      }
    }
  }
  for (int i = S.npart + 1; i <= S.ntotal; i++) A1(itype, i) = A1(itype, A1(ireal, i));  // :346
  for (int i = S.ntotal + 1; i <= S.idim; i++) { A1(rho, i) = 0.; A1(uu, i) = 0.; }        // :350-355
  return 0;
}

// src/linkND.f90:45-161 set_linklist
int set_linklist(St &S) {
  const nd_options &o = *S.o;
  const int ndim = S.ndim;
  bool allle1 = true;
  for (int d = 0; d < ndim; d++) if (o.ibound[d] > 1) allle1 = false;
  if (allle1) {                                                 // :70
    double m = A1(hh, 1);
    for (int i = 2; i <= S.npart; i++) m = std::max(m, A1(hh, i));
    S.hhmax = m;
  }
  S.dxcell = S.K.radkern * S.hhmax;                             // :72
  if (S.dxcell <= 0) return fail(S, ND_ERR_LINK, "link: max h <= 0");
  double xminpart[3], xmaxpart[3];
  for (int j = 1; j <= ndim; j++) {                             // :81-84
    double mn = X_(j, 1), mx = X_(j, 1);
    for (int i = 2; i <= S.ntotal; i++) { mn = std::min(mn, X_(j, i)); mx = std::max(mx, X_(j, i)); }
    xminpart[j - 1] = mn - 0.00001;
    xmaxpart[j - 1] = mx + 0.00001;
  }
  for (int j = 0; j < ndim; j++) {                              // :88-89
    xminpart[j] = xminpart[j] - S.dxcell - 0.00001;
    xmaxpart[j] = xmaxpart[j] + S.dxcell + 0.00001;
  }
  S.ncellsx[0] = S.ncellsx[1] = S.ncellsx[2] = 1;
  long long nc = 1;
  for (int j = 0; j < ndim; j++) {                              // :93
    S.ncellsx[j] = (int)((xmaxpart[j] - xminpart[j]) / S.dxcell) + 1;
    if (S.ncellsx[j] == 0) return fail(S, ND_ERR_LINK, "link: number of cells=0");
    nc *= S.ncellsx[j];
  }
  if (nc > 2000000000LL) return fail(S, ND_ERR_LINK, "link: too many cells");
  S.ncells = (int)nc;
  S.ncellsloop = S.ncells;
  S.ifirstincell.assign(S.ncells + 1, -1);                      // :109-114 (1-based)
  S.ll.assign(S.idim + 1, -1);
  S.iamincell.assign(S.idim + 1, 0);
  for (int i = 1; i <= S.ntotal; i++) {                         // :119
    S.ll[i] = -1;
    int icellx[3];
    for (int j = 1; j <= ndim; j++) {
      icellx[j - 1] = (int)((X_(j, i) - xminpart[j - 1]) / S.dxcell) + 1;   // :121
      if (icellx[j - 1] < 0 || icellx[j - 1] > S.ncellsx[j - 1]) return fail(S, ND_ERR_LINK, "link: particle crossed boundary");
    }
    int icell = icellx[0];
    if (ndim >= 2) {
      icell = icell + (icellx[1] - 1) * S.ncellsx[0];
      if (ndim >= 3) icell = icell + (icellx[2] - 1) * S.ncellsx[0] * S.ncellsx[1];
    }
    if (icell < 0 || icell > S.ncells) icell = S.ncells;        // :137-140
    S.ll[i] = S.ifirstincell[icell];                            // :141
    S.ifirstincell[icell] = i;
    S.iamincell[i] = icell;
  }
  return 0;
}

inline int append_chain(St &S, int cell, int *listneigh, int j) {
  // src/get_neighbour_lists.f90:124-153: walk one chain
  if (cell < 1 || cell > S.ncells) return j;   // padded empty cells guarantee this never hits a populated cell
  int ipart = S.ifirstincell[cell];
  if (ipart != -1) {
    while (S.ll[ipart] != -1) { listneigh[j++] = ipart; ipart = S.ll[ipart]; }
    listneigh[j++] = ipart;
  }
  return j;
}

// src/get_neighbour_lists.f90:37-165 (half stencil)
void get_neighbour_list(St &S, int icell, int *listneigh, int &nneigh) {
  if (S.ifirstincell[icell] <= 0) { nneigh = 0; return; }
  int neighcell[27], n = 0;
  const int nx = S.ncellsx[0];
  neighcell[n++] = icell;
  neighcell[n++] = icell + 1;
  if (S.ndim >= 2) {
    neighcell[n++] = icell + nx - 1;
    neighcell[n++] = icell + nx;
    neighcell[n++] = icell + nx + 1;
  }
  if (S.ndim >= 3) {
    const int nxy = S.ncellsx[0] * S.ncellsx[1];
    neighcell[n++] = icell + nxy - 1;
    neighcell[n++] = icell + nxy;
    neighcell[n++] = icell + nxy + 1;
    neighcell[n++] = icell + nxy + nx - 1;
    neighcell[n++] = icell + nxy + nx;
    neighcell[n++] = icell + nxy + nx + 1;
    neighcell[n++] = icell + nxy - nx - 1;
    neighcell[n++] = icell + nxy - nx;
    neighcell[n++] = icell + nxy - nx + 1;
  }
  int j = 0;
  for (int k = 0; k < n; k++) j = append_chain(S, neighcell[k], listneigh, j);
  nneigh = j;
}

// src/get_neighbour_lists.f90:176-316 (full stencil)
void get_neighbour_list_partial(St &S, int icell, int *listneigh, int &nneigh) {
  if (S.ifirstincell[icell] <= 0) { nneigh = 0; return; }
  int neighcell[27], n = 0;
  const int nx = S.ncellsx[0];
  neighcell[n++] = icell - 1;
  neighcell[n++] = icell;
  neighcell[n++] = icell + 1;
  if (S.ndim >= 2) {
    neighcell[n++] = icell + nx - 1;
    neighcell[n++] = icell + nx;
    neighcell[n++] = icell + nx + 1;
    neighcell[n++] = icell - nx - 1;
    neighcell[n++] = icell - nx;
    neighcell[n++] = icell - nx + 1;
    if (S.ndim >= 3) {
      const int nxy = S.ncellsx[0] * S.ncellsx[1];
      neighcell[n++] = icell + nxy + nx - 1;
      neighcell[n++] = icell + nxy + nx;
      neighcell[n++] = icell + nxy + nx + 1;
      neighcell[n++] = icell + nxy - 1;
      neighcell[n++] = icell + nxy;
      neighcell[n++] = icell + nxy + 1;
      neighcell[n++] = icell + nxy - nx - 1;
      neighcell[n++] = icell + nxy - nx;
      neighcell[n++] = icell + nxy - nx + 1;
      neighcell[n++] = icell - nxy + nx - 1;
      neighcell[n++] = icell - nxy + nx;
      neighcell[n++] = icell - nxy + nx + 1;
      neighcell[n++] = icell - nxy - 1;
      neighcell[n++] = icell - nxy;
      neighcell[n++] = icell - nxy + 1;
      neighcell[n++] = icell - nxy - nx - 1;
      neighcell[n++] = icell - nxy - nx;
      neighcell[n++] = icell - nxy - nx + 1;
    }
  }
  int j = 0;
  for (int k = 0; k < n; k++) j = append_chain(S, neighcell[k], listneigh, j);
  nneigh = j;
}

inline bool types_interact(int itypei, int itypej) {
  // src/density_sums.f90:169-172 == src/ratesND_mhd.f90:436-439
  return (itypej == itypei) || (itypei == ND_ITYPE_GAS && itypej == ND_ITYPE_BND) || (itypej == ND_ITYPE_GAS && itypei == ND_ITYPE_BND) ||
         (itypei == ND_ITYPE_DUST && itypej == ND_ITYPE_BNDDUST) || (itypej == ND_ITYPE_DUST && itypei == ND_ITYPE_BNDDUST);
}

// src/density_sums.f90:38-379 density (symmetric pair visit).  imhd=5, iprterm 10/12, one-fluid dust, gravity and the
// dxdx matrix are outside the supported tuple and omitted.
void density(St &S, double *rho, double *drhodt, double *densn, double *dndt, double *delsqn, double *gradh, double *gradhn,
             double *gradsoft, double *gradgradh, std::vector<int> &listneigh, std::vector<double> &h1) {
  const nd_options &o = *S.o;
  const int ndim = S.ndim, npart = S.npart, ntotal = S.ntotal;
  const Kern &K = S.K;
  for (int i = 1; i <= ntotal; i++) A1(numneigh, i) = 0;        // :94 (whole array)
  double dwdhi = 0., dwdhj = 0., dwaltdhi = 0., dwaltdhj = 0., dwdhdhi = 0., dwdhdhj = 0.;
  double dr[3] = {0., 0., 0.};
  for (int i = 1; i <= npart; i++) {                            // :105-128
    if (A1(itype, i) != ND_ITYPE_BND && A1(itype, i) != ND_ITYPE_BNDDUST) {
      rho[i - 1] = 0.; drhodt[i - 1] = 0.; densn[i - 1] = 0.; dndt[i - 1] = 0.; delsqn[i - 1] = 0.;
      gradh[i - 1] = 0.; gradhn[i - 1] = 0.; gradsoft[i - 1] = 0.; gradgradh[i - 1] = 0.;
      if (o.onef_dust) { A1(rhogas, i) = 0.; A1(rhodust, i) = 0.; }   // :123-126
    }
  }
  for (int i = 1; i <= ntotal; i++) h1[i] = 1. / A1(hh, i);    // :129-131
  int nneigh = 0;
  for (int icell = 1; icell <= S.ncellsloop; icell++) {         // :135
    get_neighbour_list(S, icell, listneigh.data(), nneigh);
    int i = S.ifirstincell[icell];
    int idone = -1;
    while (i != -1) {                                           // :148
      idone = idone + 1;
      double pmassi = A1(pmass, i);
      double dustfraci = o.onef_dust ? A1(dustfrac, i) : 0.;    // :153 (ndust = 1)
      double xi[3] = {0, 0, 0}, veli[3];
      for (int k = 1; k <= ndim; k++) xi[k - 1] = X_(k, i);
      for (int k = 1; k <= 3; k++) veli[k - 1] = V3(vel, k, i);
      double hi1 = h1[i];
      double hfacwabi = powi(hi1, ndim);
      double hi21 = hi1 * hi1;
      int itypei = A1(itype, i);
      for (int n = idone + 1; n <= nneigh; n++) {               // :165
        int j = listneigh[n - 1];
        int itypej = A1(itype, j);
        if (!types_interact(itypei, itypej)) continue;          // :169-174
        double dx[3] = {0, 0, 0};
        for (int k = 1; k <= ndim; k++) dx[k - 1] = xi[k - 1] - X_(k, j);
        double hj1 = h1[j];
        double rij2 = 0.;
        for (int k = 0; k < ndim; k++) rij2 = rij2 + dx[k] * dx[k];   // dot_product
        double q2i = rij2 * hi21;
        double q2j = rij2 * hj1 * hj1;
        // :189-190 -- .AND. binds tighter than .OR.
        if ((q2i < K.radkern2) || ((q2j < K.radkern2) && (i <= npart || j <= npart))) {
          if (i <= npart) A1(numneigh, i) = A1(numneigh, i) + 1;
          if (j <= npart && j != i) A1(numneigh, j) = A1(numneigh, j) + 1;
          double rij = std::sqrt(rij2);
          for (int k = 0; k < ndim; k++) dr[k] = dx[k] / (rij + DBL_EPSILON);   // :199
          double hfacwabj = powi(hj1, ndim);
          double weight = (j == i) ? 0.5 : 1.0;
          double pmassj = A1(pmass, j);
          double wabi, grkerni, grgrkerni, wabalti, grkernalti, wabj, grkernj, grgrkernj, wabaltj, grkernaltj;
          interpolate_kernels_dens(K, q2i, wabi, grkerni, grgrkerni, wabalti, grkernalti);   // :234-235
          interpolate_kernels_dens(K, q2j, wabj, grkernj, grgrkernj, wabaltj, grkernaltj);
          wabi = wabi * hfacwabi;
          wabalti = wabalti * hfacwabi;
          grkerni = grkerni * hfacwabi * hi1;
          grgrkerni = grgrkerni * hfacwabi * hi1 * hi1;
          grkernalti = grkernalti * hfacwabi * hi1;
          wabj = wabj * hfacwabj;
          wabaltj = wabaltj * hfacwabj;
          grkernj = grkernj * hfacwabj * hj1;
          grgrkernj = grgrkernj * hfacwabj * hj1 * hj1;
          grkernaltj = grkernaltj * hfacwabj * hj1;
          // :260-268
          dwdhi = -rij * grkerni * hi1 - ndim * wabi * hi1;
          dwdhj = -rij * grkernj * hj1 - ndim * wabj * hj1;
          dwaltdhi = -rij * grkernalti * hi1 - ndim * wabalti * hi1;
          dwaltdhj = -rij * grkernaltj * hj1 - ndim * wabaltj * hj1;
          dwdhdhi = ndim * (ndim + 1) * wabi * (hi1 * hi1) + 2. * (ndim + 1) * rij * (hi1 * hi1) * grkerni + (rij * rij) * (hi1 * hi1) * grgrkerni;
          dwdhdhj = ndim * (ndim + 1) * wabj * (hj1 * hj1) + 2. * (ndim + 1) * rij * (hj1 * hj1) * grkernj + (rij * rij) * (hj1 * hj1) * grgrkernj;
          if (itypei != ND_ITYPE_BND) {                         // :273-283
            rho[i - 1] = rho[i - 1] + pmassj * wabi * weight;
            densn[i - 1] = densn[i - 1] + wabalti * weight;
            if (o.onef_dust) {                                  // :278-282
              double dustfracj = A1(dustfrac, j);
              A1(rhodust, i) = A1(rhodust, i) + pmassj * A1(dustfrac, j) * wabi * weight;
              A1(rhogas, i) = A1(rhogas, i) + pmassj * (1. - dustfracj) * wabi * weight;
            }
          }
          if (itypej != ND_ITYPE_BND) {                         // :285-293
            rho[j - 1] = rho[j - 1] + pmassi * wabj * weight;
            densn[j - 1] = densn[j - 1] + wabaltj * weight;
            if (o.onef_dust) {                                  // :289-292
              A1(rhodust, j) = A1(rhodust, j) + pmassi * A1(dustfrac, i) * wabj * weight;
              A1(rhogas, j) = A1(rhogas, j) + pmassi * (1. - dustfraci) * wabj * weight;
            }
          }
          if (i != j) {                                         // :297-303
            double dvel[3], dvdotr = 0.;
            for (int k = 1; k <= 3; k++) dvel[k - 1] = veli[k - 1] - V3(vel, k, j);
            for (int k = 0; k < 3; k++) dvdotr = dvdotr + dvel[k] * dr[k];
            drhodt[i - 1] = drhodt[i - 1] + pmassj * dvdotr * grkerni;
            drhodt[j - 1] = drhodt[j - 1] + pmassi * dvdotr * grkernj;
            dndt[i - 1] = dndt[i - 1] + dvdotr * grkernalti;
            dndt[j - 1] = dndt[j - 1] + dvdotr * grkernaltj;
          }
          if (o.ikernav == 3) {                                 // :314-331
            if (itypei != ND_ITYPE_BND) {
              gradh[i - 1] = gradh[i - 1] + weight * pmassj * dwdhi;
              gradhn[i - 1] = gradhn[i - 1] + weight * dwaltdhi;
              gradgradh[i - 1] = gradgradh[i - 1] + weight * pmassj * dwdhdhi;
            }
            if (itypej != ND_ITYPE_BND) {
              gradh[j - 1] = gradh[j - 1] + weight * pmassi * dwdhj;
              gradhn[j - 1] = gradhn[j - 1] + weight * dwaltdhj;
              gradgradh[j - 1] = gradgradh[j - 1] + weight * pmassi * dwdhdhj;
            }
          }
        }
      }
      i = S.ll[i];                                              // :362-363
    }
  }
}

// src/density_sums.f90:396-658 density_partial (gather on a list, h_i only)
void density_partial(St &S, double *rho, double *drhodt, double *densn, double *dndt, double *delsqn, double *gradh, double *gradhn,
                     double *gradsoft, double *gradgradh, int nlist, const int *ipartlist, std::vector<int> &listneigh) {
  const int ndim = S.ndim;
  const Kern &K = S.K;
  double dr[3] = {0., 0., 0.};
  for (int ipart = 1; ipart <= nlist; ipart++) {                // :455-475
    int i = ipartlist[ipart - 1];
    rho[i - 1] = 0.; drhodt[i - 1] = 0.; densn[i - 1] = 0.; dndt[i - 1] = 0.; delsqn[i - 1] = 0.;
    gradh[i - 1] = 0.; gradhn[i - 1] = 0.; gradsoft[i - 1] = 0.; gradgradh[i - 1] = 0.;
    A1(numneigh, i) = 0;
    if (S.o->onef_dust) { A1(rhodust, i) = 0.; A1(rhogas, i) = 0.; }   // :470-473
  }
  int icellprev = 0, nneigh = 0;
  for (int ipart = 1; ipart <= nlist; ipart++) {                // :480
    int i = ipartlist[ipart - 1];
    int icell = S.iamincell[i];
    if (icell != icellprev) get_neighbour_list_partial(S, icell, listneigh.data(), nneigh);
    icellprev = icell;
    double hi = A1(hh, i);
    double hi1 = 1. / hi;
    double hi21 = hi1 * hi1;
    double hfacwabi = powi(hi1, ndim);
    double hfacgrkerni = hfacwabi * hi1;
    double xi[3] = {0, 0, 0}, veli[3];
    for (int k = 1; k <= ndim; k++) xi[k - 1] = X_(k, i);
    for (int k = 1; k <= 3; k++) veli[k - 1] = V3(vel, k, i);
    int itypei = A1(itype, i);
    if (itypei == ND_ITYPE_BND) continue;                       // :510
    for (int n = 1; n <= nneigh; n++) {                         // :514
      int j = listneigh[n - 1];
      if (A1(itype, j) != itypei && A1(itype, j) != ND_ITYPE_BND) continue;   // :517
      double dx[3] = {0, 0, 0};
      for (int k = 1; k <= ndim; k++) dx[k - 1] = xi[k - 1] - X_(k, j);
      double rij2 = 0.;
      for (int k = 0; k < ndim; k++) rij2 = rij2 + dx[k] * dx[k];
      double q2i = rij2 * hi21;
      if (q2i < K.radkern2) {                                   // :528
        double rij = std::sqrt(rij2);
        for (int k = 0; k < ndim; k++) dr[k] = dx[k] / (rij + DBL_EPSILON);
        A1(numneigh, i) = A1(numneigh, i) + 1;
        double pmassj = A1(pmass, j);
        double wabi, grkerni, grgrkerni, wabalti, grkernalti;
        interpolate_kernels_dens(K, q2i, wabi, grkerni, grgrkerni, wabalti, grkernalti);   // :547
        wabi = wabi * hfacwabi;
        wabalti = wabalti * hfacwabi;
        grkerni = grkerni * hfacgrkerni;
        grgrkerni = grgrkerni * hfacwabi * hi1 * hi1;
        grkernalti = grkernalti * hfacgrkerni;
        double dwdhi = -rij * grkerni * hi1 - ndim * wabi * hi1;                           // :557-561
        double dwaltdhi = -rij * grkernalti * hi1 - ndim * wabalti * hi1;
        double dwdhdhi = ndim * (ndim + 1) * wabi * (hi1 * hi1) + 2. * (ndim + 1) * rij * (hi1 * hi1) * grkerni + (rij * rij) * (hi1 * hi1) * grgrkerni;
        rho[i - 1] = rho[i - 1] + pmassj * wabi;                // :566-568
        densn[i - 1] = densn[i - 1] + wabalti;
        delsqn[i - 1] = delsqn[i - 1] + grgrkerni;
        if (S.o->onef_dust) {                                   // :569-573
          double dustfracj = A1(dustfrac, j);
          A1(rhodust, i) = A1(rhodust, i) + pmassj * A1(dustfrac, j) * wabi;
          A1(rhogas, i) = A1(rhogas, i) + pmassj * (1. - dustfracj) * wabi;
        }
        if (i != j) {                                           // :577-581
          double dvel[3], dvdotr = 0.;
          for (int k = 1; k <= 3; k++) dvel[k - 1] = veli[k - 1] - V3(vel, k, j);
          for (int k = 0; k < 3; k++) dvdotr = dvdotr + dvel[k] * dr[k];
          drhodt[i - 1] = drhodt[i - 1] + pmassj * dvdotr * grkerni;
          dndt[i - 1] = dndt[i - 1] + dvdotr * grkernalti;
        }
        gradh[i - 1] = gradh[i - 1] + pmassj * dwdhi;           // :594-596
        gradhn[i - 1] = gradhn[i - 1] + dwaltdhi;
        gradgradh[i - 1] = gradgradh[i - 1] + pmassj * dwdhdhi;
      }
    }
  }
}

// src/iterate_density.f90:43-360
int iterate_density(St &S) {
  const nd_options &o = *S.o;
  const int ndim = S.ndim, npart = S.npart;
  const double dndim = 1. / double(ndim);
  const double h_min = 0., rhomin = 0.;                         // initialiseND_mhd.f90:266-274, iterate_density.f90:113-117
  int itsdensitymax = ((o.ikernav == 3) && (o.ihvar != 0)) ? o.maxdensits : 0;   // :77-81
  S.itsdensity = 0;
  long long ncalctotal = 0;
  int ncalc = npart;
  bool redolink = false;
  for (int i = 1; i <= S.idim; i++)                             // :90-97 (whole arrays)
    if (A1(itype, i) != ND_ITYPE_BND) { A1(gradh, i) = 0.; A1(gradhn, i) = 0.; A1(gradsoft, i) = 0.; A1(gradgradh, i) = 0.; A1(drhodt, i) = 0.; A1(dhdt, i) = 0.; }
  std::vector<double> hhin(npart + 1), dndt(2 * (size_t)S.idim + 1, 0.), delsqn(2 * (size_t)S.idim + 1, 0.);
  for (int i = 1; i <= npart; i++) hhin[i] = A1(hh, i);
  for (int i = 1; i <= npart; i++) if (A1(hh, i) <= DBL_MIN) return fail(S, ND_ERR_H_NONPOSITIVE, "error: h <= 0 in density call");
  double dhdrhoi = 0.;
  std::vector<int> redolist(npart + 1), redolistprev(npart + 1), listneigh(S.idim + 1);
  std::vector<double> h1(S.idim + 1);
  for (int j = 1; j <= npart; j++) redolist[j - 1] = j;
  int ncalcprev = 0;
  while (ncalc > 0 && S.itsdensity <= itsdensitymax) {          // :119
    S.itsdensity = S.itsdensity + 1;
    if (redolink) {                                             // :122-126
      bool anygt1 = false;
      for (int d = 0; d < ndim; d++) if (o.ibound[d] > 1) anygt1 = true;
      if (anygt1) { if (int e = set_ghost_particles(S)) return e; }
      if (int e = set_linklist(S)) return e;
      if ((int)listneigh.size() < S.idim + 1) listneigh.resize(S.idim + 1);
    }
    if (ncalc == npart)                                         // :131-135
      density(S, S.a.rho, S.a.drhodt, S.a.rhoalt, dndt.data(), delsqn.data(), S.a.gradh, S.a.gradhn, S.a.gradsoft, S.a.gradgradh, listneigh, h1);
    else
      density_partial(S, S.a.rho, S.a.drhodt, S.a.rhoalt, dndt.data(), delsqn.data(), S.a.gradh, S.a.gradhn, S.a.gradsoft, S.a.gradgradh, ncalc, redolist.data(), listneigh);
    ncalctotal += ncalc;
    ncalcprev = ncalc;
    for (int j = 0; j < ncalcprev; j++) redolistprev[j] = redolist[j];
    ncalc = 0;
    redolink = false;
    if (o.ikernav == 3) {
      for (int j = 1; j <= ncalcprev; j++) {                    // :163
        int i = redolistprev[j - 1];
        if (A1(itype, i) != ND_ITYPE_BND && A1(itype, i) != ND_ITYPE_BNDDUST) {
          if (A1(rho, i) <= 1.e-6) {
            if (A1(rho, i) <= 0.) return fail(S, ND_ERR_RHO_NONPOSITIVE, "error: rho <= 0 in iterate_density");
          }
          double hhi = A1(hh, i);
          double rhoi = A1(pmass, i) / powi((hhi - h_min) / o.hfact, ndim) - rhomin;   // :192
          dhdrhoi = -(hhi - h_min) / (ndim * (A1(rho, i) + rhomin));                    // :193
          double dwdhsumi = A1(gradh, i);
          double omegai = 1. - dhdrhoi * A1(gradh, i);                                  // :196
          if (omegai < 1.e-5) { if (std::fabs(omegai) == 0.) omegai = 1.; }
          A1(gradh, i) = 1. / omegai;                                                   // :201
          double func = rhoi - A1(rho, i);
          double dfdh = omegai / dhdrhoi;
          A1(gradsoft, i) = A1(gradsoft, i) * dhdrhoi;
          double d2hdrho2i = hhi * (ndim + 1) / ((A1(rho, i) * ndim) * (A1(rho, i) * ndim));   // :206
          A1(gradgradh, i) = A1(rho, i) * (d2hdrho2i * dwdhsumi + (dhdrhoi * dhdrhoi) * A1(gradgradh, i));
          double hnew = hhi - func / dfdh;                                              // :212
          if (hnew > 1.2 * hhi) hnew = 1.2 * hhi;
          else if (hnew < 0.8 * hhi) hnew = 0.8 * hhi;
          if (hnew <= 0 || A1(gradh, i) <= DBL_MIN) hnew = o.hfact * std::pow(A1(pmass, i) / (A1(rho, i) + rhomin), dndim);   // :225-227
          else if (S.itsdensity > 100) hnew = o.hfact * std::pow(A1(pmass, i) / (A1(rho, i) + rhomin), dndim);              // :228-229
          if (A1(numneigh, i) <= 1) { hnew = hhi + o.psep; redolink = true; }           // :231-237
          bool converged = (std::fabs((hnew - hhi) / hhin[i]) < o.tolh && omegai > 0.) || itsdensitymax == 0;   // :242-243
          if (!converged) {
            ncalc = ncalc + 1;
            redolist[ncalc - 1] = i;
            if (S.itsdensity <= itsdensitymax && A1(itype, i) != ND_ITYPE_BND) A1(hh, i) = hnew;   // :253-255
            if (hnew > S.hhmax) redolink = true;                                        // :260-262
          } else {
            if (o.ikernav == 3) A1(drhodt, i) = A1(drhodt, i) * A1(gradh, i);           // :272-273
            A1(dhdt, i) = dhdrhoi * A1(drhodt, i);
          }
        }
      }
    } else if (o.ihvar > 0) {                                   // :289-293
      for (int i = 1; i <= npart; i++) { dhdrhoi = -A1(hh, i) / (ndim * (A1(rho, i) + rhomin)); A1(dhdt, i) = dhdrhoi * A1(drhodt, i); }
    } else {                                                    // :296-306
      for (int i = 1; i <= npart; i++) { A1(dhdt, i) = 0.; A1(gradh, i) = 1.; A1(gradhn, i) = 0.; A1(gradgradh, i) = 0.; }
    }
    bool anyeq1 = false, anygt1 = false;
    for (int d = 0; d < ndim; d++) { if (o.ibound[d] == 1) anyeq1 = true; if (o.ibound[d] > 1) anygt1 = true; }
    if (anyeq1) {                                               // :310-329
      for (int i = 1; i <= npart; i++) if (A1(itype, i) == ND_ITYPE_BND) {
        int j = A1(ireal, i);
        if (j > 0) {
          A1(rho, i) = A1(rho, j); A1(rhoalt, i) = A1(rhoalt, j); A1(drhodt, i) = A1(drhodt, j); A1(dhdt, i) = A1(dhdt, j);
          A1(hh, i) = A1(hh, j); A1(gradh, i) = A1(gradh, j); A1(gradhn, i) = A1(gradhn, j); A1(gradsoft, i) = A1(gradsoft, j);
        }
      }
    }
    if (anygt1) {                                               // :330-344
      for (int i = npart + 1; i <= S.ntotal; i++) {
        int j = A1(ireal, i);
        if (j > 0) {
          A1(rho, i) = A1(rho, j); A1(rhoalt, i) = A1(rhoalt, j); A1(drhodt, i) = A1(drhodt, j); A1(dhdt, i) = A1(dhdt, j);
          A1(hh, i) = A1(hh, j); A1(gradh, i) = A1(gradh, j); A1(gradhn, i) = A1(gradhn, j); A1(gradsoft, i) = A1(gradsoft, j);
        }
      }
    }
  }
  S.ncalctotal = ncalctotal;
  if (S.itsdensity > itsdensitymax && itsdensitymax > 0) return fail(S, ND_ERR_DENSITY_NOT_CONVERGED, "ERROR: DENSITY NOT CONVERGED");   // :349-351
  return 0;
}

// src/eos.f90:40-105 equation_of_state on rows 1..npart
void equation_of_state(St &S, const double *rho /*0-based*/) {
  const nd_options &o = *S.o;
  const double gamma1 = o.gamma - 1.;
  for (int i = 1; i <= S.npart; i++) {
    double r = rho[i - 1];
    if (o.iener == 0) {                                         // :72-89 polytropic
      int t = A1(itype, i);
      if (r > 0. && t == ND_ITYPE_GAS) { A1(pr, i) = o.polyk * std::pow(r, o.gamma); A1(spsound, i) = std::sqrt(o.gamma * A1(pr, i) / r); }
      else if (t == ND_ITYPE_GAS1) A1(pr, i) = o.polyk * (r - 1.);
      else if (t == ND_ITYPE_GAS2) A1(pr, i) = o.polyk * (r - 1.);
      else if (t != ND_ITYPE_BND) A1(pr, i) = 0.;
      if (std::fabs(gamma1) > 1.e-3) { if (r > 0.) A1(uu, i) = A1(pr, i) / (gamma1 * r); }
    } else {                                                    // :96-101 adiabatic
      if (r > 0.) { A1(pr, i) = gamma1 * A1(uu, i) * r; A1(spsound, i) = std::sqrt(o.gamma * A1(pr, i) / r); }
    }
  }
}

// src/conservative2primitive.f90:42-470, element-wise branches only
inline void cross_product3D(const double *a, const double *b, double *c) {   // src/utils.f90:60-68
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
inline double dot3(const double *a, const double *b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

// src/kernelND.f90:4643-4672 interpolate_kernel_curl
inline void interpolate_kernel_curl(const Kern &K, double q2, double &gradwalt, double &gradgradwalt) {
  int index, index1; kindex(K, q2, index, index1);
  double dxx = q2 - index * K.dq2table;
  gradwalt = K.grwijalt[index];
  double dgrwaltdx = (K.grwijalt[index1] - gradwalt) * K.ddq2table;
  gradwalt = gradwalt + dgrwaltdx * dxx;
  gradgradwalt = K.grgrwijalt[index];
  double dgrgrwaltdx = (K.grgrwijalt[index1] - gradgradwalt) * K.ddq2table;
  gradgradwalt = gradgradwalt + dgrgrwaltdx * dxx;
}

// src/get_curl.f90:64-287 get_curl, icurltype 1 (default, optional gradB) .. 4; the optional curlBgradh (vector potential) is not restated.
// Bvec, curlB: (3, idim) 1-based rows like the module arrays; gradB: (3,3,idim) column-major gradB(:,k,i) or NULL.
// Reads whatever the rows npart+1..ntotal of Bvec hold, as the reference does (the caller decides what the ghosts carry).
void get_curl(St &S, int icurltype, const double *Bvec, double *curlB, double *gradB) {
  const nd_options &o = *S.o;
  const Kern &K = S.K;
  const int ndim = S.ndim, npart = S.npart, ntotal = S.ntotal;
  auto BV = [&](int k, int i) { return Bvec[(size_t)(i - 1) * 3 + (k - 1)]; };
  auto CB = [&](int k, int i) -> double & { return curlB[(size_t)(i - 1) * 3 + (k - 1)]; };
  auto GB = [&](int l, int k, int i) -> double & { return gradB[((size_t)(i - 1) * 3 + (k - 1)) * 3 + (l - 1)]; };
  std::vector<int> listneigh(S.idim + 1);
  std::vector<double> h1(ntotal + 1);
  for (int i = 1; i <= S.idim; i++) for (int k = 1; k <= 3; k++) CB(k, i) = 0.;     // :113
  double dr[3] = {0., 0., 0.};
  const double weight = 1. / powi(o.hfact, ndim);                                   // :115
  for (int i = 1; i <= ntotal; i++) h1[i] = 1. / A1(hh, i);
  if (gradB) for (size_t q = 0; q < (size_t)S.idim * 9; q++) gradB[q] = 0.;
  int nneigh = 0;
  for (int icell = 1; icell <= S.ncellsloop; icell++) {                             // :124
    get_neighbour_list(S, icell, listneigh.data(), nneigh);
    int i = S.ifirstincell[icell];
    int idone = -1;
    while (i != -1) {                                                               // :137
      idone = idone + 1;
      double hi1 = h1[i];
      double hi21 = hi1 * hi1;
      double hfacwabi = powi(hi1, ndim);
      double pmassi = A1(pmass, i);
      double Bi[3] = {BV(1, i), BV(2, i), BV(3, i)};
      double rho21i = 1. / (A1(rho, i) * A1(rho, i));
      double rho21gradhi = rho21i * A1(gradh, i);
      double curlBi[3] = {0., 0., 0.}, gradBi[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};   // gradBi[k][l] = gradBi(l,k)
      int itypei = A1(itype, i);
      for (int n = idone + 1; n <= nneigh; n++) {                                   // :158
        int j = listneigh[n - 1];
        int itypej = A1(itype, j);
        if (!types_interact(itypei, itypej)) continue;                              // :161-167
        if (!((j != i) && !(j > npart && i > npart))) continue;                     // :168
        double dx[3] = {0., 0., 0.};
        for (int k = 1; k <= ndim; k++) dx[k - 1] = X_(k, i) - X_(k, j);
        double hj1 = h1[j];
        double rij2 = 0.;
        for (int k = 0; k < ndim; k++) rij2 = rij2 + dx[k] * dx[k];
        double q2i = rij2 * hi21;
        double q2j = rij2 * hj1 * hj1;
        if (!((q2i < K.radkern2) || (q2j < K.radkern2))) continue;                  // :183
        double hfacwabj = powi(hj1, ndim);
        double rij = std::sqrt(rij2);
        for (int k = 0; k < ndim; k++) dr[k] = dx[k] / (rij + DBL_MIN);             // :187 tiny(rij)
        double grkerni, grgrkerni, grkernj, grgrkernj;
        interpolate_kernel_curl(K, q2i, grkerni, grgrkerni);
        interpolate_kernel_curl(K, q2j, grkernj, grgrkernj);
        double dB[3], curlBterm[3], Bj[3] = {BV(1, j), BV(2, j), BV(3, j)};
        switch (icurltype) {
          case 2: {                                                                 // :202-211
            grkerni = grkerni * hfacwabi * hi1;
            grkernj = grkernj * hfacwabj * hj1 * A1(gradh, j);
            double ti_[3], tj_[3];
            cross_product3D(Bi, dr, ti_);
            cross_product3D(Bj, dr, tj_);
            for (int k = 0; k < 3; k++) curlBterm[k] = (ti_[k] * rho21gradhi * grkerni + tj_[k] / (A1(rho, j) * A1(rho, j)) * grkernj);
            for (int k = 0; k < 3; k++) curlBi[k] = curlBi[k] + A1(pmass, j) * curlBterm[k];
            for (int k = 0; k < 3; k++) CB(k + 1, j) = CB(k + 1, j) - pmassi * curlBterm[k];
          } break;
          case 3:                                                                   // :213-220
            grkerni = grkerni * hi1;
            grkernj = grkernj * hj1;
            for (int k = 0; k < 3; k++) dB[k] = Bi[k] - Bj[k];
            cross_product3D(dB, dr, curlBterm);
            for (int k = 0; k < 3; k++) curlBi[k] = curlBi[k] + curlBterm[k] * grkerni;
            for (int k = 0; k < 3; k++) CB(k + 1, j) = CB(k + 1, j) + curlBterm[k] * grkernj;
            break;
          case 4:                                                                   // :222-230
            grkerni = grkerni * hfacwabi * hi1;
            grkernj = grkernj * hfacwabj * hj1;
            for (int k = 0; k < 3; k++) dB[k] = Bi[k] - Bj[k];
            cross_product3D(dB, dr, curlBterm);
            for (int k = 0; k < 3; k++) curlBi[k] = curlBi[k] + A1(pmass, j) / (A1(rho, j) * A1(rho, j)) * curlBterm[k] * grkerni;
            for (int k = 0; k < 3; k++) CB(k + 1, j) = CB(k + 1, j) + pmassi * rho21i * curlBterm[k] * grkernj;
            break;
          default:                                                                  // :232-255
            grkerni = grkerni * hfacwabi * hi1;
            grkernj = grkernj * hfacwabj * hj1;
            for (int k = 0; k < 3; k++) dB[k] = Bi[k] - Bj[k];
            cross_product3D(dB, dr, curlBterm);
            for (int k = 0; k < 3; k++) curlBi[k] = curlBi[k] + A1(pmass, j) * curlBterm[k] * grkerni;
            for (int k = 0; k < 3; k++) CB(k + 1, j) = CB(k + 1, j) + pmassi * curlBterm[k] * grkernj;
            if (gradB) {
              for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) {
                gradBi[k][l] = gradBi[k][l] + A1(pmass, j) * dB[k] * dr[l] * grkerni;
                GB(l + 1, k + 1, j) = GB(l + 1, k + 1, j) + pmassi * dB[k] * dr[l] * grkernj;
              }
            }
        }
      }
      for (int k = 0; k < 3; k++) CB(k + 1, i) = CB(k + 1, i) + curlBi[k];          // :262
      if (gradB) for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) GB(l + 1, k + 1, i) = GB(l + 1, k + 1, i) + gradBi[k][l];
      i = S.ll[i];
    }
  }
  for (int i = 1; i <= npart; i++) {                                                // :275-290
    switch (icurltype) {
      case 4: for (int k = 1; k <= 3; k++) CB(k, i) = A1(rho, i) * CB(k, i); break;
      case 3: for (int k = 1; k <= 3; k++) CB(k, i) = weight * CB(k, i); break;
      case 2: for (int k = 1; k <= 3; k++) CB(k, i) = -A1(rho, i) * CB(k, i); break;
      default:
        for (int k = 1; k <= 3; k++) CB(k, i) = CB(k, i) * A1(gradh, i) / A1(rho, i);
        if (gradB) for (int k = 1; k <= 3; k++) for (int l = 1; l <= 3; l++) GB(l, k, i) = -GB(l, k, i) * A1(gradh, i) / A1(rho, i);
    }
  }
}

int conservative2primitive(St &S) {
  const nd_options &o = *S.o;
  const int npart = S.npart, ntotal = S.ntotal;
  for (int i = 1; i <= S.idim; i++) A1(sqrtg, i) = 1.;          // :72
  if (o.onef_dust) {                                            // :76-113 (idustevol = 0)
    for (int i = 1; i <= npart; i++) {
      A1(dustfrac, i) = A1(dustevol, i);
      if (A1(dustfrac, i) > 1.) { A1(dustfrac, i) = 1.; A1(dustevol, i) = 1.; }   // :102-108 (dustevolin is integrator state)
    }
    for (int i = 1; i <= npart; i++) A1(dens, i) = A1(rho, i) * (1. - A1(dustfrac, i));   // :112: dens is the GAS density
  } else
    for (int i = 1; i <= S.idim; i++) A1(dens, i) = A1(rho, i); // :117
  if (o.imhd >= 11) {                                           // :151-152
    for (int i = 1; i <= S.idim; i++) for (int k = 1; k <= 3; k++) V3(Bfield, k, i) = V3(Bevol, k, i);
  } else if (o.imhd >= 1 && o.imhd <= 9) {                      // :190-193
    for (int i = 1; i <= npart; i++) for (int k = 1; k <= 3; k++) V3(Bfield, k, i) = V3(Bevol, k, i) * A1(rho, i);
  } else if (o.imhd != 0) return fail(S, ND_ERR_UNSUPPORTED_OPTION, "c2p: imhd not supported");
  if (o.imhd != 0 && o.iavlim[2] == 2) {                        // :299-323 (iambipolar = 0, iresist /= 4: only the switch branch)
    // J and grad(B) by the standard (differenced) curl operator, then the Tricco & Price (2013) resistivity switch.  The ghost rows
    // of Bfield are whatever they hold at this point: Bevol's for imhd >= 11 (whole-array assignment above), the copies made when the
    // ghosts were created for imhd 1..9 (the loop above stops at npart) -- the ghost copies of :441-467 come later.
    S.gradB.assign((size_t)S.idim * 9, 0.);
    get_curl(S, 1, S.a.Bfield, S.a.curlB, S.gradB.data());
    for (int i = 1; i <= npart; i++) {
      double B2i = 0.;
      for (int k = 1; k <= 3; k++) B2i = B2i + V3(Bfield, k, i) * V3(Bfield, k, i);
      if (B2i > 1.e-8) {
        double g2 = 0.;
        for (int q = 0; q < 9; q++) g2 = g2 + S.gradB[(size_t)(i - 1) * 9 + q] * S.gradB[(size_t)(i - 1) * 9 + q];   // norm2 of the 3x3 block
        V3(alpha, 3, i) = std::min(A1(hh, i) * std::sqrt(g2) / std::sqrt(B2i + DBL_EPSILON), 1.0);
      } else V3(alpha, 3, i) = 0.;
    }
  }
  if (o.iener == 3) {                                           // :329-346
    for (int i = 1; i <= npart; i++) {
      double v2i = 0., B2i = 0.;
      for (int k = 1; k <= 3; k++) v2i = v2i + V3(vel, k, i) * V3(vel, k, i);
      for (int k = 1; k <= 3; k++) B2i = B2i + V3(Bfield, k, i) * V3(Bfield, k, i);
      B2i = B2i / A1(rho, i);
      A1(uu, i) = A1(en, i) - 0.5 * v2i - 0.5 * B2i;
      if (A1(uu, i) < 0.) A1(uu, i) = 0.;
    }
  } else if (o.iener == 1) {                                    // :348
    for (int i = 1; i <= S.idim; i++) A1(uu, i) = A1(en, i) / (o.gamma - 1.) * std::pow(A1(rho, i), o.gamma - 1.);
  } else if (o.iener == 4) {
    for (int i = 1; i <= S.idim; i++) A1(uu, i) = A1(en, i) / A1(rho, i);
  } else {                                                      // :353-368
    for (int i = 1; i <= S.idim; i++) A1(uu, i) = A1(en, i);
  }
  equation_of_state(S, o.onef_dust ? S.a.dens : S.a.rho);       // :418-424
  bool anyeq1 = false, anygt1 = false, all3 = true;
  for (int d = 0; d < S.ndim; d++) { if (o.ibound[d] == 1) anyeq1 = true; if (o.ibound[d] > 1) anygt1 = true; if (o.ibound[d] != 3) all3 = false; }
  if (anyeq1) {                                                 // :424-437
    for (int i = 1; i <= npart; i++) if (A1(itype, i) == ND_ITYPE_BND || A1(itype, i) == ND_ITYPE_BNDDUST) {
      int j = A1(ireal, i);
      if (j > 0) copy_particle(S, i, j);
    }
  }
  if (anygt1) {                                                 // :441-467
    for (int i = npart + 1; i <= ntotal; i++) {
      int j = A1(ireal, i);
      for (int k = 1; k <= 3; k++) V3(curlB, k, i) = V3(curlB, k, j);
      A1(psi, i) = A1(psi, j);
      A1(dens, i) = A1(dens, j);
      A1(uu, i) = A1(uu, j); A1(spsound, i) = A1(spsound, j); A1(pr, i) = A1(pr, j);
      for (int k = 1; k <= 3; k++) V3(Bfield, k, i) = V3(Bfield, k, j);
      if (o.onef_dust) A1(dustfrac, i) = A1(dustfrac, j);       // :464
      if (all3) copy_particle(S, i, j);
    }
  }
  return 0;
}

// src/dust.f90:77-102 get_tstop (constant drag / constant ts only)
inline double get_tstop(int idrag_nature, double rhogas, double rhodust, double /*cs*/, double Kdrag) {
  double rho = rhogas + rhodust;
  switch (idrag_nature) {
    case 1: return rhodust * rhogas / (Kdrag * rho);
    case 2: case 4: return Kdrag;
    default: return DBL_MAX;
  }
}


// src/ratesND_mhd.f90:29-979 get_rates with rates_core (:1175), artificial_dissipation (:1700),
// artificial_dissipation_phantom (:1903), mhd_terms (:2377, default tensor force), drag_forces (:1074).
int get_rates(St &S, std::vector<int> *pairs_i, std::vector<int> *pairs_j) {
  const nd_options &o = *S.o;
  const int ndim = S.ndim, npart = S.npart, ntotal = S.ntotal;
  const Kern &K = S.K;
  if (o.ikernav != 3 && o.ikernav != 2) return fail(S, ND_ERR_UNSUPPORTED_OPTION, "rates: ikernav");
  if (o.iprterm != 0) return fail(S, ND_ERR_UNSUPPORTED_OPTION, "rates: iprterm");
  if (o.imhd != 0 && !(o.imhd == 1 || o.imhd == 11)) return fail(S, ND_ERR_UNSUPPORTED_OPTION, "rates: imhd");
  if (o.imhd != 0 && o.imagforce != 2) return fail(S, ND_ERR_UNSUPPORTED_OPTION, "rates: imagforce");
  if (o.iav < 0 || o.iav > 3) return fail(S, ND_ERR_UNSUPPORTED_OPTION, "rates: iav");
  if (!(o.iener == 0 || o.iener == 2 || o.iener == 3)) return fail(S, ND_ERR_UNSUPPORTED_OPTION, "rates: iener");
  if (!(o.idust == 0 || o.idust == 1 || o.idust == 2)) return fail(S, ND_ERR_UNSUPPORTED_OPTION, "rates: idust");
  if ((o.idust == 1) != (o.onef_dust != 0)) return fail(S, ND_ERR_UNSUPPORTED_OPTION, "rates: onef_dust must be set with idust=1 only");
  if (o.idust == 1 && (o.idustevol != 0 || o.iener == 3 || o.iav == 0)) return fail(S, ND_ERR_UNSUPPORTED_OPTION, "rates: one-fluid dust needs idustevol=0, iener/=3, iav>0");
  if (o.iresist != 0 && o.iresist != 1) return fail(S, ND_ERR_UNSUPPORTED_OPTION, "rates: iresist");
  if (o.icty != 0 || o.ixsph != 0 || o.igravity != 0 || o.iexternal_force != 0 || o.damp != 0.) return fail(S, ND_ERR_UNSUPPORTED_OPTION, "rates: option");

  std::vector<int> listneigh(S.idim + 1);
  std::vector<double> h1(ntotal + 1);
  double *del2u = S.a.del2u;
  // :184-192
  S.dtcourant = 1.e6; S.dtav = DBL_MAX; S.dtvisc = DBL_MAX;
  double ts_min = DBL_MAX, h_on_csts_max = 0.;
  const double zero = 1.e-10;
  double vsigmax = 0.;
  double dr[3] = {0., 0., 0.};
  int nclumped = 0;
  long long npairs_rates = 0;
  for (int i = 1; i <= ntotal; i++) {                           // :194-226
    for (int k = 1; k <= 3; k++) { V3(force, k, i) = 0.; V3(dBevoldt, k, i) = 0.; V3(daldt, k, i) = 0.; V3(gradpsi, k, i) = 0.; V3(fmag, k, i) = 0.; V3(xsphterm, k, i) = 0.; V3(graddivv, k, i) = 0.; }
    A1(dudt, i) = 0.; A1(dendt, i) = 0.; A1(dpsidt, i) = 0.; A1(divB, i) = 0.;
    if (o.imhd > 0) for (int k = 1; k <= 3; k++) V3(curlB, k, i) = 0.;
    del2u[i - 1] = 0.;
    h1[i] = 1. / A1(hh, i);
    if (o.onef_dust) { A1(ddustevoldt, i) = 0.; for (int k = 1; k <= 3; k++) V3(ddeltavdt, k, i) = 0.; }   // :222-225
  }
  // :231-245 stressmax
  double stressmax = 0.;
  if (o.imhd != 0 && (o.imagforce == 2 || o.imagforce == 7)) {
    for (int i = 1; i <= ntotal; i++) {
      double B2i = dot3(&V3(Bfield, 1, i), &V3(Bfield, 1, i));
      double stressterm = std::max(0.5 * B2i - A1(pr, i), 0.);
      double mB = std::max(o.Bconst[0], std::max(o.Bconst[1], o.Bconst[2]));
      stressmax = std::max(stressterm, std::max(stressmax, mB));
    }
  }
  S.stressmax_out = stressmax;
  // pair-loop locals shared with the contained procedures
  double Bi[3] = {0, 0, 0}, Bj[3] = {0, 0, 0}, Brhoi[3] = {0, 0, 0}, Brhoj[3] = {0, 0, 0};
  double Brho2i = 0., Brho2j = 0., valfven2i = 0., valfven2j = 0., projBi = 0., projBj = 0., projBrhoi = 0., projBrhoj = 0., alphaBi = 0.;
  double etai = 0., etaj = 0.;
  // one-fluid dust locals (ndust = 1).  dustfraci keeps the value of the LAST particle of the pair loop when the
  // finalisation loop tests it (:566) -- the reference never reassigns it there.
  double dustfraci = 0., rhodusti = 0., rhogasi = 0., deltavi[3] = {0, 0, 0}, deltav2i = 0., rhogrhodonrhoi = 0.;
  int nneigh = 0;
  for (int icell = 1; icell <= S.ncellsloop; icell++) {         // :304
    get_neighbour_list(S, icell, listneigh.data(), nneigh);
    int i = S.ifirstincell[icell];
    int idone = -1;
    while (i != -1) {                                           // :317
      idone = idone + 1;
      double xi[3] = {0, 0, 0};
      for (int k = 1; k <= ndim; k++) xi[k - 1] = X_(k, i);
      int itypei = A1(itype, i);
      double rhoi = A1(rho, i);
      double rho1i = 1. / rhoi;
      double rho21i = rho1i * rho1i;
      double pri = std::max(A1(pr, i) - o.pext, 0.);
      double prneti = pri - 0.;                                 // pequil = 0 for iexternal_force = 0
      double pmassi = A1(pmass, i);
      double Prho2i = pri * rho21i;
      double spsoundi = A1(spsound, i);
      double uui = A1(uu, i);
      double veli[3];
      for (int k = 1; k <= 3; k++) veli[k - 1] = V3(vel, k, i);
      double alphai = V3(alpha, 1, i), alphaui = V3(alpha, 2, i);
      alphaBi = V3(alpha, 3, i);
      const double phii = 1.0, phii1 = 1. / phii;
      double sqrtgi = A1(sqrtg, i);
      if (o.onef_dust) {                                        // :344-360
        dustfraci = A1(dustfrac, i);
        if (o.use_smoothed_rhodust) { rhodusti = A1(rhodust, i); rhogasi = A1(rhogas, i); }
        else { rhodusti = dustfraci * rhoi; rhogasi = (1. - dustfraci) * rhoi; }
        for (int k = 0; k < 3; k++) deltavi[k] = V3(deltav, k + 1, i);
        deltav2i = dot3(deltavi, deltavi);
        rhogrhodonrhoi = rhogasi * rhodusti * rho1i;
      } else { rhogasi = rhoi; rhodusti = 0.; deltav2i = 0.; }
      if (o.imhd != 0) {                                        // :362-380
        for (int k = 0; k < 3; k++) Bi[k] = V3(Bfield, k + 1, i);
        for (int k = 0; k < 3; k++) Brhoi[k] = Bi[k] * rho1i;
        double B2i = dot3(Bi, Bi);                              // dot_product_gr with gdiag=1
        Brho2i = B2i * rho21i;
        valfven2i = B2i * rho1i;
        if (o.iresist > 0) etai = o.etamhd;
      }
      double gradhi = A1(gradh, i);
      double hi = A1(hh, i);
      if (hi <= 0.) return fail(S, ND_ERR_H_NONPOSITIVE, "rates: h <= 0");
      double hi1 = h1[i];
      double hi21 = hi1 * hi1;
      double hfacwabi = powi(hi1, ndim);
      double hfacgrkerni = hfacwabi * hi1;
      double forcei[3] = {0, 0, 0}, fextrai[3] = {0, 0, 0}, dBevoldti[3] = {0, 0, 0};
      for (int n = idone + 1; n <= nneigh; n++) {               // :398
        int j = listneigh[n - 1];
        if (!((j != i) && !(j > npart && i > npart))) continue; // :401
        double dx[3] = {0, 0, 0};
        for (int k = 1; k <= ndim; k++) dx[k - 1] = xi[k - 1] - X_(k, j);
        double hj = A1(hh, j);
        double hj1 = h1[j];
        double hj21 = hj1 * hj1;
        double rij2 = 0.;
        for (int k = 0; k < ndim; k++) rij2 = rij2 + dx[k] * dx[k];
        double q2i = rij2 * hi21;
        double q2j = rij2 * hj21;
        if (!((q2i < K.radkern2) || (q2j < K.radkern2))) continue;   // :415
        if (pairs_i) { pairs_i->push_back(i); pairs_j->push_back(j); }
        double rij = std::sqrt(rij2);
        if (rij <= DBL_EPSILON && A1(itype, j) == itypei) {     // :417-430
          nclumped = nclumped + 1;
          if (rij < DBL_MIN) {
            dr[0] = dr[1] = dr[2] = 0.;
            if (itypei != 2) return fail(S, ND_ERR_INVALID_ARG, "rates: dx = 0 (coincident particles of the same type)");
          }
        } else if (rij <= DBL_EPSILON) {
          dr[0] = dr[1] = dr[2] = 0.;
        } else {
          for (int k = 0; k < ndim; k++) dr[k] = dx[k] / rij;
        }
        int itypej = A1(itype, j);
        if (types_interact(itypei, itypej) || (o.idust == 2 && o.idrag_nature > 0)) npairs_rates += (i <= npart ? 1 : 0) + (j <= npart ? 1 : 0);
        if (types_interact(itypei, itypej)) {
          // ===================== rates_core :1175-1690 =====================
          double pmassj = A1(pmass, j);
          double wabi, grkerni, grkernalti, grgrkernalti, wabj, grkernj, grkernaltj, grgrkernaltj, grkern;
          interpolate_kernels(K, q2i, wabi, grkerni, grkernalti, grgrkernalti);   // :1208
          wabi = wabi * hfacwabi;
          grkerni = grkerni * hfacgrkerni;
          double hfacwabj = powi(hj1, ndim);
          double hfacgrkernj = hfacwabj * hj1;
          interpolate_kernels(K, q2j, wabj, grkernj, grkernaltj, grgrkernaltj);   // :1217
          wabj = wabj * hfacwabj;
          grkernj = grkernj * hfacgrkernj;
          if (o.ikernav == 3) {                                 // :1227-1242
            grkerni = grkerni * gradhi;
            grkernj = grkernj * A1(gradh, j);
            grkern = 0.5 * (grkerni + grkernj);
          } else {
            grkern = 0.5 * (grkerni + grkernj);
            grkerni = grkern; grkernj = grkern;
          }
          double velj[3], dvel[3];
          for (int k = 1; k <= 3; k++) velj[k - 1] = V3(vel, k, j);
          for (int k = 0; k < 3; k++) dvel[k] = veli[k] - velj[k];
          double dvdotr = dot3(dvel, dr);                       // :1250
          double rhoj = A1(rho, j);
          double rho1j = 1. / rhoj;
          double rho21j = rho1j * rho1j;
          double rhoav1 = 0.5 * (rho1i + rho1j);                // :1261
          double dustfracj = 0., rhodustj = 0., rhogasj = rhoj, rhogrhodonrhoj = 0., deltavj[3] = {0, 0, 0}, deltav2j = 0.;
          double projdvgas = dvdotr, projdeltavi = 0., projdeltavj = 0.;
          if (o.onef_dust) {                                    // :1262-1284
            dustfracj = A1(dustfrac, j);
            if (o.use_smoothed_rhodust) { rhodustj = A1(rhodust, j); rhogasj = A1(rhogas, j); }
            else { rhodustj = rhoj * dustfracj; rhogasj = (1. - dustfracj) * rhoj; }
            rhogrhodonrhoj = rhogasj * rhodustj * rho1j;
            for (int k = 0; k < 3; k++) deltavj[k] = V3(deltav, k + 1, j);
            deltav2j = dot3(deltavj, deltavj);
            double dvgas[3];
            for (int k = 0; k < 3; k++) dvgas[k] = (veli[k] - dustfraci * deltavi[k]) - (velj[k] - dustfracj * deltavj[k]);
            projdvgas = dot3(dvgas, dr);
            projdeltavi = dot3(deltavi, dr);
            projdeltavj = dot3(deltavj, dr);
          }
          double prj = std::max(A1(pr, j) - o.pext, 0.);
          double prnetj = prj - 0.;
          double Prho2j = prj * rho21j;
          double spsoundj = A1(spsound, j);
          const double phij = 1.0;
          double phii_on_phij = phii / phij;                    // :1291-1292
          double phij_on_phii = phij * phii1;
          double sqrtgj = A1(sqrtg, j);
          double dB[3] = {0, 0, 0}, projdB = 0., projBconst = 0.;
          if (o.imhd != 0) {                                    // :1295-1323
            for (int k = 0; k < 3; k++) Bj[k] = V3(Bfield, k + 1, j);
            for (int k = 0; k < 3; k++) Brhoj[k] = Bj[k] * rho1j;
            for (int k = 0; k < 3; k++) dB[k] = Bi[k] - Bj[k];
            projBi = dot3(Bi, dr);
            projBj = dot3(Bj, dr);
            projdB = dot3(dB, dr);
            projBrhoi = dot3(Brhoi, dr);
            projBrhoj = dot3(Brhoj, dr);
            double B2j = dot3(Bj, Bj);
            valfven2j = B2j * rho1j;
            Brho2j = B2j * rho21j;
            projBconst = dot3(o.Bconst, dr);
            if (o.iresist > 0) etaj = o.etamhd;
          }
          (void)projBconst;
          double forcej[3] = {0, 0, 0}, fextraj[3] = {0, 0, 0};
          double vsig = 0., vsigav = 0., vsigi, vsigj, vsigB, vsigu, vsigdtc;
          if (o.imhd != 0) {                                    // :1417-1436
            double vsig2i = spsoundi * spsoundi + valfven2i;
            double vsig2j = spsoundj * spsoundj + valfven2j;
            double vsigproji = vsig2i * vsig2i - 4. * ((spsoundi * projBi) * (spsoundi * projBi)) * rho1i;
            double vsigprojj = vsig2j * vsig2j - 4. * ((spsoundj * projBj) * (spsoundj * projBj)) * rho1j;
            if (vsigproji < 0. || vsigprojj < 0.) return fail(S, ND_ERR_VSIG_DET, "rates: vsig det < 0");
            vsigi = std::sqrt(0.5 * (vsig2i + std::sqrt(vsigproji)));
            vsigj = std::sqrt(0.5 * (vsig2j + std::sqrt(vsigprojj)));
            if (o.iavlim[2] != 2) vsigB = std::sqrt(dot3(dvel, dvel));   // norm2(dvel)
            else vsigB = 0.5 * (vsigi + vsigj) + std::fabs(dvdotr);
          } else {                                              // :1446-1450
            vsigi = spsoundi; vsigj = spsoundj; vsigB = 0.;
          }
          vsig = 0.5 * (std::max(vsigi + vsigj - o.beta * dvdotr, 0.0));   // :1452
          vsigu = std::sqrt(std::fabs(prneti - prnetj) * rhoav1);          // :1459
          vsigdtc = std::max(vsig, std::max(0.5 * (vsigi + vsigj + o.beta * std::fabs(dvdotr)), vsigB));   // :1465
          if (o.idust == 1) vsigdtc = vsigdtc + std::sqrt(deltav2i + deltav2j);                            // :1466-1468
          if (A1(itype, i) == ND_ITYPE_DUST) {                  // :1472-1482
            vsig = 0.; vsigu = 0.;
          } else {
            double dvsigdtc = 1. / vsigdtc;
            vsigmax = std::max(vsigmax, vsigdtc);
            if (vsigdtc > zero) S.dtcourant = std::min(S.dtcourant, std::min(hi * dvsigdtc, hj * dvsigdtc));
          }
          if (o.iav > 0 && o.idust == 1) {
            // ===================== artificial_dissipation_dust :1969-2148 (iav = 1, 2, 3; iener /= 3) =====================
            double alphaav = 0.5 * (alphai + V3(alpha, 1, j));
            double alphau = 0.5 * (alphaui + V3(alpha, 2, j));
            double alphaB = 0.5 * (alphaBi + V3(alpha, 3, j));
            vsigav = std::max(alphaav, alphau) * vsig;          // :1984
            double dustfracav = 0.5 * (dustfraci + dustfracj);
            double projdvgasav = dvdotr - dustfracav * (projdeltavi - projdeltavj);
            double ddeltav[3];
            for (int k = 0; k < 3; k++) ddeltav[k] = deltavi[k] - deltavj[k];
            double projddeltav = dot3(ddeltav, dr);
            double dpmomdotr = (o.iav == 3) ? projdvgasav : (o.iav == 2) ? projdvgas : dvdotr;   // :1993-1999
            double term = vsig * rhoav1 * grkern;
            double termv = term;
            const bool allpairs = (o.iav == 1);                 // :2005-2007
            if (o.iav == 2 || o.iav == 3) termv = termv * (1. - dustfracav);
            double termu = vsigu * rhoav1 * grkern * (1. - dustfracav);
            if (projdvgas < 0 || allpairs) {                    // :2023-2038
              double visc = alphaav * termv * dpmomdotr;
              if (o.iav == 1 || o.iav == 3) {
                for (int k = 0; k < 3; k++) fextrai[k] = fextrai[k] + pmassj * visc * dr[k];
                for (int k = 0; k < 3; k++) fextraj[k] = fextraj[k] - pmassi * visc * dr[k];
              } else {
                for (int k = 0; k < 3; k++) forcei[k] = forcei[k] + pmassj * visc * dr[k];
                for (int k = 0; k < 3; k++) forcej[k] = forcej[k] - pmassi * visc * dr[k];
              }
            }
            if (o.iener > 0) {                                  // :2049-2146
              double vissv, vissu, vissdv = 0., termdv = 0.;
              if (projdvgas < 0 || allpairs) {
                if (o.iav == 3) vissv = -alphaav * 0.5 * (projdvgasav * projdvgasav);
                else if (o.iav == 2) vissv = -alphaav * 0.5 * (projdvgas * projdvgas);
                else vissv = -alphaav * 0.5 * (dvdotr * dvdotr);
              } else vissv = 0.;
              if (o.iener == 1) vissu = 0.;
              else vissu = alphau * (A1(uu, i) - A1(uu, j));
              double faci = 1. / (dustfraci * (1. - dustfraci));   // :2078-2079
              double facj = 1. / (dustfracj * (1. - dustfracj));
              if (o.iav == 1 || o.iav == 3) {
                if (o.iav == 1) { vissdv = -0.5 * dot3(ddeltav, ddeltav); termdv = alphaav * vsig * rhoav1 * dustfracav * grkern; }
                else { vissdv = 0.; termdv = alphaav * vsig * rhoav1 * dustfracav * grkern * (1. - dustfracav); }
                if (o.iav == 3) {
                  for (int k = 0; k < 3; k++) V3(ddeltavdt, k + 1, i) = V3(ddeltavdt, k + 1, i) + faci * pmassj * termdv * (-projdvgasav) * dr[k];
                  for (int k = 0; k < 3; k++) V3(ddeltavdt, k + 1, j) = V3(ddeltavdt, k + 1, j) - facj * pmassi * termdv * (-projdvgasav) * dr[k];
                } else {
                  for (int k = 0; k < 3; k++) V3(ddeltavdt, k + 1, i) = V3(ddeltavdt, k + 1, i) + faci * pmassj * termdv * (ddeltav[k]);
                  for (int k = 0; k < 3; k++) V3(ddeltavdt, k + 1, j) = V3(ddeltavdt, k + 1, j) - facj * pmassi * termdv * (ddeltav[k]);
                }
              } else if (o.iav == 2) {                          // :2111-2125
                termdv = 0.; vissdv = 0.;
                if (projddeltav < 0.) {
                  double vsigdv = 0.5 * (spsoundi + spsoundj);
                  termdv = alphaav * vsigdv * rhoav1 * grkern * dustfracav * (1. - dustfracav);
                  for (int k = 0; k < 3; k++) V3(ddeltavdt, k + 1, i) = V3(ddeltavdt, k + 1, i) + faci * pmassj * termdv * (projddeltav) * dr[k];
                  for (int k = 0; k < 3; k++) V3(ddeltavdt, k + 1, j) = V3(ddeltavdt, k + 1, j) - facj * pmassi * termdv * (projddeltav) * dr[k];
                  vissdv = -0.5 * (projddeltav * projddeltav);
                }
              }
              faci = rhoi / rhogasi;                            // :2133-2138 (damp = 0)
              facj = rhoj / rhogasj;
              A1(dudt, i) = A1(dudt, i) + faci * pmassj * (termv * (vissv) + termu * vissu + termdv * vissdv);
              A1(dudt, j) = A1(dudt, j) + facj * pmassi * (termv * (vissv) - termu * vissu + termdv * vissdv);
              double vsigeps = 0.5 * (spsoundi + spsoundj);     // :2140-2145
              double diffeps = alphaB * rhoav1 * vsigeps * (dustfraci - dustfracj) * grkern;
              A1(ddustevoldt, i) = A1(ddustevoldt, i) + pmassj * diffeps;
              A1(ddustevoldt, j) = A1(ddustevoldt, j) - pmassi * diffeps;
            }
          } else if (o.iav > 0 && o.iav != 3) {
            // ===================== artificial_dissipation :1700-1894 =====================
            double alphaav = 0.5 * (alphai + V3(alpha, 1, j));
            double alphau = 0.5 * (alphaui + V3(alpha, 2, j));
            double alphaB = 0.5 * (alphaBi + V3(alpha, 3, j));
            vsigav = std::max(alphaav, std::max(alphau, alphaB)) * vsig;
            double dpmomdotr = -dvdotr;
            double term = vsig * rhoav1 * grkern;
            double termv = term;
            double termu = vsigu * rhoav1 * grkern;
            double termB = vsigB * rhoav1 * grkern;
            if (dvdotr < 0 && o.iav <= 3) {                     // :1745-1748
              double visc = alphaav * termv * dpmomdotr;
              for (int k = 0; k < 3; k++) forcei[k] = forcei[k] - pmassj * visc * dr[k];
              for (int k = 0; k < 3; k++) forcej[k] = forcej[k] + pmassi * visc * dr[k];
            }
            if (o.imhd != 0) {                                  // :1761-1784 (imhd>0)
              double Bvisc[3], dBdtvisc[3];
              if (o.iav >= 2) for (int k = 0; k < 3; k++) Bvisc[k] = dB[k] * rhoav1;
              else for (int k = 0; k < 3; k++) Bvisc[k] = (dB[k] - dr[k] * projdB) * rhoav1;
              for (int k = 0; k < 3; k++) dBdtvisc[k] = alphaB * termB * Bvisc[k];
              for (int k = 0; k < 3; k++) dBevoldti[k] = dBevoldti[k] + rhoi * pmassj * dBdtvisc[k];
              for (int k = 0; k < 3; k++) V3(dBevoldt, k + 1, j) = V3(dBevoldt, k + 1, j) - rhoj * pmassi * dBdtvisc[k];
            }
            if (o.iener == 3) {                                 // :1792-1830
              double qdiff = 0.;
              if (dvdotr < 0 && o.iav <= 3) {
                double v2i = dot3(veli, dr); v2i = v2i * v2i;
                double v2j = dot3(velj, dr); v2j = v2j * v2j;
                qdiff = qdiff + term * alphaav * 0.5 * (v2i - v2j);
              }
              qdiff = qdiff + alphau * termu * (A1(uu, i) - A1(uu, j));
              if (o.imhd > 0) {
                double B2i, B2j;
                if (o.iav >= 2) { B2i = dot3(Bi, Bi); B2j = dot3(Bj, Bj); }
                else { double pi_ = dot3(Bi, dr), pj_ = dot3(Bj, dr); B2i = (dot3(Bi, Bi) - pi_ * pi_); B2j = (dot3(Bj, Bj) - pj_ * pj_); }
                qdiff = qdiff + alphaB * termB * 0.5 * (B2i - B2j) * rhoav1;
              }
              A1(dendt, i) = A1(dendt, i) + pmassj * qdiff;
              A1(dendt, j) = A1(dendt, j) - pmassi * qdiff;
            } else if (o.iener > 0) {                           // :1835-1884
              double vissv, vissu, vissB;
              if (dvdotr < 0 && o.iav <= 3) { double t = (dot3(veli, dr) - dot3(velj, dr)); vissv = -alphaav * 0.5 * (t * t); }
              else vissv = 0.;
              vissu = alphau * (A1(uu, i) - A1(uu, j));
              if (o.imhd > 0) {
                if (o.iav >= 2) vissB = -alphaB * 0.5 * (dot3(dB, dB)) * rhoav1;
                else vissB = -alphaB * 0.5 * (dot3(dB, dB) - projdB * projdB) * rhoav1;
              } else vissB = 0.;
              A1(dudt, i) = A1(dudt, i) + pmassj * (term * (vissv) + termu * vissu + termB * (vissB));
              A1(dudt, j) = A1(dudt, j) + pmassi * (term * (vissv) - termu * vissu + termB * (vissB));
            }
          } else if (o.iav == 3) {
            // ===================== artificial_dissipation_phantom :1903-1959 =====================
            double dudti = 0., dudtj = 0.;
            if (dvdotr < 0.) {
              double vsi = std::max(alphai * spsoundi - o.beta * dvdotr, 0.);
              double vsj = std::max(V3(alpha, 1, j) * spsoundj - o.beta * dvdotr, 0.);
              double projv = dvdotr;
              double qi = -0.5 * rhoi * vsi * projv;
              double qj = -0.5 * rhoj * vsj * projv;
              double visc = (qi * rho21i * grkerni + qj * rho21j * grkernj);
              for (int k = 0; k < 3; k++) forcei[k] = forcei[k] - pmassj * visc * dr[k];
              for (int k = 0; k < 3; k++) forcej[k] = forcej[k] + pmassi * visc * dr[k];
              dudti = qi * rho21i * pmassj * dvdotr * grkerni;
              dudtj = qj * rho21j * pmassi * dvdotr * grkernj;
            }
            double du = A1(uu, i) - A1(uu, j);
            double cfaci = 0.5 * alphaui * rhoi * vsigu * du;
            double cfacj = 0.5 * V3(alpha, 2, j) * rhoj * vsigu * du;
            double diffu = cfaci * grkerni * (rho1i * rho1i) + cfacj * grkernj * (rho1j * rho1j);
            A1(dudt, i) = A1(dudt, i) + dudti + pmassj * diffu;
            A1(dudt, j) = A1(dudt, j) + dudtj - pmassi * diffu;
          }
          if (vsigav > zero) S.dtav = std::min(S.dtav, std::min(hi / vsigav, hj / vsigav));   // :1500
          {                                                     // :1538-1568 pressure term (iprterm = 0)
            double prterm = phii_on_phij * Prho2i * sqrtgi * grkerni + phij_on_phii * Prho2j * sqrtgj * grkernj;
            for (int k = 0; k < 3; k++) forcei[k] = forcei[k] - pmassj * prterm * dr[k];
            for (int k = 0; k < 3; k++) forcej[k] = forcej[k] + pmassi * prterm * dr[k];
          }
          if (o.imhd != 0) {
            // ===================== mhd_terms :2377-2720 =====================
            double faniso[3], fmagi[3];
            double fiso = 0.5 * (Brho2i * phij_on_phii * grkerni * sqrtgi + Brho2j * phii_on_phij * grkernj * sqrtgj);   // :2526
            for (int k = 0; k < 3; k++)                         // :2534-2537
              faniso[k] = (1.0 * Brhoi[k] * projBrhoi * sqrtgi - stressmax * dr[k] * rho21i) * phij_on_phii * grkerni +
                          (1.0 * Brhoj[k] * projBrhoj * sqrtgj - stressmax * dr[k] * rho21j) * phii_on_phij * grkernj;
            for (int k = 0; k < 3; k++) fmagi[k] = faniso[k] - fiso * dr[k];
            A1(divB, i) = A1(divB, i) - pmassj * projdB * grkern;                 // :2552-2553
            A1(divB, j) = A1(divB, j) - pmassi * projdB * grkern;
            double dBdtambi[3] = {0., 0., 0.};
            if (o.imhd > 0) {                                   // :2600-2603
              double curlBi[3];
              cross_product3D(dB, dr, curlBi);
              for (int k = 0; k < 3; k++) V3(curlB, k + 1, i) = V3(curlB, k + 1, i) + pmassj * curlBi[k] * grkern;
              for (int k = 0; k < 3; k++) V3(curlB, k + 1, j) = V3(curlB, k + 1, j) + pmassi * curlBi[k] * grkern;
            }
            for (int k = 0; k < 3; k++) forcei[k] = forcei[k] + pmassj * (fmagi[k]);   // :2629-2630
            for (int k = 0; k < 3; k++) forcej[k] = forcej[k] - pmassi * (fmagi[k]);
            // :2663-2667 induction, imhd = 1 / 11
            for (int k = 0; k < 3; k++) dBevoldti[k] = dBevoldti[k] - phii_on_phij * pmassj * ((dvel[k] * projBrhoi) * grkerni + rhoi * dBdtambi[k]);
            for (int k = 0; k < 3; k++) V3(dBevoldt, k + 1, j) = V3(dBevoldt, k + 1, j) - phij_on_phii * pmassi * ((dvel[k] * projBrhoj) * grkernj - rhoj * dBdtambi[k]);
            if (o.iresist == 1) {                               // :2685-2710
              double etaij = 0.5 * (etai + etaj);
              double dBdtvisc[3];
              for (int k = 0; k < 3; k++) dBdtvisc[k] = -2. * etaij * dB[k] / (rij + DBL_EPSILON);
              for (int k = 0; k < 3; k++) dBevoldti[k] = dBevoldti[k] - rhoi * pmassj * 0.5 * ((rho1i * rho1i) * grkerni + (rho1j * rho1j) * grkernj) * dBdtvisc[k];
              for (int k = 0; k < 3; k++) V3(dBevoldt, k + 1, j) = V3(dBevoldt, k + 1, j) + rhoj * pmassi * 0.5 * ((rho1i * rho1i) * grkerni + (rho1j * rho1j) * grkernj) * dBdtvisc[k];
              if (o.iener == 3 || o.iener == 1) return fail(S, ND_ERR_UNSUPPORTED_OPTION, "unimplemented iener for physical resistivity");
              else if (o.iener > 0) {
                double term = -etaij * rho1i * rho1j * dot3(dB, dB) * grkern / rij;
                A1(dudt, i) = A1(dudt, i) + pmassj * term;
                A1(dudt, j) = A1(dudt, j) + pmassi * term;
              }
            }
            if (o.idivbzero >= 2) {                             // :2712-2717
              double gradpsiterm = (A1(psi, i) * rho21i * grkerni + A1(psi, j) * rho21j * grkernj);
              for (int k = 0; k < 3; k++) V3(gradpsi, k + 1, i) = V3(gradpsi, k + 1, i) - pmassj * gradpsiterm * dr[k];
              for (int k = 0; k < 3; k++) V3(gradpsi, k + 1, j) = V3(gradpsi, k + 1, j) + pmassi * gradpsiterm * dr[k];
            }
          }
          if (o.idust == 1) {
            // ===================== dust_derivs :2726-2807 (idustevol = 0) =====================
            {
              double termi = rhogrhodonrhoi * projdeltavi * rho21i * grkerni;
              double termj = rhogrhodonrhoj * projdeltavj * rho21j * grkernj;
              double term = termi + termj;
              A1(ddustevoldt, i) = A1(ddustevoldt, i) - pmassj * term;
              A1(ddustevoldt, j) = A1(ddustevoldt, j) + pmassi * term;
            }
            double termi = (rhogasi - rhodusti) * rho1i * deltav2i;   // :2767-2769 high Mach number term
            double termj = (rhogasj - rhodustj) * rho1j * deltav2j;
            double dterm = 0.5 * (termi - termj);
            for (int k = 0; k < 3; k++)                         // :2776-2777 (the forcei term is added after the pair loop, :460)
              V3(ddeltavdt, k + 1, i) = V3(ddeltavdt, k + 1, i) + rho1i * pmassj * (dvel[k] * projdeltavi + dterm * dr[k]) * grkerni;
            for (int k = 0; k < 3; k++)
              V3(ddeltavdt, k + 1, j) = V3(ddeltavdt, k + 1, j) + rho1j * pmassi * (dvel[k] * projdeltavj + dterm * dr[k]) * grkernj - rhoj / rhogasj * forcej[k];
            double prdustterm[3];                               // :2792-2796 anisotropic pressure
            for (int k = 0; k < 3; k++)
              prdustterm[k] = rhogrhodonrhoi * deltavi[k] * projdeltavi * rho21i * grkerni + rhogrhodonrhoj * deltavj[k] * projdeltavj * rho21j * grkernj;
            for (int k = 0; k < 3; k++) fextrai[k] = fextrai[k] - pmassj * (prdustterm[k]);
            for (int k = 0; k < 3; k++) fextraj[k] = fextraj[k] + pmassi * (prdustterm[k]);
            if (o.iener > 0) {                                  // :2801-2805
              double du = A1(uu, i) - A1(uu, j);
              A1(dudt, i) = A1(dudt, i) + pmassj * (pri * rho1i / rhogasi * projdvgas - rhodusti * rho21i * du * projdeltavi) * grkerni;
              A1(dudt, j) = A1(dudt, j) + pmassi * (prj * rho1j / rhogasj * projdvgas - rhodustj * rho21j * du * projdeltavj) * grkernj;
            }
          }
          for (int k = 0; k < 3; k++) V3(force, k + 1, j) = V3(force, k + 1, j) + fextraj[k] + forcej[k];   // :1600
          if (o.iav > 0) {                                      // :1639-1656
            if (o.iavlim[1] > 0) {
              double graduterm = (A1(uu, i) - A1(uu, j)) / rij;
              del2u[i - 1] = del2u[i - 1] + pmassj * rho1j * graduterm * grkerni;
              del2u[j - 1] = del2u[j - 1] - pmassi * rho1i * graduterm * grkernj;
            }
            if (o.iavlim[0] == 3) {
              for (int k = 0; k < 3; k++) V3(graddivv, k + 1, i) = V3(graddivv, k + 1, i) + pmassj * rho1j / rij * dvdotr * grkerni * dr[k];
              for (int k = 0; k < 3; k++) V3(graddivv, k + 1, j) = V3(graddivv, k + 1, j) - pmassi * rho1i / rij * dvdotr * grkernj * dr[k];
            } else {
              for (int k = 0; k < 3; k++) V3(graddivv, k + 1, i) = V3(graddivv, k + 1, i) + pmassj * (dvel[k] - dvdotr) * grkerni;
              for (int k = 0; k < 3; k++) V3(graddivv, k + 1, j) = V3(graddivv, k + 1, j) + pmassi * (-dvel[k] - dvdotr) * grkernj;
            }
          }
        } else if (o.idust == 2 && o.idrag_nature > 0) {
          // ===================== drag_forces :1074-1169 =====================
          double velj[3], dvel[3], drdrag[3] = {0, 0, 0};
          for (int k = 1; k <= 3; k++) velj[k - 1] = V3(vel, k, j);
          for (int k = 0; k < 3; k++) dvel[k] = veli[k] - velj[k];
          double dv2 = dot3(dvel, dvel);
          bool skip = false;
          if (rij <= DBL_EPSILON) {
            if (dv2 <= DBL_EPSILON) skip = true;                // :1092-1094 return
            else { double vij = std::sqrt(dv2); for (int k = 0; k < 3; k++) drdrag[k] = dvel[k] / vij; }
          } else for (int k = 0; k < 3; k++) drdrag[k] = dr[k];
          if (!skip) {
            double wabi, wabj;
            interpolate_kerneldrag(K, q2i, wabi);               // :1111-1112
            interpolate_kerneldrag(K, q2j, wabj);
            wabi = wabi * hfacwabi;
            double hfacwabj = (1. / powi(A1(hh, j), ndim));     // :1115
            wabj = wabj * hfacwabj;
            double pmassj = A1(pmass, j);
            double rhoj = A1(rho, j);
            double projv = dot3(dvel, drdrag);
            double projvstar = projv;                           // islope_limiter < 0
            double spsoundgas, wab, ts;
            bool ret = false;
            if (itypei == ND_ITYPE_GAS || itypei == ND_ITYPE_BND) {      // :1133-1137
              spsoundgas = A1(spsound, i); wab = wabi;
              ts = get_tstop(o.idrag_nature, rhoi, rhoj, spsoundgas, o.Kdrag);
              h_on_csts_max = std::max(h_on_csts_max, A1(hh, i) / (spsoundgas * ts));
            } else {
              if (itypej != ND_ITYPE_GAS && itypej != ND_ITYPE_BND) ret = true;   // :1139
              else {
                spsoundgas = A1(spsound, j); wab = wabj;
                ts = get_tstop(o.idrag_nature, rhoj, rhoi, spsoundgas, o.Kdrag);
                h_on_csts_max = std::max(h_on_csts_max, A1(hh, j) / (spsoundgas * ts));
              }
            }
            if (!ret) {
              ts_min = std::min(ts_min, ts);
              double dragterm = ndim * wab / ((rhoi + rhoj) * ts) * projvstar;     // :1156
              double dragterm_en = dragterm * projv;
              for (int k = 0; k < 3; k++) forcei[k] = forcei[k] - dragterm * pmassj * drdrag[k];
              for (int k = 0; k < 3; k++) V3(force, k + 1, j) = V3(force, k + 1, j) + dragterm * pmassi * drdrag[k];
              if (itypei == ND_ITYPE_GAS) A1(dudt, i) = A1(dudt, i) + pmassj * dragterm_en;
              if (itypej == ND_ITYPE_GAS) A1(dudt, j) = A1(dudt, j) + pmassi * dragterm_en;
            }
          }
        }
      }
      // :458-459
      for (int k = 0; k < 3; k++) V3(force, k + 1, i) = V3(force, k + 1, i) + fextrai[k] + forcei[k];
      for (int k = 0; k < 3; k++) V3(dBevoldt, k + 1, i) = V3(dBevoldt, k + 1, i) + dBevoldti[k];
      if (o.idust == 1) for (int k = 0; k < 3; k++) V3(ddeltavdt, k + 1, i) = V3(ddeltavdt, k + 1, i) - rhoi / rhogasi * forcei[k];   // :460
      i = S.ll[i];
    }
  }
  S.nclumped = nclumped;
  S.npairs_rates = npairs_rates;
  S.vsigmax_out = vsigmax;
  S.ts_min_out = ts_min;
  S.h_on_csts_max_out = h_on_csts_max;
  if (o.imhd != 0 && o.idivbzero >= 2) S.vsig2max = vsigmax * vsigmax;   // :518-520
  else S.vsig2max = 0.;
  // ===================== finalisation loop :522-924 =====================
  double fhmax = 0.0;
  double fmean[3] = {0, 0, 0};
  S.dtdrag = DBL_MAX;
  S.dtforce = DBL_MAX;
  // iener=3: the reference reads fprev(:,i), which is only allocated when igravity/=0 (:479, :824).  We take the evident
  // intent, fprev = force as it stands after the pair loop (identical to the igravity/=0 behaviour with zero gravity).
  for (int i = 1; i <= npart; i++) {                            // :532
    double rhoi = A1(rho, i);
    double rho1i = 1. / rhoi;
    if (o.idust == 2 && o.idrag_nature != 0 && (o.Kdrag > 0. || o.idrag_nature > 1)) S.dtdrag = std::min(S.dtdrag, ts_min);   // :543-547
    else if (o.idust == 1) {                                    // :548-582 one fluid dust
      if (o.use_smoothed_rhodust) { rhodusti = A1(rhodust, i); rhogasi = A1(rhogas, i); }
      else { rhodusti = rhoi * A1(dustfrac, i); rhogasi = (1. - A1(dustfrac, i)) * rhoi; }
      double tstop = get_tstop(o.idrag_nature, rhogasi, rhodusti, A1(spsound, i), o.Kdrag);
      S.dtdrag = std::min(S.dtdrag, tstop);
      double dtstop;
      if (dustfraci > 0.) {                                     // :566 -- dustfraci is NOT particle i's (see above)
        dtstop = 1. / tstop;
        for (int k = 1; k <= 3; k++) V3(ddeltavdt, k, i) = V3(ddeltavdt, k, i) - V3(deltav, k, i) * dtstop;
      } else {
        dtstop = 0.;
        for (int k = 1; k <= 3; k++) V3(ddeltavdt, k, i) = 0.;
      }
      if (o.iener > 0) {                                        // :579-582
        deltav2i = dot3(&V3(deltav, 1, i), &V3(deltav, 1, i));
        A1(dudt, i) = A1(dudt, i) + rhodusti * rho1i * deltav2i * dtstop;
      }
    }
    if (o.imhd != 0) {                                          // :630-649
      if (o.imhd > 0) for (int k = 1; k <= 3; k++) V3(curlB, k, i) = V3(curlB, k, i) * rho1i;
      A1(divB, i) = A1(divB, i) * rho1i;
    }
    for (int k = 0; k < 3; k++) fmean[k] = fmean[k] + A1(pmass, i) * V3(force, k + 1, i);   // :678
    double forcemag = std::sqrt(dot3(&V3(force, 1, i), &V3(force, 1, i)));
    double fonh = forcemag / A1(hh, i);
    if (fonh > fhmax && A1(itype, i) != 1) fhmax = fonh;        // :681
    double valfven2i_ = 0.;
    if (o.imhd != 0) valfven2i_ = dot3(&V3(Bfield, 1, i), &V3(Bfield, 1, i)) / A1(dens, i);   // :690
    double vsig2 = A1(spsound, i) * A1(spsound, i) + valfven2i_;
    double vsig = std::sqrt(vsig2);
    if (o.imhd >= 11) {                                         // :722-730
      for (int k = 1; k <= 3; k++) V3(dBevoldt, k, i) = A1(sqrtg, i) * V3(dBevoldt, k, i) + V3(Bevol, k, i) * rho1i * A1(drhodt, i);
      if (o.idivbzero >= 2) {
        for (int k = 1; k <= 3; k++) V3(gradpsi, k, i) = V3(gradpsi, k, i) * rhoi;
        if (o.nsubsteps_divB <= 0) for (int k = 1; k <= 3; k++) V3(dBevoldt, k, i) = V3(dBevoldt, k, i) + V3(gradpsi, k, i);
      }
    } else if (o.imhd >= 1 && o.imhd <= 9) {                    // :733-752
      for (int k = 1; k <= 3; k++) V3(dBevoldt, k, i) = A1(sqrtg, i) * V3(dBevoldt, k, i) * rho1i;
      if (o.idivbzero >= 2) for (int k = 1; k <= 3; k++) V3(gradpsi, k, i) = V3(gradpsi, k, i) * (rho1i * rho1i);
    } else {
      for (int k = 1; k <= 3; k++) V3(dBevoldt, k, i) = 0.;     // :802-803
    }
    if (o.iresist > 0 && o.iresist != 2 && o.etamhd > DBL_MIN) S.dtforce = std::min(S.dtforce, A1(hh, i) * A1(hh, i) / o.etamhd);   // :808-815
    if (o.iener == 3) {                                         // :820-826
      A1(dudt, i) = A1(dudt, i) + A1(pr, i) * (rho1i * rho1i) * A1(drhodt, i);
      A1(dendt, i) = dot3(&V3(vel, 1, i), &V3(force, 1, i)) + A1(dudt, i);
    } else if (o.iener > 0 && o.iav >= 0 && o.idust != 1) {     // :832-835
      A1(dudt, i) = A1(dudt, i) + A1(pr, i) * (rho1i * rho1i) * A1(drhodt, i);
      A1(dendt, i) = A1(dudt, i);
    } else {
      A1(dendt, i) = A1(dudt, i);                               // :837
    }
    if (A1(itype, i) == ND_ITYPE_DUST) A1(dendt, i) = 0.;       // :839
    for (int k = 1; k <= 3; k++) V3(daldt, k, i) = 0.;          // :844
    if (o.iavlim[0] != 0 || o.iavlim[1] != 0 || o.iavlim[2] != 0) {
      double tdecay1 = (o.avdecayconst * vsig) / A1(hh, i);     // :846
      if (o.iavlim[0] == 1 || o.iavlim[0] == 2) {               // :850-854
        double source = std::max(A1(drhodt, i) * rho1i, 0.0);
        if (o.iavlim[0] == 2) source = source * (2.0 - V3(alpha, 1, i));
        V3(daldt, 1, i) = (o.alphamin - V3(alpha, 1, i)) * tdecay1 + o.avfact * source;
      } else if (o.iavlim[0] == 3) {                            // :855-859
        double graddivvmag = std::sqrt(dot3(&V3(graddivv, 1, i), &V3(graddivv, 1, i)));
        double source = A1(hh, i) * graddivvmag * (2.0 - V3(alpha, 1, i));
        V3(daldt, 1, i) = (o.alphamin - V3(alpha, 1, i)) * tdecay1 + o.avfact * source;
      }
      if (o.iener > 0 && o.iavlim[1] > 0) {                     // :864-873
        double sourceu;
        if (A1(uu, i) > DBL_EPSILON) sourceu = A1(hh, i) * std::fabs(del2u[i - 1]) / std::sqrt(A1(uu, i));
        else sourceu = 0.;
        V3(daldt, 2, i) = (o.alphaumin - V3(alpha, 2, i)) * tdecay1 + sourceu;
      }
      if (o.iavlim[2] != 0 && o.imhd != 0) {                    // :877-895
        double sourceJ = std::sqrt(dot3(&V3(curlB, 1, i), &V3(curlB, 1, i)) * rho1i);
        double sourcedivB = 10. * std::fabs(A1(divB, i)) * std::sqrt(rho1i);
        double sourceB = std::max(sourceJ, sourcedivB);
        if (o.iavlim[2] == 2) sourceB = sourceB * (2.0 - V3(alpha, 3, i));
        else if (o.iavlim[2] == 3) { double source = std::max(A1(drhodt, i) * rho1i, 0.0) * (2. - V3(alpha, 3, i)); sourceB = std::sqrt(source * sourceB); }
        V3(daldt, 3, i) = (o.alphaBmin - V3(alpha, 3, i)) * tdecay1 + sourceB;
      }
    }
    if (o.idivbzero >= 2 && o.idivbzero <= 7) A1(dpsidt, i) = -S.vsig2max * A1(divB, i) - o.psidecayfact * A1(psi, i) * vsigmax / A1(hh, i);   // :900-906
    else A1(dpsidt, i) = 0.;
  }
  if (fhmax < 0.) return fail(S, ND_ERR_INVALID_ARG, "rates: fhmax <= 0");
  else if (fhmax > 0.) S.dtforce = std::min(S.dtforce, std::sqrt(1. / fhmax));   // :938-943
  S.fhmax_out = fhmax;
  for (int k = 0; k < 3; k++) S.fmean[k] = fmean[k];
  for (int i = 1; i <= ntotal; i++) {                           // :949-965
    int t = A1(itype, i);
    if (t == ND_ITYPE_BND || t == ND_ITYPE_BNDDUST || i > npart) {
      for (int k = 1; k <= 3; k++) { V3(force, k, i) = 0.; V3(dBevoldt, k, i) = 0.; V3(daldt, k, i) = 0.; V3(fmag, k, i) = 0.; V3(xsphterm, k, i) = 0.; V3(gradpsi, k, i) = 0.; }
      A1(drhodt, i) = 0.; A1(dhdt, i) = 0.; A1(dudt, i) = 0.; A1(dendt, i) = 0.; A1(dpsidt, i) = 0.; A1(divB, i) = 0.;
      if (o.imhd >= 0) for (int k = 1; k <= 3; k++) V3(curlB, k, i) = 0.;
    }
  }
  return 0;
}

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

int init_state(St &S, const nd_options *o, int ndim, ndo_arrays *a, int npart, int ntotal, int idim) {
  S.o = o; S.ndim = ndim; S.npart = npart; S.ntotal = ntotal; S.idim = idim; S.a = *a; S.err = 0;
  S.hhmax = o->hhmax;
  S.itsdensity = 0; S.ncalctotal = 0;
  S.dtcourant = S.dtforce = S.dtav = S.dtdrag = S.dtvisc = S.vsig2max = 0.;
  S.vsigmax_out = S.stressmax_out = S.ts_min_out = S.h_on_csts_max_out = S.fhmax_out = 0.;
  S.fmean[0] = S.fmean[1] = S.fmean[2] = 0.; S.nclumped = 0;
  S.ncellsx[0] = S.ncellsx[1] = S.ncellsx[2] = 1; S.ncells = 0; S.ncellsloop = 0; S.dxcell = 0.;
  if (ndim < 1 || ndim > 3 || npart < 1 || ntotal < npart || idim < ntotal) return fail(S, ND_ERR_INVALID_ARG, "bad sizes");
  int ikalt = o->ikernelalt;
  return setkernels(S.K, o->ikernel, ikalt, o->idust, ndim);
}

}  // namespace

extern "C" {

const char *ndo_last_error(void) { return g_err.c_str(); }

int ndo_derivs(const nd_options *o, int ndim, ndo_arrays *a, int npart, int *ntotal, int idim, int phases, nd_scalars *s, double *ms) {
  static St S;   // large tables: keep off the stack
  g_err.clear();
  if (int e = init_state(S, o, ndim, a, npart, *ntotal, idim)) { if (g_err.empty()) g_err = "unsupported kernel"; return e; }
  double t[6];
  int e = 0;
  t[0] = now_ms();
  bool anygt1 = false;
  for (int d = 0; d < ndim; d++) if (o->ibound[d] >= 2) anygt1 = true;
  if ((phases & NDO_GHOSTS) && anygt1) e = set_ghost_particles(S);          // derivs.f90:78
  t[1] = now_ms();
  if (!e && (phases & (NDO_LINK | NDO_DENSITY | NDO_RATES))) e = set_linklist(S);   // derivs.f90:82
  t[2] = now_ms();
  if (!e && (phases & NDO_DENSITY) && o->icty <= 0) e = iterate_density(S); // derivs.f90:92
  t[3] = now_ms();
  if (!e && (phases & NDO_C2P)) e = conservative2primitive(S);              // derivs.f90:98
  t[4] = now_ms();
  if (!e && (phases & NDO_RATES)) e = get_rates(S, nullptr, nullptr);       // derivs.f90:156
  t[5] = now_ms();
  if (ms) for (int k = 0; k < 5; k++) ms[k] = t[k + 1] - t[k];
  *ntotal = S.ntotal;
  if (s) {
    memset(s, 0, sizeof(*s));
    s->dtcourant = S.dtcourant; s->dtforce = S.dtforce; s->dtav = S.dtav; s->dtdrag = S.dtdrag; s->dtvisc = S.dtvisc;
    s->vsig2max = S.vsig2max; s->vsigmax = S.vsigmax_out; s->stressmax = S.stressmax_out; s->ts_min = S.ts_min_out;
    s->h_on_csts_max = S.h_on_csts_max_out; s->fhmax = S.fhmax_out; s->hhmax = S.hhmax; s->dxcell = S.dxcell;
    for (int k = 0; k < 3; k++) { s->fmean[k] = S.fmean[k]; s->ncellsx[k] = S.ncellsx[k]; }
    s->itsdensity = S.itsdensity; s->nclumped = S.nclumped; s->ntotal = S.ntotal; s->ncells = S.ncells;
    s->ncalctotal = S.ncalctotal;
    s->npairs_rates = S.npairs_rates;
    int mn = 1 << 30, mx = 0;
    for (int i = 0; i < npart; i++) { mn = std::min(mn, a->numneigh[i]); mx = std::max(mx, a->numneigh[i]); }
    s->nneigh_min = mn; s->nneigh_max = mx;
  }
  return e;
}

int ndo_kernel_tables(int ikernel, int ikerneldrag, int ndim, double *wij, double *grwij, double *grgrwij, double *wijdrag,
                      double *radkern2, double *dq2table) {
  static Kern K;
  K.radkern = 2.;
  int e = setkerntable(K, ikernel, ndim, K.wij, K.grwij, K.grgrwij);
  if (e) return e;
  if (ikerneldrag > 0) { e = setkerntable(K, ikerneldrag, ndim, K.wijdrag, K.grwijdrag, K.grgrwijdrag); if (e) return e; }
  for (int i = 0; i <= ikern; i++) {
    if (wij) wij[i] = K.wij[i];
    if (grwij) grwij[i] = K.grwij[i];
    if (grgrwij) grgrwij[i] = K.grgrwij[i];
    if (wijdrag) wijdrag[i] = (ikerneldrag > 0) ? K.wijdrag[i] : 0.;
  }
  if (radkern2) *radkern2 = K.radkern2;
  if (dq2table) *dq2table = K.dq2table;
  return 0;
}

int ndo_interpolate(int ikernel, int ndim, double q2, double *w, double *grw, double *grgrw) {
  static Kern K;
  static int have_k = -1, have_d = -1;
  if (have_k != ikernel || have_d != ndim) {
    int e = setkernels(K, ikernel, ikernel, 0, ndim);
    if (e) return e;
    have_k = ikernel; have_d = ndim;
  }
  double walt, gwalt;
  interpolate_kernels_dens(K, q2, *w, *grw, *grgrw, walt, gwalt);
  return 0;
}

// src/random.f90:61-96 ran1
double ndo_ran1(int *iseed) {
  const int ia = 16807, im = 2147483647, iq = 127773, ir = 2836, ntab = 32, ndiv = 1 + (im - 1) / ntab;
  static int iv[32] = {0};
  static int iy = 0;
  const double am = 1. / im, eps = 1.2e-7, floatmax = 1. - eps;
  int j, k;
  if (*iseed <= 0 || iy == 0) {
    *iseed = std::max(-*iseed, 1);
    for (j = ntab + 8; j >= 1; j--) {
      k = *iseed / iq;
      *iseed = ia * (*iseed - k * iq) - ir * k;
      if (*iseed < 0) *iseed = *iseed + im;
      if (j <= ntab) iv[j - 1] = *iseed;
    }
    iy = iv[0];
  }
  k = *iseed / iq;
  *iseed = ia * (*iseed - k * iq) - ir * k;
  if (*iseed < 0) *iseed = *iseed + im;
  j = 1 + iy / ndiv;
  iy = iv[j - 1];
  iv[j - 1] = *iseed;
  return std::min(am * iy, floatmax);
}

long long ndo_bruteforce_pairs(int ndim, const double *x, const double *hh, int npart, int ntotal, double radkern2, int *pi, int *pj, long long cap) {
  // src/check_neighbourlist.f90:149-173 logic: every pair within range of either particle, at least one real
  long long n = 0;
  std::vector<double> h21(ntotal);
  for (int i = 0; i < ntotal; i++) { double h1 = 1. / hh[i]; h21[i] = h1 * h1; }
  for (int i = 0; i < ntotal; i++) {
    for (int j = i + 1; j < ntotal; j++) {
      if (i >= npart && j >= npart) continue;
      double rij2 = 0.;
      for (int k = 0; k < ndim; k++) { double d = x[(size_t)i * ndim + k] - x[(size_t)j * ndim + k]; rij2 = rij2 + d * d; }
      if ((rij2 * h21[i] < radkern2) || (rij2 * h21[j] < radkern2)) {
        if (n < cap) { pi[n] = i + 1; pj[n] = j + 1; }
        n++;
      }
    }
  }
  return n;
}

long long ndo_linklist_pairs(const nd_options *o, int ndim, ndo_arrays *a, int npart, int ntotal, int idim, int *pi, int *pj, long long cap) {
  static St S;
  g_err.clear();
  if (init_state(S, o, ndim, a, npart, ntotal, idim)) return -1;
  if (set_linklist(S)) return -1;
  // run the rates pair loop on scratch copies so the caller's arrays are untouched except rates outputs
  std::vector<int> vi, vj;
  if (get_rates(S, &vi, &vj)) return -1;
  long long n = (long long)vi.size();
  for (long long k = 0; k < n && k < cap; k++) { pi[k] = vi[k]; pj[k] = vj[k]; }
  return n;
}

/* src/get_curl.f90:64-287 on the arrays as they stand (rho, hh, gradh of a previous derivs; rows npart+1..ntotal are the ghosts of that
 * derivs): re-links and runs the operator.  Bvec, curlB (3,idim); gradB (3,3,idim) or NULL (icurltype 1 only). */
int ndo_get_curl(const nd_options *o, int ndim, ndo_arrays *a, int npart, int ntotal, int idim, int icurltype, const double *Bvec, double *curlB,
                 double *gradB) {
  static St S;
  g_err.clear();
  if (int e = init_state(S, o, ndim, a, npart, ntotal, idim)) return e;
  if (int e = set_linklist(S)) return e;
  get_curl(S, icurltype, Bvec, curlB, icurltype == 1 ? gradB : nullptr);
  return 0;
}

/* One leapfrog step: `step` (src/stepND_leapfrog_mhd.f90:39-300) with its `call derivs` (:167) and `call boundary`
 * (:216 -> src/boundaryND.f90:65-93, periodic wrap only), restated loop for loop.  On entry the rates arrays hold the
 * result of the previous derivs.  Not supported (error 2): itypebnd2 rows (cylindrical fixed particles), imhd < 0,
 * idivbzero = 10, particle splitting, idustevol /= 0. */
int ndo_step(const nd_options *o, int ndim, ndo_arrays *a, int npart, int *ntotal, int idim, double *dt_inout, double C_cour, double C_force,
             int dtfixed, nd_scalars *s) {
  g_err.clear();
  if (o->imhd < 0 || o->idivbzero == 10 || o->idustevol != 0) { g_err = "ndo_step: unsupported option"; return ND_ERR_UNSUPPORTED_OPTION; }
  const double dt = *dt_inout, hdt = 0.5 * dt;                                            // :68
  const double dndim = 1. / ndim;
  const bool onef = o->onef_dust != 0;
  const size_t N = (size_t)npart;
  std::vector<double> xin(N * ndim), velin(N * 3), Bevolin(N * 3), rhoin(N), hhin(N), enin(N), alphain(N * 3), psiin(N), forcein(N * 3),
      dBevoldtin(N * 3), drhodtin(N), dhdtin(N), dendtin(N), daldtin(N * 3), dpsidtin(N), dustevolin, ddustevoldtin, deltavin, ddeltavdtin;
  if (onef) { dustevolin.resize(N); ddustevoldtin.resize(N); deltavin.resize(N * 3); ddeltavdtin.resize(N * 3); }
  for (int i = 0; i < npart; i++) {                                                       // :70-100
    for (int d = 0; d < ndim; d++) xin[(size_t)i * ndim + d] = a->x[(size_t)i * ndim + d];
    for (int d = 0; d < 3; d++) {
      velin[(size_t)i * 3 + d] = a->vel[(size_t)i * 3 + d]; Bevolin[(size_t)i * 3 + d] = a->Bevol[(size_t)i * 3 + d];
      alphain[(size_t)i * 3 + d] = a->alpha[(size_t)i * 3 + d]; forcein[(size_t)i * 3 + d] = a->force[(size_t)i * 3 + d];
      dBevoldtin[(size_t)i * 3 + d] = a->dBevoldt[(size_t)i * 3 + d]; daldtin[(size_t)i * 3 + d] = a->daldt[(size_t)i * 3 + d];
    }
    rhoin[i] = a->rho[i]; hhin[i] = a->hh[i]; enin[i] = a->en[i]; psiin[i] = a->psi[i];
    drhodtin[i] = a->drhodt[i]; dhdtin[i] = a->dhdt[i]; dendtin[i] = a->dendt[i]; dpsidtin[i] = a->dpsidt[i];
    if (onef) {
      dustevolin[i] = a->dustevol[i]; ddustevoldtin[i] = a->ddustevoldt[i];
      if (o->idust == 1) for (int d = 0; d < 3; d++) { deltavin[(size_t)i * 3 + d] = a->deltav[(size_t)i * 3 + d]; ddeltavdtin[(size_t)i * 3 + d] = a->ddeltavdt[(size_t)i * 3 + d]; }
    }
  }
  // ---- predictor :108-163 ----
  for (int i = 0; i < npart; i++) {
    const int it = a->itype[i];
    if (it == ND_ITYPE_BND || it == ND_ITYPE_BND2 || it == ND_ITYPE_BNDDUST) {
      if (it == ND_ITYPE_BND2) { g_err = "ndo_step: itypebnd2 not supported"; return ND_ERR_UNSUPPORTED_OPTION; }
      const int j = a->ireal[i];                                                          // 1-based
      const size_t r = (j > 0) ? (size_t)(j - 1) : (size_t)i;                             // :112-117
      for (int d = 0; d < ndim; d++) a->x[(size_t)i * ndim + d] = xin[(size_t)i * ndim + d] + dt * velin[r * 3 + d] + 0.5 * dt * dt * forcein[r * 3 + d];
      for (int d = 0; d < 3; d++) { a->Bevol[(size_t)i * 3 + d] = Bevolin[(size_t)i * 3 + d]; a->alpha[(size_t)i * 3 + d] = alphain[(size_t)i * 3 + d]; }
      a->rho[i] = rhoin[i]; a->hh[i] = hhin[i]; a->en[i] = enin[i]; a->psi[i] = psiin[i];  // :135-139
      if (o->idust == 1 || o->idust == 3 || o->idust == 4) {
        if (onef) a->dustevol[i] = dustevolin[i];
        if (o->idust == 1) for (int d = 0; d < 3; d++) a->deltav[(size_t)i * 3 + d] = deltavin[(size_t)i * 3 + d];
      }
    } else {
      for (int d = 0; d < ndim; d++) a->x[(size_t)i * ndim + d] = xin[(size_t)i * ndim + d] + dt * velin[(size_t)i * 3 + d] + 0.5 * dt * dt * forcein[(size_t)i * 3 + d];   // :145
      for (int d = 0; d < 3; d++) a->vel[(size_t)i * 3 + d] = (velin[(size_t)i * 3 + d] + dt * forcein[(size_t)i * 3 + d]) / (1. + o->damp);                           // :146
      if (o->imhd != 0 && o->iresist != 2) for (int d = 0; d < 3; d++) a->Bevol[(size_t)i * 3 + d] = Bevolin[(size_t)i * 3 + d] + dt * dBevoldtin[(size_t)i * 3 + d];
      if (o->icty >= 1) a->rho[i] = rhoin[i] + dt * drhodtin[i];
      if (o->ihvar == 1) a->hh[i] = hhin[i] * pow(rhoin[i] / a->rho[i], dndim);            // :151
      else if (o->ihvar == 2 || o->ihvar == 3) a->hh[i] = hhin[i] + dt * dhdtin[i];
      if (o->iener != 0) a->en[i] = enin[i] + dt * dendtin[i];
      for (int d = 0; d < 3; d++) if (o->iavlim[d] != 0) a->alpha[(size_t)i * 3 + d] = std::min(alphain[(size_t)i * 3 + d] + dt * daldtin[(size_t)i * 3 + d], 1.0);
      if (o->idivbzero >= 2) a->psi[i] = psiin[i] + dt * dpsidtin[i];
      if (onef) {
        a->dustevol[i] = dustevolin[i] + dt * ddustevoldtin[i];
        if (o->idust == 1) for (int d = 0; d < 3; d++) a->deltav[(size_t)i * 3 + d] = deltavin[(size_t)i * 3 + d] + dt * ddeltavdtin[(size_t)i * 3 + d];
      }
    }
  }
  // ---- derivs :167; it opens with `if (any(ibound.ne.0)) call boundary` (src/derivs.f90:74), so particles the predictor moved out of a
  //      periodic domain are wrapped before the ghosts are made (set_ghost_particles makes no ghost of a particle on or over the
  //      boundary, src/ghostND_mhd.f90:204) ----
  auto boundary = [&]() {                                                                 // src/boundaryND.f90:65-93
    bool any3 = false;
    for (int d = 0; d < ndim; d++) if (o->ibound[d] == 3) any3 = true;
    if (!any3) return;
    for (int i = 0; i < npart; i++) for (int d = 0; d < ndim; d++) if (o->ibound[d] == 3) {
      double &xx = a->x[(size_t)i * ndim + d];
      if (xx > o->xmax[d]) xx = o->xmin[d] + xx - o->xmax[d];
      else if (xx < o->xmin[d]) xx = o->xmax[d] - (o->xmin[d] - xx);
    }
  };
  boundary();
  if (int e = ndo_derivs(o, ndim, a, npart, ntotal, idim, NDO_ALL, s, nullptr)) return e;
  // ---- corrector :171-209 ----
  for (int i = 0; i < npart; i++) {
    const int it = a->itype[i];
    if (it == ND_ITYPE_BND || it == ND_ITYPE_BND2 || it == ND_ITYPE_BNDDUST) {
      for (int d = 0; d < 3; d++) { a->vel[(size_t)i * 3 + d] = velin[(size_t)i * 3 + d]; a->Bevol[(size_t)i * 3 + d] = Bevolin[(size_t)i * 3 + d]; a->alpha[(size_t)i * 3 + d] = alphain[(size_t)i * 3 + d]; }
      a->rho[i] = rhoin[i]; a->hh[i] = hhin[i]; a->en[i] = enin[i]; a->psi[i] = psiin[i];
      if (o->idust == 1 || o->idust == 3 || o->idust == 4) {
        if (onef) a->dustevol[i] = dustevolin[i];
        if (o->idust == 1) for (int d = 0; d < 3; d++) a->deltav[(size_t)i * 3 + d] = deltavin[(size_t)i * 3 + d];
      }
    } else {
      for (int d = 0; d < 3; d++) a->vel[(size_t)i * 3 + d] = (velin[(size_t)i * 3 + d] + hdt * (a->force[(size_t)i * 3 + d] + forcein[(size_t)i * 3 + d])) / (1. + o->damp);   // :187
      if (o->imhd != 0) {
        for (int d = 0; d < 3; d++) {
          if (o->iresist == 2) a->Bevol[(size_t)i * 3 + d] = Bevolin[(size_t)i * 3 + d] + dt * a->dBevoldt[(size_t)i * 3 + d];
          else a->Bevol[(size_t)i * 3 + d] = Bevolin[(size_t)i * 3 + d] + hdt * (a->dBevoldt[(size_t)i * 3 + d] + dBevoldtin[(size_t)i * 3 + d]);
        }
      }
      if (o->icty >= 1) a->rho[i] = rhoin[i] + hdt * (a->drhodt[i] + drhodtin[i]);
      if (o->ihvar == 2) {
        a->hh[i] = hhin[i] + hdt * (a->dhdt[i] + dhdtin[i]);
        if (a->hh[i] <= 0.) { g_err = "step: hh -ve"; return ND_ERR_H_NONPOSITIVE; }        // :197-200
      }
      if (o->iener != 0) a->en[i] = enin[i] + hdt * (a->dendt[i] + dendtin[i]);
      for (int d = 0; d < 3; d++) if (o->iavlim[d] != 0) a->alpha[(size_t)i * 3 + d] = std::min(alphain[(size_t)i * 3 + d] + hdt * (a->daldt[(size_t)i * 3 + d] + daldtin[(size_t)i * 3 + d]), 1.0);
      if (o->idivbzero >= 2) a->psi[i] = psiin[i] + hdt * (a->dpsidt[i] + dpsidtin[i]);
      if (onef) {
        a->dustevol[i] = dustevolin[i] + hdt * (a->ddustevoldt[i] + ddustevoldtin[i]);
        if (o->idust == 1) for (int d = 0; d < 3; d++) a->deltav[(size_t)i * 3 + d] = deltavin[(size_t)i * 3 + d] + hdt * (a->ddeltavdt[(size_t)i * 3 + d] + ddeltavdtin[(size_t)i * 3 + d]);
      }
    }
  }
  boundary();                                                                             // :216
  // ---- new timestep :239-253 ----
  if (!dtfixed) *dt_inout = std::min(std::min(C_force * s->dtforce, C_cour * s->dtcourant), std::min(0.9 * s->dtdrag, C_force * s->dtvisc));
  return 0;
}

/* the particle loop and totals of `evwrite` (src/evwrite_mhd.f90:124-284), serial, in the reference's order; fmag-based columns and
 * epot are not produced (see include/ndspmhd_b200.h) */
int ndo_evwrite(const nd_options *o, int ndim, const ndo_arrays *a, int npart, nd_evwrite *ev) {
  memset(ev, 0, sizeof(*ev));
  double ekin = 0., etherm = 0., emag = 0., emagp = 0., ekiny = 0., mgas = 0., mdust = 0.;
  double mom[3] = {0, 0, 0}, dmom[3] = {0, 0, 0}, ang[3] = {0, 0, 0}, flux[3] = {0, 0, 0};
  double betaav = 0., betamin = 1.7976931348623157e308, divBmax = 0., divBav = 0., divBtot = 0., omegaav = 0., omegamax = 0., fracok = 0., crosshel = 0.;
  double rhomin = 1.7976931348623157e308, rhomax = 0., rhosum = 0.;
  for (int i = 0; i < npart; i++) {
    const double m = a->pmass[i], rhoi = a->rho[i];
    const double *v = a->vel + (size_t)i * 3, *f = a->force + (size_t)i * 3;
    double x[3] = {0, 0, 0};
    for (int d = 0; d < ndim; d++) x[d] = a->x[(size_t)i * ndim + d];
    for (int d = 0; d < 3; d++) { mom[d] += m * v[d]; dmom[d] += m * f[d]; }
    if (ndim == 3) { ang[0] += m * (x[1] * v[2] - x[2] * v[1]); ang[1] += m * (x[2] * v[0] - x[0] * v[2]); ang[2] += m * (x[0] * v[1] - x[1] * v[0]); }
    else if (ndim == 2) ang[2] += m * (x[0] * v[1] - x[1] * v[0]);
    ekin += 0.5 * m * ((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    if (o->onef_dust) {
      const double df = a->dustfrac[i], dterm = 1. - df;
      const double *dv = a->deltav + (size_t)i * 3;
      const double ekdv = 0.5 * m * df * dterm * ((dv[0] * dv[0] + dv[1] * dv[1]) + dv[2] * dv[2]);
      ekin += ekdv; ekiny += ekdv; etherm += m * a->uu[i] * dterm; mgas += m * dterm; mdust += m * df;
    } else {
      if (ndim >= 2) ekiny += 0.5 * m * v[0] * v[0];
      etherm += m * a->uu[i];
    }
    rhosum += rhoi; rhomin = std::min(rhomin, rhoi); rhomax = std::max(rhomax, rhoi);
    if (o->imhd != 0) {
      const double *B = a->Bfield + (size_t)i * 3;
      const double B2 = (B[0] * B[0] + B[1] * B[1]) + B[2] * B[2], Bmag = sqrt(B2), divBi = fabs(a->divB[i]);
      emag += 0.5 * m * B2 / rhoi; emagp += 0.5 * m * (B[0] * B[0] + B[1] * B[1]) / rhoi;
      const double beta = (B2 < 2.2250738585072014e-308) ? 0. : a->pr[i] / (0.5 * B2);
      betaav += beta; if (beta < betamin) betamin = beta;
      if (divBi > divBmax) divBmax = divBi;
      divBav += divBi; divBtot += m * divBi / rhoi;
      const double omega = (Bmag < 1e-8) ? 0. : divBi * a->hh[i] / Bmag;
      if (omega < 1.e-2) fracok += 1.;
      if (omega > omegamax) omegamax = omega;
      omegaav += omega;
      for (int d = 0; d < 3; d++) flux[d] += m * (B[d] / rhoi);
      crosshel += m * ((v[0] * (B[0] / rhoi) + v[1] * (B[1] / rhoi)) + v[2] * (B[2] / rhoi));
    }
  }
  ev->ekin = ekin; ev->etherm = etherm; ev->emag = emag; ev->emagp = emagp; ev->epot = 0.; ev->ekiny = ekiny; ev->totmassgas = mgas; ev->totmassdust = mdust;
  for (int d = 0; d < 3; d++) { ev->mom[d] = mom[d]; ev->dmom[d] = dmom[d]; ev->ang[d] = ang[d]; ev->fluxtot[d] = flux[d]; }
  ev->etot = ekin + emag + 0.;
  if (o->iprterm >= 0 || o->iprterm < -1) ev->etot += etherm;
  ev->momtot = sqrt((mom[0] * mom[0] + mom[1] * mom[1]) + mom[2] * mom[2]);
  ev->dmomtot = sqrt((dmom[0] * dmom[0] + dmom[1] * dmom[1]) + dmom[2] * dmom[2]);
  ev->angtot = sqrt((ang[0] * ang[0] + ang[1] * ang[1]) + ang[2] * ang[2]);
  ev->rhomin = rhomin; ev->rhomax = rhomax; ev->rhomean = rhosum / npart;
  if (o->imhd != 0) {
    ev->fluxtotmag = sqrt((flux[0] * flux[0] + flux[1] * flux[1]) + flux[2] * flux[2]);
    ev->betamhdav = betaav / npart; ev->betamhdmin = betamin; ev->fracdivBok = 100. * fracok / npart; ev->omegamhdav = omegaav / npart;
    ev->omegamhdmax = omegamax; ev->divBav = divBav / npart; ev->divBmax = divBmax; ev->divBtot = divBtot; ev->crosshel = crosshel;
  }
  return 0;
}

}  // extern "C"
