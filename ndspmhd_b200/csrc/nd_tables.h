// Host-side construction of the SPH kernel tables the device interpolates from.
//
// The reference tabulates W, dW/dq, d2W/dq2 on 4001 points uniform in q^2 and interpolates linearly
// (src/kernelND.f90:39-48, :4426-4469).  Parity to 1e-12 needs the same tables, including the reference's
// truncated pi (src/kernelND.f90:41) and 1D cubic normalisation 0.66666666666 (:4154).  Kernels are described here
// as piecewise data (breakpoints + evaluators) rather than as the reference's select-case.
#pragma once
#include <cmath>
#include <algorithm>

namespace ndt {

constexpr int IKERN = 4000;
constexpr double PI_TRUNC = 3.141592653589;

struct KernelTables {
  double w[IKERN + 1], grw[IKERN + 1], grgrw[IKERN + 1];
  double wdrag[IKERN + 1];
  double radkern, radkern2, dq2table, ddq2table;
  int ikernel, ikerneldrag, ndim;
};

namespace detail {
inline double p2(double x) { return x * x; }
inline double p3(double x) { return (x * x) * x; }
inline double p4(double x) { double y = x * x; return y * y; }
inline double p5(double x) { double y = x * x; return (y * x) * y; }

struct Shape {
  double radius;          // compact support in q
  double norm[3];         // normalisation in 1, 2, 3 dimensions
  // writes un-normalised w, w', w'' at (q, q2)
  void (*eval)(double q, double q2, double &w, double &g, double &gg);
};

inline void cubic(double q, double q2, double &w, double &g, double &gg) {          // kernelND.f90:4171-4184
  if (q < 1.0) { w = 1. - 1.5 * q2 + 0.75 * q * q2; g = -3. * q + 2.25 * q2; gg = -3. + 4.5 * q; }
  else if (q <= 2.0) { w = 0.25 * p3(2. - q); g = -0.75 * p2(2. - q); gg = 1.5 * (2. - q); }
  else { w = 0.; g = 0.; gg = 0.; }
}
inline void quartic(double q, double q2, double &w, double &g, double &gg) {        // kernelND.f90:171-196
  if (q < 0.5) { w = 6.0 * q2 * q2 - 15.0 * q2 + 14.375; g = q * (24.0 * q2 - 30.0); gg = 72.0 * q2 - 30.0; }
  else if (q < 1.5) { w = -5.0 * p4(-q + 1.5) + p4(-q + 2.5); g = -16.0 * q2 * q + 60.0 * q2 - 60.0 * q + 5.0; gg = -48.0 * q2 + 120.0 * q - 60.0; }
  else if (q < 2.5) { w = p4(-q + 2.5); g = -4.0 * p3(-q + 2.5); gg = 12.0 * q2 - 60.0 * q + 75.0; }
  else { w = 0.; g = 0.; gg = 0.; }
}
inline void quintic(double q, double q2, double &w, double &g, double &gg) {        // kernelND.f90:224-262
  double q4 = q2 * q2;
  double term1 = -5. * std::pow(3. - q, 4.);
  if (q < 1.0) { w = 66. - 60. * q2 + 30. * q4 - 10. * q4 * q; g = term1 + 30 * p4(2. - q) - 75. * p4(1. - q); gg = 20. * p3(3. - q) - 120. * p3(2. - q) + 300. * p3(1. - q); }
  else if (q < 2.0) { w = p5(3. - q) - 6. * p5(2. - q); g = term1 + 30 * p4(2. - q); gg = 20. * p3(3. - q) - 120. * p3(2. - q); }
  else if (q <= 3.0) { w = p5(3. - q); g = term1; gg = 20. * p3(3. - q); }
  else { w = 0.; g = 0.; gg = 0.; }
}
inline void hump_cubic(double q, double q2, double &w, double &g, double &gg) {     // kernelND.f90:1683-1696
  double q4 = q2 * q2;
  if (q < 1.0) { w = q2 - 1.5 * q4 + 0.75 * q4 * q; g = 2. * q - 6. * q2 * q + 3.75 * q4; gg = 2. - 18. * q2 + 15. * q2 * q; }
  else if (q <= 2.0) { w = 0.25 * q2 * p3(2. - q); g = 0.5 * q * p3(2. - q) - 0.75 * q2 * p2(2. - q); gg = 0.5 * p3(2. - q) - 3. * q * p2(2. - q) + 1.5 * q2 * (2. - q); }
  else { w = 0.; g = 0.; gg = 0.; }
}
inline void hump_quartic(double q, double q2, double &w, double &g, double &gg) {   // kernelND.f90:1722-1752
  double q4 = q2 * q2;
  if (q < 0.5) { w = q2 * (6.0 * q4 - 15.0 * q2 + 14.375); g = q * (36.0 * q4 - 60.0 * q2 + 28.75); gg = 180.0 * q4 - 180.0 * q2 + 28.75; }
  else if (q < 1.5) { w = q2 * (p4(q - 2.5) - 5.0 * p4(q - 1.5)); g = q * (-24.0 * q4 + 100.0 * q2 * q - 120.0 * q2 + 15.0 * q + 27.5); gg = -120.0 * q4 + 400.0 * q2 * q - 360.0 * q2 + 30.0 * q + 27.5; }
  else if (q < 2.5) { w = q2 * p4(q - 2.5); g = q * p3(q - 2.5) * (6.0 * q - 5.0); gg = 30.0 * q4 - 200.0 * q2 * q + 450.0 * q2 - 375.0 * q + 78.125; }
  else { w = 0.; g = 0.; gg = 0.; }
}

inline bool shape_for(int ikernel, Shape &s) {
  switch (ikernel) {
    case 0: s = {2.0, {0.66666666666, 10. / (7. * PI_TRUNC), 1. / PI_TRUNC}, cubic}; return true;
    case 2: s = {2.5, {1. / 24., 96. / (1199. * PI_TRUNC), 0.05 / PI_TRUNC}, quartic}; return true;
    case 3: s = {3.0, {1. / 120., 7. / (478 * PI_TRUNC), 1. / (120. * PI_TRUNC)}, quintic}; return true;
    case 41: s = {2.0, {2.0, 70. / (31. * PI_TRUNC), 10. / (9. * PI_TRUNC)}, hump_cubic}; return true;
    case 42: s = {2.5, {1. / 10., 3584. / (35783. * PI_TRUNC), 1. / (23. * PI_TRUNC)}, hump_quartic}; return true;
    default: return false;
  }
}

inline bool fill(int ikernel, int ndim, double &radkern, double &radkern2, double &dq2, double *w, double *g, double *gg) {
  Shape s;
  if (!shape_for(ikernel, s)) return false;
  radkern = (ikernel == 42) ? s.radius : std::max(radkern, s.radius);   // kernelND.f90:159, :1706
  radkern2 = radkern * radkern;
  dq2 = radkern2 / double(IKERN);
  const double cn = s.norm[ndim - 1];
  for (int i = 0; i <= IKERN; i++) {
    double q2 = i * dq2, q = std::sqrt(q2), a, b, c;
    s.eval(q, q2, a, b, c);
    w[i] = cn * a; g[i] = cn * b; gg[i] = cn * c;                       // kernelND.f90:4279-4284
  }
  return true;
}
}  // namespace detail

// setkernels + setkerndrag (src/initialiseND_mhd.f90:179-216).  Returns false for kernels outside the compiled set.
inline bool build_tables(KernelTables &T, int ikernel, int ikernelalt, int idust, int ndim) {
  if (ikernelalt != ikernel) return false;   // number-density kernel must equal the main kernel (default, initialiseND_mhd.f90:186)
  T.ikernel = ikernel; T.ndim = ndim; T.ikerneldrag = 0;
  T.radkern = 2.0;
  if (!detail::fill(ikernel, ndim, T.radkern, T.radkern2, T.dq2table, T.w, T.grw, T.grgrw)) return false;
  for (int i = 0; i <= IKERN; i++) T.wdrag[i] = 0.;
  if (idust != 0) {
    int kd = ikernel == 0 ? 41 : ikernel == 2 ? 42 : -1;
    if (kd < 0) return false;
    static double g[IKERN + 1], gg[IKERN + 1];
    double r = T.radkern, r2, dq2;
    if (!detail::fill(kd, ndim, r, r2, dq2, T.wdrag, g, gg)) return false;
    if (r != T.radkern) return false;       // "drag kernel has different support radius"
    T.ikerneldrag = kd;
  }
  T.ddq2table = 1. / T.dq2table;            // kernelND.f90:4289
  return true;
}

}  // namespace ndt
