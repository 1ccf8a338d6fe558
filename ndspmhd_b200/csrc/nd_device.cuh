// Device-side building blocks shared by the density and rates kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ndk {

constexpr int IKERN = 4000;
constexpr unsigned FULL = 0xffffffffu;

// particle types, src/variablesND.f90:165-171
constexpr int T_GAS = 0, T_BND = 1, T_DUST = 2, T_BNDDUST = 12;

// One interpolation record per table index i: value at i and slope to i+1, for W and dW/dq.
// slope = (w[i+1]-w[i])*ddq2table is exactly the `dwdx` the reference recomputes per lookup
// (src/kernelND.f90:4446-4452); at i = ikern both clamp to the same node so the slope is 0.
struct __align__(32) TabRec { double w, dw, g, dg; };
struct __align__(16) TabRec2 { double gg, dgg; };

// Everything a pair kernel needs to find neighbours; passed by value.
struct Grid {
  const int *cellStart;     // [ncells+1] first sorted slot of each cell
  const int *cellOf;        // [ntotal]   cell of each sorted slot
  const int *perm;          // [ntotal]   sorted slot -> original row
  const double4 *posh;      // [ntotal]   {x,y,z,h}
  const double4 *vm;        // [ntotal]   {vx,vy,vz,m}
  const int *typ;           // [ntotal]
  int nx, ny, nz, ncells;
  int npart, ntotal;
  double radkern2, dq2table, ddq2table;
  const TabRec *tab;        // [IKERN+1]
  const TabRec2 *tab2;      // [IKERN+1]
  const double *tabdrag;    // [2*(IKERN+1)] {w, dw}
};

__device__ __forceinline__ double4 ld4(const double4 *p) {
  // two 128-bit read-only loads of one 32-byte aligned record
  const double2 *q = reinterpret_cast<const double2 *>(p);
  double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

// rij2 = dot_product(dx,dx) in index order without FMA contraction (src/density_sums.f90:181,
// src/ratesND_mhd.f90:408; reference build has no FMA, src/Makefile:25).  Unused dimensions carry exact zeros.
__device__ __forceinline__ double dist2_exact(double dx, double dy, double dz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// index = int(q2*ddq2table), clamped (src/kernelND.f90:4435-4438)
__device__ __forceinline__ int tab_index(double q2, double ddq2table) {
  double t = __dmul_rn(q2, ddq2table);
  int idx = (t < 2147483000.0 && t >= 0.0) ? __double2int_rz(t) : IKERN;
  return idx > IKERN ? IKERN : idx;
}

// w = w[index] + dwdx*(q2 - index*dq2table)   (src/kernelND.f90:4443-4455)
__device__ __forceinline__ void interp_wg(const Grid &G, double q2, double &w, double &g) {
  int idx = tab_index(q2, G.ddq2table);
  double4 r = ld4(reinterpret_cast<const double4 *>(G.tab + idx));
  double dxx = q2 - __dmul_rn((double)idx, G.dq2table);
  w = r.x + r.y * dxx;
  g = r.z + r.w * dxx;
}
__device__ __forceinline__ void interp_wggg(const Grid &G, double q2, double &w, double &g, double &gg) {
  int idx = tab_index(q2, G.ddq2table);
  double4 r = ld4(reinterpret_cast<const double4 *>(G.tab + idx));
  double2 r2 = __ldg(reinterpret_cast<const double2 *>(G.tab2 + idx));
  double dxx = q2 - __dmul_rn((double)idx, G.dq2table);
  w = r.x + r.y * dxx;
  g = r.z + r.w * dxx;
  gg = r2.x + r2.y * dxx;
}
__device__ __forceinline__ double interp_drag(const Grid &G, double q2) {
  int idx = tab_index(q2, G.ddq2table);
  double2 r = __ldg(reinterpret_cast<const double2 *>(G.tabdrag) + idx);
  double dxx = q2 - __dmul_rn((double)idx, G.dq2table);
  return r.x + r.y * dxx;
}

template <int NDIM> __device__ __forceinline__ double powndim(double x) {
  if (NDIM == 1) return x;
  if (NDIM == 2) return x * x;
  return (x * x) * x;   // gfortran x**3
}

// type filters ----------------------------------------------------------------------------------------
// src/density_sums.f90:169-172 and src/ratesND_mhd.f90:436-439
__device__ __forceinline__ bool types_interact(int ti, int tj) {
  return (tj == ti) || (ti == T_GAS && tj == T_BND) || (tj == T_GAS && ti == T_BND) || (ti == T_DUST && tj == T_BNDDUST) ||
         (tj == T_DUST && ti == T_BNDDUST);
}

// order-preserving double <-> u64 keys for atomicMin/Max
__device__ __forceinline__ unsigned long long dkey(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double dkey_inv(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  union { unsigned long long u; double d; } c; c.u = b; return c.d;
}
__device__ __forceinline__ void atomic_min_d(unsigned long long *addr, double v) { atomicMin(addr, dkey(v)); }
__device__ __forceinline__ void atomic_max_d(unsigned long long *addr, double v) { atomicMax(addr, dkey(v)); }

__device__ __forceinline__ double warp_min(double v) {
  for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------------
// Neighbour walk.  Every lane owns one target particle.  The 3^NDIM stencil of the target's cell is visited as
// 3^(NDIM-1) x-rows; a row's three cells are contiguous in the cell-sorted arrays, so a row is one span of slots.
// Phase 1 (cull): the lane tests the span's candidates and appends accepted slots to its private column of a
// shared-memory list.  Phase 2 (flush): when any lane's column is full (and at the end) the warp runs the
// expensive pair body over the lists, so the body executes with all lanes busy instead of at the ~15 % acceptance
// rate of the raw stencil.  All lanes of a warp must call this together (warp votes inside).
// ------------------------------------------------------------------------------------------------------
template <int NDIM, int CAP, int BLOCK, class Cull, class Body>
__device__ __forceinline__ void neighbour_walk(const Grid &G, bool active, int cell, unsigned *list /* [CAP][BLOCK] */, Cull cull, Body body) {
  const int tid = threadIdx.x;
  int cnt = 0;
  int ix = 0, iy = 0, iz = 0;
  if (active) {
    ix = cell % G.nx;
    int t = cell / G.nx;
    iy = (NDIM >= 2) ? t % G.ny : 0;
    iz = (NDIM >= 3) ? t / G.ny : 0;
  }
  auto flush = [&]() {
    int k = 0;
    while (__any_sync(FULL, k < cnt)) {
      if (k < cnt) body((int)list[k * BLOCK + tid]);
      k++;
    }
    cnt = 0;
  };
  constexpr int NY = (NDIM >= 2) ? 3 : 1, NZ = (NDIM >= 3) ? 3 : 1;
#pragma unroll 1
  for (int rz = 0; rz < NZ; rz++) {
#pragma unroll 1
    for (int ry = 0; ry < NY; ry++) {
      int k = 0, e = 0;
      if (active) {
        int cy = iy + ry - (NDIM >= 2 ? 1 : 0), cz = iz + rz - (NDIM >= 3 ? 1 : 0);
        // padded empty border cells (src/linkND.f90:88-89) keep ix-1, ix+1, cy, cz inside the grid for populated cells
        if (cy >= 0 && cy < G.ny && cz >= 0 && cz < G.nz) {
          int c0 = (cz * G.ny + cy) * G.nx;
          int xa = ix > 0 ? ix - 1 : 0, xb = ix + 1 < G.nx ? ix + 1 : G.nx - 1;
          k = __ldg(G.cellStart + c0 + xa);
          e = __ldg(G.cellStart + c0 + xb + 1);
        }
      }
      while (__any_sync(FULL, k < e)) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
          if (k < e) {
            if (cull(k)) { list[cnt * BLOCK + tid] = (unsigned)k; cnt++; }
            k++;
          }
        }
        if (__any_sync(FULL, cnt > CAP - 2)) flush();
      }
    }
  }
  flush();
}

}  // namespace ndk
