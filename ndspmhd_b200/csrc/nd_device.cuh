// Device-side building blocks shared by the density and rates kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ndk {

constexpr int IKERN = 4000;
// Sorted order: (cell, fine x bin, original index).  A cell row (fixed y, z) is contiguous and its fine bins ascend in x, so a
// neighbour search can clip every row of the stencil to the chord of the search sphere instead of scanning three whole cells.
constexpr int CELL_FX = 4;
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
  const int *fineStart;     // [ncells*CELL_FX+1] first sorted slot of each fine bin (a cell is cut into CELL_FX bins along x)
  const int *cellOf;        // [ntotal]   cell of each sorted slot
  const int *perm;          // [ntotal]   sorted slot -> original row
  const double4 *posh;      // [ntotal]   {x,y,z,1/h}  (1/h correctly rounded: h1(i) = 1./hh(i), density_sums.f90:130)
  const double4 *vm;        // [ntotal]   {vx,vy,vz,m}
  const double4 *posm;      // [ntotal]   {x,y,z,m}: all the LIGHT density round needs of a neighbour (one 32-byte gather a pair)
  const int *typ;           // [ntotal]
  const float4 *p32;        // [ntotal]   FP32 screening record {(x - xminpart)/dxcell, (h/hhmax)^2}: |dX|^2 < (h/hhmax)^2  <=>  rij2/h^2 < radkern2
  double hhmax1; float screen_margin, cull_margin;
  int nx, ny, nz, ncells;
  int npart, ntotal;        // rows [0,npart) carry their own state (targets are rows < nown), [npart,ntotal) are ghosts
  int nown;
  double radkern2, dq2table, ddq2table;
  const TabRec *tab;        // [IKERN+1]
  const TabRec2 *tab2;      // [IKERN+1]
  const double2 *tabg;      // [IKERN+1] {g, dg} of tab, contiguous: the pair kernels' shared-memory copy (one TMA bulk copy)
  const double2 *tabw;      // [IKERN+1] {w, dw} of tab, likewise (density kernel)
  const double *tabdrag;    // [2*(IKERN+1)] {w, dw}
};

// screening threshold of a particle: rij2*(1/h)^2 < radkern2 with dxcell = radkern*hhmax is |dX|^2 < (h/hhmax)^2 in cell units
__device__ __forceinline__ float screen_h2(double h, double hhmax1) { const double t = h * hhmax1; return (float)(t * t); }

__device__ __forceinline__ double4 ld4(const double4 *p) {
  // one 256-bit read-only load of a 32-byte aligned record (LDG.E.256, sm_100+): one request and one sector per lane
  double4 r;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st4(double4 *p, const double4 &v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}

// Branch-free FP64 1/sqrt and sqrt from the MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~2^-22): one third-order step for 1/sqrt
// (5 FP64 instructions, <= 1 ulp), one coupled Goldschmidt step + residual correction for sqrt (7 instructions, <= 0.5 ulp
// measured).  CUDA's sqrt()/rsqrt() carry a slow-path branch per call, which stops ptxas from interleaving the six
// independent square roots of a rates pair; these do not.  Arguments <= 1e-300 give 0.
// The x > 1e-300 guard is an integer compare on the high word (a DSETP would issue on the half-rate FP64 pipe): 0x01A56E1F is the
// high word of 1e-300; the arguments are non-negative or fail the test as negatives do.
#define ND_SQRT_ARG_OK(x) (__double2hiint(x) > 0x01A56E1F)
__device__ __forceinline__ double rsqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double t = x * y, e = fma(-t, y, 1.0);          // e = 1 - x y^2
  const double q = fma(0.375, e, 0.5) * e;              // y (1 + e/2 + 3 e^2/8)
  y = fma(y, q, y);
  return ND_SQRT_ARG_OK(x) ? y : 0.;
}
__device__ __forceinline__ double sqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double s0 = x * y, h0 = 0.5 * y;
  const double r = fma(-s0, h0, 0.5);
  const double s1 = fma(s0, r, s0), h1 = fma(h0, r, h0);
  const double d = fma(-s1, s1, x);
  const double s = fma(d, h1, s1);
  return ND_SQRT_ARG_OK(x) ? s : 0.;
}
// N independent square roots advanced in lockstep: the source order interleaves the (serial, ~9-cycle per step) chains so
// that ptxas, which keeps close to source order under register pressure, issues them back to back instead of one after another.
template <int N> __device__ __forceinline__ void sqrt_n(const double (&x)[N], double (&out)[N]) {
  double y[N], s0[N], h0[N], r[N], s1[N], h1[N], d[N];
#pragma unroll
  for (int i = 0; i < N; i++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(x[i]));
#pragma unroll
  for (int i = 0; i < N; i++) { s0[i] = x[i] * y[i]; h0[i] = 0.5 * y[i]; }
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = fma(-s0[i], h0[i], 0.5);
#pragma unroll
  for (int i = 0; i < N; i++) { s1[i] = fma(s0[i], r[i], s0[i]); h1[i] = fma(h0[i], r[i], h0[i]); }
#pragma unroll
  for (int i = 0; i < N; i++) d[i] = fma(-s1[i], s1[i], x[i]);
#pragma unroll
  for (int i = 0; i < N; i++) out[i] = ND_SQRT_ARG_OK(x[i]) ? fma(d[i], h1[i], s1[i]) : 0.;
}

__global__ void k_selftest_math(const double *in, double *out_sqrt, double *out_rsqrt, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { out_sqrt[i] = sqrt_nr(in[i]); out_rsqrt[i] = rsqrt_nr(in[i]); }
}

// cp.async (LDGSTS): 16 bytes global -> shared without passing through registers; .ca keeps the line in L1 for the other lanes
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// TMA bulk copy (UBLKCP) of a contiguous global span into shared memory, completion counted on an mbarrier
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem, const void *gmem, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(smem)),
               "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(
          (unsigned)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// rij2 = dot_product(dx,dx) in index order without FMA contraction (src/density_sums.f90:181,
// src/ratesND_mhd.f90:408; reference build has no FMA, src/Makefile:25).  Unused dimensions carry exact zeros.
__device__ __forceinline__ double dist2_exact(double dx, double dy, double dz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// index = int(q2*ddq2table), clamped (src/kernelND.f90:4435-4438)
__device__ __forceinline__ int tab_index(double q2, double ddq2table) {
  const double t = __dmul_rn(q2, ddq2table);
  // cvt.rzi.s32.f64 saturates (large -> INT_MAX, negative -> INT_MIN, NaN -> 0); an unsigned minimum then clamps both ends to IKERN
  // without the two DSETP of a range test (they would issue on the half-rate FP64 pipe)
  return (int)min((unsigned)__double2int_rz(t), (unsigned)IKERN);
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

__device__ __forceinline__ int cell_begin(const Grid &G, int cell) { return __ldg(G.fineStart + (size_t)cell * CELL_FX); }

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
// Neighbour lists.  Finding neighbours (cheap test, ~85 % rejected, wants many resident warps) and evaluating pair terms
// (hundreds of FP64 instructions, wants registers) are separate kernels: build_lists_kernel walks the 3^NDIM stencil of
// the target's cell as 3^(NDIM-1) x-rows -- a row's three cells are contiguous in the cell-sorted arrays, so a row is
// one span of slots -- applies the reference's inclusion test bit for bit and appends accepted slots to the target's
// list; the pair kernels then run over the lists with every lane busy.
// Layout: target t (index within the launch) owns column (t & 31) of warp block (t >> 5):
//     nbr[((t >> 5) * lmax + n) * 32 + (t & 31)],  n < cnt[t]
// so entry n of the 32 targets of a warp is one 128-byte line, written and read coalesced.
// ------------------------------------------------------------------------------------------------------
struct NbrLists {
  unsigned *nbr;
  int *cnt;        // [targets in the launch]
  int lmax;        // column capacity
  int *overflow;   // set to the largest count seen when a column overflows (host grows lmax and repeats)
  int split;       // LIST_RATES with drag: cnt = front | back << 16 -- pairs whose types interact from the column's start, gas-dust (drag) pairs from its end
};
__device__ __forceinline__ size_t nbr_index(int t, int n, int lmax) { return ((size_t)(t >> 5) * lmax + n) * 32 + (t & 31); }

// Walks one list column.  Entries are read in batches of four, two batches (thousands of cycles) before they are needed: the
// list streams from HBM, and a load whose consumer sits in the same basic block gets its register move hoisted right behind
// it by ptxas, which exposes the full DRAM latency on every pair (measured: 20 % of all stall samples).  The inner loop is a
// real loop, so the batch rotation stays at the bottom of the outer one.  f(n, k, k1, k2): entry n and the two after it.
template <class F> __device__ __forceinline__ void walk_list(const unsigned *col, int cnt, F &&f) {
  const int last = cnt - 1;
  auto ld = [&](int n) { return (int)__ldcs(col + (size_t)min(n, last) * 32); };
  int c0 = ld(0), c1 = ld(1), c2 = ld(2), c3 = ld(3);
  int n0 = ld(4), n1 = ld(5), n2 = ld(6), n3 = ld(7);
#pragma unroll 1
  for (int nb = 0; nb < cnt; nb += 4) {
    const int m0 = ld(nb + 8), m1 = ld(nb + 9), m2 = ld(nb + 10), m3 = ld(nb + 11);
#pragma unroll 1
    for (int u = 0; u < 4; u++) {
      if (nb + u >= cnt) break;
      const int k = u == 0 ? c0 : u == 1 ? c1 : u == 2 ? c2 : c3;
      const int k1 = u == 0 ? c1 : u == 1 ? c2 : u == 2 ? c3 : n0;
      const int k2 = u == 0 ? c2 : u == 1 ? c3 : u == 2 ? n0 : n1;
      f(nb + u, k, k1, k2);
    }
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    n0 = m0; n1 = m1; n2 = m2; n3 = m3;
  }
}

enum { LIST_DENS_FIRST = 0, LIST_DENS_PARTIAL = 1, LIST_RATES = 2 };
#ifndef ND_LIST_UNROLL
#define ND_LIST_UNROLL 4
#endif
constexpr int LIST_UNROLL = ND_LIST_UNROLL;   // candidate records in flight per thread in the list builder

struct ListArgs {
  const double *hh;        // original-order current smoothing lengths (density modes)
  const int *targets;      // LIST_DENS_PARTIAL, optionally LIST_RATES: sorted slots to process; else NULL (slot = s0 + t)
  int s0, ntargets;
  int *numneigh;           // density modes: neighbour count as the reference defines it, by original row
  int drag;                // LIST_RATES: gas-dust pairs are kept (drag_forces)
  int *pair_out_i, *pair_out_j; unsigned long long *pair_count; long long pair_cap;   // parity-test hook (LIST_RATES)
};

// TYPES = false: every row has the same itype (found at link time), so the type rules always pass and typ[] is never read.
template <int NDIM, int MODE, bool TYPES>
__global__ void __launch_bounds__(128) build_lists_kernel(Grid G, ListArgs A, NbrLists L) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= A.ntargets) return;
  const int s = (MODE == LIST_DENS_PARTIAL || (MODE == LIST_RATES && A.targets)) ? A.targets[t] : A.s0 + t;
  const int orig = G.perm[s];
  const int ti = G.typ[s];
  bool active = orig < G.nown;                     // ghosts and halo rows are sources only
  if (MODE == LIST_DENS_PARTIAL && ti == T_BND) active = false;   // density_sums.f90:510
  if (!active) { L.cnt[t] = 0; return; }
  // density: current h (the sorted record holds the h of the last link); rates: the record's (h1(i) = 1./hh(i))
  const double hcur = (MODE == LIST_RATES) ? 0. : A.hh[orig];
  const int cell = G.cellOf[s];
  const int cs0 = cell_begin(G, cell), cs1 = cell_begin(G, cell + 1);
  const int ix = cell % G.nx;
  const int tq = cell / G.nx;
  const int iy = (NDIM >= 2) ? tq % G.ny : 0, iz = (NDIM >= 3) ? tq / G.ny : 0;
  int cnt = 0, cntb = 0, nneigh = 0;
  unsigned *col = L.nbr + ((size_t)(t >> 5) * L.lmax) * 32 + (t & 31);
  const bool split = (MODE == LIST_RATES) && TYPES && L.split;
  const bool hook = (MODE == LIST_RATES) && A.pair_out_i != nullptr;

  // Exact inclusion test in the reference's arithmetic (rare path, see below): sets whether the pair counts as a neighbour
  // and whether it goes on the list.
  auto exact = [&](int k, bool &counts, bool &store) {
    const double4 p = ld4(G.posh + s), pj = ld4(G.posh + k);
    const double hi1 = (MODE == LIST_RATES) ? p.w : 1.0 / hcur, hi21 = __dmul_rn(hi1, hi1);
    const double rij2 = dist2_exact(p.x - pj.x, p.y - pj.y, p.z - pj.z);
    const double hj1 = pj.w;
    if (MODE == LIST_DENS_FIRST) {
      // The reference visits a pair once; "i" is the particle met first: the one in the lower cell, or the later-inserted
      // (higher index) particle of the same chain.  Slots are sorted by (cell, index), so that is a slot comparison.
      // q2 of "i" is rij2*hi21, q2 of "j" is (rij2*hj1)*hj1 (density_sums.f90:182-183).
      // Inside a cell the slots are ordered by (fine bin, index), so the same-cell case compares the original rows.
      const bool iam_i = (k >= cs1) || (k >= cs0 && G.perm[k] <= orig);
      double q2me, q2ot;
      if (iam_i) { q2me = __dmul_rn(rij2, hi21); q2ot = __dmul_rn(__dmul_rn(rij2, hj1), hj1); }
      else { q2me = __dmul_rn(__dmul_rn(rij2, hi1), hi1); q2ot = __dmul_rn(rij2, __dmul_rn(hj1, hj1)); }
      store = q2me < G.radkern2;                                  // terms with q2me >= radkern2 are exact zeros (table end = 0)
      counts = store || q2ot < G.radkern2;                        // :189-190 with the target real
    } else if (MODE == LIST_DENS_PARTIAL) {
      counts = store = __dmul_rn(rij2, hi21) < G.radkern2;        // :528
    } else {                                                      // ratesND_mhd.f90:404-415
      const double q2i = __dmul_rn(rij2, hi21), q2j = __dmul_rn(rij2, __dmul_rn(hj1, hj1));
      counts = store = (q2i < G.radkern2) || (q2j < G.radkern2);
    }
  };
  // FP32 screening in cell units: |dX|^2 against (h/hhmax)^2.  Outside a band of +-screen_margin around the thresholds the
  // FP32 verdict provably equals the exact one (the band covers the rounding of the FP32 coordinates and thresholds); inside
  // the band -- a 1e-4 sliver of the candidates -- the exact FP64 test above decides.  The list is therefore exactly the
  // reference's neighbour set, without FP64 arithmetic on the common path.
  const float4 pf = G.p32[s];
  const float Ti = (MODE == LIST_RATES) ? pf.w : screen_h2(hcur, G.hhmax1);
  const float marg = G.screen_margin;
  // A row is scanned in groups of up to 32 consecutive slots.  The per-candidate work is branch-free: the FP32 verdicts go into
  // bit masks (listed / counted as a neighbour / inside the ambiguity band); the rare exact decisions, the type rules and the
  // stores of the accepted slots happen once per group.  (Per-candidate branches and stores cost 27 instructions a candidate
  // and kept ptxas from overlapping the loads of one candidate with the arithmetic of the next.)
  auto scan_group = [&](int g0, int gn) {
    unsigned mstore = 0u, mcount = 0u, mamb = 0u, bit = 1u;
#pragma unroll LIST_UNROLL
    for (int u = 0; u < gn; u++) {
      const float4 qj = __ldg(G.p32 + g0 + u);
      const float ddx = pf.x - qj.x, ddy = pf.y - qj.y, ddz = pf.z - qj.z;
      const float r2 = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
      const float a = r2 - Ti, b = (MODE == LIST_DENS_PARTIAL) ? a : r2 - qj.w;
      if (fminf(a, b) < 0.f) mcount |= bit;                       // :189-190 / :528 / ratesND_mhd.f90:415
      if (MODE == LIST_DENS_FIRST && a < 0.f) mstore |= bit;
      if (fminf(fabsf(a), fabsf(b)) <= marg) mamb |= bit;
      bit <<= 1;
    }
    if (MODE != LIST_DENS_FIRST) mstore = mcount;
    while (mamb) {                                                // exact FP64 decision inside the band (1e-4 of the candidates)
      const int u = __ffs(mamb) - 1;
      mamb &= mamb - 1u;
      bool counts, store;
      exact(g0 + u, counts, store);
      mcount = counts ? (mcount | (1u << u)) : (mcount & ~(1u << u));
      mstore = store ? (mstore | (1u << u)) : (mstore & ~(1u << u));
    }
    if (MODE == LIST_RATES) {
      if (s >= g0 && s < g0 + gn) { mcount &= ~(1u << (s - g0)); mstore &= ~(1u << (s - g0)); }   // j /= i (the target is real: no ghost-ghost pair)
      if (hook) {                                                 // parity-test hook: the accepted pairs before the type dispatch
        for (unsigned m = mstore; m; m &= m - 1u) {
          const int k = g0 + __ffs(m) - 1;
          unsigned long long n = atomicAdd(A.pair_count, 1ull);
          if ((long long)n < A.pair_cap) { A.pair_out_i[n] = orig + 1; A.pair_out_j[n] = G.perm[k] + 1; }
        }
      }
    }
    unsigned mback = 0u;                                          // drag runs: the pairs drag_forces takes (types do not interact)
    if (TYPES) {
      for (unsigned m = mcount; m; m &= m - 1u) {
        const int u = __ffs(m) - 1;
        const int tj = __ldg(G.typ + g0 + u);
        bool ok;
        if (MODE == LIST_DENS_FIRST) ok = types_interact(ti, tj);   // density_sums.f90:169-174
        else if (MODE == LIST_DENS_PARTIAL) ok = (tj == ti) || (tj == T_BND);   // :517
        else {                                                      // ratesND_mhd.f90:436-446
          const bool inter = types_interact(ti, tj);
          ok = A.drag || inter;
          if (split && !inter) mback |= 1u << u;
        }
        if (!ok) { mcount &= ~(1u << u); mstore &= ~(1u << u); }
      }
    }
    nneigh += __popc(mcount);                                     // :196-197 / :532
    for (unsigned m = mstore & ~mback; m; m &= m - 1u) {
      if (cnt + cntb < L.lmax) col[(size_t)cnt * 32] = (unsigned)(g0 + __ffs(m) - 1);
      cnt++;
    }
    for (unsigned m = mstore & mback; m; m &= m - 1u) {           // from the end of the column downwards
      if (cnt + cntb < L.lmax) col[(size_t)(L.lmax - 1 - cntb) * 32] = (unsigned)(g0 + __ffs(m) - 1);
      cntb++;
    }
  };

  constexpr int NY = (NDIM >= 2) ? 3 : 1, NZ = (NDIM >= 3) ? 3 : 1;
  // Chord clipping.  In cell units (dxcell = radkern*hhmax) no accepted or counted pair is farther than Rc: 1 when the
  // neighbour's h can decide (h_j <= hhmax), (h_i/hhmax) when only the target's does.  A stencil row at transverse distance d
  // from the target can only hold such pairs within |dx| <= sqrt(Rc^2 - d^2): the scan covers just the fine bins that
  // overlap that interval.  cull_margin covers the FP32 rounding of the cell-unit coordinates (bins are cut in FP64).
  const float Rc2 = ((MODE == LIST_DENS_PARTIAL) ? Ti : 1.f) + G.cull_margin;
  const float fy = (NDIM >= 2) ? pf.y - (float)iy : 0.f, fz = (NDIM >= 3) ? pf.z - (float)iz : 0.f;   // position inside the cell, [0,1)
#pragma unroll 1
  for (int rz = 0; rz < NZ; rz++) {
#pragma unroll 1
    for (int ry = 0; ry < NY; ry++) {
      const int cy = iy + ry - (NDIM >= 2 ? 1 : 0), cz = iz + rz - (NDIM >= 3 ? 1 : 0);
      // padded empty border cells (src/linkND.f90:88-89) keep ix-1, ix+1, cy, cz inside the grid for populated cells
      if (cy < 0 || cy >= G.ny || cz < 0 || cz >= G.nz) continue;
      float dty = 0.f, dtz = 0.f;
      if (NDIM >= 2) dty = ry == 1 ? 0.f : (ry == 0 ? fy : 1.f - fy);
      if (NDIM >= 3) dtz = rz == 1 ? 0.f : (rz == 0 ? fz : 1.f - fz);
      dty = fmaxf(dty, 0.f); dtz = fmaxf(dtz, 0.f);
      const float R2 = Rc2 - (dty * dty + dtz * dtz);
      if (R2 < 0.f) continue;
      const float R = sqrtf(R2) + G.cull_margin;
      const int c0 = (cz * G.ny + cy) * G.nx;
      const int xa = ix > 0 ? ix - 1 : 0, xb = ix + 1 < G.nx ? ix + 1 : G.nx - 1;
      int blo = __float2int_rd((pf.x - R) * (float)CELL_FX), bhi = __float2int_rd((pf.x + R) * (float)CELL_FX);
      blo = max(blo, xa * CELL_FX); bhi = min(bhi, xb * CELL_FX + CELL_FX - 1);
      if (bhi < blo) continue;
      int k = __ldg(G.fineStart + (size_t)c0 * CELL_FX + blo);
      const int e = __ldg(G.fineStart + (size_t)c0 * CELL_FX + bhi + 1);
#pragma unroll 1
      for (; k < e; k += 32) scan_group(k, min(32, e - k));
    }
  }
  if (cnt + cntb > L.lmax) { atomicMax(L.overflow, cnt + cntb); cnt = min(cnt, L.lmax); cntb = 0; }
  L.cnt[t] = split ? (cnt | (cntb << 16)) : cnt;
  if (MODE != LIST_RATES) A.numneigh[orig] = nneigh;
}

}  // namespace ndk
