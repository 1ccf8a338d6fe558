// nd_kernels.cuh -- the O(N) kernels of libndspmhd_b200.so and their small host helpers: integer scan and reductions, ghost generation
// (src/ghostND_mhd.f90), the cell sort of set_linklist (src/linkND.f90), conservative2primitive + EOS, the finalisation loop of get_rates
// (src/ratesND_mhd.f90:532-965), the leapfrog predictor/corrector (src/stepND_leapfrog_mhd.f90) and the evwrite sums.
// Included by nd_capi.cu INSIDE its anonymous namespace, after nd_ctx, CU(), LAUNCH() and SMALL_D2H() are defined.
#pragma once
// =====================================================================================================
// primitives: exclusive scan of ints, reductions
// =====================================================================================================
constexpr int SCAN_T = 256, SCAN_E = 8, SCAN_TILE = SCAN_T * SCAN_E;

__global__ void k_scan_tile(const int *in, int *out, int n, int *blocksums) {
  __shared__ int warp_tot[SCAN_T / 32];
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_E;
  int v[SCAN_E], sum = 0;
#pragma unroll
  for (int e = 0; e < SCAN_E; e++) { int i = base + e; v[e] = (i < n) ? in[i] : 0; sum += v[e]; }
  int incl = sum;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = (lane < SCAN_T / 32) ? warp_tot[lane] : 0, wi = w;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, wi, o); if (lane >= o) wi += t; }
    if (lane < SCAN_T / 32) warp_tot[lane] = wi - w;
    if (lane == SCAN_T / 32 - 1) blocksums[blockIdx.x] = wi;
  }
  __syncthreads();
  int run = warp_tot[wid] + incl - sum;
#pragma unroll
  for (int e = 0; e < SCAN_E; e++) { int i = base + e; if (i < n) out[i] = run; run += v[e]; }
}
__global__ void k_scan_sums(int *blocksums, int nb) {   // one block, sequential over chunks of 1024
  __shared__ int wt[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int b0 = 0; b0 < nb; b0 += 1024) {
    int i = b0 + threadIdx.x;
    int v = (i < nb) ? blocksums[i] : 0, incl = v;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wt[wid] = incl;
    __syncthreads();
    if (wid == 0) { int w = wt[lane], wi = w; for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, wi, o); if (lane >= o) wi += t; } wt[lane] = wi - w; }
    __syncthreads();
    int excl = carry_s + wt[wid] + incl - v;
    if (i < nb) blocksums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) blocksums[nb] = carry_s;   // grand total
}
__global__ void k_scan_add(int *out, int n, const int *blocksums, int nb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += blocksums[i / SCAN_TILE];
  if (i == 0) out[n] = blocksums[nb];
}

int ensure_blocksums(nd_ctx *c, int nb) {
  if (nb + 1 > c->blocksumcap) {
    if (c->blocksums) cudaFree(c->blocksums);
    c->blocksumcap = nb + 1 + 1024;
    CU(cudaMalloc(&c->blocksums, sizeof(int) * c->blocksumcap));
  }
  return 0;
}
// out[0..n] = exclusive scan of in[0..n-1]; out[n] = total
int exclusive_scan(nd_ctx *c, const int *in, int *out, int n) {
  const int nb = nblocks(n, SCAN_TILE);
  if (int e = ensure_blocksums(c, nb)) return e;
  LAUNCH(c, k_scan_tile, nb, SCAN_T, 0, in, out, n, c->blocksums);
  LAUNCH(c, k_scan_sums, 1, 1024, 0, c->blocksums, nb);
  LAUNCH(c, k_scan_add, nblocks(n, 256), 256, 0, out, n, c->blocksums, nb);
  return 0;
}

__global__ void k_max_h(const double *hh, int n, unsigned long long *key) {
  double m = -DBL_MAX;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmax(m, hh[i]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomic_max_d(key, m);
}
template <int NDIM> __global__ void k_minmax_x(const double *x, int n, unsigned long long *keys /* [0..2] min, [3..5] max */) {
  double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    for (int d = 0; d < NDIM; d++) { double v = x[(size_t)i * NDIM + d]; mn[d] = fmin(mn[d], v); mx[d] = fmax(mx[d], v); }
  for (int d = 0; d < NDIM; d++) {
    double a = warp_min(mn[d]), b = warp_max(mx[d]);
    if ((threadIdx.x & 31) == 0) { atomic_min_d(keys + d, a); atomic_max_d(keys + 3 + d, b); }
  }
}
__global__ void k_check_h(const double *hh, int n, int *flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && hh[i] <= 2.2250738585072014e-308) atomicCAS(&flags[1], 0, ND_ERR_H_NONPOSITIVE);   // iterate_density.f90:99-102
}

// =====================================================================================================
// ghosts on the device, restating src/ghostND_mhd.f90:166-346 (periodic ibound=3, reflecting 2/4/6)
// Two passes (count, scan, write) so ghost rows come out in the reference's order: by parent, then in the
// order makeghost is called for that parent.
// =====================================================================================================
struct GhostArgs {
  double *x, *vel; const double *hh; int *itype, *ireal; const int *offset; int *count;
  int npart, cap; int ibound[3]; double xmin[3], xmax[3]; double radkern, hhmax; int *flags;
};
template <int NDIM, bool WRITE> __global__ void k_ghosts(GhostArgs A) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= A.npart) return;
  double xj[3] = {0, 0, 0}, vj[3];
#pragma unroll
  for (int d = 0; d < NDIM; d++) xj[d] = A.x[(size_t)j * NDIM + d];
  double dxbound[3];
#pragma unroll
  for (int d = 0; d < 3; d++) dxbound[d] = A.radkern * A.hhmax;                                   // :80
#pragma unroll
  for (int d = 0; d < NDIM; d++) if (A.ibound[d] == 2 || A.ibound[d] == 4 || A.ibound[d] == 6) dxbound[d] = A.radkern * A.hh[j];   // :173
  // nine particles in ten are farther than the reach from every face: the test of :205 for both faces of every dimension, in registers,
  // before the general procedure below (whose run-time indexed arrays live in local memory)
  bool any = false;
#pragma unroll
  for (int d = 0; d < NDIM; d++) {
    if (A.ibound[d] <= 1) continue;
    const double dhi = A.xmax[d] - xj[d], dlo = xj[d] - A.xmin[d];
    any = any || ((dhi < dxbound[d]) && (dhi > 0)) || ((dlo < dxbound[d]) && (dlo > 0));
  }
  if (!any) { if (!WRITE) A.count[j] = 0; return; }
  for (int d = 0; d < 3; d++) vj[d] = A.vel ? A.vel[(size_t)j * 3 + d] : 0.;   // vel = NULL: velocities follow later (k_late_vel; periodic ghosts only)
  int n = 0;
  const int base = WRITE ? A.npart + A.offset[j] : 0;
  auto make = [&](const double *xp, const double *vp) {                                            // makeghost, :363-431
    if (WRITE) {
      const int r = base + n;
      if (r < A.cap) {
        for (int d = 0; d < NDIM; d++) A.x[(size_t)r * NDIM + d] = xp[d];
        if (A.vel) for (int d = 0; d < 3; d++) A.vel[(size_t)r * 3 + d] = vp[d];
        A.ireal[r] = j + 1;
        A.itype[r] = A.itype[j];                                                                   // :346
      }
    }
    n++;
  };
  bool mk[3][2];
  double xnew[3][2], xpart[3], vpart[3];
  for (int idim = 0; idim < NDIM; idim++) {                                                        // :176
    if (A.ibound[idim] <= 1) { mk[idim][0] = mk[idim][1] = false; continue; }
    const bool refl = (A.ibound[idim] == 2 || A.ibound[idim] == 4 || A.ibound[idim] == 6);
    for (int mm = 0; mm < 2; mm++) {                                                               // :184 (0: xmax, 1: xmin)
      for (int d = 0; d < 3; d++) { xpart[d] = xj[d]; vpart[d] = vj[d]; }
      double xbound, xperbound, dx;
      if (mm == 0) { xbound = A.xmax[idim]; xperbound = A.xmin[idim]; dx = A.xmax[idim] - xj[idim]; }
      else { xbound = A.xmin[idim]; xperbound = A.xmax[idim]; dx = xj[idim] - A.xmin[idim]; }
      mk[idim][mm] = (dx < dxbound[idim]) && (dx > 0);                                             // :205
      if (!mk[idim][mm]) continue;
      const double dxshift = __dsub_rn(xj[idim], xbound);                                          // :212
      if (!refl) xnew[idim][mm] = __dadd_rn(xperbound, dxshift);                                   // :225
      else { xnew[idim][mm] = __dsub_rn(xbound, dxshift); vpart[idim] = -vj[idim]; }               // :227-228
      xpart[idim] = xnew[idim][mm];
      make(xpart, vpart);                                                                          // :248
      for (int ip = 0; ip < idim; ip++) {                                                          // :257 edges
        for (int mp = 0; mp < 2; mp++) {
          if (!mk[ip][mp]) continue;
          xpart[ip] = xnew[ip][mp];                                                                // :266
          const bool reflp = (A.ibound[ip] == 2 || A.ibound[ip] == 4 || A.ibound[ip] == 6);
          if (reflp) for (int d = 0; d < 3; d++) vpart[d] = vj[d];                                 // :283-284
          make(xpart, vpart);                                                                      // :287
          if (ip >= 1) {                                                                           // :293 corners
            const int ipp = ip - 1;
            for (int mpp = 0; mpp < 2; mpp++) {
              if (!mk[ipp][mpp]) continue;
              xpart[ipp] = xnew[ipp][mpp];                                                         // :300
              const bool reflpp = (A.ibound[ipp] == 2 || A.ibound[ipp] == 4 || A.ibound[ipp] == 6);
              if (reflpp) for (int d = 0; d < 3; d++) vpart[d] = -vj[d];                           // :307-309
              make(xpart, vpart);                                                                  // :316
              xpart[ipp] = xj[ipp];                                                                // :318
            }
          }
          xpart[ip] = xj[ip];                                                                      // :327
        }
      }
    }
  }
  if (!WRITE) A.count[j] = n;
}

// =====================================================================================================
// slab halos (multi-GPU): selection, packing, unpacking.  A halo row is the copy of a neighbour rank's particle that lies
// within radkern*hhmax of the shared slab face; across the periodic wrap its x is shifted with the reference's ghost
// arithmetic, xperbound + (x - xbound) (src/ghostND_mhd.f90:212-225), so pair geometry is bitwise that of a single-GPU run.
// =====================================================================================================
struct HaloSelArgs { const double *x; int nown, ndim; double lo, hi, reach, tol; int left_on, right_on; int *flagL, *flagR, *err; };
__global__ void k_halo_flags(HaloSelArgs A) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= A.nown) return;
  const double xj = A.x[(size_t)j * A.ndim];
  if (xj < A.lo - A.tol || xj > A.hi + A.tol) atomicCAS(A.err, 0, ND_ERR_INVALID_ARG);   // outside its slab: the caller must repartition
  A.flagL[j] = (A.left_on && xj < A.lo + A.reach) ? 1 : 0;
  A.flagR[j] = (A.right_on && xj > A.hi - A.reach) ? 1 : 0;
}
__global__ void k_halo_compact(const int *flag, const int *scan, int n, int *list) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n && flag[j]) list[scan[j]] = j;
}
// exchange 1 (before the link): what the link and the density rounds read of a halo row -- [x(ndim) vel(3) pmass hh rho] as doubles,
// field-major, then itype as ints.  The evolved variables the rates read (en, Bevol, alpha, psi) travel with exchange 2 after the density
// iteration, so a pipelined upload (ndspmhd_b200_derivs_host) need not have landed them before the link starts.
struct HaloPackArgs {
  const int *list; int n, ndim;
  double *x, *vel, *pmass, *hh, *en, *Bevol, *alpha, *psi, *rho, *gradh; int *itype;
  double *buf; int row0;        // pack: buf out; unpack: rows [row0, row0+n) in
  int shift; double xbound, xperbound;   // periodic wrap: x' = xperbound + (x - xbound)
  int full;                     // exchange 2: 0 = hh, rho, gradh (mid-iteration refresh); 1 = + en, Bevol(3), alpha(3), psi (after the last round)
};
__host__ __device__ inline int halo1_nfields(int ndim) { return ndim + 6; }
__host__ __device__ inline int halo2_nfields(int full) { return full ? 11 : 3; }
__global__ void k_halo_pack1(HaloPackArgs A) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= A.n) return;
  const int j = A.list[q];
  const size_t n = A.n;
  int f = 0;
  for (int d = 0; d < A.ndim; d++) {
    double v = A.x[(size_t)j * A.ndim + d];
    if (d == 0 && A.shift) v = __dadd_rn(A.xperbound, __dsub_rn(v, A.xbound));
    A.buf[(f++) * n + q] = v;
  }
  for (int d = 0; d < 3; d++) A.buf[(f++) * n + q] = A.vel[(size_t)j * 3 + d];
  A.buf[(f++) * n + q] = A.pmass[j];
  A.buf[(f++) * n + q] = A.hh[j];
  A.buf[(f++) * n + q] = A.rho[j];
  reinterpret_cast<int *>(A.buf + (size_t)f * n)[q] = A.itype[j];
}
__global__ void k_halo_unpack1(HaloPackArgs A) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= A.n) return;
  const int r = A.row0 + q;
  const size_t n = A.n;
  int f = 0;
  for (int d = 0; d < A.ndim; d++) A.x[(size_t)r * A.ndim + d] = A.buf[(f++) * n + q];
  for (int d = 0; d < 3; d++) A.vel[(size_t)r * 3 + d] = A.buf[(f++) * n + q];
  A.pmass[r] = A.buf[(f++) * n + q];
  A.hh[r] = A.buf[(f++) * n + q];
  A.rho[r] = A.buf[(f++) * n + q];
  A.itype[r] = reinterpret_cast<const int *>(A.buf + (size_t)f * n)[q];
}
// exchange 2: the owners' hh, rho, gradh (refresh between `density` rounds); after the last round also en, Bevol, alpha, psi
__global__ void k_halo_pack2(HaloPackArgs A) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= A.n) return;
  const int j = A.list[q];
  const size_t n = A.n;
  A.buf[q] = A.hh[j]; A.buf[n + q] = A.rho[j]; A.buf[2 * n + q] = A.gradh[j];
  if (A.full) {
    int f = 3;
    A.buf[(f++) * n + q] = A.en[j];
    for (int d = 0; d < 3; d++) A.buf[(f++) * n + q] = A.Bevol[(size_t)j * 3 + d];
    for (int d = 0; d < 3; d++) A.buf[(f++) * n + q] = A.alpha[(size_t)j * 3 + d];
    A.buf[(f++) * n + q] = A.psi[j];
  }
}
__global__ void k_halo_unpack2(HaloPackArgs A) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= A.n) return;
  const int r = A.row0 + q;
  const size_t n = A.n;
  A.hh[r] = A.buf[q]; A.rho[r] = A.buf[n + q]; A.gradh[r] = A.buf[2 * n + q];
  if (A.full) {
    int f = 3;
    A.en[r] = A.buf[(f++) * n + q];
    for (int d = 0; d < 3; d++) A.Bevol[(size_t)r * 3 + d] = A.buf[(f++) * n + q];
    for (int d = 0; d < 3; d++) A.alpha[(size_t)r * 3 + d] = A.buf[(f++) * n + q];
    A.psi[r] = A.buf[(f++) * n + q];
  }
}

// =====================================================================================================
// native transport (nd_nccl.cuh): the operands of the small collectives are assembled on the device, so ONE stream synchronise serves the
// local D2H, the collective and the D2H of its result (the callback transport needs the values on the host first: two or three)
// =====================================================================================================
enum { CP_KEY = 0, CP_INT = 1, CP_INT_NONZERO = 2, CP_DOUBLE = 3, CP_U64 = 4, CP_HOST = 5 };
struct CommPack { int n; int kind[16]; const void *src[16]; double scale[16], hostval[16]; };
__global__ void k_comm_pack(CommPack P, double *dst) {
  const int k = threadIdx.x;
  if (k >= P.n) return;
  double v = 0.;
  switch (P.kind[k]) {
    case CP_KEY: { const unsigned long long key = *reinterpret_cast<const unsigned long long *>(P.src[k]); v = key ? dkey_inv(key) : 0.; break; }
    case CP_INT: v = (double)*reinterpret_cast<const int *>(P.src[k]); break;
    case CP_INT_NONZERO: v = *reinterpret_cast<const int *>(P.src[k]) != 0 ? 1. : 0.; break;
    case CP_DOUBLE: v = *reinterpret_cast<const double *>(P.src[k]); break;
    case CP_U64: v = (double)*reinterpret_cast<const unsigned long long *>(P.src[k]); break;
    default: v = P.hostval[k];
  }
  dst[k] = v * P.scale[k];
}
__global__ void k_double_to_key(const double *v, unsigned long long *key) { *key = dkey(*v); }

// =====================================================================================================
// particle migration between slabs (ndspmhd_b200_step on slab-decomposed contexts).  The reference moves every particle in the predictor
// and wraps it (src/stepND_leapfrog_mhd.f90:145, src/boundaryND.f90:65-93); with x-slabs a row whose x left [slab_lo, slab_hi) is
// re-owned by the neighbouring rank: its whole evolved state and the integrator's `*in` planes travel, the holes are filled from the
// tail of the own rows, arrivals are appended.  Rows carry a 64-bit id (ndspmhd_b200_set_row_ids) so a caller can follow them.
// =====================================================================================================
struct MigSelArgs {
  const double *x; const int *itype; int nown, ndim;
  double lo, hi, llo, lhi, rlo, rhi; int hi_closed, lhi_closed, rhi_closed, left_on, right_on, same_peer;   // [lo,hi) mine, neighbours' slabs; *_closed: x == hi belongs
  int *flagL, *flagR, *flagAny, *err;
};
__device__ __forceinline__ bool mig_inside(double x, double lo, double hi, int closed) { return x >= lo && (x < hi || (closed && x == hi)); }
__device__ __forceinline__ bool step_fixed_type(int it) { return it == T_BND || it == 11 || it == T_BNDDUST; }
__global__ void k_migrate_flags(MigSelArgs A) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= A.nown) return;
  const double xj = A.x[(size_t)j * A.ndim];
  int fl = 0, fr = 0;
  if (!mig_inside(xj, A.lo, A.hi, A.hi_closed)) {
    if (step_fixed_type(A.itype[j])) atomicCAS(A.err, 0, ND_ERR_UNSUPPORTED_OPTION);          // fixed rows are tied to their partners' row numbers
    else if (A.right_on && mig_inside(xj, A.rlo, A.rhi, A.rhi_closed)) fr = 1;
    else if (A.left_on && mig_inside(xj, A.llo, A.lhi, A.lhi_closed)) { if (A.same_peer) fr = 1; else fl = 1; }
    else atomicCAS(A.err, 0, ND_ERR_INVALID_ARG);                                              // farther than the adjacent slab in one step
  }
  A.flagL[j] = fl; A.flagR[j] = fr; A.flagAny[j] = fl | fr;
}
struct MigRows {
  double *arr[12]; int width[12]; int narr;     // state arrays, `width` doubles a row
  double *in; size_t instride; int nplanes;     // leapfrog `*in` planes
  int *itype; long long *gid;
};
__host__ __device__ inline int mig_nfields(const MigRows &R) { int f = R.nplanes + 2; for (int a = 0; a < R.narr; a++) f += R.width[a]; return f; }
struct MigPackArgs { MigRows R; const int *list; int n, row0; double *buf; };
__global__ void k_migrate_pack(MigPackArgs A) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= A.n) return;
  const int j = A.list[q];
  const size_t n = A.n;
  int f = 0;
  for (int a = 0; a < A.R.narr; a++) for (int d = 0; d < A.R.width[a]; d++) A.buf[(f++) * n + q] = A.R.arr[a][(size_t)j * A.R.width[a] + d];
  for (int pl = 0; pl < A.R.nplanes; pl++) A.buf[(f++) * n + q] = A.R.in[(size_t)pl * A.R.instride + j];
  A.buf[(f++) * n + q] = (double)A.R.itype[j];
  A.buf[(f++) * n + q] = __longlong_as_double(A.R.gid[j]);
}
__global__ void k_migrate_unpack(MigPackArgs A) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= A.n) return;
  const int r = A.row0 + q;
  const size_t n = A.n;
  int f = 0;
  for (int a = 0; a < A.R.narr; a++) for (int d = 0; d < A.R.width[a]; d++) A.R.arr[a][(size_t)r * A.R.width[a] + d] = A.buf[(f++) * n + q];
  for (int pl = 0; pl < A.R.nplanes; pl++) A.R.in[(size_t)pl * A.R.instride + r] = A.buf[(f++) * n + q];
  A.R.itype[r] = (int)A.buf[(f++) * n + q];
  A.R.gid[r] = __double_as_longlong(A.buf[(f++) * n + q]);
}
// tail rows [m, m + nl) that stay: flag them; k_migrate_fill then moves the k-th of them into the k-th hole (holes ascend, the first
// `#staying tail rows` of them lie below m)
__global__ void k_migrate_tailflags(const int *flagAny, int m, int nl, int *tailflag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nl) tailflag[k] = flagAny[m + k] ? 0 : 1;
}
__global__ void k_migrate_fill(MigRows R, const int *tailflag, const int *tailscan, const int *holes, int m, int nl) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nl || !tailflag[k]) return;
  const int src = m + k, dst = holes[tailscan[k]];
  for (int a = 0; a < R.narr; a++) for (int d = 0; d < R.width[a]; d++) R.arr[a][(size_t)dst * R.width[a] + d] = R.arr[a][(size_t)src * R.width[a] + d];
  for (int pl = 0; pl < R.nplanes; pl++) R.in[(size_t)pl * R.instride + dst] = R.in[(size_t)pl * R.instride + src];
  R.itype[dst] = R.itype[src];
  R.gid[dst] = R.gid[src];
}
__global__ void k_iota_ll(long long *a, int n, long long first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = first + i;
}

// =====================================================================================================
// cell grid (src/linkND.f90:119-145) as a counting sort: cell index, histogram, scan, scatter, per-cell ordering
// =====================================================================================================
struct CellArgs { const double *x; int ntotal; double xminpart[3], dxcell; int ncellsx[3]; int *cellOfOrig, *cellCount, *flags; };
template <int NDIM> __global__ void k_cell_index(CellArgs A) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.ntotal) return;
  int ic[3] = {0, 0, 0};
  bool bad = false;
  for (int d = 0; d < NDIM; d++) {
    ic[d] = __double2int_rz((A.x[(size_t)r * NDIM + d] - A.xminpart[d]) / A.dxcell);   // icellx - 1, :121
    if (ic[d] < 0 || ic[d] >= A.ncellsx[d]) { bad = true; ic[d] = 0; }
  }
  if (bad) atomicCAS(&A.flags[1], 0, ND_ERR_LINK);                                      // :122-125
  const int cell = ic[0] + A.ncellsx[0] * (ic[1] + A.ncellsx[1] * ic[2]);               // :127-135
  // fine bin along x inside the cell (sort key only; the cell is the reference's)
  const double tx = (A.x[(size_t)r * NDIM] - A.xminpart[0]) / A.dxcell;
  int fb = bad ? 0 : __double2int_rz((tx - (double)ic[0]) * CELL_FX);
  fb = fb < 0 ? 0 : (fb > CELL_FX - 1 ? CELL_FX - 1 : fb);
  const int fine = cell * CELL_FX + fb;
  A.cellOfOrig[r] = fine;
  atomicAdd(&A.cellCount[fine], 1);
}
__global__ void k_cell_scatter(const int *fineOfOrig, int ntotal, const int *fineStart, int *fineFill, int *permtmp) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ntotal) return;
  const int fine = fineOfOrig[r];
  permtmp[fineStart[fine] + atomicAdd(&fineFill[fine], 1)] = r;
}
// one warp per cell: order the rows of each of its fine bins by original index (rank sort) so the result is run-to-run deterministic
__global__ void k_cell_order(const int *fineStart, int ncells, const int *permtmp, int *perm) {
  const int cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (cell >= ncells) return;
  const int *fs = fineStart + (size_t)cell * CELL_FX;
  const int a = fs[0], n = fs[CELL_FX] - a;
  for (int e = lane; e < n; e += 32) {
    const int pos = a + e;
    int b = 0;
#pragma unroll
    for (int q = 1; q < CELL_FX; q++) b += (pos >= fs[q]);
    const int ba = fs[b], bn = fs[b + 1] - ba;
    const int mine = permtmp[pos];
    int rank = 0;
    for (int f = 0; f < bn; f++) rank += (permtmp[ba + f] < mine);
    perm[ba + rank] = mine;
  }
}

struct GatherArgs {
  const int *perm, *cellOfOrig, *itype, *ireal; const double *x, *vel, *pmass, *hh;
  double4 *posh, *vm, *posm; float4 *p32; int *typ, *cellOf, *inv, *mixed; int npart, ntotal;
  const double *dustfrac; double *sdf;   // one-fluid dust (NULL otherwise)
  double xminpart[3], dxcell1, hhmax1;
};
template <int NDIM> __global__ void k_gather_sorted(GatherArgs A) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= A.ntotal) return;
  const int r = A.perm[s];
  const int st = (r < A.npart) ? r : A.ireal[r] - 1;   // ghosts carry their parent's properties (makeghost -> copy_particle)
  double p[3] = {0, 0, 0};
  for (int d = 0; d < NDIM; d++) p[d] = A.x[(size_t)r * NDIM + d];
  A.posh[s] = make_double4(p[0], p[1], p[2], 1.0 / A.hh[st]);   // h1(i) = 1./hh(i)
  float q[3] = {0.f, 0.f, 0.f};
  for (int d = 0; d < NDIM; d++) q[d] = (float)((p[d] - A.xminpart[d]) * A.dxcell1);
  A.p32[s] = make_float4(q[0], q[1], q[2], screen_h2(A.hh[st], A.hhmax1));
  if (A.vel) A.vm[s] = make_double4(A.vel[(size_t)r * 3], A.vel[(size_t)r * 3 + 1], A.vel[(size_t)r * 3 + 2], A.pmass[st]);
  A.posm[s] = make_double4(p[0], p[1], p[2], A.pmass[st]);
  A.typ[s] = A.itype[r];
  if (A.sdf) A.sdf[s] = A.dustfrac[st];
  if (A.itype[r] != A.itype[0]) *A.mixed = 1;
  A.cellOf[s] = A.cellOfOrig[r] / CELL_FX;
  A.inv[r] = s;
}
// derivs_host, fast tuple with periodic ghosts only: the LIGHT density rounds never read a velocity, so the link and the density iteration
// run while vel is still on the wire; this fills what was left out -- the ghost rows' velocities (a periodic ghost carries its parent's,
// src/ghostND_mhd.f90:363-431) and the sorted {v, m} records -- before cons2prim and the rates
__global__ void k_late_vel(const int *perm, const int *ireal, double *vel, const double *pmass, double4 *vm, int npart, int ntotal) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ntotal) return;
  const int r = perm[s];
  const int st = (r < npart) ? r : ireal[r] - 1;
  const double vx = vel[(size_t)st * 3], vy = vel[(size_t)st * 3 + 1], vz = vel[(size_t)st * 3 + 2];
  if (r >= npart) { vel[(size_t)r * 3] = vx; vel[(size_t)r * 3 + 1] = vy; vel[(size_t)r * 3 + 2] = vz; }
  vm[s] = make_double4(vx, vy, vz, pmass[st]);
}
__global__ void k_refresh_h(const int *perm, const int *ireal, const double *hh, double4 *posh, float4 *p32, double hhmax1, int npart, int ntotal) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ntotal) return;
  const int r = perm[s];
  const double h = hh[(r < npart) ? r : ireal[r] - 1];
  posh[s].w = 1.0 / h;
  p32[s].w = screen_h2(h, hhmax1);
}
__global__ void k_compact(const int *redo, const int *scan, int n, int *list) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n && redo[s]) list[scan[s]] = s;
}
__global__ void k_remap_list(int *list, int n, const int *oldperm_rows, const int *inv) {   // after a relink: rows -> new slots
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) list[i] = inv[oldperm_rows[i]];
}
__global__ void k_list_rows(const int *list, int n, const int *perm, int *rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rows[i] = perm[list[i]];
}
__global__ void k_minmax_neigh(const int *numneigh, int n, int *out /* [0] min [1] max */) {
  int mn = 1 << 30, mx = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { mn = min(mn, numneigh[i]); mx = max(mx, numneigh[i]); }
  for (int o = 16; o; o >>= 1) { mn = min(mn, __shfl_xor_sync(FULL, mn, o)); mx = max(mx, __shfl_xor_sync(FULL, mx, o)); }
  if ((threadIdx.x & 31) == 0) { atomicMin(out, mn); atomicMax(out + 1, mx); }
}

// sum of the list lengths of one build (nd_scalars.npairs_rates)
// and of the pair-loop trips of the warps that will walk them (a warp owns 32 consecutive targets and loops to its longest list:
// nd_scalars.ntrips_rates, the work figure of the FP64-pipe roofline)
__global__ void k_sum_counts(const int *cnt, int n, int split, unsigned long long *out, unsigned long long *out_trips) {
  unsigned long long s = 0, tr = 0;
  for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {   // blockDim is a multiple of 32: lanes stay aligned to list columns
    const int i = i0 + threadIdx.x;
    const int v = i < n ? cnt[i] : 0;
    // split lists (drag runs): front | back << 16, two loops a warp
    int m = split ? (v & 0xffff) : v, mb = split ? (v >> 16) : 0;
    s += (unsigned long long)(m + mb);
    for (int o = 16; o; o >>= 1) { m = max(m, __shfl_xor_sync(FULL, m, o)); mb = max(mb, __shfl_xor_sync(FULL, mb, o)); }
    tr += (unsigned long long)(m + mb);
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
  if ((threadIdx.x & 31) == 0) { if (s) atomicAdd(out, s); if (tr) atomicAdd(out_trips, tr); }
}

// copies after the density iteration, src/iterate_density.f90:310-344
struct CopyArgs { double *rho, *rhoalt, *drhodt, *dhdt, *hh, *gradh, *gradhn, *gradsoft; const int *itype, *ireal; int npart, ntotal; bool aux; };
__device__ __forceinline__ void copy_density_row(const CopyArgs &A, int i, int j) {
  A.rho[i] = A.rho[j]; A.drhodt[i] = A.drhodt[j]; A.dhdt[i] = A.dhdt[j]; A.hh[i] = A.hh[j]; A.gradh[i] = A.gradh[j];
  if (A.aux) { A.rhoalt[i] = A.rhoalt[j]; A.gradhn[i] = A.gradhn[j]; A.gradsoft[i] = A.gradsoft[j]; }
}
__global__ void k_copy_fixed_density(CopyArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < A.npart && A.itype[i] == T_BND) { const int j = A.ireal[i] - 1; if (j >= 0) copy_density_row(A, i, j); }
}
__global__ void k_copy_ghost_density(CopyArgs A) {
  const int i = A.npart + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < A.ntotal) { const int j = A.ireal[i] - 1; if (j >= 0) copy_density_row(A, i, j); }
}

// =====================================================================================================
// conservative -> primitive + equation of state, element-wise branches of src/conservative2primitive.f90:42-470
// and src/eos.f90:40-105
// =====================================================================================================
struct C2PArgs {
  const double *rho, *en, *Bevol, *vel; const int *itype, *ireal;
  double *dens, *uu, *pr, *spsound, *Bfield;
  // fixed-particle replicas (copy_particle, src/copy_particle.f90:28-115) touch the state arrays too
  double *pmass, *rho_w, *rhoalt, *hh, *en_w, *Bevol_w, *alpha, *psi, *gradh, *gradhn, *gradsoft, *gradgradh;
  int npart, ntotal, imhd, iener; double gamma, polyk; bool aux;
  // one-fluid dust (NULL otherwise)
  const double *dustevol; double *dustfrac, *rhogas, *rhodust;
};
__global__ void k_c2p(C2PArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.npart) return;
  const double rhotot = A.rho[i];
  double rho = rhotot;
  if (A.dustevol) {                                                                 // :76-113 (idustevol = 0)
    double eps = A.dustevol[i];
    if (eps > 1.) eps = 1.;                                                         // :102-104 (the caller's dustevol is not touched)
    A.dustfrac[i] = eps;
    rho = rhotot * (1. - eps);                                                      // :112 dens is the GAS density; the EOS takes it (:418-424)
  }
  A.dens[i] = rho;                                                                  // :117
  double B[3] = {0, 0, 0};
  if (A.imhd >= 11) for (int k = 0; k < 3; k++) B[k] = A.Bevol[(size_t)i * 3 + k];  // :151-152
  else if (A.imhd >= 1) for (int k = 0; k < 3; k++) B[k] = A.Bevol[(size_t)i * 3 + k] * rhotot;   // :190-193
  if (A.imhd != 0) for (int k = 0; k < 3; k++) A.Bfield[(size_t)i * 3 + k] = B[k];
  double uu;
  if (A.iener == 3) {                                                               // :329-346
    const double *v = A.vel + (size_t)i * 3;
    const double v2i = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
    const double B2i = ((B[0] * B[0] + B[1] * B[1]) + B[2] * B[2]) / rhotot;
    uu = A.en[i] - 0.5 * v2i - 0.5 * B2i;
    if (uu < 0.) uu = 0.;
  } else uu = A.en[i];                                                              // :353-368
  const double gamma1 = A.gamma - 1.;
  const int t = A.itype[i];
  if (A.iener == 0) {                                                               // eos.f90:72-89
    double pr = A.pr[i], cs = A.spsound[i];
    if (rho > 0. && t == T_GAS) { pr = A.polyk * pow(rho, A.gamma); cs = sqrt(A.gamma * pr / rho); }
    else if (t == 3 || t == 4) pr = A.polyk * (rho - 1.);
    else if (t != T_BND) pr = 0.;
    if (fabs(gamma1) > 1.e-3 && rho > 0.) uu = pr / (gamma1 * rho);
    A.pr[i] = pr; A.spsound[i] = cs;
  } else if (rho > 0.) {                                                            // eos.f90:96-101
    const double pr = gamma1 * uu * rho;
    A.pr[i] = pr; A.spsound[i] = sqrt(A.gamma * pr / rho);
  }
  A.uu[i] = uu;
}
__global__ void k_c2p_fixed(C2PArgs A) {                                            // :424-437
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.npart) return;
  const int t = A.itype[i];
  if (t != T_BND && t != T_BNDDUST) return;
  const int j = A.ireal[i] - 1;
  if (j < 0) return;
  A.pmass[i] = A.pmass[j]; A.rho_w[i] = A.rho_w[j]; A.hh[i] = A.hh[j]; A.uu[i] = A.uu[j]; A.en_w[i] = A.en_w[j];
  for (int k = 0; k < 3; k++) { A.alpha[(size_t)i * 3 + k] = A.alpha[(size_t)j * 3 + k]; }
  if (A.imhd != 0) for (int k = 0; k < 3; k++) { A.Bevol_w[(size_t)i * 3 + k] = A.Bevol_w[(size_t)j * 3 + k]; A.Bfield[(size_t)i * 3 + k] = A.Bfield[(size_t)j * 3 + k]; }
  A.psi[i] = A.psi[j]; A.gradh[i] = A.gradh[j];
  if (A.aux) { A.rhoalt[i] = A.rhoalt[j]; A.gradhn[i] = A.gradhn[j]; A.gradsoft[i] = A.gradsoft[j]; A.gradgradh[i] = A.gradgradh[j]; }
  A.spsound[i] = A.spsound[j]; A.pr[i] = A.pr[j]; A.dens[i] = A.dens[j];
  if (A.dustevol) { A.dustfrac[i] = A.dustfrac[j]; A.rhogas[i] = A.rhogas[j]; A.rhodust[i] = A.rhodust[j]; }   // copy_particle.f90:84-92
}
__global__ void k_c2p_ghost(C2PArgs A) {                                            // :441-467
  const int i = A.npart + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.ntotal) return;
  const int j = A.ireal[i] - 1;
  if (j < 0) return;
  A.psi[i] = A.psi[j]; A.dens[i] = A.dens[j]; A.uu[i] = A.uu[j]; A.spsound[i] = A.spsound[j]; A.pr[i] = A.pr[j];
  if (A.imhd != 0) for (int k = 0; k < 3; k++) A.Bfield[(size_t)i * 3 + k] = A.Bfield[(size_t)j * 3 + k];
  // :464, and copy_particle for all-periodic boundaries (:465) -- the only ghost configuration accepted with one-fluid dust
  if (A.dustevol) { A.dustfrac[i] = A.dustfrac[j]; A.rhogas[i] = A.rhogas[j]; A.rhodust[i] = A.rhodust[j]; }
}

// get_curl (nd_curl.cuh): the sorted records the operator reads -- converged 1/h, mass, rho, gradh and the vector field; ghost slots carry
// their parent's values (for imhd >= 11 that is what `Bfield = Bevol` leaves in the ghost rows, conservative2primitive.f90:151-152)
struct CurlGatherArgs {
  const int *perm, *ireal; const double *hh, *pmass, *rho, *gradh, *bvec;
  double4 *posh, *vm, *gal, *bsorted; float4 *p32; double *srho; double hhmax1; int npart, ntotal;
};
__global__ void k_curl_gather(CurlGatherArgs A) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= A.ntotal) return;
  const int r = A.perm[s];
  const int st = (r < A.npart) ? r : A.ireal[r] - 1;
  const double h = A.hh[st];
  A.posh[s].w = 1.0 / h;
  A.p32[s].w = screen_h2(h, A.hhmax1);
  A.vm[s].w = A.pmass[st];
  A.srho[s] = A.rho[st];
  A.gal[s].x = A.gradh[st];
  A.bsorted[s] = make_double4(A.bvec[(size_t)st * 3], A.bvec[(size_t)st * 3 + 1], A.bvec[(size_t)st * 3 + 2], 0.);
}
__global__ void k_take_column(const double *a, int stride, int col, double *out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[(size_t)i * stride + col];
}

// =====================================================================================================
// rates: gather of the sorted inputs, finalisation loop (src/ratesND_mhd.f90:532-965)
// =====================================================================================================
struct RGatherArgs {
  const int *perm, *inv, *ireal; const double *hh, *pmass, *rho, *pr, *spsound, *uu, *gradh, *alpha, *psi, *Bfield;
  double4 *posh, *vm, *bpsi, *thermo, *gal; float4 *p32; double hhmax1; double *srho; int npart, ntotal, imhd; unsigned long long *stress_key; int imagforce; double Bconstmax, pext;
  int *err;
  // one-fluid dust (dusta NULL otherwise)
  const double *dustfrac, *deltav, *rhogas, *rhodust; double4 *dusta; double2 *dustb; int use_smoothed_rhodust;
  // iavlim(3) = 2 with ghosts that are not all periodic: the ghosts keep the alpha_B they were created with (copy_particle runs for
  // them only when every boundary is periodic, conservative2primitive.f90:465), i.e. the parent's value BEFORE the switch of this call
  const double *alphaB_ghost;
};
__global__ void k_rates_gather(RGatherArgs A) {
  // one thread per sorted slot (measured: one thread per original row with the records scattered through inv[] is 2.2 times slower)
  double stress = 0.;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < A.ntotal) {
    const int r = A.perm[s];
    const int st = (r < A.npart) ? r : A.ireal[r] - 1;
    const double h = A.hh[st], rho = A.rho[st];
    if (h <= 0.) atomicCAS(A.err, 0, ND_ERR_H_NONPOSITIVE);                         // :384-387
    A.posh[s].w = 1.0 / h;
    A.p32[s].w = screen_h2(h, A.hhmax1);
    A.srho[s] = rho;                                                                // (vm.w = m was set by k_gather_sorted / k_late_vel: masses do not change)
    A.thermo[s] = make_double4(1.0 / rho, fmax(A.pr[st] - A.pext, 0.), A.spsound[st], A.uu[st]);   // rho1i, :325; pri = max(pr - pext, 0), :328
    A.gal[s] = make_double4(A.gradh[st], A.alpha[(size_t)st * 3], A.alpha[(size_t)st * 3 + 1],
                            (A.alphaB_ghost && r >= A.npart) ? A.alphaB_ghost[st] : A.alpha[(size_t)st * 3 + 2]);
    if (A.dusta) {                                                                  // :344-356
      const double eps = A.dustfrac[st];
      A.dusta[s] = make_double4(eps, A.deltav[(size_t)st * 3], A.deltav[(size_t)st * 3 + 1], A.deltav[(size_t)st * 3 + 2]);
      A.dustb[s] = A.use_smoothed_rhodust ? make_double2(A.rhogas[st], A.rhodust[st]) : make_double2((1. - eps) * rho, eps * rho);
    }
    if (A.imhd != 0) {
      const double bx = A.Bfield[(size_t)st * 3], by = A.Bfield[(size_t)st * 3 + 1], bz = A.Bfield[(size_t)st * 3 + 2];
      A.bpsi[s] = make_double4(bx, by, bz, A.psi[st]);
      const double B2i = (bx * bx + by * by) + bz * bz;
      stress = fmax(stress, fmax(fmax(0.5 * B2i - A.pr[st], 0.), A.Bconstmax));     // :240-241
    }
  }
  if (A.imhd != 0) {                                                                // stressmax over 1..ntotal, :231-245
    stress = warp_max(stress);                                                      // one atomic per block, not per warp (same address)
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = stress;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < (int)(blockDim.x >> 5); k++) stress = fmax(stress, red[k]);
      if (stress > 0.) atomic_max_d(A.stress_key, stress);
    }
  }
}

struct FinalArgs {
  const int *perm, *typ; const double4 *posh, *vm, *bpsi, *thermo, *gal; RatesSums S; RatesOpts O;
  const double *drhodt_in, *Bevol, *dens, *hh, *rho, *pr; const unsigned long long *vsigmax_key;
  double *force, *dudt, *dendt, *dBevoldt, *daldt, *dpsidt, *gradpsi, *divB, *curlB, *graddivv, *del2u, *drhodt, *dhdt;
  RatesRed R; int npart, ntotal;
  int drho_from_pairs, ndim;          // fast tuple: drho/dt comes from the pair kernel's sum (S.V.w), dh/dt is made here
  // One thread per ORIGINAL row in [row0,row1) (a row chunk, or every own row): the particle's own state comes from the original-order
  // arrays and the 13 output arrays are written coalesced; only the five 32-byte sum records of the pair kernel are gathered (whole
  // sectors) through inv[].
  const int *inv; int row0, row1; const double *vel, *pmass, *spsound, *uu, *alpha, *psi, *Bfield; const int *itype;
  double *partial;                    // [gridDim.x][6] block partials of the scalar reductions
  // one-fluid dust (dusta NULL otherwise)
  const double4 *dusta; const double2 *dustb; const int *fineStart, *cellOf; double *ddustevoldt, *ddeltavdt;
};
__global__ void k_rates_final(FinalArgs A) {
  double fhmax = 0., dtforce = DBL_MAX, fm0 = 0., fm1 = 0., fm2 = 0., tsmin = DBL_MAX;
  {
    for (int i = A.row0 + blockIdx.x * blockDim.x + threadIdx.x; i < A.row1; i += gridDim.x * blockDim.x) {
      const RatesOpts &O = A.O;
      const int s = A.inv[i];
      const double4 F = ld4(A.S.F + s), dB4 = ld4(A.S.dB + s), C = ld4(A.S.C + s), P = ld4(A.S.P + s), V = ld4(A.S.V + s);
      const double rhoi = A.rho[i], hi = A.hh[i], pri = A.pr[i];
      const double rho1i = 1.0 / rhoi;                                                               // rho1i = 1./rhoi, :325 (thermo.x of the pair kernel)
      const double4 v = make_double4(A.vel[(size_t)i * 3], A.vel[(size_t)i * 3 + 1], A.vel[(size_t)i * 3 + 2], A.pmass[i]);
      const double4 th = make_double4(rho1i, 0., A.spsound[i], A.uu[i]);
      const double4 g = make_double4(0., A.alpha[(size_t)i * 3], A.alpha[(size_t)i * 3 + 1], A.alpha[(size_t)i * 3 + 2]);
      const double vsigmax = dkey_inv(*A.vsigmax_key);
      const double vsig2max = (O.imhd != 0 && O.idivbzero >= 2) ? vsigmax * vsigmax : 0.;           // :518-520
      double fx = F.x, fy = F.y, fz = F.z, dudt = F.w;
      double divB = dB4.w, cbx = C.x, cby = C.y, cbz = C.z;
      double bx = 0, by = 0, bz = 0, psii = 0;
      if (O.imhd != 0) {
        bx = A.Bfield[(size_t)i * 3]; by = A.Bfield[(size_t)i * 3 + 1]; bz = A.Bfield[(size_t)i * 3 + 2]; psii = A.psi[i];
        if (O.imhd > 0) { cbx *= rho1i; cby *= rho1i; cbz *= rho1i; }                                // :640
        divB *= rho1i;                                                                               // :643
      }
      fm0 += v.w * fx; fm1 += v.w * fy; fm2 += v.w * fz;                                             // :678
      const double forcemag = sqrt((fx * fx + fy * fy) + fz * fz);
      const double fonh = forcemag / hi;
      const int ti = A.itype[i];
      if (ti != 1) fhmax = fmax(fhmax, fonh);                                                                     // :681
      double valfven2i = 0.;
      if (O.imhd != 0) valfven2i = ((bx * bx + by * by) + bz * bz) / A.dens[i];                      // :690
      const double vsig = sqrt(th.z * th.z + valfven2i);                                             // :695-696
      // drho/dt: the density iteration's (iterate_density.f90:272), or -- fast tuple -- the pair kernel's sum over the same pairs with the
      // same grad W (already times gradh); dh/dt = dhdrho * drho/dt (:273) is then made here
      const double drhodti = A.drho_from_pairs ? V.w : A.drhodt_in[i];
      if (A.drho_from_pairs) { A.drhodt[i] = drhodti; A.dhdt[i] = (-hi / (A.ndim * rhoi)) * drhodti; }
      double dbx = dB4.x, dby = dB4.y, dbz = dB4.z, gpx = P.x, gpy = P.y, gpz = P.z;
      if (O.imhd >= 11) {                                                                            // :722-730
        const double *Be = A.Bevol + (size_t)i * 3;
        dbx = dbx + Be[0] * rho1i * drhodti; dby = dby + Be[1] * rho1i * drhodti; dbz = dbz + Be[2] * rho1i * drhodti;
        if (O.idivbzero >= 2) {
          gpx *= rhoi; gpy *= rhoi; gpz *= rhoi;
          if (O.nsubsteps_divB <= 0) { dbx += gpx; dby += gpy; dbz += gpz; }
        }
      } else if (O.imhd >= 1) {                                                                      // :733-752
        dbx *= rho1i; dby *= rho1i; dbz *= rho1i;
        if (O.idivbzero >= 2) { const double r2 = rho1i * rho1i; gpx *= r2; gpy *= r2; gpz *= r2; }
      } else { dbx = dby = dbz = 0.; }
      if (O.iresist > 0 && O.iresist != 2 && O.etamhd > DBL_MIN) dtforce = fmin(dtforce, hi * hi / O.etamhd);       // :808-815
      double ddv0 = 0., ddv1 = 0., ddv2 = 0., ddust = 0., tstop = DBL_MAX;
      if (A.dusta) {                                                                                 // :548-582 one fluid dust
        const double4 D = A.S.D[s], da = A.dusta[s];
        const double2 db = A.dustb[s];
        ddv0 = D.x; ddv1 = D.y; ddv2 = D.z; ddust = D.w;
        tstop = get_tstop(O.idrag_nature, db.x, db.y, O.Kdrag);
        // :566 tests `dustfraci`, which the reference last assigned in the pair loop: it belongs to the LAST particle that loop
        // visited -- the lowest-index row of the last non-empty cell, i.e. the first slot of the last slot's cell
        // (slots inside a cell are ordered by (fine bin, index): search the cell for its lowest row)
        const int lastcell = A.cellOf[A.ntotal - 1];
        int sl = A.fineStart[(size_t)lastcell * CELL_FX];
        for (int q = sl + 1, qe = A.fineStart[(size_t)(lastcell + 1) * CELL_FX]; q < qe; q++) if (A.perm[q] < A.perm[sl]) sl = q;
        const double dustfrac_stale = A.dusta[sl].x;
        double dtstop;
        if (dustfrac_stale > 0.) { dtstop = 1. / tstop; ddv0 -= da.y * dtstop; ddv1 -= da.z * dtstop; ddv2 -= da.w * dtstop; }
        else { dtstop = 0.; ddv0 = ddv1 = ddv2 = 0.; }
        if (O.iener > 0) dudt = dudt + db.y * rho1i * ((da.y * da.y + da.z * da.z) + da.w * da.w) * dtstop;   // :579-582
      }
      double dendt;
      if (O.iener == 3) {                                                                            // :820-826 (+ pair part :1829)
        dudt = dudt + pri * (rho1i * rho1i) * drhodti;
        dendt = ((v.x * fx + v.y * fy) + v.z * fz) + dudt;
        // NOTE: the reference overwrites the pair-summed dendt here (:824); P.w is therefore discarded
      } else if (O.iener > 0 && O.iav >= 0 && O.idust != 1) {                                        // :832-835
        dudt = dudt + pri * (rho1i * rho1i) * drhodti;
        dendt = dudt;
      } else dendt = dudt;                                                                           // :837
      if (ti == T_DUST) dendt = 0.;                                                                  // :839
      double da0 = 0., da1 = 0., da2 = 0.;
      if (O.iavlim0 != 0 || O.iavlim1 != 0 || O.iavlim2 != 0) {                                      // :845-896
        const double tdecay1 = (O.avdecayconst * vsig) / hi;
        if (O.iavlim0 == 1 || O.iavlim0 == 2) {
          double source = fmax(drhodti * rho1i, 0.0);
          if (O.iavlim0 == 2) source = source * (2.0 - g.y);
          da0 = (O.alphamin - g.y) * tdecay1 + O.avfact * source;
        } else if (O.iavlim0 == 3) {
          const double graddivvmag = sqrt((V.x * V.x + V.y * V.y) + V.z * V.z);
          da0 = (O.alphamin - g.y) * tdecay1 + O.avfact * (hi * graddivvmag * (2.0 - g.y));
        }
        if (O.iener > 0 && O.iavlim1 > 0) {
          const double sourceu = (th.w > 2.220446049250313e-16) ? hi * fabs(C.w) / sqrt(th.w) : 0.;
          da1 = (O.alphaumin - g.z) * tdecay1 + sourceu;
        }
        if (O.iavlim2 != 0 && O.imhd != 0) {
          const double sourceJ = sqrt(((cbx * cbx + cby * cby) + cbz * cbz) * rho1i);
          const double sourcedivB = 10. * fabs(divB) * sqrt(rho1i);
          double sourceB = fmax(sourceJ, sourcedivB);
          if (O.iavlim2 == 2) sourceB = sourceB * (2.0 - g.w);
          else if (O.iavlim2 == 3) { const double source = fmax(drhodti * rho1i, 0.0) * (2. - g.w); sourceB = sqrt(source * sourceB); }
          da2 = (O.alphaBmin - g.w) * tdecay1 + sourceB;
        }
      }
      double dpsidt = 0.;
      if (O.idivbzero >= 2 && O.idivbzero <= 7) dpsidt = -vsig2max * divB - O.psidecayfact * psii * vsigmax / hi;   // :902
      // zero rates on fixed particles, :949-965
      const bool fixed = (ti == T_BND || ti == T_BNDDUST);
      if (fixed) {
        fx = fy = fz = dudt = dendt = dbx = dby = dbz = da0 = da1 = da2 = dpsidt = divB = cbx = cby = cbz = gpx = gpy = gpz = 0.;
        A.drhodt[i] = 0.; A.dhdt[i] = 0.;
      }
      double *o3;
      o3 = A.force + (size_t)i * 3; o3[0] = fx; o3[1] = fy; o3[2] = fz;
      A.dudt[i] = dudt; A.dendt[i] = dendt;
      if (A.dBevoldt) { o3 = A.dBevoldt + (size_t)i * 3; o3[0] = dbx; o3[1] = dby; o3[2] = dbz; }
      o3 = A.daldt + (size_t)i * 3; o3[0] = da0; o3[1] = da1; o3[2] = da2;
      A.dpsidt[i] = dpsidt;
      o3 = A.gradpsi + (size_t)i * 3; o3[0] = gpx; o3[1] = gpy; o3[2] = gpz;
      A.divB[i] = divB;
      o3 = A.curlB + (size_t)i * 3; o3[0] = cbx; o3[1] = cby; o3[2] = cbz;
      o3 = A.graddivv + (size_t)i * 3; o3[0] = V.x; o3[1] = V.y; o3[2] = V.z;
      A.del2u[i] = C.w;
      if (A.dusta) {
        A.ddustevoldt[i] = ddust;
        o3 = A.ddeltavdt + (size_t)i * 3; o3[0] = ddv0; o3[1] = ddv1; o3[2] = ddv2;
        tsmin = fmin(tsmin, tstop);
      }
    }
  }
  // block partials {fhmax, dtforce, sum m f (3), min tstop} in a fixed order, no atomics: k_final_reduce folds them (a set of atomics per
  // warp of 32 rows was 2.6 M same-address atomics at 16.8 M particles -- 2.5 of this kernel's 4.3 ms)
  fhmax = warp_max(fhmax); dtforce = warp_min(dtforce); tsmin = warp_min(tsmin);
  fm0 = warp_sum(fm0); fm1 = warp_sum(fm1); fm2 = warp_sum(fm2);
  __shared__ double red[8][6];
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[w][0] = fhmax; red[w][1] = dtforce; red[w][2] = fm0; red[w][3] = fm1; red[w][4] = fm2; red[w][5] = tsmin; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < nw; k++) {
      fhmax = fmax(fhmax, red[k][0]); dtforce = fmin(dtforce, red[k][1]); fm0 += red[k][2]; fm1 += red[k][3]; fm2 += red[k][4]; tsmin = fmin(tsmin, red[k][5]);
    }
    double *o = A.partial + (size_t)blockIdx.x * 6;
    o[0] = fhmax; o[1] = dtforce; o[2] = fm0; o[3] = fm1; o[4] = fm2; o[5] = tsmin;
  }
}
// one block: the partials of k_rates_final in block order (run-to-run deterministic force sum), then into the reduction keys
__global__ void k_final_reduce(const double *partial, int nb, RatesRed R, int dust) {
  double fhmax = 0., dtforce = DBL_MAX, fm0 = 0., fm1 = 0., fm2 = 0., tsmin = DBL_MAX;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    const double *o = partial + (size_t)b * 6;
    fhmax = fmax(fhmax, o[0]); dtforce = fmin(dtforce, o[1]); fm0 += o[2]; fm1 += o[3]; fm2 += o[4]; tsmin = fmin(tsmin, o[5]);
  }
  fhmax = warp_max(fhmax); dtforce = warp_min(dtforce); tsmin = warp_min(tsmin);
  fm0 = warp_sum(fm0); fm1 = warp_sum(fm1); fm2 = warp_sum(fm2);
  __shared__ double red[32][6];
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[w][0] = fhmax; red[w][1] = dtforce; red[w][2] = fm0; red[w][3] = fm1; red[w][4] = fm2; red[w][5] = tsmin; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < nw; k++) {
      fhmax = fmax(fhmax, red[k][0]); dtforce = fmin(dtforce, red[k][1]); fm0 += red[k][2]; fm1 += red[k][3]; fm2 += red[k][4]; tsmin = fmin(tsmin, red[k][5]);
    }
    atomic_max_d(R.fhmax_max, fhmax);
    atomic_min_d(R.dtforce_min, dtforce);
    R.fmean[0] += fm0; R.fmean[1] += fm1; R.fmean[2] += fm2;      // the row chunks of one get_rates run one after the other on the stream
    if (dust) atomic_min_d(R.ts_min, tsmin);
  }
}
// Row-chunked rates: dpsidt needs the maximum signal velocity over ALL pairs (:518-520, :902), known only after the last chunk.
__global__ void k_rates_dpsidt(const double *divB, const double *psi, const double *hh, const int *itype, const unsigned long long *vsigmax_key,
                               double psidecayfact, int imhd, int idivbzero, double *dpsidt, int npart) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npart) return;
  const double vsigmax = dkey_inv(*vsigmax_key);
  const double vsig2max = (imhd != 0 && idivbzero >= 2) ? vsigmax * vsigmax : 0.;
  double v = 0.;
  if (idivbzero >= 2 && idivbzero <= 7) v = -vsig2max * divB[i] - psidecayfact * psi[i] * vsigmax / hh[i];
  const int ti = itype[i];
  if (ti == T_BND || ti == T_BNDDUST) v = 0.;
  dpsidt[i] = v;
}
// flags the sorted slots whose original row lies in [row0,row1) (rows below nown only): the chunk's targets, in slot order
__global__ void k_chunk_flags(const int *perm, int ntotal, int row0, int row1, int *flag) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < ntotal) { const int r = perm[s]; flag[s] = (r >= row0 && r < row1) ? 1 : 0; }
}

struct ZeroArgs { double *force, *dudt, *dendt, *dBevoldt, *daldt, *dpsidt, *gradpsi, *divB, *curlB, *graddivv, *del2u, *drhodt, *dhdt; int npart, ntotal; };
__global__ void k_rates_zero_ghosts(ZeroArgs A) {                                    // :949-965 for rows > npart
  const int i = A.npart + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.ntotal) return;
  for (int k = 0; k < 3; k++) {
    A.force[(size_t)i * 3 + k] = 0.; A.daldt[(size_t)i * 3 + k] = 0.; A.gradpsi[(size_t)i * 3 + k] = 0.; A.curlB[(size_t)i * 3 + k] = 0.;
    A.graddivv[(size_t)i * 3 + k] = 0.;
    if (A.dBevoldt) A.dBevoldt[(size_t)i * 3 + k] = 0.;
  }
  A.dudt[i] = 0.; A.dendt[i] = 0.; A.dpsidt[i] = 0.; A.divB[i] = 0.; A.del2u[i] = 0.; A.drhodt[i] = 0.; A.dhdt[i] = 0.;
}

// =====================================================================================================
// per-step diagnostics: the particle loop of `evwrite`, src/evwrite_mhd.f90:124-266, as a two-stage
// fixed-order reduction (EV_BLOCKS partial rows, then one block): deterministic run to run.
// =====================================================================================================
enum { EV_EKIN = 0, EV_ETHERM, EV_EMAG, EV_EMAGP, EV_MOM0, EV_MOM1, EV_MOM2, EV_DMOM0, EV_DMOM1, EV_DMOM2, EV_ANG0, EV_ANG1, EV_ANG2, EV_EKINY,
       EV_MGAS, EV_MDUST, EV_BETAAV, EV_DIVBAV, EV_DIVBTOT, EV_OMEGAAV, EV_FRACOK, EV_FLUX0, EV_FLUX1, EV_FLUX2, EV_CROSSHEL, EV_RHOSUM, EV_NSUM,
       EV_BETAMIN = EV_NSUM, EV_RHOMIN, EV_NMIN_END, EV_DIVBMAX = EV_NMIN_END, EV_OMEGAMAX, EV_RHOMAX, EV_NQ };
constexpr int EV_BLOCKS = 592, EV_THREADS = 256;
struct EvArgs {
  const double *x, *vel, *pmass, *rho, *uu, *Bfield, *pr, *divB, *hh, *force, *dustfrac, *deltav;
  int npart, ndim, imhd, onef;
  double *partial;   // [EV_BLOCKS][EV_NQ]
};
__device__ __forceinline__ double ev_combine(int q, double a, double b) { return q < EV_NSUM ? a + b : (q < EV_NMIN_END ? fmin(a, b) : fmax(a, b)); }
__device__ __forceinline__ double ev_identity(int q) { return q < EV_NSUM ? 0. : (q < EV_NMIN_END ? 1.7976931348623157e308 : 0.); }

__global__ void __launch_bounds__(EV_THREADS) k_evwrite_partial(EvArgs A) {
  double acc[EV_NQ];
#pragma unroll
  for (int q = 0; q < EV_NQ; q++) acc[q] = ev_identity(q);
  for (int i = blockIdx.x * EV_THREADS + threadIdx.x; i < A.npart; i += EV_BLOCKS * EV_THREADS) {
    const double m = A.pmass[i], rhoi = A.rho[i];
    const double v0 = A.vel[(size_t)i * 3], v1 = A.vel[(size_t)i * 3 + 1], v2 = A.vel[(size_t)i * 3 + 2];
    double xi[3] = {0., 0., 0.};
    for (int d = 0; d < A.ndim; d++) xi[d] = A.x[(size_t)i * A.ndim + d];
    acc[EV_MOM0] += m * v0; acc[EV_MOM1] += m * v1; acc[EV_MOM2] += m * v2;                                   // :129
    acc[EV_DMOM0] += m * A.force[(size_t)i * 3]; acc[EV_DMOM1] += m * A.force[(size_t)i * 3 + 1]; acc[EV_DMOM2] += m * A.force[(size_t)i * 3 + 2];
    if (A.ndim == 3) {                                                                                        // :131-136
      acc[EV_ANG0] += m * (xi[1] * v2 - xi[2] * v1); acc[EV_ANG1] += m * (xi[2] * v0 - xi[0] * v2); acc[EV_ANG2] += m * (xi[0] * v1 - xi[1] * v0);
    } else if (A.ndim == 2) acc[EV_ANG2] += m * (xi[0] * v1 - xi[1] * v0);
    acc[EV_EKIN] += 0.5 * m * ((v0 * v0 + v1 * v1) + v2 * v2);                                                // :137
    if (A.onef) {                                                                                             // :139-147
      const double df = A.dustfrac[i], dterm = 1. - df;
      const double d0 = A.deltav[(size_t)i * 3], d1 = A.deltav[(size_t)i * 3 + 1], d2 = A.deltav[(size_t)i * 3 + 2];
      const double ekdv = 0.5 * m * df * dterm * ((d0 * d0 + d1 * d1) + d2 * d2);
      acc[EV_EKIN] += ekdv; acc[EV_EKINY] += ekdv;
      acc[EV_ETHERM] += m * A.uu[i] * dterm;
      acc[EV_MGAS] += m * dterm; acc[EV_MDUST] += m * df;
    } else {
      if (A.ndim >= 2) acc[EV_EKINY] += 0.5 * m * v0 * v0;                                                    // :149 (sic: vel(1,i))
      acc[EV_ETHERM] += m * A.uu[i];
    }
    acc[EV_RHOSUM] += rhoi; acc[EV_RHOMIN] = fmin(acc[EV_RHOMIN], rhoi); acc[EV_RHOMAX] = fmax(acc[EV_RHOMAX], rhoi);   // minmaxave, :269
    if (A.imhd != 0) {                                                                                        // :172-263
      const double b0 = A.Bfield[(size_t)i * 3], b1 = A.Bfield[(size_t)i * 3 + 1], b2 = A.Bfield[(size_t)i * 3 + 2];
      const double B2 = (b0 * b0 + b1 * b1) + b2 * b2, Bmag = sqrt(B2), divBi = fabs(A.divB[i]);
      acc[EV_EMAG] += 0.5 * m * B2 / rhoi;
      acc[EV_EMAGP] += 0.5 * m * (b0 * b0 + b1 * b1) / rhoi;
      const double beta = (B2 < 2.2250738585072014e-308) ? 0. : A.pr[i] / (0.5 * B2);
      acc[EV_BETAAV] += beta; acc[EV_BETAMIN] = fmin(acc[EV_BETAMIN], beta);
      acc[EV_DIVBMAX] = fmax(acc[EV_DIVBMAX], divBi); acc[EV_DIVBAV] += divBi; acc[EV_DIVBTOT] += m * divBi / rhoi;
      const double omega = (Bmag < 1e-8) ? 0. : divBi * A.hh[i] / Bmag;                                       // :240-247
      if (omega < 1.e-2) acc[EV_FRACOK] += 1.;
      acc[EV_OMEGAMAX] = fmax(acc[EV_OMEGAMAX], omega); acc[EV_OMEGAAV] += omega;
      const double br0 = b0 / rhoi, br1 = b1 / rhoi, br2 = b2 / rhoi;
      acc[EV_FLUX0] += m * br0; acc[EV_FLUX1] += m * br1; acc[EV_FLUX2] += m * br2;                           // :255
      acc[EV_CROSSHEL] += m * ((v0 * br0 + v1 * br1) + v2 * br2);                                             // :259
    }
  }
  __shared__ double sh[EV_THREADS / 32][EV_NQ];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < EV_NQ; q++) {
    double v = acc[q];
    for (int o = 16; o; o >>= 1) v = ev_combine(q, v, __shfl_xor_sync(FULL, v, o));
    if (lane == 0) sh[w][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < EV_NQ) {
    double v = sh[0][threadIdx.x];
    for (int k = 1; k < EV_THREADS / 32; k++) v = ev_combine(threadIdx.x, v, sh[k][threadIdx.x]);
    A.partial[(size_t)blockIdx.x * EV_NQ + threadIdx.x] = v;
  }
}
__global__ void k_evwrite_final(const double *partial, double *out_pinned) {
  const int q = threadIdx.x;
  if (q >= EV_NQ) return;
  double v = partial[q];
  for (int b = 1; b < EV_BLOCKS; b++) v = ev_combine(q, v, partial[(size_t)b * EV_NQ + q]);
  out_pinned[q] = v;
  __threadfence_system();
}

// =====================================================================================================
// leapfrog integrator: `step`, src/stepND_leapfrog_mhd.f90:39-300, and the periodic wrap of
// `boundary`, src/boundaryND.f90:65-93.  Element-wise over rows [0,npart); the `*in` copies live in one buffer of
// STEP_NIN(+dust) planes of npart doubles.
// =====================================================================================================
struct StepArgs {
  double *x, *vel, *Bevol, *rho, *hh, *en, *alpha, *psi, *dustevol, *deltav;                     // state (in place)
  const double *force, *dBevoldt, *drhodt, *dhdt, *dendt, *daldt, *dpsidt, *ddustevoldt, *ddeltavdt;   // rates of the last derivs
  const int *itype, *ireal;
  double *in;            // `*in` planes
  size_t n;              // plane stride = npart
  int npart, ndim;
  int imhd, iresist, icty, ihvar, iener, idivbzero, idust, onef, iavlim[3], ibound[3];
  double dt, damp, xmin[3], xmax[3];
  int *flags;
};
// plane offsets (in units of n doubles)
enum { SP_X = 0, SP_VEL = 3, SP_BEVOL = 6, SP_RHO = 9, SP_HH = 10, SP_EN = 11, SP_ALPHA = 12, SP_PSI = 15, SP_FORCE = 16, SP_DBEVOL = 19, SP_DRHO = 22,
       SP_DH = 23, SP_DEN = 24, SP_DAL = 25, SP_DPSI = 28, STEP_NIN = 29, SP_DUSTEVOL = 29, SP_DDUSTEVOL = 30, SP_DELTAV = 31, SP_DDELTAV = 34, STEP_NIN_DUST = 37 };

__global__ void k_step_save(StepArgs A) {                                                 // :70-100
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.npart) return;
  double *in = A.in + i; const size_t n = A.n;
  for (int d = 0; d < A.ndim; d++) in[(SP_X + d) * n] = A.x[(size_t)i * A.ndim + d];
  for (int d = 0; d < 3; d++) {
    in[(SP_VEL + d) * n] = A.vel[(size_t)i * 3 + d]; in[(SP_BEVOL + d) * n] = A.Bevol[(size_t)i * 3 + d]; in[(SP_ALPHA + d) * n] = A.alpha[(size_t)i * 3 + d];
    in[(SP_FORCE + d) * n] = A.force[(size_t)i * 3 + d]; in[(SP_DBEVOL + d) * n] = A.dBevoldt[(size_t)i * 3 + d]; in[(SP_DAL + d) * n] = A.daldt[(size_t)i * 3 + d];
  }
  in[SP_RHO * n] = A.rho[i]; in[SP_HH * n] = A.hh[i]; in[SP_EN * n] = A.en[i]; in[SP_PSI * n] = A.psi[i];
  in[SP_DRHO * n] = A.drhodt[i]; in[SP_DH * n] = A.dhdt[i]; in[SP_DEN * n] = A.dendt[i]; in[SP_DPSI * n] = A.dpsidt[i];
  if (A.onef) {
    in[SP_DUSTEVOL * n] = A.dustevol[i]; in[SP_DDUSTEVOL * n] = A.ddustevoldt[i];
    if (A.idust == 1) for (int d = 0; d < 3; d++) { in[(SP_DELTAV + d) * n] = A.deltav[(size_t)i * 3 + d]; in[(SP_DDELTAV + d) * n] = A.ddeltavdt[(size_t)i * 3 + d]; }
  }
}
__device__ __forceinline__ bool step_fixed(int it) { return it == T_BND || it == 11 /*itypebnd2*/ || it == T_BNDDUST; }

__global__ void k_step_predict(StepArgs A) {                                              // :108-163
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.npart) return;
  const double *in = A.in + i; const size_t n = A.n;
  const double dt = A.dt;
  const int it = A.itype[i];
  if (step_fixed(it)) {
    if (it == 11) { atomicCAS(&A.flags[1], 0, ND_ERR_UNSUPPORTED_OPTION); return; }       // itypebnd2: cylindrical fixed particles
    const int j = A.ireal[i];
    const double *pin = (j > 0) ? A.in + (j - 1) : in;                                    // :112-117
    for (int d = 0; d < A.ndim; d++) A.x[(size_t)i * A.ndim + d] = in[(SP_X + d) * n] + dt * pin[(SP_VEL + d) * n] + 0.5 * dt * dt * pin[(SP_FORCE + d) * n];
    // Bevol, rho, hh, en, alpha, psi (and the dust variables) are reset to their `in` values, i.e. left as they are (:130-142)
  } else {
    for (int d = 0; d < A.ndim; d++) A.x[(size_t)i * A.ndim + d] = in[(SP_X + d) * n] + dt * in[(SP_VEL + d) * n] + 0.5 * dt * dt * in[(SP_FORCE + d) * n];   // :145
    for (int d = 0; d < 3; d++) A.vel[(size_t)i * 3 + d] = (in[(SP_VEL + d) * n] + dt * in[(SP_FORCE + d) * n]) / (1. + A.damp);                              // :146
    if (A.imhd != 0 && A.iresist != 2) for (int d = 0; d < 3; d++) A.Bevol[(size_t)i * 3 + d] = in[(SP_BEVOL + d) * n] + dt * in[(SP_DBEVOL + d) * n];
    double rho = in[SP_RHO * n];
    if (A.icty >= 1) { rho = in[SP_RHO * n] + dt * in[SP_DRHO * n]; A.rho[i] = rho; }
    if (A.ihvar == 1) A.hh[i] = in[SP_HH * n] * pow(in[SP_RHO * n] / rho, 1. / A.ndim);   // :151
    else if (A.ihvar == 2 || A.ihvar == 3) A.hh[i] = in[SP_HH * n] + dt * in[SP_DH * n];
    if (A.iener != 0) A.en[i] = in[SP_EN * n] + dt * in[SP_DEN * n];
    for (int d = 0; d < 3; d++) if (A.iavlim[d] != 0) A.alpha[(size_t)i * 3 + d] = fmin(in[(SP_ALPHA + d) * n] + dt * in[(SP_DAL + d) * n], 1.0);
    if (A.idivbzero >= 2) A.psi[i] = in[SP_PSI * n] + dt * in[SP_DPSI * n];
    if (A.onef) {
      A.dustevol[i] = in[SP_DUSTEVOL * n] + dt * in[SP_DDUSTEVOL * n];
      if (A.idust == 1) for (int d = 0; d < 3; d++) A.deltav[(size_t)i * 3 + d] = in[(SP_DELTAV + d) * n] + dt * in[(SP_DDELTAV + d) * n];
    }
  }
}
__global__ void k_step_correct(StepArgs A) {                                              // :171-209, then boundaryND.f90:65-93
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.npart) return;
  const double *in = A.in + i; const size_t n = A.n;
  const double dt = A.dt, hdt = 0.5 * A.dt;
  const int it = A.itype[i];
  if (step_fixed(it)) {
    if (it == T_BND || it == T_BNDDUST) for (int d = 0; d < 3; d++) A.vel[(size_t)i * 3 + d] = in[(SP_VEL + d) * n];
    for (int d = 0; d < 3; d++) { A.Bevol[(size_t)i * 3 + d] = in[(SP_BEVOL + d) * n]; A.alpha[(size_t)i * 3 + d] = in[(SP_ALPHA + d) * n]; }
    A.rho[i] = in[SP_RHO * n]; A.hh[i] = in[SP_HH * n]; A.en[i] = in[SP_EN * n]; A.psi[i] = in[SP_PSI * n];   // derivs rewrote rho, hh of fixed rows
    if (A.idust == 1 || A.idust == 3 || A.idust == 4) {
      if (A.onef) A.dustevol[i] = in[SP_DUSTEVOL * n];
      if (A.idust == 1) for (int d = 0; d < 3; d++) A.deltav[(size_t)i * 3 + d] = in[(SP_DELTAV + d) * n];
    }
  } else {
    for (int d = 0; d < 3; d++) A.vel[(size_t)i * 3 + d] = (in[(SP_VEL + d) * n] + hdt * (A.force[(size_t)i * 3 + d] + in[(SP_FORCE + d) * n])) / (1. + A.damp);   // :187
    if (A.imhd != 0) for (int d = 0; d < 3; d++) {
      if (A.iresist == 2) A.Bevol[(size_t)i * 3 + d] = in[(SP_BEVOL + d) * n] + dt * A.dBevoldt[(size_t)i * 3 + d];
      else A.Bevol[(size_t)i * 3 + d] = in[(SP_BEVOL + d) * n] + hdt * (A.dBevoldt[(size_t)i * 3 + d] + in[(SP_DBEVOL + d) * n]);
    }
    if (A.icty >= 1) A.rho[i] = in[SP_RHO * n] + hdt * (A.drhodt[i] + in[SP_DRHO * n]);
    if (A.ihvar == 2) {
      const double h = in[SP_HH * n] + hdt * (A.dhdt[i] + in[SP_DH * n]);
      A.hh[i] = h;
      if (h <= 0.) atomicCAS(&A.flags[1], 0, ND_ERR_H_NONPOSITIVE);                        // :197-200
    }
    if (A.iener != 0) A.en[i] = in[SP_EN * n] + hdt * (A.dendt[i] + in[SP_DEN * n]);
    for (int d = 0; d < 3; d++) if (A.iavlim[d] != 0) A.alpha[(size_t)i * 3 + d] = fmin(in[(SP_ALPHA + d) * n] + hdt * (A.daldt[(size_t)i * 3 + d] + in[(SP_DAL + d) * n]), 1.0);
    if (A.idivbzero >= 2) A.psi[i] = in[SP_PSI * n] + hdt * (A.dpsidt[i] + in[SP_DPSI * n]);
    if (A.onef) {
      A.dustevol[i] = in[SP_DUSTEVOL * n] + hdt * (A.ddustevoldt[i] + in[SP_DDUSTEVOL * n]);
      if (A.idust == 1) for (int d = 0; d < 3; d++) A.deltav[(size_t)i * 3 + d] = in[(SP_DELTAV + d) * n] + hdt * (A.ddeltavdt[(size_t)i * 3 + d] + in[(SP_DDELTAV + d) * n]);
    }
  }
}
// particles cross the periodic domain: `boundary`, src/boundaryND.f90:65-93 -- called at the top of derivs (src/derivs.f90:74, before
// the ghosts: set_ghost_particles makes no ghost of a particle on or over the boundary) and after the corrector (:216)
__global__ void k_step_boundary(StepArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.npart) return;
  for (int d = 0; d < A.ndim; d++) if (A.ibound[d] == 3) {
    double xx = A.x[(size_t)i * A.ndim + d];
    if (xx > A.xmax[d]) xx = A.xmin[d] + xx - A.xmax[d];
    else if (xx < A.xmin[d]) xx = A.xmax[d] - (A.xmin[d] - xx);
    A.x[(size_t)i * A.ndim + d] = xx;
  }
}

