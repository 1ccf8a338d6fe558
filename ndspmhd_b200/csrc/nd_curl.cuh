// get_curl on the pair engine.
//
// Replaces `get_curl` (src/get_curl.f90:64-287): the reference visits each pair once and updates both particles; here every real
// particle gathers its own side over the same neighbour list the rates use (build_lists_kernel<LIST_RATES>: j /= i, not both ghosts,
// q2i < radkern2 or q2j < radkern2, the type rule of :161-167).  For icurltype 1, 3, 4 a target's sum only involves grad W(h_target)
// and (B_i - B_j) x dr_ij, which is unchanged under the role swap, so the gather reproduces each of the reference's terms; icurltype 2
// (symmetric operator) carries both kernels and flips sign with dr, as the reference's `curlB(:,j) = curlB(:,j) - pmassi*curlBterm`.
// Call site on the path: conservative2primitive.f90:299-323 (iavlim(3) = 2, the Tricco & Price 2013 resistivity switch), which needs
// grad B as well: alpha_B = min(h |grad B| / |B|, 1).
#pragma once
#include "nd_device.cuh"

namespace ndk {

struct CurlArgs {
  const double4 *bvec;     // [ntotal] sorted {Bx, By, Bz, -} of the vector whose curl is taken (ghost slots: the parent's)
  const double *srho;      // [ntotal] sorted rho
  const double4 *gal;      // [ntotal] sorted, .x = gradh
  double *curlB;           // (3, rows) original order, out
  double *gradB;           // (3,3, rows) original order, gradB(l,k,i) at [(i*3 + k)*3 + l], out; NULL = not wanted (icurltype 1 only)
  double *alpha;           // (3, rows) original order: alpha(3,i) <- the resistivity switch; NULL = operator only
  const double *hh;        // original order (switch only)
  int icurltype;
  double weight;           // 1/hfact**ndim, :115
  int s0, ntargets; const int *targets;
};

template <int NDIM>
__global__ void __launch_bounds__(128) curl_pair_kernel(Grid G, CurlArgs A, NbrLists L) {
  const int tix = blockIdx.x * 128 + threadIdx.x;
  if (tix >= A.ntargets) return;
  const int s = A.targets ? A.targets[tix] : A.s0 + tix;
  const int orig = G.perm[s];
  if (orig >= G.nown) return;                                   // :275 `do i=1,npart`: ghost rows take their parent's values afterwards
  const double4 p = ld4(G.posh + s), bi = ld4(A.bvec + s);
  const double hi1 = p.w, hi21 = __dmul_rn(hi1, hi1);
  const double hfacwabi = powndim<NDIM>(hi1);
  const double rhoi = A.srho[s], gradhi = A.gal[s].x;
  const double rho21i = 1. / (rhoi * rhoi), rho21gradhi = rho21i * gradhi;       // :146-147
  double cx = 0, cy = 0, cz = 0;
  double g[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};            // g[k][l] = gradBi(l,k)
  const int cnt = L.cnt[tix];
  const unsigned *col = L.nbr + ((size_t)(tix >> 5) * L.lmax) * 32 + (tix & 31);
  const bool wantg = (A.gradB != nullptr || A.alpha != nullptr) && A.icurltype == 1;
  for (int n = 0; n < cnt; n++) {
    const int k = (int)col[(size_t)n * 32];
    const double4 pj = ld4(G.posh + k), bj = ld4(A.bvec + k);
    const double dx = p.x - pj.x, dy = p.y - pj.y, dz = p.z - pj.z;
    const double rij2 = dist2_exact(dx, dy, dz);
    const double hj1 = pj.w;
    const double q2i = __dmul_rn(rij2, hi21), q2j = __dmul_rn(__dmul_rn(rij2, hj1), hj1);   // :179-180
    // dr = dx/(rij + tiny(rij)), :187: coincident particles (rij = 0) give dr = 0
    const double rinv = rsqrt_nr(rij2);
    const double drx = dx * rinv, dry = dy * rinv, drz = dz * rinv;
    double grkerni, grkernj = 0.;
    {
      // interpolate_kernel_curl (src/kernelND.f90:4643-4672): grwijalt = grwij (ikernelalt = ikernel)
      const int idxi = tab_index(q2i, G.ddq2table);
      const double2 ri = __ldg(G.tabg + idxi);
      grkerni = ri.x + ri.y * (q2i - __dmul_rn((double)idxi, G.dq2table));
      if (A.icurltype == 2) {
        const int idxj = tab_index(q2j, G.ddq2table);
        const double2 rj = __ldg(G.tabg + idxj);
        grkernj = rj.x + rj.y * (q2j - __dmul_rn((double)idxj, G.dq2table));
      }
    }
    const double pmassj = __ldg(&G.vm[k].w);
    if (A.icurltype == 2) {                                       // :202-211
      const double rhoj = __ldg(A.srho + k), gradhj = __ldg(&A.gal[k].x);
      grkerni = grkerni * hfacwabi * hi1;
      grkernj = grkernj * powndim<NDIM>(hj1) * hj1 * gradhj;
      const double tix_ = bi.y * drz - bi.z * dry, tiy_ = bi.z * drx - bi.x * drz, tiz_ = bi.x * dry - bi.y * drx;
      const double tjx_ = bj.y * drz - bj.z * dry, tjy_ = bj.z * drx - bj.x * drz, tjz_ = bj.x * dry - bj.y * drx;
      const double wi = rho21gradhi * grkerni, wj = grkernj / (rhoj * rhoj);
      cx += pmassj * (tix_ * wi + tjx_ * wj); cy += pmassj * (tiy_ * wi + tjy_ * wj); cz += pmassj * (tiz_ * wi + tjz_ * wj);
    } else {
      const double dBx = bi.x - bj.x, dBy = bi.y - bj.y, dBz = bi.z - bj.z;
      const double tx = dBy * drz - dBz * dry, ty = dBz * drx - dBx * drz, tz = dBx * dry - dBy * drx;   // cross_product3D(dB,dr)
      double w;
      if (A.icurltype == 3) w = grkerni * hi1;                    // :214 (weights applied after the loop)
      else if (A.icurltype == 4) { const double rhoj = __ldg(A.srho + k); w = pmassj / (rhoj * rhoj) * (grkerni * hfacwabi * hi1); }   // :223-228
      else w = pmassj * (grkerni * hfacwabi * hi1);               // :233-239
      cx += tx * w; cy += ty * w; cz += tz * w;
      if (wantg) {                                                // :250-255 gradBi(:,k) += pmass(j)*dB(k)*dr(:)*grkerni
        const double dB[3] = {dBx, dBy, dBz}, dr[3] = {drx, dry, drz};
#pragma unroll
        for (int kk = 0; kk < 3; kk++)
#pragma unroll
          for (int l = 0; l < 3; l++) g[kk][l] += w * dB[kk] * dr[l];
      }
    }
  }
  // :275-290
  if (A.icurltype == 4) { cx = rhoi * cx; cy = rhoi * cy; cz = rhoi * cz; }
  else if (A.icurltype == 3) { cx = A.weight * cx; cy = A.weight * cy; cz = A.weight * cz; }
  else if (A.icurltype == 2) { cx = -rhoi * cx; cy = -rhoi * cy; cz = -rhoi * cz; }
  else { cx = cx * gradhi / rhoi; cy = cy * gradhi / rhoi; cz = cz * gradhi / rhoi; }
  A.curlB[(size_t)orig * 3] = cx; A.curlB[(size_t)orig * 3 + 1] = cy; A.curlB[(size_t)orig * 3 + 2] = cz;
  if (wantg) {
    double g2 = 0.;
#pragma unroll
    for (int kk = 0; kk < 3; kk++)
#pragma unroll
      for (int l = 0; l < 3; l++) {
        const double v = -g[kk][l] * gradhi / rhoi;               // :287
        if (A.gradB) A.gradB[((size_t)orig * 3 + kk) * 3 + l] = v;
        g2 += v * v;
      }
    if (A.alpha) {                                                // conservative2primitive.f90:304-311
      const double B2i = (bi.x * bi.x + bi.y * bi.y) + bi.z * bi.z;
      A.alpha[(size_t)orig * 3 + 2] = B2i > 1.e-8 ? fmin(A.hh[orig] * sqrt(g2) / sqrt(B2i + 2.220446049250313e-16), 1.0) : 0.;
    }
  }
}

}  // namespace ndk
