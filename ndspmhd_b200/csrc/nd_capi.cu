// libndspmhd_b200.so -- the context, and the C-ABI of include/ndspmhd_b200.h; one translation unit with nd_kernels.cuh (O(N) kernels) and
// nd_host.cuh (host orchestration), which it includes below.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include "../../include/ndspmhd_b200.h"
#include "nd_tables.h"
#include <chrono>
#include <functional>
#include "nd_device.cuh"
#include "nd_density.cuh"
#include "nd_rates.cuh"
#include "nd_curl.cuh"
#include "nd_nccl.cuh"

#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cfloat>
#include <string>
#include <vector>
#include <algorithm>

using namespace ndk;

// =====================================================================================================
// context
// =====================================================================================================
struct RowBuf { void **p; size_t rowbytes; };   // a per-particle device array that grows with capacity

struct nd_ctx {
  nd_options o;
  int ndim = 3, device = 0;
  cudaStream_t stream = nullptr, stream_h2d = nullptr, stream_d2h = nullptr;   // compute; copy-in / copy-out of derivs_host
  cudaEvent_t ev_in[3] = {nullptr, nullptr, nullptr}, ev_out[3] = {nullptr, nullptr, nullptr};
  bool defer_vel = false;   // derivs_host, LIGHT rounds, periodic ghosts only: vel arrives during the density iteration (k_late_vel)
  bool wait_in2 = false;    // derivs_host on a slab context: en, Bevol, alpha, psi (ev_in[1]) are first read by the halo exchange after the density iteration
  bool wait_in1b = false;   // derivs_host: vel, pmass, rho are still on their way (ev_in[2]); the link waits for them where it first reads them
  std::string err;
  long long launches = 0;
  ndt::KernelTables *T = nullptr;
  TabRec *d_tab = nullptr; TabRec2 *d_tab2 = nullptr; double *d_tabdrag = nullptr; double2 *d_tabg = nullptr, *d_tabw = nullptr;
  int num_sms = 148;
  // sizes
  int npart = 0, ntotal = 0, cap = 0, nown = 0;   // nown <= npart: rows this context computes (the rest of [0,npart) are halo copies)
  bool mixed_types = true;   // false when the last link found one itype only (the list builder then skips the type rules)
  bool uploaded = false, linked = false, density_done = false, prim_done = false, rates_done = false;
  // ---- original-order arrays (row r = Fortran index r+1) ----
  double *x = nullptr, *vel = nullptr, *pmass = nullptr, *hh = nullptr, *en = nullptr, *Bevol = nullptr, *alpha = nullptr, *psi = nullptr;
  int *itype = nullptr, *ireal = nullptr;
  double *hhin = nullptr, *hh0 = nullptr, *alphaB_in = nullptr;
  double *rho = nullptr, *gradh = nullptr, *drhodt = nullptr, *dhdt = nullptr, *rhoalt = nullptr, *gradhn = nullptr, *gradsoft = nullptr, *gradgradh = nullptr;
  int *numneigh = nullptr;
  double *dens = nullptr, *uu = nullptr, *pr = nullptr, *spsound = nullptr, *Bfield = nullptr;
  double *force = nullptr, *dudt = nullptr, *dendt = nullptr, *dBevoldt = nullptr, *daldt = nullptr, *dpsidt = nullptr, *gradpsi = nullptr, *divB = nullptr,
         *curlB = nullptr, *graddivv = nullptr, *del2u = nullptr;
  // ---- sorted-order arrays ----
  double4 *posh = nullptr, *vm = nullptr, *posm = nullptr, *bpsi = nullptr, *thermo = nullptr, *gal = nullptr;
  bool dens_light = false;   // set by the fused entry points: the rates kernel of the same derivs makes drho/dt (fast tuple)
  bool slab_light = true;    // LIGHT rounds in slab-decomposed contexts too (NDSPMHD_B200_SLAB_LIGHT=0 turns them off: the A/B of tools/gpu_slab_ab.sh)
  bool drho_pairs = false;   // the density rounds of this derivs ran LIGHT: k_rates_final takes drho/dt from the pair sums and makes dh/dt
  float4 *p32 = nullptr;   // FP32 screening records of the list builder (nd_device.cuh)
  double *srho = nullptr;
  // ---- one-fluid dust (idust=1; allocated only then) ----
  double *dustevol = nullptr, *dustfrac = nullptr, *deltav = nullptr, *rhogas = nullptr, *rhodust = nullptr, *ddustevoldt = nullptr, *ddeltavdt = nullptr;
  double *sdf = nullptr;            // sorted: entry dust fraction of the row's parent (density sums)
  double4 *dusta = nullptr, *sD = nullptr; double2 *dustb = nullptr;   // sorted rates records / sums (nd_rates.cuh)
  double4 *sF = nullptr, *sdB = nullptr, *sC = nullptr, *sP = nullptr, *sV = nullptr;
  int *typ = nullptr, *perm = nullptr, *permtmp = nullptr, *inv = nullptr, *cellOf = nullptr, *cellOfOrig = nullptr, *redo = nullptr, *list = nullptr,
      *scanout = nullptr, *ghostcount = nullptr;
  std::vector<RowBuf> rowbufs;
  // ---- cell grid ----
  int *cellStart = nullptr, *cellCount = nullptr; int cellcap = 0;
  int *blocksums = nullptr; int blocksumcap = 0;
  // ---- neighbour lists (nd_device.cuh): one chunk of targets at a time ----
  unsigned *nbr = nullptr; int *lcnt = nullptr; size_t nbrcap = 0; int lcntcap = 0, lmax = 0;
  int ncellsx[3] = {1, 1, 1}, ncells = 0;
  double xminpart[3] = {0, 0, 0}, dxcell = 0, hhmax = 0;
  // ---- small device scratch: reduction keys, flags ----
  unsigned long long *red = nullptr;   // [16]
  double *fmean = nullptr;             // [4]
  int *flags = nullptr;                // [16]
  unsigned long long *h_red = nullptr; int *h_flags = nullptr; double *h_fmean = nullptr;   // pinned mirrors
  // ---- iterate_density state (kept across ND_NEED_RELINK) ----
  int itsdensity = 0, ncalc = 0, nrelink = 0; long long ncalctotal = 0, ncalc_g = 0; bool redolink = false;
  nd_scalars sc;
  // ---- slab decomposition (nd_comm): halo send lists and staging buffers ----
  nd_comm comm; bool has_comm = false; bool slab_too_narrow = false;
  // native NCCL transport (nd_nccl.cuh): communicator, a 64-double device block and its pinned mirror for the small collectives
  nd_ncclComm_t nccl = nullptr; NcclApi *nccl_api = nullptr; double *d_comm = nullptr, *h_comm = nullptr;
  long long n_allreduce = 0, halo_bytes_sent = 0, migrated_bytes_sent = 0, nmigrated_out = 0, nmigrated_in = 0;
  std::vector<double> edges;   // every rank's slab faces (migration)
  long long *gid = nullptr;    // row ids (ndspmhd_b200_set_row_ids; rows keep theirs when they migrate between slabs)
  int *sendlist[2] = {nullptr, nullptr}; int sendcap[2] = {0, 0}, nsend[2] = {0, 0}, nrecv[2] = {0, 0};
  void *sendbuf[2] = {nullptr, nullptr}, *recvbuf[2] = {nullptr, nullptr}; size_t sendbufcap[2] = {0, 0}, recvbufcap[2] = {0, 0};
  cudaEvent_t ev[8];
  // rates in row chunks (ndspmhd_b200_derivs_host): chunk q = original rows [q*rows, (q+1)*rows) so that its results can be
  // downloaded while the next chunk's pair kernel runs
  int rate_chunks = 1, rate_chunks_used = 1, list_overflows = 0; int *rlist = nullptr; size_t rlistcap = 0;
  std::function<int(int, int, int)> on_rates_chunk;   // (chunk, row0, row1) after the chunk's finalisation is enqueued
  std::vector<cudaEvent_t> chunk_events;
  double *evpartial = nullptr, *h_ev = nullptr;        // evwrite reductions
  double *finalpart = nullptr; size_t finalpartcap = 0;   // block partials of k_rates_final's scalar reductions
  double *stepbuf = nullptr; size_t stepbufrows = 0;   // leapfrog `*in` copies (ndspmhd_b200_step), rows [0,npart)
  cudaEvent_t ev_pair[2] = {nullptr, nullptr};   // around the rates pair kernel alone (the roofline's kernel time)
  double ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

namespace {

int set_err(nd_ctx *c, int code, const std::string &m) { if (c) c->err = m; return code; }

#define CU(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) return set_err(c, ND_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

inline int nblocks(long long n, int b) { return (int)((n + b - 1) / b); }

#define LAUNCH(c, kern, grid, block, smem, ...)                      \
  do {                                                               \
    if ((grid) > 0) {                                                \
      auto kfn_ = kern;                                              \
      kfn_<<<(grid), (block), (smem), (c)->stream>>>(__VA_ARGS__);   \
      (c)->launches++;                                               \
    }                                                                \
  } while (0)

// Scalars travel device -> host as stores of a one-warp kernel into page-locked host memory (directly addressable under
// unified addressing), not as cudaMemcpy: a memcpy queues on the D2H copy engine behind whatever bulk download
// ndspmhd_b200_derivs_host has in flight (measured: a 27 ms stall per step at 16.8M particles), a store does not.
__global__ void k_to_host(const int *src, int *dst_pinned, int nwords) {
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst_pinned[i] = src[i];
  __threadfence_system();
}
#define SMALL_D2H(c, dst_pinned, src, bytes) LAUNCH(c, k_to_host, 1, 32, 0, reinterpret_cast<const int *>(src), reinterpret_cast<int *>(dst_pinned), (int)((bytes) / sizeof(int)))

#include "nd_kernels.cuh"   // O(N) kernels: scan/reductions, ghosts, cell sort, cons2prim, rates finalisation, leapfrog, evwrite sums

#include "nd_host.cuh"      // host orchestration: buffers, halos, ghosts, link, lists, iterate_density, cons2prim, get_rates

}  // namespace

// =====================================================================================================
// C-ABI
// =====================================================================================================
extern "C" {

int ndspmhd_b200_version(void) { return 100; }

int ndspmhd_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int ndspmhd_b200_default_options(nd_options *o) {
  if (!o) return ND_ERR_INVALID_ARG;
  memset(o, 0, sizeof(*o));
  o->psep = 0.01; o->gamma = 5. / 3.; o->iener = 2; o->polyk = 1.0; o->icty = 0; o->maxdensits = 250; o->iprterm = 0; o->iav = 2;   // defaults.f90:47-118
  o->alphamin = 0.1; o->alphaumin = 0.0; o->alphaBmin = 1.0; o->beta = 2.0; o->iavlim[0] = 2; o->iavlim[1] = 1; o->iavlim[2] = 0;
  o->avdecayconst = 0.1; o->ikernav = 3; o->ihvar = 2; o->hfact = 1.2; o->tolh = 1.e-3; o->imhd = 0; o->imagforce = 2; o->idivbzero = 0;
  o->psidecayfact = 0.1; o->use_smoothed_rhodust = 1; o->islope_limiter = -1;
  o->avfact = std::log(4.) / (std::log((o->gamma + 1.) / (o->gamma - 1.)));            // initialiseND_mhd.f90:168-172
  o->want_aux = 1;
  return 0;
}

const char *ndspmhd_b200_last_error(const nd_ctx *c) { return c ? c->err.c_str() : "null context"; }

int ndspmhd_b200_create(const nd_options *o, int ndim, int device, nd_ctx **out) {
  if (!o || !out) return ND_ERR_INVALID_ARG;
  *out = nullptr;
  nd_ctx *c = new nd_ctx();
  *out = c;   // returned even on failure so the caller can read the message; destroy() is always safe
  c->o = *o; c->ndim = ndim; c->device = device;
  memset(&c->sc, 0, sizeof(c->sc));
  for (int k = 0; k < 8; k++) c->ev[k] = nullptr;
  if (int e = check_options(c, *o, ndim)) return e;
  c->T = new ndt::KernelTables();
  if (!ndt::build_tables(*c->T, o->ikernel, o->ikernelalt, o->idust, ndim)) return set_err(c, ND_ERR_UNSUPPORTED_OPTION, "unsupported kernel (ikernel must be 0, 2 or 3; drag needs 0 or 2)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_err(c, ND_ERR_NO_DEVICE, "no CUDA device: the NDSPMHD hot path has no CPU fallback in this library");
  if (device < 0 || device >= ndev) return set_err(c, ND_ERR_INVALID_ARG, "bad device ordinal");
  CU(cudaSetDevice(device));
  CU(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
  if (const char *ev = getenv("NDSPMHD_B200_SLAB_LIGHT")) c->slab_light = atoi(ev) != 0;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (int k = 0; k < 8; k++) CU(cudaEventCreate(&c->ev[k]));
  for (int k = 0; k < 2; k++) CU(cudaEventCreate(&c->ev_pair[k]));
  CU(cudaStreamCreateWithFlags(&c->stream_h2d, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->stream_d2h, cudaStreamNonBlocking));
  for (int k = 0; k < 3; k++) CU(cudaEventCreateWithFlags(&c->ev_in[k], cudaEventDisableTiming));
  for (int k = 0; k < 3; k++) CU(cudaEventCreateWithFlags(&c->ev_out[k], cudaEventDisableTiming));
  // kernel tables -> interpolation records
  std::vector<TabRec> tab(IKERN + 1);
  std::vector<TabRec2> tab2(IKERN + 1);
  std::vector<double> tabd(2 * (IKERN + 1));
  for (int i = 0; i <= IKERN; i++) {
    const int i1 = std::min(i + 1, IKERN);
    tab[i].w = c->T->w[i]; tab[i].dw = (c->T->w[i1] - c->T->w[i]) * c->T->ddq2table;
    tab[i].g = c->T->grw[i]; tab[i].dg = (c->T->grw[i1] - c->T->grw[i]) * c->T->ddq2table;
    tab2[i].gg = c->T->grgrw[i]; tab2[i].dgg = (c->T->grgrw[i1] - c->T->grgrw[i]) * c->T->ddq2table;
    tabd[2 * i] = c->T->wdrag[i]; tabd[2 * i + 1] = (c->T->wdrag[i1] - c->T->wdrag[i]) * c->T->ddq2table;
  }
  CU(cudaMalloc(&c->d_tab, sizeof(TabRec) * (IKERN + 1)));
  CU(cudaMalloc(&c->d_tab2, sizeof(TabRec2) * (IKERN + 1)));
  CU(cudaMalloc(&c->d_tabdrag, sizeof(double) * 2 * (IKERN + 1)));
  {
    std::vector<double2> tabg(IKERN + 1);
    for (int i = 0; i <= IKERN; i++) tabg[i] = make_double2(tab[i].g, tab[i].dg);
    CU(cudaMalloc(&c->d_tabg, sizeof(double2) * (IKERN + 1)));
    CU(cudaMemcpy(c->d_tabg, tabg.data(), sizeof(double2) * (IKERN + 1), cudaMemcpyHostToDevice));
    for (int i = 0; i <= IKERN; i++) tabg[i] = make_double2(tab[i].w, tab[i].dw);
    CU(cudaMalloc(&c->d_tabw, sizeof(double2) * (IKERN + 1)));
    CU(cudaMemcpy(c->d_tabw, tabg.data(), sizeof(double2) * (IKERN + 1), cudaMemcpyHostToDevice));
  }
  CU(cudaMemcpy(c->d_tab, tab.data(), sizeof(TabRec) * (IKERN + 1), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(c->d_tab2, tab2.data(), sizeof(TabRec2) * (IKERN + 1), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(c->d_tabdrag, tabd.data(), sizeof(double) * 2 * (IKERN + 1), cudaMemcpyHostToDevice));
  CU(cudaMalloc(&c->red, sizeof(unsigned long long) * 16));
  CU(cudaMalloc(&c->fmean, sizeof(double) * 4));
  CU(cudaMalloc(&c->flags, sizeof(int) * 16));
  CU(cudaMemset(c->flags, 0, sizeof(int) * 16));
  CU(cudaMallocHost(&c->h_red, sizeof(unsigned long long) * 16));
  CU(cudaMallocHost(&c->h_flags, sizeof(int) * 32));
  CU(cudaMallocHost(&c->h_fmean, sizeof(double) * 4));
  register_rows(c);
  return 0;
}

int ndspmhd_b200_set_options(nd_ctx *c, const nd_options *o) {
  if (!c || !o) return ND_ERR_INVALID_ARG;
  if (int e = check_options(c, *o, c->ndim)) return e;
  if (o->ikernel != c->o.ikernel || (o->idust != 0) != (c->o.idust != 0)) return set_err(c, ND_ERR_INVALID_ARG, "kernel choice is fixed at create()");
  c->o = *o;
  return 0;
}

int ndspmhd_b200_destroy(nd_ctx *c) {
  if (!c) return 0;
  if (c->stream) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
  for (auto &rb : c->rowbufs) if (*rb.p) { cudaFree(*rb.p); *rb.p = nullptr; }
  void *singles[] = {c->sendlist[0], c->sendlist[1], c->sendbuf[0], c->sendbuf[1], c->recvbuf[0], c->recvbuf[1], c->nbr, c->lcnt, c->scanout, c->cellStart, c->cellCount, c->blocksums, c->red, c->fmean, c->flags, c->d_tab, c->d_tab2, c->d_tabdrag, c->d_tabg, c->d_tabw};
  for (void *p : singles) if (p) cudaFree(p);
  if (c->h_red) cudaFreeHost(c->h_red);
  if (c->h_flags) cudaFreeHost(c->h_flags);
  if (c->h_fmean) cudaFreeHost(c->h_fmean);
  if (c->nccl && c->nccl_api) c->nccl_api->CommDestroy(c->nccl);
  if (c->d_comm) cudaFree(c->d_comm);
  if (c->h_comm) cudaFreeHost(c->h_comm);
  if (c->stepbuf) cudaFree(c->stepbuf);
  if (c->finalpart) cudaFree(c->finalpart);
  if (c->evpartial) cudaFree(c->evpartial);
  if (c->h_ev) cudaFreeHost(c->h_ev);
  if (c->rlist) cudaFree(c->rlist);
  for (auto x : c->chunk_events) cudaEventDestroy(x);
  for (int k = 0; k < 8; k++) if (c->ev[k]) cudaEventDestroy(c->ev[k]);
  for (int k = 0; k < 2; k++) if (c->ev_pair[k]) cudaEventDestroy(c->ev_pair[k]);
  for (int k = 0; k < 3; k++) if (c->ev_in[k]) cudaEventDestroy(c->ev_in[k]);
  for (int k = 0; k < 3; k++) if (c->ev_out[k]) cudaEventDestroy(c->ev_out[k]);
  if (c->stream_h2d) cudaStreamDestroy(c->stream_h2d);
  if (c->stream_d2h) cudaStreamDestroy(c->stream_d2h);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c->T;
  delete c;
  return 0;
}

int ndspmhd_b200_get_kernel_tables(const nd_ctx *c, double *wij, double *grwij, double *grgrwij, double *wijdrag, double *radkern2, double *dq2table) {
  if (!c || !c->T) return ND_ERR_INVALID_ARG;
  for (int i = 0; i <= IKERN; i++) {
    if (wij) wij[i] = c->T->w[i];
    if (grwij) grwij[i] = c->T->grw[i];
    if (grgrwij) grgrwij[i] = c->T->grgrw[i];
    if (wijdrag) wijdrag[i] = c->T->wdrag[i];
  }
  if (radkern2) *radkern2 = c->T->radkern2;
  if (dq2table) *dq2table = c->T->dq2table;
  return 0;
}

namespace {
int check_upload_args(nd_ctx *c, const nd_arrays *a, int npart, int &ntotal, int idim) {
  const nd_options &o = c->o;
  if (o.device_ghosts) ntotal = npart;
  if (c->has_comm && !o.device_ghosts) return set_err(c, ND_ERR_UNSUPPORTED_OPTION, "slab decomposition needs device_ghosts = 1");
  if (npart < 1 || ntotal < npart || idim < ntotal) return set_err(c, ND_ERR_INVALID_ARG, "upload: need 1 <= npart <= ntotal <= idim");
  if (!a->x || !a->vel || !a->pmass || !a->hh_in || !a->itype) return set_err(c, ND_ERR_INVALID_ARG, "upload: x, vel, pmass, hh_in, itype are required");
  if (o.imhd != 0 && !a->Bevol) return set_err(c, ND_ERR_INVALID_ARG, "upload: Bevol required with imhd /= 0");
  if (!a->en || !a->alpha) return set_err(c, ND_ERR_INVALID_ARG, "upload: en and alpha are required");
  if (o.onef_dust && (!a->dustevol || !a->dustfrac_in || !a->deltav)) return set_err(c, ND_ERR_INVALID_ARG, "upload: dustevol, dustfrac_in, deltav required with idust = 1");
  if (o.onef_dust && c->has_comm) return set_err(c, ND_ERR_UNSUPPORTED_OPTION, "one-fluid dust is not available with the slab decomposition");
  if ((ntotal > npart || any_fixed_bound(c)) && !a->ireal) return set_err(c, ND_ERR_INVALID_ARG, "upload: ireal required with ghosts/fixed particles");
  return 0;
}
// group 1: what link + density read; group 2: what cons2prim + rates read in addition
int upload_group(nd_ctx *c, const nd_arrays *a, size_t n, int group, cudaStream_t st) {
  auto up = [&](void *dst, const void *src, size_t bytes) -> cudaError_t { return src ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) : cudaMemsetAsync(dst, 0, bytes, st); };
  if (group == 1 || group == 10) {   // 10 = first half of group 1: what the link and the LIGHT density rounds read (56 of the 80 bytes a row)
    CU(up(c->x, a->x, sizeof(double) * c->ndim * n));
    CU(up(c->hh, a->hh_in, sizeof(double) * n));
    CU(cudaMemcpyAsync(c->hh0, c->hh, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    CU(up(c->itype, a->itype, sizeof(int) * n));
    CU(up(c->ireal, a->ireal, sizeof(int) * n));
    CU(up(c->pmass, a->pmass, sizeof(double) * n));
    CU(up(c->rho, a->rho_in, sizeof(double) * n));   // fixed particles without a parent keep their density (must land before the rounds write rho)
  }
  if (group == 1 || group == 11) {   // 11 = second half: first read when the ghost rows are written / the sorted records are gathered
    CU(up(c->vel, a->vel, sizeof(double) * 3 * n));
    if (c->o.onef_dust) CU(up(c->dustfrac, a->dustfrac_in, sizeof(double) * n));   // read by the density sums (density_sums.f90:280-282)
  }
  if (group == 2) {
    CU(up(c->en, a->en, sizeof(double) * n));
    CU(up(c->Bevol, a->Bevol, sizeof(double) * 3 * n));
    CU(up(c->alpha, a->alpha, sizeof(double) * 3 * n));
    CU(up(c->psi, a->psi, sizeof(double) * n));
    if (c->o.onef_dust) { CU(up(c->dustevol, a->dustevol, sizeof(double) * n)); CU(up(c->deltav, a->deltav, sizeof(double) * 3 * n)); }
  }
  return 0;
}
// phase 1: density outputs that get_rates does not touch; 2: primitives; 3: rates (+ drhodt, dhdt, zeroed on ghosts/fixed by get_rates)
// rows [r0,r1) of the arrays get_rates writes (download group 3); dpsidt separately when the rates ran in row chunks
int download_rates_rows(nd_ctx *c, nd_arrays *a, size_t r0, size_t r1, unsigned mask, cudaStream_t st, bool with_dpsidt) {
  if (r1 <= r0) return 0;
  const size_t D = sizeof(double), n = r1 - r0;
  auto dn = [&](double *dst, const double *src, size_t w) -> cudaError_t {
    return (dst && src) ? cudaMemcpyAsync(dst + r0 * w, src + r0 * w, D * w * n, cudaMemcpyDeviceToHost, st) : cudaSuccess;
  };
  if (mask & (ND_DL_DENSITY | ND_DL_RATES)) { CU(dn(a->drhodt, c->drhodt, 1)); CU(dn(a->dhdt, c->dhdt, 1)); }
  if (mask & ND_DL_RATES) {
    CU(dn(a->force, c->force, 3)); CU(dn(a->dudt, c->dudt, 1)); CU(dn(a->dendt, c->dendt, 1));
    if (c->o.imhd != 0) {
      CU(dn(a->dBevoldt, c->dBevoldt, 3)); if (with_dpsidt) CU(dn(a->dpsidt, c->dpsidt, 1)); CU(dn(a->gradpsi, c->gradpsi, 3)); CU(dn(a->divB, c->divB, 1));
      CU(dn(a->curlB, c->curlB, 3));
    }
    CU(dn(a->daldt, c->daldt, 3));
    if (c->o.want_aux || c->o.iavlim[0] == 3) CU(dn(a->graddivv, c->graddivv, 3));
    if (c->o.want_aux) CU(dn(a->del2u, c->del2u, 1));
    if (c->o.onef_dust) { CU(dn(a->ddustevoldt, c->ddustevoldt, 1)); CU(dn(a->ddeltavdt, c->ddeltavdt, 3)); }
  }
  return 0;
}
int download_group(nd_ctx *c, nd_arrays *a, size_t n, int group, unsigned mask, cudaStream_t st) {
  auto dn = [&](void *dst, const void *src, size_t bytes) -> cudaError_t { return (dst && src) ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st) : cudaSuccess; };
  const size_t D = sizeof(double);
  if (group == 1 && (mask & ND_DL_DENSITY)) {
    CU(dn(a->hh, c->hh, D * n)); CU(dn(a->rho, c->rho, D * n)); CU(dn(a->gradh, c->gradh, D * n)); CU(dn(a->numneigh, c->numneigh, sizeof(int) * n));
    if (c->o.want_aux) { CU(dn(a->rhoalt, c->rhoalt, D * n)); CU(dn(a->gradhn, c->gradhn, D * n)); CU(dn(a->gradsoft, c->gradsoft, D * n)); CU(dn(a->gradgradh, c->gradgradh, D * n)); }
  }
  if (group == 2 && (mask & ND_DL_DENSITY) && c->o.onef_dust) { CU(dn(a->rhogas, c->rhogas, D * n)); CU(dn(a->rhodust, c->rhodust, D * n)); }   // ghost rows are final after c2p
  if (group == 2 && (mask & ND_DL_PRIM)) {
    CU(dn(a->dens, c->dens, D * n)); CU(dn(a->uu, c->uu, D * n)); CU(dn(a->pr, c->pr, D * n)); CU(dn(a->spsound, c->spsound, D * n));
    if (c->o.imhd != 0) CU(dn(a->Bfield, c->Bfield, D * 3 * n));
    if (c->o.imhd != 0 && c->o.iavlim[2] == 2) CU(dn(a->alpha_out, c->alpha, D * 3 * n));   // alpha(3,:) rewritten by the resistivity switch
    if (c->o.onef_dust) CU(dn(a->dustfrac, c->dustfrac, D * n));
  }
  if (group == 3) {
    if (mask & (ND_DL_DENSITY | ND_DL_RATES)) { CU(dn(a->drhodt, c->drhodt, D * n)); CU(dn(a->dhdt, c->dhdt, D * n)); }
    if (mask & ND_DL_RATES) {
      CU(dn(a->force, c->force, D * 3 * n)); CU(dn(a->dudt, c->dudt, D * n)); CU(dn(a->dendt, c->dendt, D * n));
      if (c->o.imhd != 0) {
        CU(dn(a->dBevoldt, c->dBevoldt, D * 3 * n)); CU(dn(a->dpsidt, c->dpsidt, D * n)); CU(dn(a->gradpsi, c->gradpsi, D * 3 * n)); CU(dn(a->divB, c->divB, D * n));
        CU(dn(a->curlB, c->curlB, D * 3 * n));
      }
      CU(dn(a->daldt, c->daldt, D * 3 * n));
      // graddivv holds dead "curl v" sums unless iavlim(1)=3, del2u is a local of the reference: shipped only on request
      if (c->o.want_aux || c->o.iavlim[0] == 3) CU(dn(a->graddivv, c->graddivv, D * 3 * n));
      if (c->o.want_aux) CU(dn(a->del2u, c->del2u, D * n));
      if (c->o.onef_dust) { CU(dn(a->ddustevoldt, c->ddustevoldt, D * n)); CU(dn(a->ddeltavdt, c->ddeltavdt, D * 3 * n)); }
    }
  }
  return 0;
}
}  // namespace

int ndspmhd_b200_upload(nd_ctx *c, const nd_arrays *a, int npart, int ntotal, int idim) {
  if (!c || !a || !c->stream) return c ? set_err(c, ND_ERR_STATE, "context not initialised") : ND_ERR_INVALID_ARG;
  if (int e = check_upload_args(c, a, npart, ntotal, idim)) return e;
  CU(cudaSetDevice(c->device));
  int want = ntotal;
  if (c->o.device_ghosts && any_ghost_bound(c)) want = npart + npart / 4 + 1024;   // first guess; make_ghosts grows it if needed
  if (int e = ensure_capacity(c, want, 0)) return e;
  if (int e = upload_group(c, a, (size_t)ntotal, 1, c->stream)) return e;
  if (int e = upload_group(c, a, (size_t)ntotal, 2, c->stream)) return e;
  LAUNCH(c, k_iota_ll, nblocks(npart, 256), 256, 0, c->gid, npart, 0LL);   // default row ids: the row number
  CU(cudaStreamSynchronize(c->stream));
  c->npart = npart; c->ntotal = ntotal; c->nown = npart;
  c->uploaded = true; c->linked = c->density_done = c->prim_done = c->rates_done = false;
  return 0;
}

int ndspmhd_b200_set_row_ids(nd_ctx *c, const long long *ids, int n) {
  if (!c || !ids) return ND_ERR_INVALID_ARG;
  if (!c->uploaded || n != c->nown) return set_err(c, ND_ERR_STATE, "set_row_ids: call after upload with n = the rows uploaded");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(c->gid, ids, sizeof(long long) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int ndspmhd_b200_get_row_ids(nd_ctx *c, long long *ids, int cap) {
  if (!c || !ids) return ND_ERR_INVALID_ARG;
  if (!c->uploaded || cap < c->nown) return set_err(c, ND_ERR_INVALID_ARG, "get_row_ids: cap < own rows");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(ids, c->gid, sizeof(long long) * (size_t)c->nown, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int ndspmhd_b200_migration_stats(const nd_ctx *c, long long *rows_out, long long *rows_in, long long *bytes_sent) {
  if (!c) return ND_ERR_INVALID_ARG;
  if (rows_out) *rows_out = c->nmigrated_out;
  if (rows_in) *rows_in = c->nmigrated_in;
  if (bytes_sent) *bytes_sent = c->migrated_bytes_sent;
  return 0;
}

int ndspmhd_b200_set_comm(nd_ctx *c, const nd_comm *comm) {
  if (!c) return ND_ERR_INVALID_ARG;
  if (c->nccl && c->nccl_api) { c->nccl_api->CommDestroy(c->nccl); c->nccl = nullptr; }   // the callback transport replaces a native one
  c->edges.clear();   // the slab faces of every rank are gathered again at the next migration (a re-attach is how a caller rebalances)
  if (!comm || comm->nranks <= 1) { c->has_comm = false; return 0; }
  if (!comm->allreduce || !comm->sendrecv_counts || !comm->sendrecv) return set_err(c, ND_ERR_INVALID_ARG, "set_comm: all three callbacks are required");
  if (comm->rank < 0 || comm->rank >= comm->nranks || !(comm->slab_hi > comm->slab_lo) || comm->nglobal < 1) return set_err(c, ND_ERR_INVALID_ARG, "set_comm: bad rank / slab / nglobal");
  if (!c->o.device_ghosts) return set_err(c, ND_ERR_UNSUPPORTED_OPTION, "slab decomposition needs device_ghosts = 1");
  if (!(c->o.ibound[0] == 0 || c->o.ibound[0] == 1 || c->o.ibound[0] == 3)) return set_err(c, ND_ERR_UNSUPPORTED_OPTION, "slab decomposition needs ibound(1) in {0,1,3}");
  c->comm = *comm; c->has_comm = true;
  c->uploaded = c->linked = c->density_done = c->prim_done = c->rates_done = false;
  return 0;
}

int ndspmhd_b200_nccl_unique_id(unsigned char id[128]) {
  if (!id) return ND_ERR_INVALID_ARG;
  NcclApi *api = nccl_api();
  if (!api) return ND_ERR_COMM;
  nd_ncclUniqueId u;
  if (api->GetUniqueId(&u) != 0) return ND_ERR_COMM;
  memcpy(id, u.internal, 128);
  return 0;
}

int ndspmhd_b200_set_comm_nccl(nd_ctx *c, const unsigned char id[128], int rank, int nranks, double slab_lo, double slab_hi, long long nglobal) {
  if (!c || !id) return ND_ERR_INVALID_ARG;
  c->edges.clear();   // see ndspmhd_b200_set_comm
  if (nranks <= 1) { c->has_comm = false; return 0; }
  if (rank < 0 || rank >= nranks || !(slab_hi > slab_lo) || nglobal < 1) return set_err(c, ND_ERR_INVALID_ARG, "set_comm_nccl: bad rank / slab / nglobal");
  if (nranks > 31) return set_err(c, ND_ERR_INVALID_ARG, "set_comm_nccl: at most 31 ranks");
  if (!c->o.device_ghosts) return set_err(c, ND_ERR_UNSUPPORTED_OPTION, "slab decomposition needs device_ghosts = 1");
  if (!(c->o.ibound[0] == 0 || c->o.ibound[0] == 1 || c->o.ibound[0] == 3)) return set_err(c, ND_ERR_UNSUPPORTED_OPTION, "slab decomposition needs ibound(1) in {0,1,3}");
  NcclApi *api = nccl_api();
  if (!api) return set_err(c, ND_ERR_COMM, "libnccl could not be opened (set NDSPMHD_B200_NCCL_LIB, or use the callback transport ndspmhd_b200_set_comm)");
  CU(cudaSetDevice(c->device));
  if (c->nccl) { api->CommDestroy(c->nccl); c->nccl = nullptr; }
  nd_ncclUniqueId u;
  memcpy(u.internal, id, 128);
  c->nccl_api = api;
  NCCLCHK(api->CommInitRank(&c->nccl, nranks, u, rank));
  if (!c->d_comm) { CU(cudaMalloc(&c->d_comm, sizeof(double) * 64)); CU(cudaMallocHost(&c->h_comm, sizeof(double) * 96)); }
  memset(&c->comm, 0, sizeof(c->comm));
  c->comm.rank = rank; c->comm.nranks = nranks; c->comm.slab_lo = slab_lo; c->comm.slab_hi = slab_hi; c->comm.nglobal = nglobal;
  c->has_comm = true;
  c->uploaded = c->linked = c->density_done = c->prim_done = c->rates_done = false;
  return 0;
}

int ndspmhd_b200_comm_stats(const nd_ctx *c, long long *n_allreduce, long long *halo_bytes_sent) {
  if (!c) return ND_ERR_INVALID_ARG;
  if (n_allreduce) *n_allreduce = c->n_allreduce;
  if (halo_bytes_sent) *halo_bytes_sent = c->halo_bytes_sent;
  return 0;
}

int ndspmhd_b200_row_counts(const nd_ctx *c, int *nown, int *nsrc, int *ntotal) {
  if (!c) return ND_ERR_INVALID_ARG;
  if (nown) *nown = c->nown;
  if (nsrc) *nsrc = c->npart;
  if (ntotal) *ntotal = c->ntotal;
  return 0;
}

int ndspmhd_b200_link(nd_ctx *c) {
  if (!c) return ND_ERR_INVALID_ARG;
  if (!c->uploaded) return set_err(c, ND_ERR_STATE, "link before upload");
  CU(cudaSetDevice(c->device));
  int e = DISPATCH_NDIM(c, do_link<1>(c), do_link<2>(c), do_link<3>(c));
  if (!e) fill_link_scalars(c);
  return e;
}

int ndspmhd_b200_iterate_density(nd_ctx *c, int resume, nd_scalars *s) {
  if (!c) return ND_ERR_INVALID_ARG;
  if (!c->linked) return set_err(c, ND_ERR_STATE, "iterate_density before link");
  CU(cudaSetDevice(c->device));
  int e = DISPATCH_NDIM(c, do_iterate_density<1>(c, resume), do_iterate_density<2>(c, resume), do_iterate_density<3>(c, resume));
  if (e) return e;
  if (int e2 = fill_density_scalars(c)) return e2;
  if (s) *s = c->sc;
  return 0;
}

int ndspmhd_b200_cons2prim(nd_ctx *c) {
  if (!c) return ND_ERR_INVALID_ARG;
  if (!c->density_done) return set_err(c, ND_ERR_STATE, "cons2prim before iterate_density");
  CU(cudaSetDevice(c->device));
  return do_cons2prim(c);
}

int ndspmhd_b200_get_rates(nd_ctx *c, nd_scalars *s) {
  if (!c) return ND_ERR_INVALID_ARG;
  if (!c->prim_done) return set_err(c, ND_ERR_STATE, "get_rates before cons2prim");
  CU(cudaSetDevice(c->device));
  int e = DISPATCH_NDIM(c, do_get_rates<1>(c, nullptr, nullptr, nullptr, 0), do_get_rates<2>(c, nullptr, nullptr, nullptr, 0), do_get_rates<3>(c, nullptr, nullptr, nullptr, 0));
  if (e) return e;
  if (s) *s = c->sc;
  return 0;
}

int ndspmhd_b200_derivs(nd_ctx *c, nd_scalars *s) {
  if (!c) return ND_ERR_INVALID_ARG;
  if (!c->uploaded) return set_err(c, ND_ERR_STATE, "derivs before upload");
  CU(cudaSetDevice(c->device));
  CU(cudaEventRecord(c->ev[0], c->stream));
  int e = DISPATCH_NDIM(c, do_link<1>(c), do_link<2>(c), do_link<3>(c));
  if (e) return e;
  CU(cudaEventRecord(c->ev[1], c->stream));
  c->dens_light = ND_DENS_LIGHT && fast_tuple(c->o) && (!c->has_comm || c->slab_light);   // get_rates follows in this call and makes drho/dt itself
  const bool light = c->dens_light;
  e = DISPATCH_NDIM(c, do_iterate_density<1>(c, 0), do_iterate_density<2>(c, 0), do_iterate_density<3>(c, 0));
  c->dens_light = false;
  if (e) return e;
  CU(cudaEventRecord(c->ev[2], c->stream));
  e = do_cons2prim(c);
  if (e) return e;
  c->drho_pairs = light;
  e = DISPATCH_NDIM(c, do_get_rates<1>(c, nullptr, nullptr, nullptr, 0), do_get_rates<2>(c, nullptr, nullptr, nullptr, 0), do_get_rates<3>(c, nullptr, nullptr, nullptr, 0));
  c->drho_pairs = false;
  if (e) return e;
  if (int e2 = fill_density_scalars(c)) return e2;
  CU(cudaEventSynchronize(c->ev[5]));
  float t;
  for (int k = 0; k < 5; k++) { cudaEventElapsedTime(&t, c->ev[k], c->ev[k + 1]); c->ms[k] = t; }   // link, density, c2p+gather, pair, final
  cudaEventElapsedTime(&t, c->ev_pair[0], c->ev_pair[1]); c->ms[5] = t;                              // rates_pair_kernel alone
  if (s) *s = c->sc;
  return 0;
}

int ndspmhd_b200_download(nd_ctx *c, nd_arrays *a, unsigned mask, int idim) {
  if (!c || !a) return ND_ERR_INVALID_ARG;
  if (!c->uploaded) return set_err(c, ND_ERR_STATE, "download before upload");
  if (idim < c->ntotal) return set_err(c, ND_ERR_INVALID_ARG, "download: idim < ntotal (re-allocate the host arrays, src/ghostND_mhd.f90:383-386)");
  CU(cudaSetDevice(c->device));
  const size_t n = (size_t)(c->has_comm ? c->nown : ((mask & ND_DL_REAL_ROWS) ? c->npart : c->ntotal));   // with slabs only this rank's own rows go back
  for (int g = 1; g <= 3; g++) if (int e = download_group(c, a, n, g, mask, c->stream)) return e;
  if ((mask & ND_DL_GHOSTS) && !c->has_comm) {
    const size_t g0 = (size_t)c->npart, ng = (size_t)c->ntotal - g0, D = sizeof(double);
    auto dn = [&](void *dst, const void *src, size_t bytes) -> cudaError_t { return (dst && src) ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream) : cudaSuccess; };
    if (ng > 0) {
      if (a->x_out) CU(dn(a->x_out + g0 * c->ndim, c->x + g0 * c->ndim, D * c->ndim * ng));
      if (a->vel_out) CU(dn(a->vel_out + g0 * 3, c->vel + g0 * 3, D * 3 * ng));
      if (a->ireal_out) CU(dn(a->ireal_out + g0, c->ireal + g0, sizeof(int) * ng));
      if (a->itype_out) CU(dn(a->itype_out + g0, c->itype + g0, sizeof(int) * ng));
    }
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// upload + derivs + download in one call, with the copies overlapped with the kernels: the inputs of cons2prim/rates arrive
// while the density iteration runs, the density results and primitives leave while the rates run.  Host arrays should be
// page-locked (ndspmhd_b200_host_alloc) for the copies to be asynchronous.
int ndspmhd_b200_derivs_host(nd_ctx *c, nd_arrays *a, int npart, int ntotal, int idim, unsigned mask, nd_scalars *s) {
  if (!c || !a || !c->stream) return c ? set_err(c, ND_ERR_STATE, "context not initialised") : ND_ERR_INVALID_ARG;
  if (int e = check_upload_args(c, a, npart, ntotal, idim)) return e;
  const auto t_enter = std::chrono::steady_clock::now();
  CU(cudaSetDevice(c->device));
  int want = std::max(ntotal, c->ntotal);
  if (c->o.device_ghosts && any_ghost_bound(c)) want = std::max(want, npart + npart / 4 + 1024);
  if (int e = ensure_capacity(c, want, 0)) return e;
  const size_t nin = (size_t)ntotal;
  if (int e = upload_group(c, a, nin, 10, c->stream_h2d)) return e;
  CU(cudaEventRecord(c->ev_in[0], c->stream_h2d));
  if (int e = upload_group(c, a, nin, 11, c->stream_h2d)) return e;
  CU(cudaEventRecord(c->ev_in[2], c->stream_h2d));
  if (int e = upload_group(c, a, nin, 2, c->stream_h2d)) return e;
  CU(cudaEventRecord(c->ev_in[1], c->stream_h2d));
  c->npart = npart; c->ntotal = ntotal; c->nown = npart;
  c->uploaded = true; c->linked = c->density_done = c->prim_done = c->rates_done = false;
  CU(cudaStreamWaitEvent(c->stream, c->ev_in[0], 0));
  // slab contexts pack x, vel, pmass, hh, rho into the halo records during the link (k_halo_pack1): the whole first group must have landed;
  // en, Bevol, alpha, psi travel after the density iteration (halo_exchange_density waits for them)
  if (c->has_comm) { CU(cudaStreamWaitEvent(c->stream, c->ev_in[2], 0)); c->wait_in2 = true; }
  else c->wait_in1b = true;   // hhmax, the ghost count and its scan run while vel, rho are still on the wire (wait_second_half)
  const bool will_light = ND_DENS_LIGHT && fast_tuple(c->o) && (!c->has_comm || c->slab_light) && (mask & ND_DL_RATES);   // the rates of this call make drho/dt
  {
    // LIGHT rounds read {x, m, h} only: with periodic ghosts (a ghost's velocity is its parent's) and no fixed particles (they keep the
    // uploaded rho) the link and the whole density iteration run before vel has landed; k_late_vel then completes the records
    bool plain_ghosts = true;
    for (int d = 0; d < c->ndim; d++) if (!(c->o.ibound[d] == 0 || c->o.ibound[d] == 3)) plain_ghosts = false;
    c->defer_vel = will_light && !c->has_comm && plain_ghosts && c->o.device_ghosts && !c->o.onef_dust && !getenv("NDSPMHD_B200_NO_DEFER");
  }
  CU(cudaEventRecord(c->ev[0], c->stream));
  int e = DISPATCH_NDIM(c, do_link<1>(c), do_link<2>(c), do_link<3>(c));
  bool light = false;
  if (!e) {
    CU(cudaEventRecord(c->ev[1], c->stream));
    c->dens_light = will_light;
    light = c->dens_light;
    e = DISPATCH_NDIM(c, do_iterate_density<1>(c, 0), do_iterate_density<2>(c, 0), do_iterate_density<3>(c, 0));
    c->dens_light = false;
  }
  if (c->wait_in2) { c->wait_in2 = false; cudaStreamWaitEvent(c->stream, c->ev_in[1], 0); }   // the iteration left early (error): join anyway
  if (c->wait_in1b || c->defer_vel) {   // join the second half of the upload (also when the link or the iteration left early)
    c->wait_in1b = false;
    cudaStreamWaitEvent(c->stream, c->ev_in[2], 0);
    if (c->defer_vel && !e) LAUNCH(c, k_late_vel, nblocks(c->ntotal, 256), 256, 0, c->perm, c->ireal, c->vel, c->pmass, c->vm, c->npart, c->ntotal);
    c->defer_vel = false;
  }
  if (e) { cudaStreamSynchronize(c->stream_h2d); return e; }
  if (idim < c->ntotal) { cudaStreamSynchronize(c->stream_h2d); return set_err(c, ND_ERR_INVALID_ARG, "derivs_host: idim < ntotal after ghost generation"); }
  const size_t nout = (size_t)(c->has_comm ? c->nown : ((mask & ND_DL_REAL_ROWS) ? c->npart : c->ntotal));
  // the density outputs leave while cons2prim and the rates run -- unless fixed particles take their partner's gradgradh in cons2prim
  // (copy_particle, conservative2primitive.f90:424-437): then that group goes after cons2prim
  const bool dens_after_c2p = any_fixed_bound(c) && c->o.want_aux;
  CU(cudaEventRecord(c->ev_out[0], c->stream));
  CU(cudaStreamWaitEvent(c->stream_d2h, c->ev_out[0], 0));
  if (!dens_after_c2p) { if (int e2 = download_group(c, a, nout, 1, mask, c->stream_d2h)) return e2; }
  CU(cudaStreamWaitEvent(c->stream, c->ev_in[1], 0));
  CU(cudaEventRecord(c->ev[2], c->stream));
  e = do_cons2prim(c);
  if (e) { cudaStreamSynchronize(c->stream_h2d); cudaStreamSynchronize(c->stream_d2h); return e; }
  CU(cudaEventRecord(c->ev_out[1], c->stream));
  CU(cudaStreamWaitEvent(c->stream_d2h, c->ev_out[1], 0));
  if (dens_after_c2p) { if (int e2 = download_group(c, a, nout, 1, mask, c->stream_d2h)) return e2; }
  if (int e2 = download_group(c, a, nout, 2, mask, c->stream_d2h)) return e2;
  // The rates run in row chunks (do_get_rates) so that a chunk's rows go down the wire while the next chunk's pair kernel
  // runs; what is left after the last kernel is one chunk, dpsidt and the (zero) ghost rows instead of the whole 168 B/row.
  int nchunk = 1;
  if (mask & ND_DL_RATES) {   // slab contexts too: a rank's own rows go down its own PCIe link while its next chunk runs
    if (const char *ev = getenv("NDSPMHD_B200_RATE_CHUNKS")) nchunk = std::max(1, std::min(64, atoi(ev)));
    else if (c->nown >= (1 << 20)) nchunk = 4;
  }
  std::vector<cudaEvent_t> &cev = c->chunk_events;
  while ((int)cev.size() < nchunk) { cudaEvent_t x; CU(cudaEventCreateWithFlags(&x, cudaEventDisableTiming)); cev.push_back(x); }
  c->rate_chunks = nchunk;
  int cb_err = 0;
  c->on_rates_chunk = [&](int q, int r0, int r1) -> int {
    if (cudaEventRecord(cev[q], c->stream) != cudaSuccess || cudaStreamWaitEvent(c->stream_d2h, cev[q], 0) != cudaSuccess) { cb_err = 1; return set_err(c, ND_ERR_CUDA, "derivs_host: chunk event"); }
    return download_rates_rows(c, a, (size_t)r0, (size_t)r1, mask, c->stream_d2h, false);
  };
  c->drho_pairs = light;
  e = DISPATCH_NDIM(c, do_get_rates<1>(c, nullptr, nullptr, nullptr, 0), do_get_rates<2>(c, nullptr, nullptr, nullptr, 0), do_get_rates<3>(c, nullptr, nullptr, nullptr, 0));
  c->drho_pairs = false;
  c->rate_chunks = 1; c->on_rates_chunk = nullptr;
  if (e || cb_err) { cudaStreamSynchronize(c->stream_h2d); cudaStreamSynchronize(c->stream_d2h); return e ? e : ND_ERR_CUDA; }
  if (nchunk == 1) {
    if (int e2 = download_group(c, a, nout, 3, mask, c->stream)) return e2;   // stream order: after the final kernel
  } else {                                                                    // stream order: after k_rates_dpsidt / k_rates_zero_ghosts
    if ((mask & ND_DL_RATES) && c->o.imhd != 0 && a->dpsidt) CU(cudaMemcpyAsync(a->dpsidt, c->dpsidt, sizeof(double) * nout, cudaMemcpyDeviceToHost, c->stream));
    if (int e2 = download_rates_rows(c, a, (size_t)c->npart, nout, mask, c->stream, false)) return e2;
  }
  if ((mask & ND_DL_GHOSTS) && !c->has_comm && c->ntotal > c->npart) {
    const size_t g0 = (size_t)c->npart, ng = (size_t)c->ntotal - g0, D = sizeof(double);
    if (a->x_out) CU(cudaMemcpyAsync(a->x_out + g0 * c->ndim, c->x + g0 * c->ndim, D * c->ndim * ng, cudaMemcpyDeviceToHost, c->stream_d2h));
    if (a->vel_out) CU(cudaMemcpyAsync(a->vel_out + g0 * 3, c->vel + g0 * 3, D * 3 * ng, cudaMemcpyDeviceToHost, c->stream_d2h));
    if (a->ireal_out) CU(cudaMemcpyAsync(a->ireal_out + g0, c->ireal + g0, sizeof(int) * ng, cudaMemcpyDeviceToHost, c->stream_d2h));
    if (a->itype_out) CU(cudaMemcpyAsync(a->itype_out + g0, c->itype + g0, sizeof(int) * ng, cudaMemcpyDeviceToHost, c->stream_d2h));
  }
  if (int e2 = fill_density_scalars(c)) return e2;
  CU(cudaStreamSynchronize(c->stream_d2h));
  CU(cudaStreamSynchronize(c->stream));
  float t;
  for (int k = 0; k < 5; k++) { cudaEventElapsedTime(&t, c->ev[k], c->ev[k + 1]); c->ms[k] = t; }
  cudaEventElapsedTime(&t, c->ev_pair[0], c->ev_pair[1]); c->ms[5] = t;
  if (getenv("NDSPMHD_B200_DEBUG")) {
    const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
    fprintf(stderr, "derivs_host: wall %.1f ms; kernels link %.1f density %.1f c2p %.1f rates_pair %.1f final %.1f\n", wall, c->ms[0], c->ms[1], c->ms[2], c->ms[3], c->ms[4]);
  }
  if (s) *s = c->sc;
  return 0;
}

int ndspmhd_b200_update_ghosts(nd_ctx *c, const nd_arrays *a, int ntotal, int idim, double hhmax) {
  if (!c || !a) return ND_ERR_INVALID_ARG;
  if (!c->uploaded) return set_err(c, ND_ERR_STATE, "update_ghosts before upload");
  if (c->o.device_ghosts) return set_err(c, ND_ERR_STATE, "update_ghosts with device_ghosts=1");
  if (ntotal < c->npart || idim < ntotal) return set_err(c, ND_ERR_INVALID_ARG, "update_ghosts: need npart <= ntotal <= idim");
  if (ntotal > c->npart && (!a->x || !a->vel || !a->itype || !a->ireal)) return set_err(c, ND_ERR_INVALID_ARG, "update_ghosts: x, vel, itype, ireal required");
  CU(cudaSetDevice(c->device));
  if (int e = ensure_capacity(c, ntotal, c->npart)) return e;
  const size_t g0 = (size_t)c->npart, ng = (size_t)ntotal - g0;
  if (ng > 0) {
    CU(cudaMemcpyAsync(c->x + g0 * c->ndim, a->x + g0 * c->ndim, sizeof(double) * c->ndim * ng, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->vel + g0 * 3, a->vel + g0 * 3, sizeof(double) * 3 * ng, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->itype + g0, a->itype + g0, sizeof(int) * ng, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->ireal + g0, a->ireal + g0, sizeof(int) * ng, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  c->ntotal = ntotal;
  c->o.hhmax = hhmax;
  return 0;
}

void *ndspmhd_b200_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
  return p;
}
void ndspmhd_b200_host_free(void *p) { if (p) cudaFreeHost(p); }

int ndspmhd_b200_last_timings(const nd_ctx *c, double ms[8]) {
  if (!c || !ms) return ND_ERR_INVALID_ARG;
  for (int k = 0; k < 8; k++) ms[k] = c->ms[k];
  return 0;
}

int ndspmhd_b200_rewind(nd_ctx *c) {
  if (!c) return ND_ERR_INVALID_ARG;
  if (!c->uploaded) return set_err(c, ND_ERR_STATE, "rewind before upload");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(c->hh, c->hh0, sizeof(double) * c->npart, cudaMemcpyDeviceToDevice, c->stream));
  c->linked = c->density_done = c->prim_done = c->rates_done = false;
  return 0;
}

int ndspmhd_b200_selftest_math(nd_ctx *c, const double *in, double *out_sqrt, double *out_rsqrt, int n) {
  if (!c || !in || !out_sqrt || !out_rsqrt || n < 1) return ND_ERR_INVALID_ARG;
  CU(cudaSetDevice(c->device));
  double *d = nullptr;
  CU(cudaMalloc(&d, sizeof(double) * 3 * (size_t)n));
  CU(cudaMemcpyAsync(d, in, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, k_selftest_math, nblocks(n, 256), 256, 0, d, d + n, d + 2 * (size_t)n, n);
  CU(cudaMemcpyAsync(out_sqrt, d + n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(out_rsqrt, d + 2 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  cudaFree(d);
  return 0;
}

long long ndspmhd_b200_launch_count(const nd_ctx *c) { return c ? c->launches : 0; }
void *ndspmhd_b200_stream(const nd_ctx *c) { return c ? (void *)c->stream : nullptr; }

int ndspmhd_b200_rates_pairs(nd_ctx *c, int *pair_i, int *pair_j, long long cap, long long *npairs) {
  if (!c || !npairs) return ND_ERR_INVALID_ARG;
  if (!c->prim_done) return set_err(c, ND_ERR_STATE, "rates_pairs before cons2prim");
  CU(cudaSetDevice(c->device));
  int *di = nullptr, *dj = nullptr; unsigned long long *dc = nullptr;
  CU(cudaMalloc(&di, sizeof(int) * std::max<long long>(cap, 1)));
  CU(cudaMalloc(&dj, sizeof(int) * std::max<long long>(cap, 1)));
  CU(cudaMalloc(&dc, sizeof(unsigned long long)));
  CU(cudaMemset(dc, 0, sizeof(unsigned long long)));
  int e = DISPATCH_NDIM(c, do_get_rates<1>(c, di, dj, dc, cap), do_get_rates<2>(c, di, dj, dc, cap), do_get_rates<3>(c, di, dj, dc, cap));
  unsigned long long n = 0;
  cudaMemcpy(&n, dc, sizeof(n), cudaMemcpyDeviceToHost);
  const long long m = std::min<long long>((long long)n, cap);
  if (!e && m > 0 && pair_i && pair_j) { cudaMemcpy(pair_i, di, sizeof(int) * m, cudaMemcpyDeviceToHost); cudaMemcpy(pair_j, dj, sizeof(int) * m, cudaMemcpyDeviceToHost); }
  cudaFree(di); cudaFree(dj); cudaFree(dc);
  *npairs = (long long)n;
  return e;
}

/* ---- get_curl as an operator on the resident state ---- */
int ndspmhd_b200_get_curl(nd_ctx *c, int icurltype, const double *Bvec, double *curlB, double *gradB, int idim) {
  if (!c || !Bvec || !curlB) return ND_ERR_INVALID_ARG;
  if (!c->density_done) return set_err(c, ND_ERR_STATE, "get_curl needs the linked state and the converged rho, h, gradh of iterate_density");
  if (icurltype < 1 || icurltype > 4) return set_err(c, ND_ERR_UNSUPPORTED_OPTION, "get_curl: icurltype must be 1..4");
  if (gradB && icurltype != 1) return set_err(c, ND_ERR_UNSUPPORTED_OPTION, "get_curl: gradB comes with icurltype = 1 only (src/get_curl.f90:250-255)");
  if (c->has_comm) return set_err(c, ND_ERR_UNSUPPORTED_OPTION, "get_curl is not available on slab-decomposed contexts");
  if (idim < c->npart) return set_err(c, ND_ERR_INVALID_ARG, "get_curl: idim < npart");
  CU(cudaSetDevice(c->device));
  const size_t n = (size_t)c->npart;
  double *d = nullptr;
  CU(cudaMalloc(&d, sizeof(double) * n * (3 + 3 + (gradB ? 9 : 0))));
  double *dB = d, *dC = d + 3 * n, *dG = gradB ? d + 6 * n : nullptr;
  CU(cudaMemcpyAsync(dB, Bvec, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
  int e = DISPATCH_NDIM(c, do_get_curl<1>(c, icurltype, dB, dC, dG, nullptr), do_get_curl<2>(c, icurltype, dB, dC, dG, nullptr), do_get_curl<3>(c, icurltype, dB, dC, dG, nullptr));
  if (!e) {
    CU(cudaMemcpyAsync(curlB, dC, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    if (gradB) CU(cudaMemcpyAsync(gradB, dG, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  cudaFree(d);
  // the operator re-used the sorted rates records (bpsi, gal, posh.w): a following get_rates regathers them (k_rates_gather) anyway
  return e;
}

/* ---- leapfrog step on the resident state ---- */
int ndspmhd_b200_step(nd_ctx *c, const nd_step_opts *so, double *dt_inout, nd_scalars *s) {
  if (!c || !so || !dt_inout) return ND_ERR_INVALID_ARG;
  if (!c->uploaded || !c->rates_done) return set_err(c, ND_ERR_STATE, "step needs a prior upload + derivs (the reference enters `step` with the rates of the previous call)");
  const nd_options &o = c->o;
  if (o.imhd < 0 || o.idivbzero == 10 || o.idustevol != 0) return set_err(c, ND_ERR_UNSUPPORTED_OPTION, "step: imhd < 0, idivbzero = 10 and idustevol /= 0 are not supported");
  bool ghost_bound = false;
  for (int d = 0; d < c->ndim; d++) if (o.ibound[d] >= 2) ghost_bound = true;
  if (ghost_bound && !o.device_ghosts) return set_err(c, ND_ERR_STATE, "step on the resident state needs device_ghosts = 1 (derivs regenerates the ghosts, src/derivs.f90:78)");
  if (c->has_comm && any_fixed_bound(c)) return set_err(c, ND_ERR_UNSUPPORTED_OPTION, "step: slab-decomposed contexts with fixed-particle boundaries (rows tied to partner rows cannot migrate)");
  CU(cudaSetDevice(c->device));
  int np = c->nown;                                            // own rows (= npart without slabs; halo rows are remade by derivs)
  if (c->stepbufrows < (size_t)np) {
    if (c->stepbuf) cudaFree(c->stepbuf);
    c->stepbuf = nullptr; c->stepbufrows = 0;
    const size_t rows = c->has_comm ? (size_t)np + (size_t)np / 8 + 1024 : (size_t)np;   // room for arrivals
    CU(cudaMalloc(&c->stepbuf, sizeof(double) * (size_t)STEP_NIN_DUST * rows));
    c->stepbufrows = rows;
  }
  auto args = [&]() {
    StepArgs A;
    A.x = c->x; A.vel = c->vel; A.Bevol = c->Bevol; A.rho = c->rho; A.hh = c->hh; A.en = c->en; A.alpha = c->alpha; A.psi = c->psi;
    A.dustevol = c->dustevol; A.deltav = c->deltav;
    A.force = c->force; A.dBevoldt = c->dBevoldt; A.drhodt = c->drhodt; A.dhdt = c->dhdt; A.dendt = c->dendt; A.daldt = c->daldt; A.dpsidt = c->dpsidt;
    A.ddustevoldt = c->ddustevoldt; A.ddeltavdt = c->ddeltavdt;
    A.itype = c->itype; A.ireal = c->ireal; A.in = c->stepbuf; A.n = c->stepbufrows; A.npart = np; A.ndim = c->ndim;
    A.imhd = o.imhd; A.iresist = o.iresist; A.icty = o.icty; A.ihvar = o.ihvar; A.iener = o.iener; A.idivbzero = o.idivbzero; A.idust = o.idust; A.onef = o.onef_dust;
    for (int d = 0; d < 3; d++) { A.iavlim[d] = o.iavlim[d]; A.ibound[d] = d < c->ndim ? o.ibound[d] : 0; A.xmin[d] = o.xmin[d]; A.xmax[d] = o.xmax[d]; }
    A.dt = *dt_inout; A.damp = o.damp; A.flags = c->flags;
    return A;
  };
  StepArgs A = args();
  LAUNCH(c, k_step_save, nblocks(np, 256), 256, 0, A);
  LAUNCH(c, k_step_predict, nblocks(np, 256), 256, 0, A);
  LAUNCH(c, k_step_boundary, nblocks(np, 256), 256, 0, A);     // src/derivs.f90:74
  c->linked = c->density_done = c->prim_done = c->rates_done = false;
  if (c->has_comm) {                                           // rows that left the slab change owner before the link (the corrector does not move x)
    c->npart = c->ntotal = np;
    if (int e = migrate_rows(c)) return e;
    np = c->nown;
  }
  c->ntotal = (ghost_bound || c->has_comm) ? np : c->ntotal;   // the ghost rows are stale: derivs makes new ones from rows [0,npart)
  if (int e = ndspmhd_b200_derivs(c, s)) return e;
  A = args();                                                  // derivs may have re-allocated the arrays (more ghosts)
  LAUNCH(c, k_step_correct, nblocks(np, 256), 256, 0, A);
  LAUNCH(c, k_step_boundary, nblocks(np, 256), 256, 0, A);     // :216
  if (int e = sync_flags(c)) return e;
  double bad = c->h_flags[1] ? 1. : 0.;
  if (int e = comm_allreduce(c, &bad, 1, 0)) return e;         // every rank leaves together
  if (bad != 0.) {
    const int code = c->h_flags[1] ? c->h_flags[1] : ND_ERR_COMM;
    CU(cudaMemsetAsync(c->flags, 0, sizeof(int) * 16, c->stream));
    return set_err(c, code, code == ND_ERR_H_NONPOSITIVE ? "step: hh -ve" : code == ND_ERR_COMM ? "step: another rank reported an error" : "step: itypebnd2 (cylindrical fixed particles) is not supported");
  }
  // new timestep, :239-253
  if (!so->dtfixed) *dt_inout = std::min(std::min(so->C_force * c->sc.dtforce, so->C_cour * c->sc.dtcourant), std::min(0.9 * c->sc.dtdrag, so->C_force * c->sc.dtvisc));
  return 0;
}

/* evolved state of rows [0,npart) back to the host arrays (NULL pointers skipped) */
int ndspmhd_b200_download_state(nd_ctx *c, const nd_state_out *st, int idim) {
  if (!c || !st) return ND_ERR_INVALID_ARG;
  if (!c->uploaded) return set_err(c, ND_ERR_STATE, "download_state before upload");
  if (idim < c->nown) return set_err(c, ND_ERR_INVALID_ARG, "download_state: idim < npart");
  CU(cudaSetDevice(c->device));
  const size_t n = (size_t)c->nown, D = sizeof(double);   // own rows (slab contexts: ndspmhd_b200_row_counts tells how many there are now)
  auto dn = [&](void *dst, const void *src, size_t bytes) -> cudaError_t { return (dst && src) ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream) : cudaSuccess; };
  CU(dn(st->x, c->x, D * c->ndim * n)); CU(dn(st->vel, c->vel, D * 3 * n)); CU(dn(st->hh, c->hh, D * n)); CU(dn(st->en, c->en, D * n));
  CU(dn(st->Bevol, c->Bevol, D * 3 * n)); CU(dn(st->alpha, c->alpha, D * 3 * n)); CU(dn(st->psi, c->psi, D * n)); CU(dn(st->rho, c->rho, D * n));
  if (c->o.onef_dust) { CU(dn(st->dustevol, c->dustevol, D * n)); CU(dn(st->deltav, c->deltav, D * 3 * n)); }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int ndspmhd_b200_evwrite(nd_ctx *c, nd_evwrite *ev) {
  if (!c || !ev) return ND_ERR_INVALID_ARG;
  if (!c->uploaded || !c->rates_done) return set_err(c, ND_ERR_STATE, "evwrite needs the primitives and rates of a derivs / step");
  CU(cudaSetDevice(c->device));
  if (!c->evpartial) {
    CU(cudaMalloc(&c->evpartial, sizeof(double) * EV_BLOCKS * EV_NQ));
    CU(cudaMallocHost(&c->h_ev, sizeof(double) * 64));
  }
  EvArgs A;
  A.x = c->x; A.vel = c->vel; A.pmass = c->pmass; A.rho = c->rho; A.uu = c->uu; A.Bfield = c->Bfield; A.pr = c->pr; A.divB = c->divB; A.hh = c->hh;
  A.force = c->force; A.dustfrac = c->dustfrac; A.deltav = c->deltav; A.npart = c->npart; A.ndim = c->ndim; A.imhd = c->o.imhd; A.onef = c->o.onef_dust;
  A.partial = c->evpartial;
  LAUNCH(c, k_evwrite_partial, EV_BLOCKS, EV_THREADS, 0, A);
  LAUNCH(c, k_evwrite_final, 1, 64, 0, c->evpartial, c->h_ev);
  CU(cudaStreamSynchronize(c->stream));
  const double *h = c->h_ev;
  const double np = (double)c->npart;
  memset(ev, 0, sizeof(*ev));
  ev->ekin = h[EV_EKIN]; ev->etherm = h[EV_ETHERM]; ev->emag = h[EV_EMAG]; ev->emagp = h[EV_EMAGP]; ev->epot = 0.;
  for (int k = 0; k < 3; k++) { ev->mom[k] = h[EV_MOM0 + k]; ev->dmom[k] = h[EV_DMOM0 + k]; ev->ang[k] = h[EV_ANG0 + k]; ev->fluxtot[k] = h[EV_FLUX0 + k]; }
  ev->etot = ev->ekin + ev->emag + ev->epot;                                                        // :263
  if (c->o.iprterm >= 0 || c->o.iprterm < -1) ev->etot = ev->etot + ev->etherm;
  ev->momtot = std::sqrt((ev->mom[0] * ev->mom[0] + ev->mom[1] * ev->mom[1]) + ev->mom[2] * ev->mom[2]);
  ev->dmomtot = std::sqrt((ev->dmom[0] * ev->dmom[0] + ev->dmom[1] * ev->dmom[1]) + ev->dmom[2] * ev->dmom[2]);
  ev->angtot = std::sqrt((ev->ang[0] * ev->ang[0] + ev->ang[1] * ev->ang[1]) + ev->ang[2] * ev->ang[2]);
  ev->rhomin = h[EV_RHOMIN]; ev->rhomax = h[EV_RHOMAX]; ev->rhomean = h[EV_RHOSUM] / np;           // minmaxave
  ev->ekiny = h[EV_EKINY]; ev->totmassgas = h[EV_MGAS]; ev->totmassdust = h[EV_MDUST];
  if (c->o.imhd != 0) {                                                                             // :277-284
    ev->fluxtotmag = std::sqrt((ev->fluxtot[0] * ev->fluxtot[0] + ev->fluxtot[1] * ev->fluxtot[1]) + ev->fluxtot[2] * ev->fluxtot[2]);
    ev->betamhdav = h[EV_BETAAV] / np; ev->betamhdmin = h[EV_BETAMIN];
    ev->fracdivBok = 100. * h[EV_FRACOK] / np; ev->omegamhdav = h[EV_OMEGAAV] / np; ev->omegamhdmax = h[EV_OMEGAMAX];
    ev->divBav = h[EV_DIVBAV] / np; ev->divBmax = h[EV_DIVBMAX]; ev->divBtot = h[EV_DIVBTOT]; ev->crosshel = h[EV_CROSSHEL];
  }
  return 0;
}

}  // extern "C"
