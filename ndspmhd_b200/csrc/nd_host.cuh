// nd_host.cuh -- host orchestration of the hot path on one context: device buffers and their growth, halo exchange through the nd_comm
// callbacks, ghost generation, set_linklist, neighbour-list capacity, iterate_density, conservative2primitive and get_rates as sequences of
// kernel launches on the context's stream (the do_* functions the C entry points of nd_capi.cu call).
// Included by nd_capi.cu INSIDE its anonymous namespace, after nd_kernels.cuh.
#pragma once
// =====================================================================================================
// host orchestration
// =====================================================================================================
#define DISPATCH_NDIM(c, expr1, expr2, expr3) ((c)->ndim == 1 ? (expr1) : (c)->ndim == 2 ? (expr2) : (expr3))

template <class T> int dev_alloc(nd_ctx *c, T **p, size_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  if (n == 0) n = 1;
  CU(cudaMalloc((void **)p, n * sizeof(T)));
  return 0;
}

void register_rows(nd_ctx *c) {
  auto &v = c->rowbufs;
  v.clear();
  const size_t D = sizeof(double), I = sizeof(int), D4 = sizeof(double4);
  v.push_back({(void **)&c->x, D * c->ndim});
#define R3(a) v.push_back({(void **)&c->a, D * 3})
#define R1(a) v.push_back({(void **)&c->a, D})
#define RI(a) v.push_back({(void **)&c->a, I})
#define R4(a) v.push_back({(void **)&c->a, D4})
  R3(vel); R1(pmass); R1(hh); R1(en); R3(Bevol); R3(alpha); R1(psi); RI(itype); RI(ireal); R1(hhin); R1(hh0); R1(alphaB_in);
  R1(rho); R1(gradh); R1(drhodt); R1(dhdt); R1(rhoalt); R1(gradhn); R1(gradsoft); R1(gradgradh); RI(numneigh);
  R1(dens); R1(uu); R1(pr); R1(spsound); R3(Bfield);
  R3(force); R1(dudt); R1(dendt); R3(dBevoldt); R3(daldt); R1(dpsidt); R3(gradpsi); R1(divB); R3(curlB); R3(graddivv); R1(del2u);
  v.push_back({(void **)&c->p32, sizeof(float4)});
  R1(srho); R4(posh); R4(vm); R4(posm); R4(bpsi); R4(thermo); R4(gal); R4(sF); R4(sdB); R4(sC); R4(sP); R4(sV);
  v.push_back({(void **)&c->gid, sizeof(long long)});
  RI(typ); RI(perm); RI(permtmp); RI(inv); RI(cellOf); RI(cellOfOrig); RI(redo); RI(list); RI(ghostcount);
  if (c->o.onef_dust) {
    R1(dustevol); R1(dustfrac); R3(deltav); R1(rhogas); R1(rhodust); R1(ddustevoldt); R3(ddeltavdt); R1(sdf); R4(dusta); R4(sD);
    v.push_back({(void **)&c->dustb, sizeof(double2)});
  }
#undef R3
#undef R1
#undef RI
#undef R4
}

// grow every per-particle array to `rows` rows, keeping the first `keep` rows
int ensure_capacity(nd_ctx *c, int rows, int keep) {
  if (rows <= c->cap) return 0;
  const int newcap = (int)std::min<long long>(2000000000LL, (long long)rows + rows / 8 + 1024);
  if (c->stream_h2d) CU(cudaStreamSynchronize(c->stream_h2d));   // a pipelined upload may still be writing the old buffers
  if (c->stream_d2h) CU(cudaStreamSynchronize(c->stream_d2h));
  for (auto &rb : c->rowbufs) {
    void *np_ = nullptr;
    CU(cudaMalloc(&np_, rb.rowbytes * (size_t)newcap));
    CU(cudaMemsetAsync(np_, 0, rb.rowbytes * (size_t)newcap, c->stream));
    if (*rb.p && keep > 0) CU(cudaMemcpyAsync(np_, *rb.p, rb.rowbytes * (size_t)keep, cudaMemcpyDeviceToDevice, c->stream));
    if (*rb.p) { CU(cudaStreamSynchronize(c->stream)); cudaFree(*rb.p); }
    *rb.p = np_;
  }
  // scan output needs rows+1 ints
  if (c->scanout) cudaFree(c->scanout);
  CU(cudaMalloc(&c->scanout, sizeof(int) * ((size_t)newcap + 1)));
  c->cap = newcap;
  return 0;
}

int check_options(nd_ctx *c, const nd_options &o, int ndim) {
  auto bad = [&](const char *m) { return set_err(c, ND_ERR_UNSUPPORTED_OPTION, std::string("unsupported option: ") + m); };
  if (ndim < 1 || ndim > 3) return set_err(c, ND_ERR_INVALID_ARG, "ndim must be 1, 2 or 3");
  if (o.ikernav != 3) return bad("ikernav /= 3");
  if (o.iprterm != 0) return bad("iprterm /= 0");
  if (!(o.imhd == 0 || o.imhd == 1 || o.imhd == 11)) return bad("imhd not in {0,1,11}");
  if (o.imhd != 0 && o.imagforce != 2) return bad("imagforce /= 2");
  if (o.iav < 0 || o.iav > 3) return bad("iav not in 0..3");
  if (!(o.iener == 0 || o.iener == 2 || o.iener == 3)) return bad("iener not in {0,2,3}");
  if (!(o.idust == 0 || o.idust == 1 || o.idust == 2)) return bad("idust not in {0,1,2}");
  if (!(o.iresist == 0 || o.iresist == 1)) return bad("iresist not in {0,1}");
  if (o.icty != 0 || o.ixsph != 0 || o.igravity != 0 || o.iexternal_force != 0 || o.damp != 0.) return bad("icty/ixsph/igravity/iexternal_force/damp");
  if (o.usenumdens || o.ibiascorrection || o.iuse_exact_derivs || o.iambipolar || o.ivisc || o.iquantum || o.ind_timesteps || o.islope_limiter >= 0)
    return bad("usenumdens/ibiascorrection/iuse_exact_derivs/iambipolar/ivisc/iquantum/ind_timesteps/islope_limiter");
  if ((o.idust == 1) != (o.onef_dust != 0)) return bad("onef_dust must be set exactly when idust = 1 (initialiseND_mhd.f90:148; idust = 3, 4 unsupported)");
  if (o.idust == 1) {
    // one-fluid dust (Laibe & Price 2014): dust fraction evolved directly, thermal energy, AV 1-3.  Ghost rows of rhogas/rhodust
    // are only refreshed by copy_particle when every boundary is periodic (conservative2primitive.f90:465); with reflecting or
    // mixed ghosts the reference reads stale partial sums there, which is not reproduced.
    if (o.idustevol != 0) return bad("idustevol /= 0 with one-fluid dust");
    if (o.iener == 3 || o.iener == 1) return bad("iener = 1, 3 with one-fluid dust (the reference stops, ratesND_mhd.f90:2044)");
    if (o.iav < 1 || o.iav > 3) return bad("one-fluid dust needs iav in 1..3");
    bool ghosts = false, all3 = true;
    for (int d = 0; d < ndim; d++) { if (o.ibound[d] >= 2) ghosts = true; if (o.ibound[d] != 3) all3 = false; }
    if (ghosts && !all3) return bad("one-fluid dust with ghost boundaries needs ibound = 3 in every dimension");
  }
  if (o.ikernelalt != o.ikernel) return bad("ikernelalt /= ikernel");
  // iavlim(3) = 2 (get_curl inside conservative2primitive): with B/rho evolved the reference's get_curl reads the ghosts' Bfield of the
  // PREVIOUS derivs (the loop of conservative2primitive.f90:190-193 stops at npart); only imhd >= 11 (`Bfield = Bevol`, whole array) is
  // a function of this call's inputs
  if (o.iavlim[2] == 2 && o.imhd != 0 && o.imhd < 11) return bad("iavlim(3) = 2 needs imhd >= 11 (B evolved)");
  for (int d = 0; d < ndim; d++) {
    const int b = o.ibound[d];
    if (!(b == 0 || b == 1 || b == 2 || b == 3)) return bad("ibound not in {0,1,2,3}");
  }
  return 0;
}

Grid make_grid(nd_ctx *c) {
  Grid G;
  G.fineStart = c->cellStart; G.cellOf = c->cellOf; G.perm = c->perm; G.posh = c->posh; G.vm = c->vm; G.posm = c->posm; G.typ = c->typ; G.p32 = c->p32;
  G.hhmax1 = 1.0 / c->hhmax;
  // FP32 screening band (scaled units, cell = 1): 4 * 2^-23 * largest scaled coordinate + rounding of the thresholds
  G.screen_margin = 4.f * 1.1920929e-7f * (float)std::max(c->ncellsx[0], std::max(c->ncellsx[1], c->ncellsx[2])) + 2.e-6f;
  G.cull_margin = 16.f * G.screen_margin + 1.e-5f;
  G.nx = c->ncellsx[0]; G.ny = c->ncellsx[1]; G.nz = c->ncellsx[2]; G.ncells = c->ncells;
  G.npart = c->npart; G.ntotal = c->ntotal; G.nown = c->nown;
  G.radkern2 = c->T->radkern2; G.dq2table = c->T->dq2table; G.ddq2table = c->T->ddq2table;
  G.tab = c->d_tab; G.tab2 = c->d_tab2; G.tabdrag = c->d_tabdrag; G.tabg = c->d_tabg; G.tabw = c->d_tabw;
  return G;
}

void fill_link_scalars(nd_ctx *c);
// with slabs the x faces are halo faces: no ghosts are made in x on any rank
int local_ibound(const nd_ctx *c, int d) { return (c->has_comm && d == 0 && c->o.ibound[0] >= 2) ? 0 : c->o.ibound[d]; }
bool any_ghost_bound(const nd_ctx *c) { for (int d = 0; d < c->ndim; d++) if (local_ibound(c, d) >= 2) return true; return false; }
bool any_fixed_bound(const nd_ctx *c) { for (int d = 0; d < c->ndim; d++) if (c->o.ibound[d] == 1) return true; return false; }
bool has_copies(const nd_ctx *c) { return any_ghost_bound(c) || c->has_comm; }

#define NCCLCHK(call)                                                                                                   \
  do {                                                                                                                  \
    int r_ = (call);                                                                                                    \
    if (r_ != 0) return set_err(c, ND_ERR_COMM, std::string(#call) + ": " + (c->nccl_api->GetErrorString ? c->nccl_api->GetErrorString(r_) : "nccl error")); \
  } while (0)

int comm_allreduce(nd_ctx *c, double *v, int n, int op) {
  if (!c->has_comm) return 0;
  if (c->nccl) {   // native: H2D of the n doubles, ncclAllReduce on the compute stream, store to pinned memory, one synchronise
    if (n > 32) return set_err(c, ND_ERR_INVALID_ARG, "comm_allreduce: at most 32 values");
    for (int k = 0; k < n; k++) c->h_comm[k] = v[k];
    CU(cudaMemcpyAsync(c->d_comm, c->h_comm, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    NCCLCHK(c->nccl_api->AllReduce(c->d_comm, c->d_comm, (size_t)n, ND_NCCL_FLOAT64, op == 0 ? ND_NCCL_MAX : op == 1 ? ND_NCCL_MIN : ND_NCCL_SUM, c->nccl, c->stream));
    SMALL_D2H(c, c->h_comm + 32, c->d_comm, sizeof(double) * n);
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < n; k++) v[k] = c->h_comm[32 + k];
    c->n_allreduce++;
    return 0;
  }
  if (c->comm.allreduce(c->comm.user, v, n, op)) return set_err(c, ND_ERR_COMM, "allreduce callback failed");
  c->n_allreduce++;
  return 0;
}

// several maxima and minima in ONE all-reduce (min x = -max(-x)): every callback round trip costs ~0.1 ms of host latency
int comm_allreduce_maxmin(nd_ctx *c, double *mx, int nmx, double *mn, int nmn) {
  if (!c->has_comm) return 0;
  double v[16];
  if (nmx + nmn > 16) return set_err(c, ND_ERR_INVALID_ARG, "comm_allreduce_maxmin: too many values");
  for (int k = 0; k < nmx; k++) v[k] = mx[k];
  for (int k = 0; k < nmn; k++) v[nmx + k] = -mn[k];
  if (int e = comm_allreduce(c, v, nmx + nmn, 0)) return e;
  for (int k = 0; k < nmx; k++) mx[k] = v[k];
  for (int k = 0; k < nmn; k++) mn[k] = -v[nmx + k];
  return 0;
}

// native transport: pack n_max + n_sum device/host values into d_comm[off..), all-reduce the first n_max with MAX and the rest with SUM on
// the compute stream, and queue the D2H of the results to h_comm[32 + off ..).  No synchronise: the caller's next one covers it.
struct PackList {
  CommPack P;
  PackList() { P.n = 0; }
  void add(int kind, const void *src, double scale = 1., double hostval = 0.) { const int k = P.n++; P.kind[k] = kind; P.src[k] = src; P.scale[k] = scale; P.hostval[k] = hostval; }
};
int comm_reduce_packed(nd_ctx *c, const PackList &L, int off, int n_max, int n_sum) {
  if (L.P.n != n_max + n_sum || off + L.P.n > 32) return set_err(c, ND_ERR_INVALID_ARG, "comm_reduce_packed: bad sizes");
  LAUNCH(c, k_comm_pack, 1, 32, 0, L.P, c->d_comm + off);
  if (n_max > 0) { NCCLCHK(c->nccl_api->AllReduce(c->d_comm + off, c->d_comm + off, (size_t)n_max, ND_NCCL_FLOAT64, ND_NCCL_MAX, c->nccl, c->stream)); c->n_allreduce++; }
  if (n_sum > 0) { NCCLCHK(c->nccl_api->AllReduce(c->d_comm + off + n_max, c->d_comm + off + n_max, (size_t)n_sum, ND_NCCL_FLOAT64, ND_NCCL_SUM, c->nccl, c->stream)); c->n_allreduce++; }
  SMALL_D2H(c, c->h_comm + 32 + off, c->d_comm + off, sizeof(double) * L.P.n);
  return 0;
}

int sync_flags(nd_ctx *c) {   // D2H of the flag block; returns a pending device-side error code
  SMALL_D2H(c, c->h_flags, c->flags, sizeof(int) * 16);
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- hhmax = maxval(hh(1:npart)), ghostND_mhd.f90:79 (over all ranks with slabs) ----
int compute_hhmax(nd_ctx *c) {
  CU(cudaMemsetAsync(c->red, 0, sizeof(unsigned long long) * 16, c->stream));
  LAUNCH(c, k_max_h, std::min(nblocks(c->nown, 256), 1184), 256, 0, c->hh, c->nown, c->red);
  if (c->has_comm && c->nccl) {   // key -> double, all-reduce and D2H under one synchronise
    PackList L; L.add(CP_KEY, c->red);
    if (int e = comm_reduce_packed(c, L, 0, 1, 0)) return e;
    CU(cudaStreamSynchronize(c->stream));
    c->hhmax = c->h_comm[32];
    return 0;
  }
  SMALL_D2H(c, c->h_red, c->red, sizeof(unsigned long long));
  CU(cudaStreamSynchronize(c->stream));
  double m = c->nown > 0 ? dkey_inv(c->h_red[0]) : 0.;
  if (int e = comm_allreduce(c, &m, 1, 0)) return e;
  c->hhmax = m;
  return 0;
}

template <class T> int grow_buf(nd_ctx *c, T **p, size_t *cap, size_t need) {
  if (need <= *cap) return 0;
  if (*p) { CU(cudaStreamSynchronize(c->stream)); cudaFree(*p); *p = nullptr; *cap = 0; }
  const size_t n = need + need / 4 + 1024;
  CU(cudaMalloc((void **)p, n * sizeof(T)));
  *cap = n;
  return 0;
}

int halo_sendrecv(nd_ctx *c, const long long sb[2], const long long rb[2]) {
  void *const sbuf[2] = {c->sendbuf[0], c->sendbuf[1]};
  void *const rbuf[2] = {c->recvbuf[0], c->recvbuf[1]};
  c->halo_bytes_sent += sb[0] + sb[1];
  if (c->nccl) {
    // Grouped point-to-point on the compute stream, after the pack kernels and before the unpack kernels.  Posting order
    // [send->right, send->left, recv<-left, recv<-right]: on a ring of two ranks both neighbours are the same peer, and NCCL matches the
    // sends and receives between one pair of ranks in posting order -- the peer's first send (to ITS right = my left side) meets my first
    // receive (from my left).
    const int nr = c->comm.nranks, left = (c->comm.rank - 1 + nr) % nr, right = (c->comm.rank + 1) % nr;
    NCCLCHK(c->nccl_api->GroupStart());
    if (sb[1] > 0) NCCLCHK(c->nccl_api->Send(sbuf[1], (size_t)sb[1], ND_NCCL_CHAR, right, c->nccl, c->stream));
    if (sb[0] > 0) NCCLCHK(c->nccl_api->Send(sbuf[0], (size_t)sb[0], ND_NCCL_CHAR, left, c->nccl, c->stream));
    if (rb[0] > 0) NCCLCHK(c->nccl_api->Recv(rbuf[0], (size_t)rb[0], ND_NCCL_CHAR, left, c->nccl, c->stream));
    if (rb[1] > 0) NCCLCHK(c->nccl_api->Recv(rbuf[1], (size_t)rb[1], ND_NCCL_CHAR, right, c->nccl, c->stream));
    NCCLCHK(c->nccl_api->GroupEnd());
    return 0;
  }
  if (c->comm.sendrecv(c->comm.user, sbuf, sb, rbuf, rb, (void *)c->stream)) return set_err(c, ND_ERR_COMM, "sendrecv callback failed");
  return 0;
}

// byte counts of a neighbour exchange: what I send to {left, right} -> what I receive from {left, right}
int exchange_counts(nd_ctx *c, const long long sb[2], long long rb[2]) {
  if (c->nccl) {   // byte counts of every rank by one all-gather; mine come from my left neighbour's right side and vice versa
    const int nr_ = c->comm.nranks, rank = c->comm.rank;
    long long *h = reinterpret_cast<long long *>(c->h_comm);          // pinned: [0,2) mine, [2, 2 + 2 nranks) everybody's
    h[0] = sb[0]; h[1] = sb[1];
    CU(cudaMemcpyAsync(c->d_comm, h, sizeof(long long) * 2, cudaMemcpyHostToDevice, c->stream));
    NCCLCHK(c->nccl_api->AllGather(c->d_comm, c->d_comm + 2, 2, ND_NCCL_INT64, c->nccl, c->stream));
    SMALL_D2H(c, h + 2, c->d_comm + 2, sizeof(long long) * 2 * nr_);
    CU(cudaStreamSynchronize(c->stream));
    const int left = (rank - 1 + nr_) % nr_, right = (rank + 1) % nr_;
    rb[0] = h[2 + 2 * left + 1]; rb[1] = h[2 + 2 * right + 0];
    return 0;
  }
  long long sbc[2] = {sb[0], sb[1]};
  if (c->comm.sendrecv_counts(c->comm.user, sbc, rb)) return set_err(c, ND_ERR_COMM, "sendrecv_counts callback failed");
  return 0;
}

HaloPackArgs halo_args(nd_ctx *c) {
  HaloPackArgs A;
  A.ndim = c->ndim; A.x = c->x; A.vel = c->vel; A.pmass = c->pmass; A.hh = c->hh; A.en = c->en; A.Bevol = c->Bevol; A.alpha = c->alpha; A.psi = c->psi;
  A.rho = c->rho; A.gradh = c->gradh; A.itype = c->itype; A.full = 0; A.shift = 0; A.xbound = A.xperbound = 0.; A.row0 = 0; A.list = nullptr; A.n = 0; A.buf = nullptr;
  return A;
}

// ---- halo exchange 1: select the rows within reach of the slab faces, ship the inputs, append them as rows [nown, npart) ----
int halo_exchange_inputs(nd_ctx *c) {
  const nd_options &o = c->o;
  const int np = c->nown, rank = c->comm.rank, nr = c->comm.nranks;
  const bool periodic = (o.ibound[0] == 3);
  HaloSelArgs SA;
  SA.x = c->x; SA.nown = np; SA.ndim = c->ndim; SA.lo = c->comm.slab_lo; SA.hi = c->comm.slab_hi;
  SA.reach = c->T->radkern * c->hhmax * (1.0 + 1.e-10);
  // Halos come from the two adjacent ranks only: a slab narrower than the reach would miss neighbours two ranks away, and on a
  // periodic ring of two ranks a pair could be within reach through both faces (two copies of one particle).  Flagged here,
  // raised by do_link after the ranks' error flags have been all-reduced, so that every rank leaves together.
  c->slab_too_narrow = (SA.hi - SA.lo) < ((periodic && nr == 2) ? 2. : 1.) * SA.reach;
  SA.tol = 1.e-9 * std::max(1.0, std::fabs(SA.hi - SA.lo));
  SA.left_on = (periodic || rank > 0) ? 1 : 0; SA.right_on = (periodic || rank < nr - 1) ? 1 : 0;
  SA.flagL = c->cellOfOrig; SA.flagR = c->ghostcount; SA.err = c->flags + 1;   // scratch: both are rewritten later in this link
  LAUNCH(c, k_halo_flags, nblocks(np, 256), 256, 0, SA);
  for (int side = 0; side < 2; side++) {
    const int *flag = side == 0 ? c->cellOfOrig : c->ghostcount;
    if (int e = exclusive_scan(c, flag, c->scanout, np)) return e;
    SMALL_D2H(c, c->h_flags + 24, c->scanout + np, sizeof(int));
    CU(cudaStreamSynchronize(c->stream));
    const int n = c->h_flags[24];
    size_t cap = (size_t)c->sendcap[side];
    if (int e = grow_buf(c, &c->sendlist[side], &cap, (size_t)n)) return e;
    c->sendcap[side] = (int)cap;
    LAUNCH(c, k_halo_compact, nblocks(np, 256), 256, 0, flag, c->scanout, np, c->sendlist[side]);
    c->nsend[side] = n;
  }
  const long long rec = 8LL * halo1_nfields(c->ndim) + 4;
  long long sb[2] = {c->nsend[0] * rec, c->nsend[1] * rec}, rb[2] = {0, 0};
  if (int e = exchange_counts(c, sb, rb)) return e;
  if (rb[0] % rec || rb[1] % rec) return set_err(c, ND_ERR_COMM, "halo record size mismatch between ranks");
  c->nrecv[0] = (int)(rb[0] / rec); c->nrecv[1] = (int)(rb[1] / rec);
  for (int side = 0; side < 2; side++) {
    char **sp = (char **)&c->sendbuf[side], **rp = (char **)&c->recvbuf[side];
    if (int e = grow_buf(c, sp, &c->sendbufcap[side], (size_t)sb[side] + 64)) return e;
    if (int e = grow_buf(c, rp, &c->recvbufcap[side], (size_t)rb[side] + 64)) return e;
  }
  const int nsrc = np + c->nrecv[0] + c->nrecv[1];
  if (int e = ensure_capacity(c, nsrc + nsrc / 4 + 1024, np)) return e;
  for (int side = 0; side < 2; side++) {
    HaloPackArgs A = halo_args(c);
    A.list = c->sendlist[side]; A.n = c->nsend[side]; A.buf = (double *)c->sendbuf[side];
    if (periodic && side == 0 && rank == 0) { A.shift = 1; A.xbound = o.xmin[0]; A.xperbound = o.xmax[0]; }          // ghostND_mhd.f90:212-225, xmin face
    if (periodic && side == 1 && rank == nr - 1) { A.shift = 1; A.xbound = o.xmax[0]; A.xperbound = o.xmin[0]; }     // xmax face
    LAUNCH(c, k_halo_pack1, nblocks(A.n, 256), 256, 0, A);
  }
  if (int e = halo_sendrecv(c, sb, rb)) return e;
  int row0 = np;
  for (int side = 0; side < 2; side++) {
    HaloPackArgs A = halo_args(c);
    A.n = c->nrecv[side]; A.buf = (double *)c->recvbuf[side]; A.row0 = row0;
    LAUNCH(c, k_halo_unpack1, nblocks(A.n, 256), 256, 0, A);
    row0 += A.n;
  }
  c->npart = nsrc;
  return 0;
}

// ---- halo exchange 2: the owners' converged hh, rho, gradh for the halo rows (the rates test needs h_j: ratesND_mhd.f90:404-415) ----
int halo_exchange_density(nd_ctx *c, int full) {
  const long long recb = 8LL * halo2_nfields(full);
  long long sb[2] = {c->nsend[0] * recb, c->nsend[1] * recb}, rb[2] = {c->nrecv[0] * recb, c->nrecv[1] * recb};
  for (int side = 0; side < 2; side++) {   // the full record can be larger than exchange 1's
    char **sp = (char **)&c->sendbuf[side], **rp = (char **)&c->recvbuf[side];
    if (int e = grow_buf(c, sp, &c->sendbufcap[side], (size_t)sb[side] + 64)) return e;
    if (int e = grow_buf(c, rp, &c->recvbufcap[side], (size_t)rb[side] + 64)) return e;
  }
  if (full && c->wait_in2) { CU(cudaStreamWaitEvent(c->stream, c->ev_in[1], 0)); c->wait_in2 = false; }   // derivs_host: en, Bevol, alpha, psi have landed
  for (int side = 0; side < 2; side++) {
    HaloPackArgs A = halo_args(c);
    A.full = full;
    A.list = c->sendlist[side]; A.n = c->nsend[side]; A.buf = (double *)c->sendbuf[side];
    LAUNCH(c, k_halo_pack2, nblocks(A.n, 256), 256, 0, A);
  }
  if (int e = halo_sendrecv(c, sb, rb)) return e;
  int row0 = c->nown;
  for (int side = 0; side < 2; side++) {
    HaloPackArgs A = halo_args(c);
    A.full = full;
    A.n = c->nrecv[side]; A.buf = (double *)c->recvbuf[side]; A.row0 = row0;
    LAUNCH(c, k_halo_unpack2, nblocks(A.n, 256), 256, 0, A);
    row0 += A.n;
  }
  return 0;
}

// ---- particle migration between slabs (ndspmhd_b200_step): rows whose x left [slab_lo, slab_hi) go to the adjacent rank ----
int grow_stepbuf(nd_ctx *c, size_t rows, int keep) {
  if (rows <= c->stepbufrows) return 0;
  const size_t newrows = rows + rows / 8 + 1024;
  double *nb = nullptr;
  CU(cudaMalloc(&nb, sizeof(double) * (size_t)STEP_NIN_DUST * newrows));
  if (c->stepbuf && keep > 0)   // planes keep their contents: the stride changes
    CU(cudaMemcpy2DAsync(nb, sizeof(double) * newrows, c->stepbuf, sizeof(double) * c->stepbufrows, sizeof(double) * (size_t)keep, STEP_NIN_DUST, cudaMemcpyDeviceToDevice, c->stream));
  if (c->stepbuf) { CU(cudaStreamSynchronize(c->stream)); cudaFree(c->stepbuf); }
  c->stepbuf = nb; c->stepbufrows = newrows;
  return 0;
}

MigRows mig_rows(nd_ctx *c) {
  MigRows R;
  int a = 0;
  auto put = [&](double *p, int w) { R.arr[a] = p; R.width[a] = w; a++; };
  put(c->x, c->ndim); put(c->vel, 3); put(c->pmass, 1); put(c->hh, 1); put(c->en, 1); put(c->Bevol, 3); put(c->alpha, 3); put(c->psi, 1); put(c->rho, 1);
  if (c->o.onef_dust) { put(c->dustevol, 1); put(c->deltav, 3); }
  R.narr = a;
  for (; a < 12; a++) { R.arr[a] = nullptr; R.width[a] = 0; }
  R.in = c->stepbuf; R.instride = c->stepbufrows; R.nplanes = c->o.onef_dust ? STEP_NIN_DUST : STEP_NIN;
  R.itype = c->itype; R.gid = c->gid;
  return R;
}

int migrate_rows(nd_ctx *c) {
  const nd_options &o = c->o;
  const int np = c->nown, rank = c->comm.rank, nr = c->comm.nranks;
  const bool periodic = (o.ibound[0] == 3);
  if ((int)c->edges.size() != nr + 1) {        // every rank's faces, once: a one-hot sum (the transport has no all-gather of doubles)
    double v[32];
    if (nr + 1 > 32) return set_err(c, ND_ERR_INVALID_ARG, "migration: at most 31 ranks");
    for (int k = 0; k <= nr; k++) v[k] = 0.;
    v[rank] = c->comm.slab_lo;
    if (rank == nr - 1) v[nr] = c->comm.slab_hi;
    if (int e = comm_allreduce(c, v, nr + 1, 2)) return e;
    c->edges.assign(v, v + nr + 1);
  }
  const int left = (rank - 1 + nr) % nr, right = (rank + 1) % nr;
  MigSelArgs SA;
  SA.x = c->x; SA.itype = c->itype; SA.nown = np; SA.ndim = c->ndim;
  SA.lo = c->edges[rank]; SA.hi = c->edges[rank + 1]; SA.hi_closed = rank == nr - 1;
  SA.llo = c->edges[left]; SA.lhi = c->edges[left + 1]; SA.lhi_closed = left == nr - 1;
  SA.rlo = c->edges[right]; SA.rhi = c->edges[right + 1]; SA.rhi_closed = right == nr - 1;
  SA.left_on = (periodic || rank > 0) ? 1 : 0; SA.right_on = (periodic || rank < nr - 1) ? 1 : 0; SA.same_peer = (left == right) ? 1 : 0;
  SA.flagL = c->cellOfOrig; SA.flagR = c->ghostcount; SA.flagAny = c->redo; SA.err = c->flags + 1;   // scratch: all rewritten by the link that follows
  LAUNCH(c, k_migrate_flags, nblocks(np, 256), 256, 0, SA);
  int nside[2] = {0, 0};
  for (int side = 0; side < 2; side++) {
    const int *flag = side == 0 ? c->cellOfOrig : c->ghostcount;
    if (int e = exclusive_scan(c, flag, c->scanout, np)) return e;
    SMALL_D2H(c, c->h_flags + 24, c->scanout + np, sizeof(int));
    CU(cudaStreamSynchronize(c->stream));
    const int n = c->h_flags[24];
    size_t cap = (size_t)c->sendcap[side];
    if (int e = grow_buf(c, &c->sendlist[side], &cap, (size_t)n)) return e;
    c->sendcap[side] = (int)cap;
    LAUNCH(c, k_halo_compact, nblocks(np, 256), 256, 0, flag, c->scanout, np, c->sendlist[side]);
    nside[side] = n;
  }
  const int nl = nside[0] + nside[1], m = np - nl;
  if (int e = exclusive_scan(c, c->redo, c->scanout, np)) return e;                       // the holes, ascending
  LAUNCH(c, k_halo_compact, nblocks(np, 256), 256, 0, c->redo, c->scanout, np, c->list);
  MigRows R = mig_rows(c);
  const long long rec = 8LL * mig_nfields(R);
  long long sb[2] = {nside[0] * rec, nside[1] * rec}, rb[2] = {0, 0};
  if (int e = exchange_counts(c, sb, rb)) return e;
  if (rb[0] % rec || rb[1] % rec) return set_err(c, ND_ERR_COMM, "migration record size mismatch between ranks");
  const int na[2] = {(int)(rb[0] / rec), (int)(rb[1] / rec)};
  for (int side = 0; side < 2; side++) {
    char **sp = (char **)&c->sendbuf[side], **rp = (char **)&c->recvbuf[side];
    if (int e = grow_buf(c, sp, &c->sendbufcap[side], (size_t)sb[side] + 64)) return e;
    if (int e = grow_buf(c, rp, &c->recvbufcap[side], (size_t)rb[side] + 64)) return e;
  }
  for (int side = 0; side < 2; side++) {
    MigPackArgs A; A.R = R; A.list = c->sendlist[side]; A.n = nside[side]; A.row0 = 0; A.buf = (double *)c->sendbuf[side];
    LAUNCH(c, k_migrate_pack, nblocks(A.n, 128), 128, 0, A);
  }
  if (int e = halo_sendrecv(c, sb, rb)) return e;
  c->halo_bytes_sent -= sb[0] + sb[1];                       // counted separately
  c->migrated_bytes_sent += sb[0] + sb[1];
  if (nl > 0) {                                              // fill the holes below m from the staying rows of the tail [m, np)
    LAUNCH(c, k_migrate_tailflags, nblocks(nl, 256), 256, 0, c->redo, m, nl, c->permtmp);
    if (int e = exclusive_scan(c, c->permtmp, c->scanout, nl)) return e;
    LAUNCH(c, k_migrate_fill, nblocks(nl, 128), 128, 0, R, c->permtmp, c->scanout, c->list, m, nl);
  }
  const int nnew = m + na[0] + na[1];
  if (int e = ensure_capacity(c, nnew + nnew / 8 + 1024, m)) return e;
  if (int e = grow_stepbuf(c, (size_t)nnew, m)) return e;
  R = mig_rows(c);                                           // the arrays may have moved
  int row0 = m;
  for (int side = 0; side < 2; side++) {
    MigPackArgs A; A.R = R; A.list = nullptr; A.n = na[side]; A.row0 = row0; A.buf = (double *)c->recvbuf[side];
    LAUNCH(c, k_migrate_unpack, nblocks(A.n, 128), 128, 0, A);
    row0 += A.n;
  }
  c->nown = c->npart = c->ntotal = nnew;
  c->nmigrated_out += nl; c->nmigrated_in += na[0] + na[1];
  return 0;
}

// derivs_host uploads {x, hh, itype, ireal} first and {vel, pmass, rho} behind them: the compute stream joins the second half here,
// right before its first reader (the ghost rows' velocities, else the sorted records)
int wait_second_half(nd_ctx *c) {
  if (c->defer_vel) return 0;   // nothing before the end of the density iteration reads them (derivs_host)
  if (c->wait_in1b) { CU(cudaStreamWaitEvent(c->stream, c->ev_in[2], 0)); c->wait_in1b = false; }
  return 0;
}

// ---- ghosts (device_ghosts=1): rows [npart, ntotal) from rows [0, npart) ----
template <int NDIM> int make_ghosts(nd_ctx *c) {
  if (int e = compute_hhmax(c)) return e;
  c->npart = c->nown;
  if (c->has_comm) { if (int e = halo_exchange_inputs(c)) return e; }
  const int np = c->npart;
  c->ntotal = np;
  if (!any_ghost_bound(c)) return 0;
  GhostArgs A;
  A.x = c->x; A.vel = c->defer_vel ? nullptr : c->vel; A.hh = c->hh; A.itype = c->itype; A.ireal = c->ireal; A.offset = c->scanout; A.count = c->ghostcount;
  A.npart = np; A.cap = c->cap; A.radkern = c->T->radkern; A.hhmax = c->hhmax; A.flags = c->flags;
  for (int d = 0; d < 3; d++) { A.ibound[d] = d < NDIM ? local_ibound(c, d) : 0; A.xmin[d] = c->o.xmin[d]; A.xmax[d] = c->o.xmax[d]; }
  LAUNCH(c, (k_ghosts<NDIM, false>), nblocks(np, 256), 256, 0, A);
  if (int e = exclusive_scan(c, c->ghostcount, c->scanout, np)) return e;
  SMALL_D2H(c, c->h_flags + 25, c->scanout + np, sizeof(int));
  CU(cudaStreamSynchronize(c->stream));
  const int nghost = c->h_flags[25];
  const int cap_before = c->cap;
  if (int e = ensure_capacity(c, np + nghost, np)) return e;
  A.x = c->x; A.vel = c->defer_vel ? nullptr : c->vel; A.hh = c->hh; A.itype = c->itype; A.ireal = c->ireal; A.offset = c->scanout; A.count = c->ghostcount; A.cap = c->cap;
  // ensure_capacity reallocates scanout when it grows the arrays: redo the scan in that case only
  if (c->cap != cap_before) { if (int e = exclusive_scan(c, c->ghostcount, c->scanout, np)) return e; }
  A.offset = c->scanout;
  if (int e = wait_second_half(c)) return e;
  LAUNCH(c, (k_ghosts<NDIM, true>), nblocks(np, 256), 256, 0, A);
  c->ntotal = np + nghost;
  return 0;
}

// ---- set_linklist (src/linkND.f90:45-161): bounds, grid, counting sort, sorted SoA ----
template <int NDIM> int build_cells(nd_ctx *c) {
  const int nt = c->ntotal;
  const nd_options &o = c->o;
  if (!o.device_ghosts) {
    bool allle1 = !any_ghost_bound(c);
    if (allle1) {                                                                   // :70
      CU(cudaMemsetAsync(c->red, 0, sizeof(unsigned long long) * 16, c->stream));
      LAUNCH(c, k_max_h, std::min(nblocks(c->npart, 256), 1184), 256, 0, c->hh, c->npart, c->red);
      SMALL_D2H(c, c->h_red, c->red, sizeof(unsigned long long));
      CU(cudaStreamSynchronize(c->stream));
      c->hhmax = dkey_inv(c->h_red[0]);
    } else if (!o.device_ghosts) c->hhmax = o.hhmax;                                // set by the host's set_ghost_particles
  }
  c->dxcell = c->T->radkern * c->hhmax;                                             // :72
  if (!(c->dxcell > 0)) return set_err(c, ND_ERR_LINK, "link: max h <= 0");
  // :81-89 min/max of the particle distribution including ghosts
  unsigned long long init[16];
  for (int k = 0; k < 16; k++) init[k] = 0;
  for (int k = 0; k < 3; k++) init[k] = ~0ull;
  CU(cudaMemcpyAsync(c->red, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, (k_minmax_x<NDIM>), std::min(nblocks(nt, 256), 1184), 256, 0, c->x, nt, c->red);
  SMALL_D2H(c, c->h_red, c->red, sizeof(unsigned long long) * 6);
  CU(cudaStreamSynchronize(c->stream));
  long long nc = 1;
  c->ncellsx[0] = c->ncellsx[1] = c->ncellsx[2] = 1;
  for (int d = 0; d < NDIM; d++) {
    double xminpart = dkey_inv(c->h_red[d]) - 0.00001, xmaxpart = dkey_inv(c->h_red[3 + d]) + 0.00001;
    xminpart = xminpart - c->dxcell - 0.00001;
    xmaxpart = xmaxpart + c->dxcell + 0.00001;
    c->xminpart[d] = xminpart;
    const double q = (xmaxpart - xminpart) / c->dxcell;
    if (!(q < 2.0e9)) return set_err(c, ND_ERR_LINK, "link: too many cells");
    c->ncellsx[d] = (int)q + 1;                                                     // :93
    nc *= c->ncellsx[d];
  }
  if (nc * CELL_FX > 1500000000LL) return set_err(c, ND_ERR_LINK, "link: too many cells");
  c->ncells = (int)nc;
  const int nfine = c->ncells * CELL_FX;   // cellStart / cellCount are indexed by fine bin (cell * CELL_FX + bin along x)
  if (nfine + 2 > c->cellcap) {
    c->cellcap = nfine + nfine / 4 + 1024;
    if (int e = dev_alloc(c, &c->cellStart, (size_t)c->cellcap + 1)) return e;
    if (int e = dev_alloc(c, &c->cellCount, (size_t)c->cellcap + 1)) return e;
  }
  CU(cudaMemsetAsync(c->cellCount, 0, sizeof(int) * ((size_t)nfine + 1), c->stream));
  CellArgs CA;
  CA.x = c->x; CA.ntotal = nt; CA.dxcell = c->dxcell; CA.cellOfOrig = c->cellOfOrig; CA.cellCount = c->cellCount; CA.flags = c->flags;
  for (int d = 0; d < 3; d++) { CA.xminpart[d] = c->xminpart[d]; CA.ncellsx[d] = c->ncellsx[d]; }
  LAUNCH(c, (k_cell_index<NDIM>), nblocks(nt, 256), 256, 0, CA);
  if (int e = exclusive_scan(c, c->cellCount, c->cellStart, nfine)) return e;
  CU(cudaMemsetAsync(c->cellCount, 0, sizeof(int) * ((size_t)nfine + 1), c->stream));
  LAUNCH(c, k_cell_scatter, nblocks(nt, 256), 256, 0, c->cellOfOrig, nt, c->cellStart, c->cellCount, c->permtmp);
  LAUNCH(c, k_cell_order, nblocks((long long)c->ncells * 32, 256), 256, 0, c->cellStart, c->ncells, c->permtmp, c->perm);
  GatherArgs GA;
  GA.perm = c->perm; GA.cellOfOrig = c->cellOfOrig; GA.itype = c->itype; GA.ireal = c->ireal; GA.x = c->x; GA.vel = c->defer_vel ? nullptr : c->vel; GA.pmass = c->pmass; GA.hh = c->hh;
  GA.posh = c->posh; GA.vm = c->vm; GA.posm = c->posm; GA.p32 = c->p32; GA.typ = c->typ; GA.cellOf = c->cellOf; GA.inv = c->inv; GA.npart = c->npart; GA.ntotal = nt;
  for (int d = 0; d < 3; d++) GA.xminpart[d] = c->xminpart[d];
  GA.dxcell1 = 1.0 / c->dxcell; GA.hhmax1 = 1.0 / c->hhmax; GA.mixed = c->flags + 6;
  GA.dustfrac = c->dustfrac; GA.sdf = c->o.onef_dust ? c->sdf : nullptr;
  CU(cudaMemsetAsync(c->flags + 6, 0, sizeof(int), c->stream));
  if (int e = wait_second_half(c)) return e;
  LAUNCH(c, (k_gather_sorted<NDIM>), nblocks(nt, 256), 256, 0, GA);
  return 0;
}

template <int NDIM> int do_link(nd_ctx *c) {
  if (c->o.device_ghosts) { if (int e = make_ghosts<NDIM>(c)) return e; }
  if (int e = build_cells<NDIM>(c)) return e;
  const bool packed = c->has_comm && c->nccl;
  if (packed) {
    PackList L; L.add(CP_INT_NONZERO, c->flags + 1); L.add(CP_HOST, nullptr, 1., c->slab_too_narrow ? 1. : 0.);
    if (int e = comm_reduce_packed(c, L, 0, 2, 0)) return e;
  }
  if (int e = sync_flags(c)) return e;
  c->mixed_types = c->h_flags[6] != 0;
  double ef = (c->h_flags[1] != 0 || (c->has_comm && c->slab_too_narrow)) ? 1. : 0.;
  if (packed) ef = std::max(c->h_comm[32], c->h_comm[33]);
  else if (int e = comm_allreduce(c, &ef, 1, 0)) return e;   // every rank leaves together
  if (ef != 0.) {
    CU(cudaMemsetAsync(c->flags, 0, sizeof(int) * 16, c->stream));
    if (c->h_flags[1]) return set_err(c, c->h_flags[1], "link: particle outside the boundary / its slab");
    if (c->has_comm && c->slab_too_narrow)
      return set_err(c, ND_ERR_INVALID_ARG, "link: slab narrower than the interaction reach radkern*hhmax (twice that on a periodic ring of 2 ranks): repartition with fewer ranks");
    return set_err(c, ND_ERR_COMM, "link: another rank reported an error");
  }
  c->linked = true;
  return 0;
}

// ---- neighbour lists: capacity, chunking, overflow ----
constexpr int LIST_CHUNK = 32 << 20;   // targets per list build: bounds the list buffer to chunk * lmax * 4 bytes (12 GB at lmax = 96)

int ensure_lists(nd_ctx *c, int ntargets) {
  if (c->lmax == 0) {   // first guess: ~2.2x the mean neighbour number of the kernel at hfact = 1.2; grown on overflow
    const double v = c->ndim == 1 ? 2. : c->ndim == 2 ? 3.141592653589793 : 4.1887902047863905;
    double r = c->T->radkern * std::max(c->o.hfact, 1.0);
    double nn = v * (c->ndim == 1 ? r : c->ndim == 2 ? r * r : r * r * r);
    c->lmax = std::max(16, (int)(2.2 * nn) + 8);
    if (const char *ev = getenv("NDSPMHD_B200_LMAX0")) c->lmax = std::max(1, atoi(ev));   // test hook: start small, exercise the overflow retry
  }
  const size_t need = ((size_t)(ntargets + 31) / 32) * 32 * (size_t)c->lmax;
  if (need > c->nbrcap) {
    if (c->nbr) cudaFree(c->nbr);
    c->nbr = nullptr; c->nbrcap = 0;
    CU(cudaMalloc(&c->nbr, sizeof(unsigned) * need));
    c->nbrcap = need;
  }
  if (ntargets > c->lcntcap) {
    if (c->lcnt) cudaFree(c->lcnt);
    c->lcnt = nullptr; c->lcntcap = 0;
    CU(cudaMalloc(&c->lcnt, sizeof(int) * ((size_t)ntargets + 32)));
    c->lcntcap = ntargets + 32;
  }
  return 0;
}

// builds the lists of one chunk; on overflow grows lmax and repeats.  flags[5] is the overflow word.
template <int NDIM, int MODE> int build_lists(nd_ctx *c, const Grid &G, ListArgs LA, NbrLists &L) {
  for (int attempt = 0; attempt < 8; attempt++) {
    if (int e = ensure_lists(c, LA.ntargets)) return e;
    L.nbr = c->nbr; L.cnt = c->lcnt; L.lmax = c->lmax; L.overflow = c->flags + 5;
    L.split = (MODE == LIST_RATES && LA.drag && c->mixed_types && c->lmax < 65536) ? 1 : 0;   // drag runs: hydro pairs in front, drag pairs at the back
    if (c->mixed_types) LAUNCH(c, (build_lists_kernel<NDIM, MODE, true>), nblocks(LA.ntargets, 128), 128, 0, G, LA, L);
    else LAUNCH(c, (build_lists_kernel<NDIM, MODE, false>), nblocks(LA.ntargets, 128), 128, 0, G, LA, L);
    SMALL_D2H(c, c->h_flags + 20, c->flags + 5, sizeof(int));
    CU(cudaStreamSynchronize(c->stream));
    const int big = c->h_flags[20];
    if (big == 0) return 0;
    c->lmax = big + big / 4 + 8;
    c->list_overflows++;
    CU(cudaMemsetAsync(c->flags + 5, 0, sizeof(int), c->stream));
  }
  return set_err(c, ND_ERR_NEIGHBOUR_OVERFLOW, "neighbour list overflow");
}

// the first-class option tuple (FAST instantiations of the rates pair kernel): no run-time option tests, and the kernel also
// makes drho/dt (so the density rounds of a fused derivs can run LIGHT)
static bool fast_tuple(const nd_options &o) {
  return (o.idust != 2 || o.imhd == 0) && o.idust != 1 && !o.want_aux && o.iav == 2 && (o.iener == 0 || o.iener == 2) && o.ikernav == 3 && o.iresist == 0 && o.iavlim[0] != 3 &&
         o.iavlim[2] != 2;
}

template <int NDIM, bool FIRST> int launch_density_round(nd_ctx *c, DensityArgs A, int n) {
  Grid G = make_grid(c);
  for (int c0 = 0; c0 < n; c0 += LIST_CHUNK) {
    const int m = std::min(LIST_CHUNK, n - c0);
    ListArgs LA;
    LA.hh = c->hh; LA.targets = FIRST ? nullptr : c->list + c0; LA.s0 = c0; LA.ntargets = m; LA.numneigh = c->numneigh; LA.drag = 0;
    LA.pair_out_i = LA.pair_out_j = nullptr; LA.pair_count = nullptr; LA.pair_cap = 0;
    NbrLists L;
    if (int e = build_lists<NDIM, FIRST ? LIST_DENS_FIRST : LIST_DENS_PARTIAL>(c, G, LA, L)) return e;
    A.list = FIRST ? nullptr : c->list + c0; A.nlist = m; A.s0 = c0;
    A.sched = c->flags + 9;
    const bool aux = c->o.want_aux || c->o.onef_dust;
    // persistent blocks, one per SM (128 KB of shared-memory tables each); warps draw 32-target units from flags[9]
    auto kaux = density_round_kernel<NDIM, FIRST, true, false>;
    auto kfast = density_round_kernel<NDIM, FIRST, false, false>;
    auto klight = density_round_kernel<NDIM, FIRST, false, true>;
    CU(cudaFuncSetAttribute(kaux, cudaFuncAttributeMaxDynamicSharedMemorySize, DENS_SMEM_BYTES));
    CU(cudaFuncSetAttribute(kfast, cudaFuncAttributeMaxDynamicSharedMemorySize, DENS_SMEM_BYTES));
    CU(cudaFuncSetAttribute(klight, cudaFuncAttributeMaxDynamicSharedMemorySize, DENS_SMEM_BYTES));
    CU(cudaMemsetAsync(c->flags + 9, 0, sizeof(int), c->stream));
    const int grid = std::min(nblocks(m, DENS_BLOCK), c->num_sms);
    if (aux) LAUNCH(c, kaux, grid, DENS_BLOCK, DENS_SMEM_BYTES, G, A, L);
    else if (c->dens_light) LAUNCH(c, klight, std::min(nblocks(m, DENS_BLOCK_LIGHT), c->num_sms), DENS_BLOCK_LIGHT, DENS_SMEM_BYTES, G, A, L);
    else LAUNCH(c, kfast, grid, DENS_BLOCK, DENS_SMEM_BYTES, G, A, L);
  }
  return 0;
}

// ---- iterate_density (src/iterate_density.f90:43-360) ----
template <int NDIM> int do_iterate_density(nd_ctx *c, int resume) {
  const nd_options &o = c->o;
  const int np = c->nown;
  const long long nglobal = c->has_comm ? c->comm.nglobal : (long long)np;
  const int itsdensitymax = (o.ikernav == 3 && o.ihvar != 0) ? o.maxdensits : 0;     // :77-81
  if (!resume) {
    c->itsdensity = 0; c->ncalctotal = 0; c->ncalc = np; c->ncalc_g = nglobal; c->redolink = false; c->nrelink = 0;
    CU(cudaMemcpyAsync(c->hhin, c->hh, sizeof(double) * np, cudaMemcpyDeviceToDevice, c->stream));   // :98
    CU(cudaMemsetAsync(c->flags, 0, sizeof(int) * 16, c->stream));
    LAUNCH(c, k_check_h, nblocks(np, 256), 256, 0, c->hh, np, c->flags);
    CU(cudaMemsetAsync(c->dhdt, 0, sizeof(double) * c->ntotal, c->stream));          // :90-97
    CU(cudaMemsetAsync(c->numneigh, 0, sizeof(int) * c->ntotal, c->stream));
  }
  while (c->ncalc_g > 0 && c->itsdensity <= itsdensitymax) {                         // :119
    const bool first = (c->ncalc_g == nglobal);                                      // :131 `density` when every particle is (re)done
    if (c->redolink) {                                                               // :122-126
      // host-made ghosts: the caller re-runs set_ghost_particles, calls update_ghosts() and comes back with resume=1
      if (any_ghost_bound(c) && !o.device_ghosts && !resume) { fill_link_scalars(c); c->sc.itsdensity = c->itsdensity; return ND_NEED_RELINK; }
      resume = 0;
      // remember which ROWS are still to be done (slots change with the re-sort), rebuild halos + ghosts + grid, map back
      const int n = c->ncalc;
      if (!first) LAUNCH(c, k_list_rows, nblocks(n, 256), 256, 0, c->list, n, c->perm, c->redo);   // redo[] doubles as row scratch
      if (int e = do_link<NDIM>(c)) return e;
      if (!first) LAUNCH(c, k_remap_list, nblocks(n, 256), 256, 0, c->list, n, c->redo, c->inv);
      c->nrelink++;
      c->redolink = false;
    }
    c->itsdensity++;
    DensityArgs A;
    A.hh = c->hh; A.hhin = c->hhin; A.rho = c->rho; A.gradh = c->gradh; A.drhodt = c->drhodt; A.dhdt = c->dhdt; A.numneigh = c->numneigh;
    A.rhoalt = c->rhoalt; A.gradhn = c->gradhn; A.gradsoft = c->gradsoft; A.gradgradh = c->gradgradh;
    A.sdf = o.onef_dust ? c->sdf : nullptr; A.rhogas = c->rhogas; A.rhodust = c->rhodust;
    A.list = c->list; A.nlist = c->ncalc; A.s0 = 0; A.redo = c->redo; A.flags = c->flags;
    A.itsdensity = c->itsdensity; A.itsdensitymax = itsdensitymax; A.hfact = o.hfact; A.psep = o.psep; A.tolh = o.tolh; A.hhmax = c->hhmax;
    CU(cudaMemsetAsync(c->redo, 0, sizeof(int) * c->ntotal, c->stream));
    if (first) {                                                                     // :131-132 symmetric `density`
      if (c->itsdensity > 1) {   // the neighbour count of `density` also looks at h_j (:189-190): refresh the sources' 1/h
        if (c->has_comm) { if (int e = halo_exchange_density(c, 0)) return e; }
        LAUNCH(c, k_refresh_h, nblocks(c->ntotal, 256), 256, 0, c->perm, c->ireal, c->hh, c->posh, c->p32, 1.0 / c->hhmax, c->npart, c->ntotal);
      }
      if (int e = launch_density_round<NDIM, true>(c, A, c->ntotal)) return e;
    } else {                                                                         // :133-134 `density_partial`
      if (int e = launch_density_round<NDIM, false>(c, A, c->ncalc)) return e;
    }
    c->ncalctotal += c->ncalc;                                                       // :154
    if (int e = exclusive_scan(c, c->redo, c->scanout, c->ntotal)) return e;
    LAUNCH(c, k_compact, nblocks(c->ntotal, 256), 256, 0, c->redo, c->scanout, c->ntotal, c->list);
    SMALL_D2H(c, &c->h_flags[16], c->scanout + c->ntotal, sizeof(int));
    const bool packed = c->has_comm && c->nccl;
    if (packed) {   // {unconverged count, relink flag, error flag} summed over the ranks without a host round trip in between
      PackList L; L.add(CP_INT, c->scanout + c->ntotal); L.add(CP_INT_NONZERO, c->flags); L.add(CP_INT_NONZERO, c->flags + 1);
      if (int e = comm_reduce_packed(c, L, 0, 0, 3)) return e;
    }
    if (int e = sync_flags(c)) return e;
    c->ncalc = c->h_flags[16];
    double v[3] = {(double)c->ncalc, (double)(c->h_flags[0] != 0), (double)(c->h_flags[1] != 0)};
    if (packed) { v[0] = c->h_comm[32]; v[1] = c->h_comm[33]; v[2] = c->h_comm[34]; }
    else if (int e = comm_allreduce(c, v, 3, 2)) return e;
    c->ncalc_g = (long long)(v[0] + 0.5);
    if (v[2] != 0.) {
      const int code = c->h_flags[1] ? c->h_flags[1] : ND_ERR_COMM;
      return set_err(c, code, code == ND_ERR_RHO_NONPOSITIVE ? "error: rho <= 0 in iterate_density" : code == ND_ERR_COMM ? "iterate_density: another rank reported an error" : "error: h <= 0 in density call");
    }
    c->redolink = v[1] != 0.;
    CU(cudaMemsetAsync(c->flags, 0, sizeof(int) * 4, c->stream));
  }
  if (c->itsdensity > itsdensitymax && itsdensitymax > 0) return set_err(c, ND_ERR_DENSITY_NOT_CONVERGED, "ERROR: DENSITY NOT CONVERGED");   // :349-351
  // halo rows take the owners' converged values, then :310-344 copies to fixed particles and ghosts
  if (c->has_comm) { if (int e = halo_exchange_density(c, 1)) return e; }
  CopyArgs CA;
  CA.rho = c->rho; CA.rhoalt = c->rhoalt; CA.drhodt = c->drhodt; CA.dhdt = c->dhdt; CA.hh = c->hh; CA.gradh = c->gradh; CA.gradhn = c->gradhn;
  CA.gradsoft = c->gradsoft; CA.itype = c->itype; CA.ireal = c->ireal; CA.npart = c->npart; CA.ntotal = c->ntotal; CA.aux = o.want_aux != 0;
  if (any_fixed_bound(c)) { CopyArgs CF = CA; CF.npart = np; LAUNCH(c, k_copy_fixed_density, nblocks(np, 256), 256, 0, CF); }
  if (any_ghost_bound(c)) LAUNCH(c, k_copy_ghost_density, nblocks(c->ntotal - c->npart, 256), 256, 0, CA);
  c->density_done = true;
  return 0;
}

// ---- get_curl (src/get_curl.f90:64-287) on the linked, density-converged state: curl (and grad) of `bvec` (device, original order,
//      3 doubles a row, rows [0,npart)) into curlB / gradB (device, original order); with `alpha` the Tricco & Price switch as well ----
template <int NDIM> int do_get_curl(nd_ctx *c, int icurltype, const double *bvec, double *curlB, double *gradB, double *alpha) {
  const int nt = c->ntotal;
  CurlGatherArgs GA;
  GA.perm = c->perm; GA.ireal = c->ireal; GA.hh = c->hh; GA.pmass = c->pmass; GA.rho = c->rho; GA.gradh = c->gradh; GA.bvec = bvec;
  GA.posh = c->posh; GA.vm = c->vm; GA.gal = c->gal; GA.bsorted = c->bpsi; GA.p32 = c->p32; GA.srho = c->srho; GA.hhmax1 = 1.0 / c->hhmax;
  GA.npart = c->npart; GA.ntotal = nt;
  LAUNCH(c, k_curl_gather, nblocks(nt, 256), 256, 0, GA);
  Grid G = make_grid(c);
  for (int c0 = 0; c0 < nt; c0 += LIST_CHUNK) {
    const int m = std::min(LIST_CHUNK, nt - c0);
    ListArgs LA;
    LA.hh = c->hh; LA.targets = nullptr; LA.s0 = c0; LA.ntargets = m; LA.numneigh = nullptr; LA.drag = 0;   // type rule of get_curl.f90:161-167
    LA.pair_out_i = LA.pair_out_j = nullptr; LA.pair_count = nullptr; LA.pair_cap = 0;
    NbrLists L;
    if (int e = build_lists<NDIM, LIST_RATES>(c, G, LA, L)) return e;
    CurlArgs A;
    A.bvec = c->bpsi; A.srho = c->srho; A.gal = c->gal; A.curlB = curlB; A.gradB = gradB; A.alpha = alpha; A.hh = c->hh; A.icurltype = icurltype;
    A.weight = 1.0 / std::pow(c->o.hfact, NDIM); A.s0 = c0; A.ntargets = m; A.targets = nullptr;
    LAUNCH(c, (curl_pair_kernel<NDIM>), nblocks(m, 128), 128, 0, G, A, L);
  }
  return 0;
}

int do_cons2prim(nd_ctx *c) {
  const nd_options &o = c->o;
  C2PArgs A;
  A.rho = c->rho; A.en = c->en; A.Bevol = c->Bevol; A.vel = c->vel; A.itype = c->itype; A.ireal = c->ireal;
  A.dens = c->dens; A.uu = c->uu; A.pr = c->pr; A.spsound = c->spsound; A.Bfield = c->Bfield;
  A.pmass = c->pmass; A.rho_w = c->rho; A.rhoalt = c->rhoalt; A.hh = c->hh; A.en_w = c->en; A.Bevol_w = c->Bevol; A.alpha = c->alpha; A.psi = c->psi;
  A.gradh = c->gradh; A.gradhn = c->gradhn; A.gradsoft = c->gradsoft; A.gradgradh = c->gradgradh;
  A.npart = c->npart; A.ntotal = c->ntotal; A.imhd = o.imhd; A.iener = o.iener; A.gamma = o.gamma; A.polyk = o.polyk; A.aux = o.want_aux != 0;
  A.dustevol = o.onef_dust ? c->dustevol : nullptr; A.dustfrac = c->dustfrac; A.rhogas = c->rhogas; A.rhodust = c->rhodust;
  LAUNCH(c, k_c2p, nblocks(c->npart, 256), 256, 0, A);
  if (o.imhd != 0 && o.iavlim[2] == 2 && c->has_comm)
    return set_err(c, ND_ERR_UNSUPPORTED_OPTION, "iavlim(3) = 2 (resistivity switch) is not available on slab-decomposed contexts: halo rows would need the owners' alpha_B");
  if (o.imhd != 0 && o.iavlim[2] == 2) {
    // conservative2primitive.f90:299-311: J and grad B by the differenced curl operator, then alpha_B = min(h |grad B| / |B|, 1).  It sits
    // between the B-field assignment and the copies to fixed particles and ghosts, like the reference's.  The old alpha_B is kept for
    // ghost rows that do not get copy_particle (k_rates_gather).
    LAUNCH(c, k_take_column, nblocks(c->npart, 256), 256, 0, c->alpha, 3, 2, c->alphaB_in, c->npart);
    if (int e = DISPATCH_NDIM(c, do_get_curl<1>(c, 1, c->Bfield, c->curlB, nullptr, c->alpha), do_get_curl<2>(c, 1, c->Bfield, c->curlB, nullptr, c->alpha),
                              do_get_curl<3>(c, 1, c->Bfield, c->curlB, nullptr, c->alpha))) return e;
  }
  if (any_fixed_bound(c)) LAUNCH(c, k_c2p_fixed, nblocks(c->npart, 256), 256, 0, A);
  if (any_ghost_bound(c)) LAUNCH(c, k_c2p_ghost, nblocks(c->ntotal - c->npart, 256), 256, 0, A);
  c->prim_done = true;
  return 0;
}

RatesOpts make_rates_opts(const nd_ctx *c) {
  const nd_options &o = c->o;
  RatesOpts O;
  O.iener = o.iener; O.iav = o.iav; O.imhd = o.imhd; O.idivbzero = o.idivbzero; O.iresist = o.iresist; O.idust = o.idust; O.idrag_nature = o.idrag_nature;
  O.ikernav = o.ikernav; O.iavlim0 = o.iavlim[0]; O.iavlim1 = o.iavlim[1]; O.iavlim2 = o.iavlim[2]; O.nsubsteps_divB = o.nsubsteps_divB;
  O.beta = o.beta; O.pext = o.pext; O.etamhd = o.etamhd; O.Kdrag = o.Kdrag; O.stressmax = 0.; O.gamma = o.gamma;
  O.alphamin = o.alphamin; O.alphaumin = o.alphaumin; O.alphaBmin = o.alphaBmin; O.avdecayconst = o.avdecayconst; O.avfact = o.avfact; O.psidecayfact = o.psidecayfact;
  for (int d = 0; d < 3; d++) O.Bconst[d] = o.Bconst[d];
  return O;
}

enum { RED_DTC = 0, RED_VSIG = 1, RED_DTAV = 2, RED_TS = 3, RED_HCS = 4, RED_FH = 5, RED_DTF = 6, RED_STRESS = 7, RED_NPAIRS = 8, RED_NTRIPS = 9 /* plain sums, not keys */ };

template <int NDIM, bool MHD, bool DRAG, int FAST, bool ONEF> int launch_rates_pair(nd_ctx *c, const RatesIn &I, const RatesOpts &O, const RatesSums &S, const RatesRed &R, int *pi, int *pj,
                                                                          unsigned long long *pc, long long cap, const int *targets, int ntargets) {
  Grid G = make_grid(c);
  const int n = targets ? ntargets : c->ntotal;
  for (int c0 = 0; c0 < n; c0 += LIST_CHUNK) {
    const int m = std::min(LIST_CHUNK, n - c0);
    ListArgs LA;
    LA.hh = c->hh; LA.targets = targets ? targets + c0 : nullptr; LA.s0 = c0; LA.ntargets = m; LA.numneigh = nullptr; LA.drag = (DRAG && O.idrag_nature > 0) ? 1 : 0;
    LA.pair_out_i = pi; LA.pair_out_j = pj; LA.pair_count = pc; LA.pair_cap = cap;
    NbrLists L;
    if (int e = build_lists<NDIM, LIST_RATES>(c, G, LA, L)) return e;
    LAUNCH(c, k_sum_counts, std::min(nblocks(m, 256), 1184), 256, 0, L.cnt, m, L.split, c->red + RED_NPAIRS, c->red + RED_NTRIPS);
    // persistent blocks: one per resident slot (the 64 KB shared-memory table is loaded once per block)
    static int resident = 0, carveout = 100;
    auto kfn = rates_pair_kernel<NDIM, MHD, DRAG, FAST, ONEF>;
    CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, RATES_SMEM_BYTES));   // attributes are per device
    if (!resident) {
      CU(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      int per_sm = 0;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, RATES_BLOCK, RATES_SMEM_BYTES));
      if (per_sm < 1) return set_err(c, ND_ERR_CUDA, "rates_pair_kernel does not fit on an SM");
      resident = per_sm * c->num_sms;
      // leave the rest of the 256 KB L1/shared array to L1: the neighbour gather lives on its hit rate
      carveout = std::min(100, (int)((per_sm * (size_t)(RATES_SMEM_BYTES + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)));
      if (getenv("NDSPMHD_B200_DEBUG")) fprintf(stderr, "rates_pair_kernel: %d blocks/SM x %d SMs, %d B dynamic smem, carveout %d%%\n", per_sm, c->num_sms, RATES_SMEM_BYTES, carveout);
    }
    CU(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
    CU(cudaMemsetAsync(c->flags + 8, 0, sizeof(int), c->stream));
    if (c0 == 0) CU(cudaEventRecord(c->ev_pair[0], c->stream));   // the first chunk's launch is the one timed (the only one below 32 Mi rows)
    LAUNCH(c, (rates_pair_kernel<NDIM, MHD, DRAG, FAST, ONEF>), std::min(nblocks(m, RATES_BLOCK), resident), RATES_BLOCK, RATES_SMEM_BYTES, G, I, O, S, R, L, c0, m, targets ? targets + c0 : nullptr);
    if (c0 == 0) CU(cudaEventRecord(c->ev_pair[1], c->stream));
  }
  return 0;
}

// ---- get_rates (src/ratesND_mhd.f90:29-979) ----
template <int NDIM> int do_get_rates(nd_ctx *c, int *pi, int *pj, unsigned long long *pc, long long cap) {
  const nd_options &o = c->o;
  const int nt = c->ntotal, np = c->nown;
  // reduction keys: minima start at +huge (key of DBL_MAX), maxima at 0
  unsigned long long init[16];
  for (int k = 0; k < 16; k++) init[k] = 0x8000000000000000ull;                       // key(+0.0)
  union { double d; unsigned long long u; } cv;
  auto keyof = [&](double v) { cv.d = v; return cv.u | 0x8000000000000000ull; };     // v >= 0
  init[RED_NPAIRS] = init[RED_NTRIPS] = 0ull;
  init[RED_DTC] = keyof(1.e6); init[RED_DTAV] = keyof(DBL_MAX); init[RED_TS] = keyof(DBL_MAX); init[RED_DTF] = keyof(DBL_MAX);
  CU(cudaMemcpyAsync(c->red, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(c->fmean, 0, sizeof(double) * 4, c->stream));
  CU(cudaMemsetAsync(c->flags, 0, sizeof(int) * 16, c->stream));
  RGatherArgs GA;
  GA.perm = c->perm; GA.inv = c->inv; GA.ireal = c->ireal; GA.hh = c->hh; GA.pmass = c->pmass; GA.rho = c->rho; GA.pr = c->pr; GA.spsound = c->spsound; GA.uu = c->uu;
  GA.gradh = c->gradh; GA.alpha = c->alpha; GA.psi = c->psi; GA.Bfield = c->Bfield;
  GA.p32 = c->p32; GA.hhmax1 = 1.0 / c->hhmax;
  GA.dustfrac = c->dustfrac; GA.deltav = c->deltav; GA.rhogas = c->rhogas; GA.rhodust = c->rhodust;
  GA.dusta = o.onef_dust ? c->dusta : nullptr; GA.dustb = c->dustb; GA.use_smoothed_rhodust = o.use_smoothed_rhodust;
  {
    bool all3 = true;
    for (int d = 0; d < c->ndim; d++) if (o.ibound[d] != 3) all3 = false;
    GA.alphaB_ghost = (o.imhd != 0 && o.iavlim[2] == 2 && any_ghost_bound(c) && !all3) ? c->alphaB_in : nullptr;
  }
  GA.posh = c->posh; GA.vm = c->vm; GA.bpsi = c->bpsi; GA.thermo = c->thermo; GA.gal = c->gal; GA.npart = c->npart; GA.ntotal = nt; GA.imhd = o.imhd;
  GA.stress_key = c->red + RED_STRESS; GA.imagforce = o.imagforce; GA.srho = c->srho; GA.pext = o.pext; GA.err = c->flags + 1;
  GA.Bconstmax = std::max(o.Bconst[0], std::max(o.Bconst[1], o.Bconst[2]));
  LAUNCH(c, k_rates_gather, nblocks(nt, 256), 256, 0, GA);
  RatesOpts O = make_rates_opts(c);
  if (o.imhd != 0) {   // stressmax feeds the pair kernel by value: one 8-byte D2H
    if (c->has_comm && c->nccl) {
      PackList L; L.add(CP_KEY, c->red + RED_STRESS);
      if (int e = comm_reduce_packed(c, L, 0, 1, 0)) return e;
      CU(cudaStreamSynchronize(c->stream));
      O.stressmax = c->h_comm[32];
    } else {
      SMALL_D2H(c, c->h_red, c->red + RED_STRESS, sizeof(unsigned long long));
      CU(cudaStreamSynchronize(c->stream));
      O.stressmax = dkey_inv(c->h_red[0]);
      if (int e = comm_allreduce(c, &O.stressmax, 1, 0)) return e;
    }
  }
  RatesIn I; I.bpsi = c->bpsi; I.thermo = c->thermo; I.gal = c->gal; I.srho = c->srho; I.dusta = c->dusta; I.dustb = c->dustb;
  RatesSums S; S.F = c->sF; S.dB = c->sdB; S.C = c->sC; S.P = c->sP; S.V = c->sV; S.D = c->sD;
  RatesRed R;
  R.dtcourant_min = c->red + RED_DTC; R.vsigmax_max = c->red + RED_VSIG; R.dtav_min = c->red + RED_DTAV; R.ts_min = c->red + RED_TS;
  R.h_on_csts_max = c->red + RED_HCS; R.fhmax_max = c->red + RED_FH; R.dtforce_min = c->red + RED_DTF; R.fmean = c->fmean;
  R.nclumped = c->flags + 4; R.err = c->flags + 1; R.sched = c->flags + 8;
  CU(cudaEventRecord(c->ev[3], c->stream));
  const bool mhd = o.imhd != 0, drag = (o.idust == 2);
  // first-class tuple without run-time option tests (and without the dead graddivv "curl v" sums: want_aux = 0)
  const bool fast = fast_tuple(o);
  auto pair = [&](const int *targets, int ntargets) -> int {
    if (o.idust == 1 && mhd) return launch_rates_pair<NDIM, true, false, 0, true>(c, I, O, S, R, pi, pj, pc, cap, targets, ntargets);
    if (o.idust == 1) return launch_rates_pair<NDIM, false, false, 0, true>(c, I, O, S, R, pi, pj, pc, cap, targets, ntargets);
    if (!mhd && drag && fast && o.iener != 0) return launch_rates_pair<NDIM, false, true, 2, false>(c, I, O, S, R, pi, pj, pc, cap, targets, ntargets);   // two-fluid dust, hydro
    if (!mhd && drag && fast) return launch_rates_pair<NDIM, false, true, 1, false>(c, I, O, S, R, pi, pj, pc, cap, targets, ntargets);
    if (mhd && fast && o.iener != 0) return launch_rates_pair<NDIM, true, false, 2, false>(c, I, O, S, R, pi, pj, pc, cap, targets, ntargets);
    if (mhd && fast) return launch_rates_pair<NDIM, true, false, 1, false>(c, I, O, S, R, pi, pj, pc, cap, targets, ntargets);
    if (!mhd && fast && o.iener != 0) return launch_rates_pair<NDIM, false, false, 2, false>(c, I, O, S, R, pi, pj, pc, cap, targets, ntargets);
    if (!mhd && fast) return launch_rates_pair<NDIM, false, false, 1, false>(c, I, O, S, R, pi, pj, pc, cap, targets, ntargets);
    if (mhd && !drag) return launch_rates_pair<NDIM, true, false, 0, false>(c, I, O, S, R, pi, pj, pc, cap, targets, ntargets);
    if (!mhd && !drag) return launch_rates_pair<NDIM, false, false, 0, false>(c, I, O, S, R, pi, pj, pc, cap, targets, ntargets);
    if (!mhd && drag) return launch_rates_pair<NDIM, false, true, 0, false>(c, I, O, S, R, pi, pj, pc, cap, targets, ntargets);
    return launch_rates_pair<NDIM, true, true, 0, false>(c, I, O, S, R, pi, pj, pc, cap, targets, ntargets);
  };
  FinalArgs FA;
  FA.perm = c->perm; FA.typ = c->typ; FA.posh = c->posh; FA.vm = c->vm; FA.bpsi = c->bpsi; FA.thermo = c->thermo; FA.gal = c->gal; FA.S = S; FA.O = O;
  FA.drhodt_in = c->drhodt; FA.Bevol = c->Bevol; FA.dens = c->dens; FA.hh = c->hh; FA.rho = c->rho; FA.pr = c->pr; FA.vsigmax_key = c->red + RED_VSIG;
  FA.force = c->force; FA.dudt = c->dudt; FA.dendt = c->dendt; FA.dBevoldt = c->dBevoldt; FA.daldt = c->daldt; FA.dpsidt = c->dpsidt; FA.gradpsi = c->gradpsi;
  FA.divB = c->divB; FA.curlB = c->curlB; FA.graddivv = c->graddivv; FA.del2u = c->del2u; FA.drhodt = c->drhodt; FA.dhdt = c->dhdt;
  FA.R = R; FA.npart = np; FA.ntotal = nt; FA.inv = c->inv; FA.row0 = 0; FA.row1 = np;
  FA.vel = c->vel; FA.pmass = c->pmass; FA.spsound = c->spsound; FA.uu = c->uu; FA.alpha = c->alpha; FA.psi = c->psi; FA.Bfield = c->Bfield; FA.itype = c->itype; FA.drho_from_pairs = (c->drho_pairs && fast) ? 1 : 0; FA.ndim = NDIM;
  FA.dusta = o.onef_dust ? c->dusta : nullptr; FA.dustb = c->dustb; FA.fineStart = c->cellStart; FA.cellOf = c->cellOf;
  FA.ddustevoldt = c->ddustevoldt; FA.ddeltavdt = c->ddeltavdt;
  {
    const size_t need = (size_t)nblocks(np, 256) * 6 + 6;
    if (c->finalpartcap < need) {
      if (c->finalpart) cudaFree(c->finalpart);
      c->finalpart = nullptr; c->finalpartcap = 0;
      CU(cudaMalloc(&c->finalpart, sizeof(double) * (need + need / 8)));
      c->finalpartcap = need + need / 8;
    }
    FA.partial = c->finalpart;
  }
  const int nchunk = (c->rate_chunks > 1 && !pi) ? c->rate_chunks : 1;
  c->rate_chunks_used = nchunk;
  // vsigmax feeds dpsidt (:518-520, :902): with slabs all ranks need the global maximum of the pair kernels before it is made
  auto reduce_vsigmax = [&]() -> int {
    if (c->has_comm && c->nccl) {   // consumed on the device (k_rates_final / k_rates_dpsidt): no host round trip at all
      PackList L; L.add(CP_KEY, c->red + RED_VSIG);
      LAUNCH(c, k_comm_pack, 1, 32, 0, L.P, c->d_comm + 8);
      NCCLCHK(c->nccl_api->AllReduce(c->d_comm + 8, c->d_comm + 8, 1, ND_NCCL_FLOAT64, ND_NCCL_MAX, c->nccl, c->stream));
      c->n_allreduce++;
      LAUNCH(c, k_double_to_key, 1, 1, 0, c->d_comm + 8, c->red + RED_VSIG);
    } else if (c->has_comm) {
      SMALL_D2H(c, c->h_red, c->red + RED_VSIG, sizeof(unsigned long long));
      CU(cudaStreamSynchronize(c->stream));
      double vs = dkey_inv(c->h_red[0]);
      if (int e = comm_allreduce(c, &vs, 1, 0)) return e;
      union { double d; unsigned long long u; } kv; kv.d = vs;
      c->h_red[0] = kv.u | 0x8000000000000000ull;   // key of a non-negative double
      CU(cudaMemcpyAsync(c->red + RED_VSIG, c->h_red, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
    }
    return 0;
  };
  if (nchunk == 1) {
    if (int e = pair(nullptr, 0)) return e;
    CU(cudaEventRecord(c->ev[4], c->stream));
    if (int e = reduce_vsigmax()) return e;
    const int fgrid = std::min(nblocks(np, 256), 8 * c->num_sms);   // persistent: 1.87 ms against 2.15 ms with one row a thread at 16.8 M rows
    LAUNCH(c, k_rates_final, fgrid, 256, 0, FA);
    LAUNCH(c, k_final_reduce, 1, 1024, 0, c->finalpart, fgrid, R, o.onef_dust ? 1 : 0);
  } else {
    // Row chunks: chunk q gathers and finalises the targets whose ORIGINAL row lies in [q*rows, (q+1)*rows); its output rows are
    // then complete (except dpsidt) and contiguous in the caller's arrays, so on_rates_chunk can start their download while the
    // next chunk's pair kernel runs.  Results are those of the single launch bit for bit: every target's sums are its own.
    if (c->rlistcap < (size_t)nt) {
      if (c->rlist) cudaFree(c->rlist);
      c->rlist = nullptr; c->rlistcap = 0;
      CU(cudaMalloc(&c->rlist, sizeof(int) * ((size_t)c->cap + 1)));
      c->rlistcap = (size_t)c->cap + 1;
    }
    const int rows = (np + nchunk - 1) / nchunk;
    for (int q = 0; q < nchunk; q++) {
      const int r0 = q * rows, r1 = std::min(np, r0 + rows);
      if (r1 <= r0) continue;
      LAUNCH(c, k_chunk_flags, nblocks(nt, 256), 256, 0, c->perm, nt, r0, r1, c->redo);
      if (int e = exclusive_scan(c, c->redo, c->scanout, nt)) return e;
      LAUNCH(c, k_compact, nblocks(nt, 256), 256, 0, c->redo, c->scanout, nt, c->rlist);
      const int m = r1 - r0;                         // every row below nown has exactly one slot
      if (int e = pair(c->rlist, m)) return e;
      FA.row0 = r0; FA.row1 = r1;
      const int fgrid = std::min(nblocks(m, 256), 8 * c->num_sms);
      LAUNCH(c, k_rates_final, fgrid, 256, 0, FA);
      LAUNCH(c, k_final_reduce, 1, 1024, 0, c->finalpart, fgrid, R, o.onef_dust ? 1 : 0);
      if (c->on_rates_chunk) { if (int e = c->on_rates_chunk(q, r0, r1)) return e; }
    }
    CU(cudaEventRecord(c->ev[4], c->stream));
    if (int e = reduce_vsigmax()) return e;
    LAUNCH(c, k_rates_dpsidt, nblocks(np, 256), 256, 0, c->divB, c->psi, c->hh, c->itype, c->red + RED_VSIG, o.psidecayfact, o.imhd, o.idivbzero, c->dpsidt, np);
  }
  ZeroArgs ZA;
  ZA.force = c->force; ZA.dudt = c->dudt; ZA.dendt = c->dendt; ZA.dBevoldt = c->dBevoldt; ZA.daldt = c->daldt; ZA.dpsidt = c->dpsidt; ZA.gradpsi = c->gradpsi;
  ZA.divB = c->divB; ZA.curlB = c->curlB; ZA.graddivv = c->graddivv; ZA.del2u = c->del2u; ZA.drhodt = c->drhodt; ZA.dhdt = c->dhdt; ZA.npart = np; ZA.ntotal = nt;
  LAUNCH(c, k_rates_zero_ghosts, nblocks(nt - np, 256), 256, 0, ZA);
  CU(cudaEventRecord(c->ev[5], c->stream));
  // scalars back to the host (module timestep)
  SMALL_D2H(c, c->h_red, c->red, sizeof(unsigned long long) * 16);
  SMALL_D2H(c, c->h_fmean, c->fmean, sizeof(double) * 4);
  const bool packed = c->has_comm && c->nccl;
  if (packed) {   // {error flag, maxima, -minima} and the sums: two all-reduces queued behind the kernels, one synchronise for everything
    PackList L;
    L.add(CP_INT, c->flags + 1); L.add(CP_KEY, c->red + RED_VSIG); L.add(CP_KEY, c->red + RED_HCS); L.add(CP_KEY, c->red + RED_FH);
    L.add(CP_KEY, c->red + RED_DTC, -1.); L.add(CP_KEY, c->red + RED_DTAV, -1.); L.add(CP_KEY, c->red + RED_TS, -1.); L.add(CP_KEY, c->red + RED_DTF, -1.);
    L.add(CP_DOUBLE, c->fmean); L.add(CP_DOUBLE, c->fmean + 1); L.add(CP_DOUBLE, c->fmean + 2); L.add(CP_INT, c->flags + 4); L.add(CP_U64, c->red + RED_NPAIRS);
    if (int e3 = comm_reduce_packed(c, L, 0, 8, 5)) return e3;
  }
  if (int e2 = sync_flags(c)) return e2;
  nd_scalars &s = c->sc;
  s.dtcourant = dkey_inv(c->h_red[RED_DTC]);
  s.vsigmax = dkey_inv(c->h_red[RED_VSIG]);
  s.dtav = dkey_inv(c->h_red[RED_DTAV]);
  s.ts_min = dkey_inv(c->h_red[RED_TS]);
  s.h_on_csts_max = dkey_inv(c->h_red[RED_HCS]);
  s.fhmax = dkey_inv(c->h_red[RED_FH]);
  double ef = c->h_flags[1];
  if (c->has_comm) {   // two all-reduces: {error flag, maxima, minima} and the sums
    double mx[4] = {ef, s.vsigmax, s.h_on_csts_max, s.fhmax}, mn[4] = {s.dtcourant, s.dtav, s.ts_min, dkey_inv(c->h_red[RED_DTF])};
    double sm[5] = {c->h_fmean[0], c->h_fmean[1], c->h_fmean[2], (double)c->h_flags[4], (double)c->h_red[RED_NPAIRS] /* < 2^53 */};
    if (packed) {
      const double *r = c->h_comm + 32;
      for (int k = 0; k < 4; k++) { mx[k] = r[k]; mn[k] = -r[4 + k]; }
      for (int k = 0; k < 5; k++) sm[k] = r[8 + k];
    } else {
    if (int e3 = comm_allreduce_maxmin(c, mx, 4, mn, 4)) return e3;
    if (int e3 = comm_allreduce(c, sm, 5, 2)) return e3;
    }
    c->h_red[RED_NPAIRS] = (unsigned long long)(sm[4] + 0.5);
    ef = mx[0]; s.vsigmax = mx[1]; s.h_on_csts_max = mx[2]; s.fhmax = mx[3];
    s.dtcourant = mn[0]; s.dtav = mn[1]; s.ts_min = mn[2];
    c->h_fmean[0] = sm[0]; c->h_fmean[1] = sm[1]; c->h_fmean[2] = sm[2]; c->h_flags[4] = (int)(sm[3] + 0.5);
    union { double d; unsigned long long u; } kv; kv.d = mn[3]; c->h_red[RED_DTF] = kv.u | 0x8000000000000000ull;
  }
  if (ef != 0.) {
    const int code = c->h_flags[1] ? c->h_flags[1] : ND_ERR_COMM;
    if (code == ND_ERR_COMM) return set_err(c, code, "rates: another rank reported an error");
    return set_err(c, code, code == ND_ERR_VSIG_DET ? "rates: vsig det < 0" : code == ND_ERR_H_NONPOSITIVE ? "rates: h <= 0" : "rates: dx = 0 (coincident particles of the same type)");
  }
  s.stressmax = O.stressmax;
  s.vsig2max = (o.imhd != 0 && o.idivbzero >= 2) ? s.vsigmax * s.vsigmax : 0.;
  s.dtvisc = DBL_MAX;
  s.dtforce = dkey_inv(c->h_red[RED_DTF]);
  if (s.fhmax > 0.) s.dtforce = std::min(s.dtforce, std::sqrt(1. / s.fhmax));        // :938-943
  s.dtdrag = DBL_MAX;
  if (o.idust == 2 && o.idrag_nature != 0 && (o.Kdrag > 0. || o.idrag_nature > 1)) s.dtdrag = std::min(s.dtdrag, s.ts_min);   // :543-547
  else if (o.idust == 1) { s.dtdrag = std::min(s.dtdrag, s.ts_min); s.ts_min = DBL_MAX; }   // :561: the key carried min tstop; module ts_min is two-fluid only
  for (int k = 0; k < 3; k++) s.fmean[k] = c->h_fmean[k];
  s.nclumped = c->h_flags[4];
  s.npairs_rates = (long long)c->h_red[RED_NPAIRS];
  s.ntrips_rates = (long long)c->h_red[RED_NTRIPS];
  c->rates_done = true;
  return 0;
}

void fill_link_scalars(nd_ctx *c) {
  nd_scalars &s = c->sc;
  s.hhmax = c->hhmax; s.dxcell = c->dxcell; s.ntotal = c->ntotal; s.ncells = c->ncells;
  s.lmax = c->lmax; s.list_overflows = c->list_overflows; s.rate_chunks = c->rate_chunks_used;
  for (int d = 0; d < 3; d++) s.ncellsx[d] = c->ncellsx[d];
}

int fill_density_scalars(nd_ctx *c) {
  nd_scalars &s = c->sc;
  s.itsdensity = c->itsdensity; s.ncalctotal = c->ncalctotal; s.nrelink = c->nrelink;
  int init[2] = {1 << 30, 0};
  CU(cudaMemcpyAsync(c->flags + 12, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, k_minmax_neigh, std::min(nblocks(c->nown, 256), 1184), 256, 0, c->numneigh, c->nown, c->flags + 12);
  if (int e = sync_flags(c)) return e;
  double nmn = c->h_flags[12], nmx = c->h_flags[13], cnt[2] = {(double)c->ncalctotal, 0.};
  if (int e = comm_allreduce_maxmin(c, &nmx, 1, &nmn, 1)) return e;
  if (int e = comm_allreduce(c, cnt, 1, 2)) return e;
  s.nneigh_min = (int)nmn; s.nneigh_max = (int)nmx;
  if (c->has_comm) s.ncalctotal = (long long)(cnt[0] + 0.5);
  fill_link_scalars(c);
  return 0;
}


