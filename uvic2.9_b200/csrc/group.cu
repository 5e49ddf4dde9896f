// group.cu -- several devices driven by ONE host thread (the reference's host is a single serial Fortran
// process, source/mom/mom.F:289-407; "multi-GPU is driven from this single host thread: one process, 8 devices").
//
// A group is N contexts, one latitude slab each, created from the GLOBAL host arrays exactly as the
// COMMON blocks hold them.  uvic_b200_group_step queues one ocean step on every device (the per-context entry points
// are asynchronous on their streams), then every device PULLS the two halo rows of t(tau+1) on each side from its
// neighbours' memory with one strided peer copy per side (all nt tracers: cudaMemcpy2DAsync over NVLink, no host
// staging, no NCCL -- a serial host has no second process to rendezvous with) on its own exchange stream; the copy
// waits for the two producers' steps through events and the consumer's next step waits for it right before its first
// advection kernel (uvic_b200_wait_before_advection), so the exchange runs beside the next step's coefficient and
// diffusion kernels.  Global inventories are the per-device partial sums added on the host in device order: a fixed
// order, reproducible to the bit for a given decomposition (north_star: fp64 fixed-order reduction).
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>
#include "ctx.h"

struct uvic_b200_group {
  int n = 0;
  uvic_b200_dims gdims;                 // global sizes
  std::vector<uvic_b200_ctx *> ctx;
  std::vector<int> dev, jlo, jhi, jbase, jl;
  std::vector<cudaStream_t> stream, xstream;   // launch stream and exchange stream per device
  std::vector<cudaEvent_t> done, halo;         // step queued / halo rows in place
  bool need_first_exchange = false;
  std::string err;
};

static std::string g_group_err;
static int gfail(uvic_b200_group *g, const std::string &m) {
  if (g) g->err = m; else g_group_err = m;
  return 1;
}
#define GCK(call)                                                                                   \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) {                                                                        \
      char b_[512];                                                                                 \
      snprintf(b_, sizeof b_, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return gfail(g, b_);                                                                          \
    }                                                                                               \
  } while (0)
#define GCTX(r, call)                                                                               \
  do {                                                                                              \
    if ((call) != 0) return gfail(g, std::string("device ") + std::to_string(g->dev[r]) + ": " + uvic_b200_last_error(g->ctx[r])); \
  } while (0)

// rows jb..jb+jl-1 of a host array whose j extent is the middle dimension: (imt, jmt, n3rd) -> (imt, jl, n3rd)
template <typename T>
static std::vector<T> gather_mid(const T *a, int imt, int jmt, int n3, int jb, int jl) {
  std::vector<T> out((size_t)imt * jl * n3);
  for (int q = 0; q < n3; q++)
    memcpy(out.data() + (size_t)q * imt * jl, a + (size_t)q * imt * jmt + (size_t)(jb - 1) * imt, sizeof(T) * (size_t)imt * jl);
  return out;
}

extern "C" {

const char *uvic_b200_group_last_error(const uvic_b200_group *g) { return g ? g->err.c_str() : g_group_err.c_str(); }

int uvic_b200_group_destroy(uvic_b200_group *g) {
  if (!g) return 0;
  for (int r = 0; r < (int)g->ctx.size(); r++) {
    if (g->ctx[r]) uvic_b200_destroy(g->ctx[r]);
    cudaSetDevice(g->dev[r]);
    if (r < (int)g->done.size() && g->done[r]) cudaEventDestroy(g->done[r]);
    if (r < (int)g->halo.size() && g->halo[r]) cudaEventDestroy(g->halo[r]);
    if (r < (int)g->stream.size() && g->stream[r]) cudaStreamDestroy(g->stream[r]);
    if (r < (int)g->xstream.size() && g->xstream[r]) cudaStreamDestroy(g->xstream[r]);
  }
  delete g;
  return 0;
}

// Slabs of equal estimated work (wet cells + land_cost of a cell for every cell of a row: the kernels skip land), at least
// two rows each -- the rule of uvic2.9_b200/slab.py:partition_rows_balanced, so that the single-process and the
// one-process-per-GPU drivers cut the same grid the same way.
static void partition_rows(const int32_t *kmt, int imt, int jmt, int km, int n, std::vector<int> &jlo, std::vector<int> &jhi) {
  const int nrows = jmt - 2;
  std::vector<double> cum(nrows + 1, 0.0);
  for (int q = 0; q < nrows; q++) {
    double w = 0.0;
    for (int i = 1; i < imt - 1; i++) w += kmt[(size_t)(q + 1) * imt + i];
    cum[q + 1] = cum[q] + w + 0.5 * (double)(imt - 2) * km;
  }
  std::vector<int> cuts(n + 1, 0);
  for (int r = 1; r < n; r++) {
    const double target = cum[nrows] * r / n;
    int c = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
    c = std::max(c, cuts[r - 1] + 2);
    c = std::min(c, nrows - 2 * (n - r));
    cuts[r] = c;
  }
  cuts[n] = nrows;
  jlo.resize(n);
  jhi.resize(n);
  for (int r = 0; r < n; r++) { jlo[r] = 2 + cuts[r]; jhi[r] = 2 + cuts[r + 1] - 1; }
}

int uvic_b200_group_create(const uvic_b200_dims *d, const uvic_b200_grid *grid, const uvic_b200_params *par, const uvic_b200_static *st,
                           int ndev, const int32_t *devices, uvic_b200_group **out) {
  uvic_b200_group *g = nullptr;
  if (!d || !grid || !par || !st || !out || ndev < 1) return gfail(nullptr, "uvic_b200_group_create: null argument or ndev < 1");
  *out = nullptr;
  if (d->jmt - 2 < 2 * ndev) return gfail(nullptr, "uvic_b200_group_create: fewer than two rows per device");
  g = new uvic_b200_group();
  g->n = ndev;
  g->gdims = *d;
  g->dev.resize(ndev);
  for (int r = 0; r < ndev; r++) g->dev[r] = devices ? devices[r] : r;
  partition_rows(st->kmt, d->imt, d->jmt, d->km, ndev, g->jlo, g->jhi);
  g->ctx.assign(ndev, nullptr);
  g->stream.assign(ndev, nullptr);
  g->xstream.assign(ndev, nullptr);
  g->done.assign(ndev, nullptr);
  g->halo.assign(ndev, nullptr);
  g->jbase.resize(ndev);
  g->jl.resize(ndev);
  const int imt = d->imt, jmt = d->jmt, km = d->km;
  for (int r = 0; r < ndev; r++) {
    const int jb = std::max(1, g->jlo[r] - 2), jt = std::min(jmt, g->jhi[r] + 2), jl = jt - jb + 1;
    g->jbase[r] = jb;
    g->jl[r] = jl;
    uvic_b200_dims dd = *d;
    dd.jrow_lo = g->jlo[r];
    dd.jrow_hi = g->jhi[r];
    // arrays with j as the LAST extent are contiguous per slab: an offset into the caller's array is enough
    uvic_b200_grid gg = *grid;
    gg.tlat = grid->tlat ? grid->tlat + (size_t)(jb - 1) * imt : nullptr;
    uvic_b200_static ss = *st;
    const size_t o2 = (size_t)(jb - 1) * imt, o3 = (size_t)(jb - 1) * imt * km;
    ss.kmt = st->kmt + o2;
    ss.mskhr = st->mskhr ? st->mskhr + o2 : nullptr;
    ss.addisop = st->addisop ? st->addisop + o3 : nullptr;
    ss.edrm2 = st->edrm2 ? st->edrm2 + o3 : nullptr;
    ss.edrs2 = st->edrs2 ? st->edrs2 + o3 : nullptr;
    ss.edrk1 = st->edrk1 ? st->edrk1 + o3 : nullptr;
    ss.edro1 = st->edro1 ? st->edro1 + o3 : nullptr;
    // ... those with j in the middle, (imt,jmt,km) / (imt,jmt,12), are gathered
    std::vector<double> fisop, sgb, feh, fea;
    if (st->fisop) { fisop = gather_mid(st->fisop, imt, jmt, km, jb, jl); ss.fisop = fisop.data(); }
    if (st->sg_bathy) { sgb = gather_mid(st->sg_bathy, imt, jmt, km, jb, jl); ss.sg_bathy = sgb.data(); }
    if (st->fe_hydr) { feh = gather_mid(st->fe_hydr, imt, jmt, km, jb, jl); ss.fe_hydr = feh.data(); }
    if (st->fe_atmdep) { fea = gather_mid(st->fe_atmdep, imt, jmt, 12, jb, jl); ss.fe_atmdep = fea.data(); }
    if (uvic_b200_create(&dd, &gg, par, &ss, g->dev[r], &g->ctx[r]) != 0) {
      g_group_err = std::string("uvic_b200_group_create: device ") + std::to_string(g->dev[r]) + ": " + uvic_b200_last_error(nullptr);
      uvic_b200_group_destroy(g);
      return 1;
    }
    cudaSetDevice(g->dev[r]);
    if (cudaStreamCreateWithFlags(&g->stream[r], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&g->xstream[r], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&g->done[r], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&g->halo[r], cudaEventDisableTiming) != cudaSuccess) {
      g_group_err = "uvic_b200_group_create: stream / event creation failed";
      uvic_b200_group_destroy(g);
      return 1;
    }
    uvic_b200_set_stream(g->ctx[r], g->stream[r]);
  }
  // peer access between neighbours (NVLink / NVSwitch on the 8 x B200 box); without it the runtime stages the copies
  for (int r = 0; r + 1 < ndev; r++) {
    if (g->dev[r] == g->dev[r + 1]) continue;
    int can = 0;
    cudaDeviceCanAccessPeer(&can, g->dev[r], g->dev[r + 1]);
    if (can) {
      cudaSetDevice(g->dev[r]);
      cudaError_t e = cudaDeviceEnablePeerAccess(g->dev[r + 1], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      cudaSetDevice(g->dev[r + 1]);
      e = cudaDeviceEnablePeerAccess(g->dev[r], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      cudaGetLastError();
    }
  }
  *out = g;
  return 0;
}

int uvic_b200_group_size(const uvic_b200_group *g) { return g ? g->n : 0; }
uvic_b200_ctx *uvic_b200_group_ctx(uvic_b200_group *g, int r) { return (g && r >= 0 && r < g->n) ? g->ctx[r] : nullptr; }
int uvic_b200_group_rows(const uvic_b200_group *g, int r, int32_t *jrow_lo, int32_t *jrow_hi) {
  if (!g || r < 0 || r >= g->n) return 1;
  if (jrow_lo) *jrow_lo = g->jlo[r];
  if (jrow_hi) *jrow_hi = g->jhi[r];
  return 0;
}

int uvic_b200_group_synchronize(uvic_b200_group *g) {
  if (!g) return 1;
  for (int r = 0; r < g->n; r++) {
    GCTX(r, uvic_b200_synchronize(g->ctx[r]));
    cudaSetDevice(g->dev[r]);
    GCK(cudaStreamSynchronize(g->xstream[r]));
  }
  return 0;
}

// a GLOBAL host field (imt, [km | 0:km | 1], jmt, nf) <-> the slabs: per device one strided copy of its jl rows of every
// one of the nf fields (pitch = the global field size)
static int move_rows(uvic_b200_group *g, int r, double *dev, double *host, size_t row_elems, int nf, bool to_device) {
  cudaSetDevice(g->dev[r]);
  const size_t gpitch = row_elems * g->gdims.jmt * sizeof(double), lpitch = row_elems * g->jl[r] * sizeof(double);
  double *h = host + row_elems * (size_t)(g->jbase[r] - 1);
  if (to_device)
    GCK(cudaMemcpy2DAsync(dev, lpitch, h, gpitch, lpitch, nf, cudaMemcpyHostToDevice, g->stream[r]));
  else
    GCK(cudaMemcpy2DAsync(h, gpitch, dev, lpitch, lpitch, nf, cudaMemcpyDeviceToHost, g->stream[r]));
  return 0;
}

int uvic_b200_group_upload_t(uvic_b200_group *g, int level, const double *t_global) {
  if (!g || !t_global || level < -1 || level > 1) return gfail(g, "group_upload_t: bad argument");
  const size_t re = (size_t)g->gdims.imt * g->gdims.km;
  for (int r = 0; r < g->n; r++) {
    GCTX(r, uvic_b200_invalidate_lookahead(g->ctx[r]));
    if (move_rows(g, r, (double *)uvic_b200_t_ptr(g->ctx[r], level), const_cast<double *>(t_global), re, g->gdims.nt, true)) return 1;
  }
  return 0;
}
// every device writes the rows it OWNS (plus the closed wall rows 1 and jmt at the two ends) into the global array
int uvic_b200_group_download_t(uvic_b200_group *g, int level, double *t_global) {
  if (!g || !t_global || level < -1 || level > 1) return gfail(g, "group_download_t: bad argument");
  const size_t re = (size_t)g->gdims.imt * g->gdims.km;
  const int jmt = g->gdims.jmt;
  for (int r = 0; r < g->n; r++) {
    cudaSetDevice(g->dev[r]);
    const int j0 = (r == 0) ? 1 : g->jlo[r], j1 = (r == g->n - 1) ? jmt : g->jhi[r];
    const double *dv = (const double *)uvic_b200_t_ptr(g->ctx[r], level) + re * (size_t)(j0 - g->jbase[r]);
    GCK(cudaMemcpy2DAsync(t_global + re * (size_t)(j0 - 1), re * jmt * sizeof(double), dv, re * g->jl[r] * sizeof(double),
                          re * (size_t)(j1 - j0 + 1) * sizeof(double), g->gdims.nt, cudaMemcpyDeviceToHost, g->stream[r]));
  }
  for (int r = 0; r < g->n; r++) {
    cudaSetDevice(g->dev[r]);
    GCK(cudaStreamSynchronize(g->stream[r]));
  }
  return 0;
}
int uvic_b200_group_upload_adv_vel(uvic_b200_group *g, const double *vet, const double *vnt, const double *vbt) {
  if (!g) return 1;
  const size_t re = (size_t)g->gdims.imt * g->gdims.km, rez = (size_t)g->gdims.imt * (g->gdims.km + 1);
  for (int r = 0; r < g->n; r++)
    GCTX(r, uvic_b200_upload_adv_vel(g->ctx[r], vet ? vet + re * (size_t)(g->jbase[r] - 1) : nullptr,
                                     vnt ? vnt + re * (size_t)(g->jbase[r] - 1) : nullptr, vbt ? vbt + rez * (size_t)(g->jbase[r] - 1) : nullptr));
  return 0;
}
int uvic_b200_group_upload_vbc(uvic_b200_group *g, const double *stf, const double *btf) {
  if (!g) return 1;
  const size_t re = (size_t)g->gdims.imt;
  for (int r = 0; r < g->n; r++) {
    size_t n = 0;
    double *dstf = (double *)uvic_b200_device_ptr(g->ctx[r], "stf", &n), *dbtf = (double *)uvic_b200_device_ptr(g->ctx[r], "btf", &n);
    if (!dstf || !dbtf) return gfail(g, "group_upload_vbc: stf / btf not resident");
    if (stf && move_rows(g, r, dstf, const_cast<double *>(stf), re, g->gdims.nt, true)) return 1;
    if (btf && move_rows(g, r, dbtf, const_cast<double *>(btf), re, g->gdims.nt, true)) return 1;
  }
  return 0;
}
int uvic_b200_group_upload_forcing(uvic_b200_group *g, const double *dnswr, const double *aice, const double *hice, const double *hsno) {
  if (!g) return 1;
  const size_t re = (size_t)g->gdims.imt;
  for (int r = 0; r < g->n; r++) {
    const size_t o = re * (size_t)(g->jbase[r] - 1);
    GCTX(r, uvic_b200_upload_forcing(g->ctx[r], dnswr ? dnswr + o : nullptr, aice ? aice + o : nullptr, hice ? hice + o : nullptr,
                                     hsno ? hsno + o : nullptr));
  }
  return 0;
}

// halo rows of time level `level` of every device pulled from its neighbours, ordered behind `done[]` of the producers
static int exchange(uvic_b200_group *g, int level) {
  const size_t re = (size_t)g->gdims.imt * g->gdims.km;
  const int nt = g->gdims.nt;
  for (int r = 0; r < g->n; r++) {
    if (g->n == 1) break;
    cudaSetDevice(g->dev[r]);
    double *mine = (double *)uvic_b200_t_ptr(g->ctx[r], level);
    const size_t mypitch = re * g->jl[r] * sizeof(double);
    GCK(cudaStreamWaitEvent(g->xstream[r], g->done[r], 0));
    if (r > 0) {      // rows jlo-2, jlo-1 <- the southern neighbour's owned rows jhi-1, jhi
      const int s = r - 1;
      GCK(cudaStreamWaitEvent(g->xstream[r], g->done[s], 0));
      const double *src = (const double *)uvic_b200_t_ptr(g->ctx[s], level) + re * (size_t)(g->jhi[s] - 1 - g->jbase[s]);
      GCK(cudaMemcpy2DAsync(mine + re * (size_t)(g->jlo[r] - 2 - g->jbase[r]), mypitch, src, re * g->jl[s] * sizeof(double),
                            2 * re * sizeof(double), nt, cudaMemcpyDefault, g->xstream[r]));
    }
    if (r + 1 < g->n) {   // rows jhi+1, jhi+2 <- the northern neighbour's owned rows jlo, jlo+1
      const int s = r + 1;
      GCK(cudaStreamWaitEvent(g->xstream[r], g->done[s], 0));
      const double *src = (const double *)uvic_b200_t_ptr(g->ctx[s], level) + re * (size_t)(g->jlo[s] - g->jbase[s]);
      GCK(cudaMemcpy2DAsync(mine + re * (size_t)(g->jhi[r] + 1 - g->jbase[r]), mypitch, src, re * g->jl[s] * sizeof(double),
                            2 * re * sizeof(double), nt, cudaMemcpyDefault, g->xstream[r]));
    }
    GCK(cudaEventRecord(g->halo[r], g->xstream[r]));
  }
  return 0;
}

// One ocean step on every device (isopyc -> vmixc -> MOBI -> tracer, source/mom/mom.F:340-389), then the halo exchange of
// t(tau+1).  `next` (may be NULL) is the stepinfo of the step after this one: the MOBI look-ahead hint.  Asynchronous.
int uvic_b200_group_step(uvic_b200_group *g, const uvic_b200_stepinfo *si, const uvic_b200_stepinfo *next) {
  if (!g || !si) return gfail(g, "group_step: null argument");
  for (int r = 0; r < g->n; r++) {
    if (next) GCTX(r, uvic_b200_hint_next_step(g->ctx[r], next));
    GCTX(r, uvic_b200_step(g->ctx[r], si));
    cudaSetDevice(g->dev[r]);
    GCK(cudaEventRecord(g->done[r], g->stream[r]));
  }
  if (exchange(g, +1)) return 1;
  if (g->n > 1)
    for (int r = 0; r < g->n; r++) GCTX(r, uvic_b200_wait_before_advection(g->ctx[r], g->halo[r]));
  return 0;
}

int uvic_b200_group_rotate(uvic_b200_group *g) {
  if (!g) return 1;
  for (int r = 0; r < g->n; r++) GCTX(r, uvic_b200_rotate(g->ctx[r]));
  return 0;
}

// global tracer inventories sum(t dV) of a time level: partial sums per device, added in device order on the host
int uvic_b200_group_inventory(uvic_b200_group *g, int level, double *out_nt) {
  if (!g || !out_nt) return gfail(g, "group_inventory: null argument");
  const int nt = g->gdims.nt;
  std::vector<double> part((size_t)nt);
  for (int n = 0; n < nt; n++) out_nt[n] = 0.0;
  for (int r = 0; r < g->n; r++) {
    GCTX(r, uvic_b200_inventory(g->ctx[r], level, part.data()));
    for (int n = 0; n < nt; n++) out_nt[n] += part[n];
  }
  return 0;
}

}  // extern "C"
