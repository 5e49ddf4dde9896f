// abi.cu -- the C ABI of include/uvic_b200.h: context life cycle, state movement and the
// entry points that replace the reference's call sites in `mom` (source/mom/mom.F:340-389).
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include "ctx.h"
#include <stdlib.h>

static std::string g_create_err;

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      char b_[512];                                                                               \
      snprintf(b_, sizeof b_, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      if (ctx) ctx->err = b_; else g_create_err = b_;                                             \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)

static int fail(uvic_b200_ctx *ctx, const std::string &msg) {
  if (ctx) ctx->err = msg; else g_create_err = msg;
  return 1;
}

template <typename T>
static int dev_alloc(uvic_b200_ctx *ctx, const char *name, T **slot, size_t n, const T *host_init) {
  T *p = nullptr;
  CK(cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T)));
  if (host_init) CK(cudaMemcpy(p, host_init, n * sizeof(T), cudaMemcpyHostToDevice));
  else CK(cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T)));
  *slot = p;
  ctx->owned.push_back((void *)p);
  NamedArr a;
  a.name = name; a.slot = (void **)slot; a.nelem = n; a.is_int = sizeof(T) == 4;
  ctx->arrs.push_back(a);
  return 0;
}
#define DALLOC(field, n, init) \
  if (dev_alloc(ctx, #field, const_cast<double **>(&ctx->v.field), (size_t)(n), (const double *)(init))) return 1
#define IALLOC(field, n, init) \
  if (dev_alloc(ctx, #field, const_cast<int **>(&ctx->v.field), (size_t)(n), (const int *)(init))) return 1

static void set_levels(uvic_b200_ctx *c, bool leapfrog) {
  c->v.t_0 = c->t_slot[c->lev[1]];
  c->v.t_p1 = c->t_slot[c->lev[2]];
  // on mixing steps both tau-1 and tau are read from the tau slot (09/mom/loadmw.F:109-111)
  c->v.t_m1 = leapfrog ? c->t_slot[c->lev[0]] : c->t_slot[c->lev[1]];
}


// H2D copy of a host velocity array whose first row is global row ctx->host_jfirst (uvic_b200_set_host_window): the
// reference dimensions adv_vet / adv_vnt / adv_vbt (imt,km,jsmw:jmw), 09/mom/mw.h:246-263, i.e. WITHOUT row 1
#define H2D_VEL(dst, src, rowelems, total, strm)                                                                       \
  do {                                                                                                                 \
    const long long skip_ = (long long)std::max(0, ctx->host_jfirst - ctx->v.jbase) * (long long)(rowelems);          \
    CK(cudaMemcpyAsync((dst) + skip_, (src), ((size_t)(total) - (size_t)skip_) * sizeof(double), cudaMemcpyHostToDevice, strm)); \
  } while (0)

extern "C" {

static void halo_wait_now(uvic_b200_ctx *ctx);

const char *uvic_b200_version(void) { return "uvic_b200 0.1 (sm_100a)"; }

const char *uvic_b200_last_error(const uvic_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

static int create_impl(const uvic_b200_dims *d, const uvic_b200_grid *g, const uvic_b200_params *par, const uvic_b200_static *st,
                       int device, uvic_b200_ctx **out, uvic_b200_ctx *&ctx);
int uvic_b200_create(const uvic_b200_dims *d, const uvic_b200_grid *g, const uvic_b200_params *par, const uvic_b200_static *st,
                     int device, uvic_b200_ctx **out) {
  uvic_b200_ctx *ctx = nullptr;
  if (out) *out = nullptr;
  const int rc = create_impl(d, g, par, st, device, out, ctx);
  if (rc != 0 && ctx) {
    // a failed create leaves nothing behind: the cause goes to the create-error slot (uvic_b200_last_error(NULL)), the
    // device memory, streams and events of the half-built context are released
    g_create_err = ctx->err;
    if (out) *out = nullptr;
    uvic_b200_destroy(ctx);
  }
  return rc;
}
static int create_impl(const uvic_b200_dims *d, const uvic_b200_grid *g, const uvic_b200_params *par, const uvic_b200_static *st,
                       int device, uvic_b200_ctx **out, uvic_b200_ctx *&ctx) {
  if (!d || !g || !par || !st || !out) return fail(nullptr, "uvic_b200_create: null argument");
  if (d->imt < 4 || d->jmt < 4 || d->km < 2 || d->nt < 2) return fail(nullptr, "uvic_b200_create: bad dims");
  if (d->jrow_lo < 2 || d->jrow_hi > d->jmt - 1 || d->jrow_lo > d->jrow_hi)
    return fail(nullptr, "uvic_b200_create: rows must satisfy 2 <= jrow_lo <= jrow_hi <= jmt-1");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, "uvic_b200_create: no CUDA device (this library has no CPU fallback)");
  {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  }
  ctx = new uvic_b200_ctx();
  ctx->device = device;
  ctx->stream = 0;
  ctx->launches = 0;
  ctx->par = *par;
  ctx->red_partial = ctx->red_out = nullptr;
  ctx->pin_buf = nullptr; ctx->pin_bytes = 0;
  ctx->prof_on = false;
  ctx->stream3 = nullptr; ctx->ev_elem = ctx->ev_gm = nullptr; ctx->gm_inflight = false;
  ctx->stream2 = nullptr; ctx->fork_event = nullptr; ctx->mobi_event = nullptr; ctx->src_ready[0] = ctx->src_ready[1] = nullptr; ctx->mobi_inflight = false;
  ctx->mobi_dtnpzd = 0.0;
  ctx->src_buf[0] = ctx->src_buf[1] = nullptr; ctx->src_cur = 0;
  ctx->hint_valid = ctx->ahead_valid = false; ctx->ahead_tm1 = nullptr; ctx->ahead_buf = 0;
  ctx->main_done_event = nullptr; ctx->copy_in = ctx->copy_out = nullptr; ctx->h2d_event = nullptr; ctx->h2d_vbc_event = nullptr;
  ctx->rho_dev = nullptr;
  ctx->halo_event = nullptr;
  ctx->clinic = nullptr;
  ctx->filtu_items = nullptr; ctx->filtu_mats = nullptr; ctx->filtu_rows = nullptr; ctx->filtu_nitems = ctx->filtu_maxim = ctx->filtu_nrows = 0;
  ctx->tavg_t = ctx->tavg_stf = ctx->tavg_tmp = ctx->tavg_vflux = ctx->tavg_gaost = nullptr; ctx->navgts = 0; ctx->d2h_dst = nullptr; ctx->d2h_ntr = 0;
  DevView &v = ctx->v;
  memset(&v, 0, sizeof v);
  v.imt = d->imt; v.jmt = d->jmt; v.km = d->km; v.nt = d->nt; v.nsrc = d->nsrc;
  v.jlo = d->jrow_lo; v.jhi = d->jrow_hi;
  v.jbase = std::max(1, v.jlo - 2);
  int jtop = std::min(v.jmt, v.jhi + 2);
  v.jl = jtop - v.jbase + 1;
  const int imt = v.imt, km = v.km, jl = v.jl, jmt = v.jmt, nt = v.nt;
  v.n2 = (long long)imt * jl;
  v.n3 = v.n2 * km;
  v.n3z = v.n2 * (km + 1);
  v.aidif = par->aidif; v.kappa_h = par->kappa_h; v.ahisop = par->ahisop; v.athkdf = par->athkdf; v.slmxr = par->slmxr;
  v.diff_cet = par->diff_cet; v.diff_cnt = par->diff_cnt; v.zetar = par->zetar; v.ogamma = par->ogamma;
  v.gravrho0r = par->gravrho0r;
  v.fct = par->fct; v.isopycmix = par->isopycmix; v.tidal_kv = par->tidal_kv;

  // closed walls: rows 1 and jmt must be land (the FCT boundary rules rely on it,
  // 09/mom/tracer_adv_flx.F:467-482, 554-556)
  for (int jj = 0; jj < jl; jj++) {
    int jglob = v.jbase + jj;
    if (jglob == 1 || jglob == jmt)
      for (int i = 0; i < imt; i++)
        if (st->kmt[i + (size_t)imt * jj] != 0) { return fail(ctx, "uvic_b200_create: rows 1 and jmt must be land (kmt=0)"); }
  }

  // ---- grid ----
#define G1(name, n) DALLOC(name, n, g->name)
  G1(dxt, imt); G1(dxtr, imt); G1(dxt2r, imt); G1(dxt4r, imt); G1(dxu, imt); G1(dxur, imt);
  G1(dyt, jmt); G1(dytr, jmt); G1(dyt2r, jmt); G1(dyt4r, jmt); G1(dyu, jmt); G1(dyur, jmt);
  G1(cst, jmt); G1(cstr, jmt); G1(csu, jmt); G1(csur, jmt); G1(cstdytr, jmt); G1(cstdyt2r, jmt); G1(csu_dyur, jmt);
  G1(dzt, km); G1(dztr, km); G1(dzt2r, km); G1(dztur, km); G1(dztlr, km); G1(zt, km); G1(zw, km);
  G1(dzw, km + 1); G1(dzwr, km + 1);
  G1(dtxcel, km); G1(dtxsqr, km); G1(dztxcl, km); G1(dzwxcl, km);
  G1(duw, imt); G1(due, imt); G1(dus, jmt); G1(dun, jmt);
  G1(eosc, km * 9); G1(to, km); G1(so, km);
  G1(tlat, v.n2);
#undef G1
  // tidal geometry tables (09/mom/vmixc.F:100-104): host libm exp, time invariant
  {
    std::vector<double> e1((size_t)km * km, 0.0), den(km, 1.0);
    for (int k1 = 1; k1 <= km; k1++) {
      den[k1 - 1] = 1 - exp(-par->zetar * g->zw[k1 - 1]);
      for (int k = 1; k <= km; k++) {
        double hab = g->zw[k - 1] - g->zw[k1 - 1];
        e1[(k - 1) + (size_t)km * (k1 - 1)] = exp(hab * par->zetar);
      }
    }
    DALLOC(edr_e1, (size_t)km * km, e1.data());
    DALLOC(edr_den, km, den.data());
  }
  // ---- static inputs ----
  IALLOC(kmt, v.n2, st->kmt);
  if (st->mskhr) IALLOC(mskhr, v.n2, st->mskhr);
  ctx->itrc_h.assign(nt, 0);
  if (par->itrc) for (int n = 0; n < nt; n++) ctx->itrc_h[n] = par->itrc[n];
  for (int n = 0; n < nt; n++)
    if (ctx->itrc_h[n] < 0 || ctx->itrc_h[n] > v.nsrc) { return fail(ctx, "uvic_b200_create: itrc out of range"); }
  IALLOC(itrc, nt, ctx->itrc_h.data());
  DALLOC(fisop, v.n3, st->fisop);
  DALLOC(addisop, v.n3, st->addisop);
  {
    // latitude-weighted tidal constituent sum (09/mom/vmixc.F:73-84,101-102)
    std::vector<double> es((size_t)v.n3, 0.0);
    if (par->tidal_kv && st->edrm2 && st->edrs2 && st->edrk1 && st->edro1) {
      for (int jj = 0; jj < jl; jj++)
        for (int i = 0; i < imt; i++) {
          double lat = g->tlat[i + (size_t)imt * jj];
          double qk1, qo1, q2;
          if (fabs(lat) < 30.) { qk1 = 0.33; qo1 = 0.33; } else { qk1 = 1.; qo1 = 1.; }
          if (fabs(lat) < 70.) q2 = 0.33; else q2 = 1.;
          for (int k = 0; k < km; k++) {
            size_t x = i + (size_t)imt * (k + (size_t)km * jj);
            es[x] = (q2 * (st->edrm2[x] + st->edrs2[x]) + qk1 * st->edrk1[x] + qo1 * st->edro1[x]);
          }
        }
    }
    DALLOC(edrsum, v.n3, es.data());
  }
  // ---- state and work arrays ----
  // MOBI option subsets (O_carbon_13 / O_carbon_14 / O_mobi_nitrogen_15 off in run/mk.in: BASELINE config 2, nt = 21): an
  // index of 0 in the MOBI maps says "this tracer does not exist".  The kernels keep their 32-variable state vector; the
  // absent variables are read from one scratch field behind the nt tracers of every time level (constant 1, so every
  // isotope ratio stays finite) and their sources land in one scratch slot behind the nsrc sources that nothing reads.
  // The equations of the present variables never read an isotope variable (09/mom/mobi.F: isotope terms only appear in
  // the isotope equations), so the present tracers are those of the reference built without the options.
  bool subset = false;
  if (par->mobi && par->mobi_index)
    for (int m = 0; m < IX_N && m < par->n_mobi_index; m++) subset = subset || par->mobi_index[m] == 0;
  const int nt_alloc = nt + (subset ? 1 : 0);
  for (int s = 0; s < 3; s++) {
    double *p = nullptr;
    CK(cudaMalloc((void **)&p, (size_t)v.n3 * nt_alloc * sizeof(double)));
    CK(cudaMemset(p, 0, (size_t)v.n3 * nt_alloc * sizeof(double)));
    if (subset) {
      std::vector<double> ones((size_t)v.n3, 1.0);
      CK(cudaMemcpy(p + (size_t)v.n3 * nt, ones.data(), (size_t)v.n3 * sizeof(double), cudaMemcpyHostToDevice));
    }
    ctx->t_slot[s] = p;
    ctx->owned.push_back(p);
    ctx->lev[s] = s;
  }
  set_levels(ctx, true);
  DALLOC(u, v.n3 * 2, nullptr);
  DALLOC(adv_vet, v.n3, nullptr); DALLOC(adv_vnt, v.n3, nullptr); DALLOC(adv_vbt, v.n3z, nullptr);
  DALLOC(ue, v.n3, nullptr); DALLOC(vn, v.n3, nullptr); DALLOC(wb, v.n3z, nullptr);
  DALLOC(adv_vetiso, v.n3, nullptr); DALLOC(adv_vntiso, v.n3, nullptr); DALLOC(adv_vbtiso, v.n3z, nullptr);
  DALLOC(alphai, v.n3, nullptr); DALLOC(betai, v.n3, nullptr);
  DALLOC(ddxt, v.n3 * 2, nullptr); DALLOC(ddyt, v.n3 * 2, nullptr); DALLOC(ddzt, v.n3z * 2, nullptr);
  DALLOC(ce, v.n3 * 4, nullptr); DALLOC(cn, v.n3 * 4, nullptr); DALLOC(cbx, v.n3 * 4, nullptr); DALLOC(cby, v.n3 * 4, nullptr);
  DALLOC(K11, v.n3, nullptr); DALLOC(K22, v.n3, nullptr); DALLOC(K33, v.n3, nullptr);
  DALLOC(diff_cbt, v.n3, nullptr); DALLOC(tri_a, v.n3, nullptr); DALLOC(tri_e, v.n3, nullptr); DALLOC(tri_bet, v.n3, nullptr);
  DALLOC(stf, v.n2 * nt, nullptr); DALLOC(btf, v.n2 * nt, nullptr);
  DALLOC(src, v.n3 * (std::max(v.nsrc, 1) + (subset ? 1 : 0)), nullptr);
  ctx->src_buf[0] = v.src;
  if (par->mobi) {
    if (!par->mobi_par || !par->mobi_index || par->n_mobi_par < (int)(sizeof(MobiPar) / sizeof(double)) || par->n_mobi_index < IX_N ||
        !st->sg_bathy || !st->fe_hydr || !st->fe_atmdep) {
      return fail(ctx, "uvic_b200_create: O_mobi needs mobi_par, mobi_index, sg_bathy, fe_hydr and fe_atmdep");
    }
    if (km > MOBI_KMAX) { return fail(ctx, "uvic_b200_create: km exceeds MOBI_KMAX"); }
    for (int m = 0; m < IX_N; m++) {
      int x = par->mobi_index[m];
      bool is_src = (m >= IX_SRC && m < IX_SRC + MOBI_NVAR) || m >= IX_ISALK;
      if (x < 0 || x > (is_src ? v.nsrc : nt)) { return fail(ctx, "uvic_b200_create: MOBI index map out of range"); }
      // the variables every MOBI configuration has (09/mom/mobi.h:104-142 with O_mobi alone)
      const bool required = m == IX_TR + V_PO4 || m == IX_TR + V_PHYT || m == IX_TR + V_ZOOP || m == IX_TR + V_DETR || m == IX_TR + V_DIC ||
                            m == IX_ITEMP || m == IX_ISALT || m == IX_IALK || m == IX_IO2;
      if (x == 0 && required) { return fail(ctx, "uvic_b200_create: MOBI index map: a required tracer is marked absent"); }
    }
    double *mp = nullptr;
    if (dev_alloc(ctx, "mobi_par", &mp, sizeof(MobiPar) / sizeof(double), par->mobi_par)) return 1;
    v.mobi_par = (const MobiPar *)mp;
    ctx->mobi_dtnpzd = ((const MobiPar *)par->mobi_par)->dtnpzd;
    IALLOC(mobi_idx, MOBI_NIDX, nullptr);
    {
      // absent entries -> the scratch tracer (nt + 1) / the scratch source slot (nsrc + 1)
      std::vector<int> mi(MOBI_NIDX, 0);
      for (int m = 0; m < std::min(par->n_mobi_index, MOBI_NIDX); m++) mi[m] = par->mobi_index[m];
      for (int m = 0; m < IX_N; m++) {
        const bool is_src = (m >= IX_SRC && m < IX_SRC + MOBI_NVAR) || m >= IX_ISALK;
        if (mi[m] == 0) mi[m] = is_src ? v.nsrc + 1 : nt + 1;
      }
      CK(cudaMemcpy(const_cast<int *>(v.mobi_idx), mi.data(), sizeof(int) * MOBI_NIDX, cudaMemcpyHostToDevice));
    }
    DALLOC(sg_bathy, v.n3, st->sg_bathy);
    DALLOC(fe_hydr, v.n3, st->fe_hydr);
    DALLOC(fe_atmdep, v.n2 * 12, st->fe_atmdep);
    DALLOC(dnswr, v.n2, nullptr); DALLOC(aice, v.n2, nullptr); DALLOC(hice, v.n2, nullptr); DALLOC(hsno, v.n2, nullptr);
    DALLOC(mobi_pre, (size_t)v.n3 * MOBI_NPRE, nullptr);
    DALLOC(mobi_day, v.n2, nullptr);
    {
      // depth dependence of the benthic denitrification fractionation (09/mom/mobi.F:1062): host libm exp, time invariant
      std::vector<double> eb(km);
      for (int k = 0; k < km; k++) eb[k] = ((const MobiPar *)par->mobi_par)->eps_bdeni0 * exp(-2.5e-6 * g->zt[k]);
      DALLOC(mobi_epsbd, km, eb.data());
      // water columns of the owned rows, deepest first (stable, so equal depths keep the i-fastest order)
      std::vector<int> cols;
      for (int jj = v.jlo - v.jbase; jj <= v.jhi - v.jbase; jj++)
        for (int i = 1; i < imt - 1; i++)
          if (st->kmt[i + (size_t)imt * jj] > 0) cols.push_back(i + imt * jj);
      std::stable_sort(cols.begin(), cols.end(), [&](int a, int b) { return st->kmt[a] > st->kmt[b]; });
      v.mobi_ncols = (int)cols.size();
      IALLOC(mobi_cols, cols.size(), cols.data());
    }
    CK(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->mobi_event, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->src_ready[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->src_ready[1], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->main_done_event, cudaEventDisableTiming));
  }
  CK(cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&ctx->stream3, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&ctx->ev_elem, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->ev_gm, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->vel_free, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->h2d_event, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->h2d_vbc_event, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
  {
    int *ip = nullptr;
    if (dev_alloc(ctx, "conv_n", &ip, (size_t)v.n2, (const int *)nullptr)) return 1;
    v.conv_n = ip;
    if (dev_alloc(ctx, "conv_kt", &ip, (size_t)v.n2 * (km / 2 + 1), (const int *)nullptr)) return 1;
    v.conv_kt = ip;
    DALLOC(conv_zsm, (size_t)v.n2 * (km / 2 + 1), nullptr);
  }
  {
    // The marching FCT kernel needs no scratch and takes all tracers in one batch.  The two-pass reference variants
    // (UVIC_B200_FCT=split|merged) park six ratios per tracer of a group in HBM; keep the group within ~12 GiB.
    v.ngroup = nt;
    v.Rfac = nullptr;
    if (fct_variant() != 0) {
      size_t per = (size_t)v.n3 * 6 * sizeof(double);
      size_t cap = (size_t)12 << 30;
      v.ngroup = (int)std::max<size_t>(1, std::min<size_t>((size_t)nt, cap / per));
      DALLOC(Rfac, (size_t)v.n3 * 6 * v.ngroup, nullptr);
    }
  }
  {
    size_t ntb = (size_t)km * nt * jl;
    CK(cudaMalloc((void **)&ctx->tbar, ntb * sizeof(double)));
    CK(cudaMemset(ctx->tbar, 0, ntb * sizeof(double)));
    ctx->owned.push_back(ctx->tbar);
    CK(cudaMalloc((void **)&ctx->travar, ntb * sizeof(double)));
    CK(cudaMalloc((void **)&ctx->dtabs, ntb * sizeof(double)));
    CK(cudaMemset(ctx->travar, 0, ntb * sizeof(double)));
    CK(cudaMemset(ctx->dtabs, 0, ntb * sizeof(double)));
    ctx->owned.push_back(ctx->travar);
    ctx->owned.push_back(ctx->dtabs);
    CK(cudaMalloc((void **)&ctx->sumbk, (size_t)3 * km * nt * sizeof(double)));
    ctx->owned.push_back(ctx->sumbk);
    CK(cudaMalloc((void **)&ctx->red_out, (size_t)nt * sizeof(double)));
    ctx->owned.push_back(ctx->red_out);
  }
  ctx->filt_items = nullptr; ctx->filt_mats = nullptr; ctx->filt_nitems = 0; ctx->filt_maxim = 0;
  if (par->fourfil) {
    if (par->jfrst < 1 || par->jft0 < 1 || par->jft0 > jmt || par->jft1 < par->jfrst || par->jft2 <= par->jft1 || par->jft2 > jmt) {
      return fail(ctx, "uvic_b200_create: O_fourfil needs 1 <= jfrst <= jft1 < jft2 <= jmt and a valid jft0");
    }
    if (filter_setup(ctx, st->kmt, g->cst, g->cstr)) return fail(ctx, "uvic_b200_create: filter set-up failed");
  }
  CK(cudaDeviceSynchronize());
  *out = ctx;
  return 0;
}

int uvic_b200_destroy(uvic_b200_ctx *ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (void *p : ctx->owned) cudaFree(p);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream3) cudaStreamDestroy(ctx->stream3);
  if (ctx->ev_elem) cudaEventDestroy(ctx->ev_elem);
  if (ctx->ev_gm) cudaEventDestroy(ctx->ev_gm);
  if (ctx->vel_free) cudaEventDestroy(ctx->vel_free);
  if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
  if (ctx->mobi_event) cudaEventDestroy(ctx->mobi_event);
  for (auto e : ctx->src_ready) if (e) cudaEventDestroy(e);
  if (ctx->main_done_event) cudaEventDestroy(ctx->main_done_event);
  if (ctx->h2d_event) cudaEventDestroy(ctx->h2d_event);
  if (ctx->h2d_vbc_event) cudaEventDestroy(ctx->h2d_vbc_event);
  if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
  if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
  for (auto e : ctx->ev_batch) cudaEventDestroy(e);
  delete ctx->clinic;
  delete ctx;
  return 0;
}

int uvic_b200_set_stream(uvic_b200_ctx *ctx, void *s) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx) return 1;
  ctx->stream = (cudaStream_t)s;
  return 0;
}
int uvic_b200_synchronize(uvic_b200_ctx *ctx) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  gm_join(ctx);
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  return 0;
}

static int lev_index(int level) { return level + 1; }

int uvic_b200_upload_t(uvic_b200_ctx *ctx, int level, const double *h) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (level < -1 || level > 1) return fail(ctx, "upload_t: level must be -1, 0 or 1");
  if (ctx->ahead_valid) {   // a look-ahead MOBI may be reading this slot; its result is void now
    CK(cudaStreamWaitEvent(ctx->stream, ctx->mobi_event, 0));
    ctx->ahead_valid = false;
  }
  CK(cudaMemcpyAsync(ctx->t_slot[ctx->lev[lev_index(level)]], h, (size_t)ctx->v.n3 * ctx->v.nt * sizeof(double),
                     cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
int uvic_b200_download_t(uvic_b200_ctx *ctx, int level, double *h) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (level < -1 || level > 1) return fail(ctx, "download_t: level must be -1, 0 or 1");
  CK(cudaMemcpyAsync(h, ctx->t_slot[ctx->lev[lev_index(level)]], (size_t)ctx->v.n3 * ctx->v.nt * sizeof(double),
                     cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int uvic_b200_download_tracer(uvic_b200_ctx *ctx, int level, int n, double *h) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (level < -1 || level > 1 || n < 1 || n > ctx->v.nt) return fail(ctx, "download_tracer: bad level or tracer index");
  CK(cudaMemcpyAsync(h, ctx->t_slot[ctx->lev[lev_index(level)]] + (size_t)(n - 1) * ctx->v.n3, (size_t)ctx->v.n3 * sizeof(double),
                     cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int uvic_b200_upload_adv_vel(uvic_b200_ctx *ctx, const double *vet, const double *vnt, const double *vbt) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  DevView &v = ctx->v;
  if (vet) H2D_VEL(v.adv_vet, vet, (long long)v.imt * v.km, v.n3, ctx->stream);
  if (vnt) H2D_VEL(v.adv_vnt, vnt, (long long)v.imt * v.km, v.n3, ctx->stream);
  if (vbt) H2D_VEL(v.adv_vbt, vbt, (long long)v.imt * (v.km + 1), v.n3z, ctx->stream);
  return 0;
}
int uvic_b200_upload_u(uvic_b200_ctx *ctx, const double *u) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  CK(cudaMemcpyAsync(ctx->v.u, u, (size_t)ctx->v.n3 * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
int uvic_b200_adv_vel(uvic_b200_ctx *ctx) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (ctx) ctx->vel_free_valid = false;
  launch_adv_vel(ctx);
  CK(cudaGetLastError());
  return 0;
}
// state (source/mom/state.F) of a time level: rho_host(imt,km,jl) = dens(T - to, S - so, k); the scratch field of the
// time averages doubles as the device buffer
int uvic_b200_state(uvic_b200_ctx *ctx, int level, double *rho_host) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  DevView &v = ctx->v;
  if (level < -1 || level > 1 || !rho_host) return fail(ctx, "state: level must be -1, 0 or 1 and rho non-null");
  if (!ctx->rho_dev) {
    CK(cudaMalloc((void **)&ctx->rho_dev, (size_t)v.n3 * sizeof(double)));
    ctx->owned.push_back(ctx->rho_dev);
  }
  if (level >= 0) halo_wait_now(ctx);
  launch_state(ctx, ctx->t_slot[ctx->lev[level + 1]], ctx->rho_dev);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(rho_host, ctx->rho_dev, (size_t)v.n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int uvic_b200_upload_vbc(uvic_b200_ctx *ctx, const double *stf, const double *btf) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  DevView &v = ctx->v;
  if (stf) CK(cudaMemcpyAsync(v.stf, stf, (size_t)v.n2 * v.nt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (btf) CK(cudaMemcpyAsync(v.btf, btf, (size_t)v.n2 * v.nt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
int uvic_b200_upload_forcing(uvic_b200_ctx *ctx, const double *dnswr, const double *aice, const double *hice, const double *hsno) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  DevView &v = ctx->v;
  if (!ctx->par.mobi) return fail(ctx, "upload_forcing: context was created without O_mobi");
  if (ctx->ahead_valid) {
    CK(cudaStreamWaitEvent(ctx->stream, ctx->mobi_event, 0));
    ctx->ahead_valid = false;
  }
  size_t nb = (size_t)v.n2 * sizeof(double);
  if (dnswr) CK(cudaMemcpyAsync(v.dnswr, dnswr, nb, cudaMemcpyHostToDevice, ctx->stream));
  if (aice) CK(cudaMemcpyAsync(v.aice, aice, nb, cudaMemcpyHostToDevice, ctx->stream));
  if (hice) CK(cudaMemcpyAsync(v.hice, hice, nb, cudaMemcpyHostToDevice, ctx->stream));
  if (hsno) CK(cudaMemcpyAsync(v.hsno, hsno, nb, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
// ---- surface boundary conditions on the device ----
int uvic_b200_sbc_setup(uvic_b200_ctx *ctx, int numsbc, const int32_t *flx_index, const int32_t *acc_index) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx || numsbc < 1 || !flx_index || !acc_index) return fail(ctx, "sbc_setup: bad argument");
  if (ctx->v.sbc) return fail(ctx, "sbc_setup: already set up");
  for (int n = 0; n < ctx->v.nt; n++)
    if (flx_index[n] < 0 || flx_index[n] > numsbc || acc_index[n] < 0 || acc_index[n] > numsbc)
      return fail(ctx, "sbc_setup: slot index out of range");
  ctx->v.numsbc = numsbc;
  DALLOC(sbc, (size_t)ctx->v.n2 * numsbc, nullptr);
  DALLOC(bhf, ctx->v.n2, nullptr);
  IALLOC(sbc_flx, ctx->v.nt, flx_index);
  IALLOC(sbc_acc, ctx->v.nt, acc_index);
  return 0;
}
int uvic_b200_upload_sbc(uvic_b200_ctx *ctx, const double *sbc, const double *bhf) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  DevView &v = ctx->v;
  if (!v.sbc) return fail(ctx, "upload_sbc: call uvic_b200_sbc_setup first");
  if (sbc) CK(cudaMemcpyAsync(v.sbc, sbc, (size_t)v.n2 * v.numsbc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (bhf) CK(cudaMemcpyAsync(v.bhf, bhf, (size_t)v.n2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
int uvic_b200_upload_sbc_slot(uvic_b200_ctx *ctx, int slot, const double *field) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  DevView &v = ctx->v;
  if (!v.sbc || slot < 1 || slot > v.numsbc || !field) return fail(ctx, "upload_sbc_slot: bad slot or no sbc_setup");
  CK(cudaMemcpyAsync(v.sbc + (size_t)(slot - 1) * v.n2, field, (size_t)v.n2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
int uvic_b200_download_sbc_slot(uvic_b200_ctx *ctx, int slot, double *field) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  DevView &v = ctx->v;
  if (!v.sbc || slot < 1 || slot > v.numsbc || !field) return fail(ctx, "download_sbc_slot: bad slot or no sbc_setup");
  CK(cudaMemcpyAsync(field, v.sbc + (size_t)(slot - 1) * v.n2, (size_t)v.n2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int uvic_b200_download_sbc(uvic_b200_ctx *ctx, double *sbc) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  DevView &v = ctx->v;
  if (!v.sbc || !sbc) return fail(ctx, "download_sbc: call uvic_b200_sbc_setup first");
  CK(cudaMemcpyAsync(sbc, v.sbc, (size_t)v.n2 * v.numsbc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int uvic_b200_gasbc(uvic_b200_ctx *ctx, const uvic_b200_gasbc_par *gp) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  DevView &v = ctx->v;
  if (!v.sbc || !gp) return fail(ctx, "gasbc: call uvic_b200_sbc_setup first");
  if (!ctx->par.mobi || !v.aice) return fail(ctx, "gasbc: needs a context with O_mobi (ice fraction, carbon tracers)");
  const int32_t rd[] = {gp->isst, gp->isss, gp->issdic, gp->issalk, gp->issdic13, gp->issc14, gp->isso2, gp->iws,
                        gp->idicflx, gp->idic13flx, gp->ic14flx, gp->io2flx};
  for (int32_t x : rd)
    if (x < 1 || x > v.numsbc) return fail(ctx, "gasbc: sbc slot out of range");
  if (gp->inpp > 0 && (gp->inpp > v.numsbc || gp->isr < 1 || gp->isr > v.numsbc || gp->iburn < 1 || gp->iburn > v.numsbc))
    return fail(ctx, "gasbc: land carbon slot out of range");
  launch_gasbc(ctx, gp);
  CK(cudaGetLastError());
  return 0;
}
int uvic_b200_setvbc(uvic_b200_ctx *ctx) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx->v.sbc) return fail(ctx, "setvbc: call uvic_b200_sbc_setup first");
  launch_setvbc(ctx);
  CK(cudaGetLastError());
  return 0;
}
int uvic_b200_set_sbc(uvic_b200_ctx *ctx, int eots, int osegs, int osege, int ntspos) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx->v.sbc) return fail(ctx, "set_sbc: call uvic_b200_sbc_setup first");
  if (ntspos < 1) return fail(ctx, "set_sbc: ntspos must be >= 1");
  launch_set_sbc(ctx, eots, osegs, osege, ntspos);
  CK(cudaGetLastError());
  return 0;
}

// ---- time averages of the tracers on the device ----
int uvic_b200_tavg_accumulate(uvic_b200_ctx *ctx, const double *vflux, const double *gaost) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  DevView &v = ctx->v;
  if (!ctx->tavg_t) {
    CK(cudaMalloc((void **)&ctx->tavg_t, (size_t)v.n3 * v.nt * sizeof(double)));
    ctx->owned.push_back(ctx->tavg_t);
    CK(cudaMalloc((void **)&ctx->tavg_stf, (size_t)v.n2 * v.nt * sizeof(double)));
    ctx->owned.push_back(ctx->tavg_stf);
    CK(cudaMalloc((void **)&ctx->tavg_tmp, (size_t)v.n3 * sizeof(double)));
    ctx->owned.push_back(ctx->tavg_tmp);
    CK(cudaMalloc((void **)&ctx->tavg_vflux, (size_t)v.n2 * sizeof(double)));
    ctx->owned.push_back(ctx->tavg_vflux);
    CK(cudaMalloc((void **)&ctx->tavg_gaost, (size_t)v.nt * sizeof(double)));
    ctx->owned.push_back(ctx->tavg_gaost);
    CK(cudaMemsetAsync(ctx->tavg_t, 0, (size_t)v.n3 * v.nt * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->tavg_stf, 0, (size_t)v.n2 * v.nt * sizeof(double), ctx->stream));
    ctx->navgts = 0;
  }
  if (vflux) CK(cudaMemcpyAsync(ctx->tavg_vflux, vflux, (size_t)v.n2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (gaost) CK(cudaMemcpyAsync(ctx->tavg_gaost, gaost, (size_t)v.nt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  launch_tavg_accumulate(ctx, vflux ? ctx->tavg_vflux : nullptr, gaost ? ctx->tavg_gaost : nullptr);
  CK(cudaGetLastError());
  ctx->navgts += 1;
  return 0;
}
int uvic_b200_tavg_fetch(uvic_b200_ctx *ctx, double *avg_t, double *avg_stf, int32_t *navgts, int reset) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  DevView &v = ctx->v;
  if (!ctx->tavg_t || ctx->navgts < 1) return fail(ctx, "tavg_fetch: nothing accumulated");
  const double rnavgt = 1.0 / (double)ctx->navgts;   // 09/mom/timeavgs.F:407
  if (navgts) *navgts = ctx->navgts;
  if (avg_t)
    for (int n = 0; n < v.nt; n++) {   // one tracer at a time through a one-field scratch: no second nt-sized array
      launch_tavg_mean(ctx, ctx->tavg_t + (size_t)n * v.n3, ctx->tavg_tmp, v.n3, rnavgt);
      CK(cudaMemcpyAsync(avg_t + (size_t)n * v.n3, ctx->tavg_tmp, (size_t)v.n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
  if (avg_stf)
    for (int n = 0; n < v.nt; n++) {
      launch_tavg_mean(ctx, ctx->tavg_stf + (size_t)n * v.n2, ctx->tavg_tmp, v.n2, rnavgt);
      CK(cudaMemcpyAsync(avg_stf + (size_t)n * v.n2, ctx->tavg_tmp, (size_t)v.n2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
  if (reset) {
    CK(cudaMemsetAsync(ctx->tavg_t, 0, (size_t)v.n3 * v.nt * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->tavg_stf, 0, (size_t)v.n2 * v.nt * sizeof(double), ctx->stream));
    ctx->navgts = 0;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  return 0;
}

// ---- baroclinic momentum step on the device (09/mom/clinic.F) ----
static int clinic_setup_impl(uvic_b200_ctx *ctx, const uvic_b200_clinic_static *cs);
int uvic_b200_clinic_setup(uvic_b200_ctx *ctx, const uvic_b200_clinic_static *cs) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  const size_t narrs = ctx ? ctx->arrs.size() : 0;
  const int rc = clinic_setup_impl(ctx, cs);
  if (rc != 0 && ctx && ctx->clinic && ctx->err.find("already set up") == std::string::npos) {
    ctx->arrs.resize(narrs);   // their slots point into the ClinicView deleted below (uvic_b200_device_ptr / _fetch by name)
    // a half-built momentum state must not be usable: the device arrays stay owned by the context (freed at destroy)
    delete ctx->clinic;
    ctx->clinic = nullptr;
    ctx->filtu_nitems = ctx->filtu_nrows = 0;
  }
  return rc;
}
static int clinic_setup_impl(uvic_b200_ctx *ctx, const uvic_b200_clinic_static *cs) {
  if (!ctx || !cs) return fail(ctx, "clinic_setup: null argument");
  if (ctx->clinic) return fail(ctx, "clinic_setup: already set up");
  if (!cs->kmu || !cs->hr || !cs->cori || !cs->advmet || !cs->am3 || !cs->am4 || !cs->dxmetr || !cs->dxu2r || !cs->dyu2r ||
      !cs->dyu4r || !cs->csudyu2r || !cs->visc_ceu || !cs->amc_north || !cs->amc_south)
    return fail(ctx, "clinic_setup: every array of uvic_b200_clinic_static is required");
  DevView &v = ctx->v;
  for (long long x = 0; x < v.n2; x++)
    if (cs->kmu[x] < 0 || cs->kmu[x] > v.km) return fail(ctx, "clinic_setup: kmu out of range");
  ClinicView *cv = new ClinicView();
  memset(cv, 0, sizeof *cv);
  ctx->clinic = cv;
#define CALLOC(field, n, init) \
  if (dev_alloc(ctx, #field, const_cast<double **>(&cv->field), (size_t)(n), (const double *)(init))) return 1
  if (dev_alloc(ctx, "kmu", const_cast<int **>(&cv->kmu), (size_t)v.n2, (const int *)cs->kmu)) return 1;
  CALLOC(hr, v.n2, cs->hr); CALLOC(cori, v.n2 * 2, cs->cori);
  CALLOC(advmet, v.jmt * 2, cs->advmet); CALLOC(am3, v.jmt, cs->am3); CALLOC(am4, v.jmt * 2, cs->am4);
  CALLOC(dxmetr, v.imt, cs->dxmetr); CALLOC(dxu2r, v.imt, cs->dxu2r);
  CALLOC(dyu2r, v.jmt, cs->dyu2r); CALLOC(dyu4r, v.jmt, cs->dyu4r); CALLOC(csudyu2r, v.jmt, cs->csudyu2r);
  CALLOC(visc_ceu, v.n3, cs->visc_ceu); CALLOC(amc_north, v.n3, cs->amc_north); CALLOC(amc_south, v.n3, cs->amc_south);
  CALLOC(u_m1, v.n3 * 2, nullptr); CALLOC(u_p1, v.n3 * 2, nullptr);
  CALLOC(adv_veu, v.n3, nullptr); CALLOC(adv_vnu, v.n3, nullptr); CALLOC(adv_vbu, v.n3z, nullptr);
  CALLOC(smf, v.n2 * 2, nullptr); CALLOC(bmf, v.n2 * 2, nullptr); CALLOC(zu, v.n2 * 2, nullptr);
  CALLOC(grad_p, v.n3 * 2, nullptr); CALLOC(rho, v.n3, nullptr);
#undef CALLOC
  cv->kappa_m = cs->kappa_m; cv->cdbot = cs->cdbot; cv->grav_rho0r = cs->grav_rho0r;
  cv->jc0 = std::max(2, v.jlo);
  cv->jc1 = std::min(v.jmt - 1, v.jhi);
  if (cs->fourfil) {
    // O_fourfil: filuv (source/common/filuv.F) on the rows poleward of jfu1 / jfu2
    if (!cs->spsin || !cs->spcos || !cs->phi) return fail(ctx, "clinic_setup: fourfil needs spsin, spcos and phi");
    if (cs->jfrst < 1 || cs->jfu0 < 1 || cs->jfu0 > v.jmt || cs->jfu1 < 1 || cs->jfu2 > v.jmt || cs->jfu1 >= cs->jfu2)
      return fail(ctx, "clinic_setup: bad filter rows jfrst / jfu0 / jfu1 / jfu2");
    {
      double *p = nullptr;
      if (dev_alloc(ctx, "spsin", &p, (size_t)v.imt, cs->spsin)) return 1;
      cv->spsin = p;
      if (dev_alloc(ctx, "spcos", &p, (size_t)v.imt, cs->spcos)) return 1;
      cv->spcos = p;
    }
    std::vector<double> csu(v.jmt), csur(v.jmt);
    CK(cudaMemcpy(csu.data(), v.csu, sizeof(double) * v.jmt, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(csur.data(), v.csur, sizeof(double) * v.jmt, cudaMemcpyDeviceToHost));
    if (filuv_setup(ctx, cs->kmu, csu.data(), csur.data(), cs->phi, cs->jfrst, cs->jfu0, cs->jfu1, cs->jfu2, cv->jc0, cv->jc1))
      return fail(ctx, "clinic_setup: filuv_setup failed (device allocation)");
  }
  return 0;
}
int uvic_b200_upload_u_level(uvic_b200_ctx *ctx, int level, const double *u) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx || !u) return fail(ctx, "upload_u_level: null argument");
  double *dst = nullptr;
  if (level == 0) dst = ctx->v.u;
  else if (level == -1 && ctx->clinic) dst = ctx->clinic->u_m1;
  else return fail(ctx, "upload_u_level: level must be 0, or -1 after uvic_b200_clinic_setup");
  CK(cudaMemcpyAsync(dst, u, (size_t)ctx->v.n3 * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
int uvic_b200_download_u(uvic_b200_ctx *ctx, int level, double *u) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx || !u) return fail(ctx, "download_u: null argument");
  const double *src = nullptr;
  if (level == 0) src = ctx->v.u;
  else if (ctx->clinic && (level == -1 || level == 1)) src = level < 0 ? ctx->clinic->u_m1 : ctx->clinic->u_p1;
  else return fail(ctx, "download_u: level must be 0, or -1 / +1 after uvic_b200_clinic_setup");
  CK(cudaMemcpyAsync(u, src, (size_t)ctx->v.n3 * 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int uvic_b200_upload_smf(uvic_b200_ctx *ctx, const double *smf) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx || !ctx->clinic || !smf) return fail(ctx, "upload_smf: call uvic_b200_clinic_setup first");
  CK(cudaMemcpyAsync(ctx->clinic->smf, smf, (size_t)ctx->v.n2 * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
int uvic_b200_clinic(uvic_b200_ctx *ctx, double c2dtuv, int itaux, int itauy) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx || !ctx->clinic) return fail(ctx, "clinic: call uvic_b200_clinic_setup first");
  DevView &v = ctx->v;
  if (itaux != 0 || itauy != 0) {
    if (!v.sbc) return fail(ctx, "clinic: wind stress slots need uvic_b200_sbc_setup (pass 0, 0 to use uvic_b200_upload_smf)");
    if (itaux < 1 || itaux > v.numsbc || itauy < 1 || itauy > v.numsbc) return fail(ctx, "clinic: wind stress slot out of range");
  }
  ctx->clinic->c2dtuv = c2dtuv;
  ctx->vel_free_valid = false;   // the U-cell advective velocities below read adv_vet / adv_vnt / adv_vbt
  halo_wait_now(ctx);   // rho and grad_p read the halo rows of t(tau): a pending asynchronous exchange must have landed
  // 09/mom/loadmw.F:150-155: rho of t(tau).  The tau slot is addressed directly: the time-level view of the tracer step
  // (which maps tau-1 onto tau on mixing steps) is left as the tracer entry points set it
  launch_state(ctx, ctx->t_slot[ctx->lev[1]], ctx->clinic->rho);
  launch_setvbc_mom(ctx, itaux, itauy);
  launch_clinic(ctx);
  CK(cudaGetLastError());
  return 0;
}
int uvic_b200_download_zu(uvic_b200_ctx *ctx, double *zu) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx || !ctx->clinic || !zu) return fail(ctx, "download_zu: call uvic_b200_clinic_setup first");
  CK(cudaMemcpyAsync(zu, ctx->clinic->zu, (size_t)ctx->v.n2 * 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int uvic_b200_rotate_u(uvic_b200_ctx *ctx) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx || !ctx->clinic) return fail(ctx, "rotate_u: call uvic_b200_clinic_setup first");
  ClinicView *cv = ctx->clinic;
  double *old_m1 = cv->u_m1;
  cv->u_m1 = ctx->v.u;
  ctx->v.u = cv->u_p1;
  cv->u_p1 = old_m1;
  return 0;
}

int uvic_b200_rotate(uvic_b200_ctx *ctx) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  // tau+1 overwrites the old tau-1 slot next step (source/mom/mom.F:210-212)
  int old_m1 = ctx->lev[0];
  ctx->lev[0] = ctx->lev[1];
  ctx->lev[1] = ctx->lev[2];
  ctx->lev[2] = old_m1;
  set_levels(ctx, true);
  return 0;
}

static bool same_step(const uvic_b200_stepinfo &a, const uvic_b200_stepinfo &b) {
  // relyr: a host that forms "the next step's model time" itself (relyr + dtts / year length) and the model's own clock
  // differ in the last bits; 1e-9 years = 0.03 s of a 1.25 day step is the same step (the sources are then those of the
  // hinted time: the declination differs by 1e-9 relative, far below the 1e-10 gate of the sources)
  return a.dtts == b.dtts && a.leapfrog == b.leapfrog && fabs(a.relyr - b.relyr) <= 1e-9 && a.co2ccn == b.co2ccn;
}

// MOBI of this step.  If the look-ahead of the previous step already produced the sources for exactly this step
// (same stepinfo, same t(tau-1) slot, nothing uploaded in between) they are adopted; otherwise MOBI is launched on the
// side stream now.  Either way the main stream waits for mobi_event right before the first sourced k_invtri.
static int lookahead_mobi(uvic_b200_ctx *ctx);
static void begin_mobi_now(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *si);

static void begin_mobi(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *si) {
  if (!ctx->par.mobi || ctx->mobi_inflight) return;
  begin_mobi_now(ctx, si);
  // A hinted leapfrog step reads today's t(tau) as its t(tau-1): that field is final already, so its MOBI is queued
  // right behind this step's MOBI instead of behind this step's tracer kernels (a hinted mixing step needs t(tau+1)
  // and is launched at the end of the step, uvic_b200_tracer).  The side stream then runs a whole step ahead and the
  // implicit solve never waits for it.
  static const bool early = !(getenv("UVIC_B200_MOBI_EARLY") && atoi(getenv("UVIC_B200_MOBI_EARLY")) == 0);   // A/B switch
  if (early && ctx->hint_valid && ctx->hint_si.leapfrog && !ctx->prof_on) lookahead_mobi(ctx);
}

static void begin_mobi_now(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *si) {
  if (ctx->ahead_valid) {
    const bool hit = !ctx->prof_on && same_step(*si, ctx->ahead_si) && ctx->ahead_tm1 == ctx->v.t_m1;
    ctx->ahead_valid = false;
    if (!ctx->prof_on) (hit ? ctx->la_hits : ctx->la_misses)++;
    if (hit) {
      ctx->src_cur = ctx->ahead_buf;
      ctx->v.src = ctx->src_buf[ctx->src_cur];
      ctx->mobi_inflight = true;
      return;
    }
    // mispredicted: the side stream may still be writing the other buffer; order this step's MOBI behind it (same stream)
  }
  if (ctx->prof_on) {
    // per-kernel timing: run MOBI in line so that no two kernels share the SMs and every
    // CUDA-event interval is the duration of exactly one kernel
    cudaStreamWaitEvent(ctx->stream, ctx->mobi_event, 0);   // a look-ahead still in flight uses the same scratch
    launch_mobi(ctx, ctx->v, si);
    cudaEventRecord(ctx->mobi_event, ctx->stream);
    cudaEventRecord(ctx->src_ready[ctx->src_cur], ctx->stream);
    ctx->mobi_inflight = true;
    return;
  }
  cudaEventRecord(ctx->fork_event, ctx->stream);
  cudaStreamWaitEvent(ctx->stream2, ctx->fork_event, 0);
  cudaStream_t main_stream = ctx->stream;
  ctx->stream = ctx->stream2;
  launch_mobi(ctx, ctx->v, si);
  ctx->stream = main_stream;
  cudaEventRecord(ctx->mobi_event, ctx->stream2);
  cudaEventRecord(ctx->src_ready[ctx->src_cur], ctx->stream2);
  ctx->mobi_inflight = true;
}

// MOBI of the NEXT step (hinted), queued on the side stream behind everything the main stream holds at this point
static int lookahead_mobi(uvic_b200_ctx *ctx) {
  if (!ctx->hint_valid) return 0;
  ctx->hint_valid = false;
  if (!ctx->par.mobi || ctx->prof_on) return 0;
  if (!ctx->src_buf[1]) {
    double *p = nullptr;
    CK(cudaMalloc((void **)&p, (size_t)ctx->v.n3 * (std::max(ctx->v.nsrc, 1) + 1) * sizeof(double)));   // + the scratch slot of a MOBI subset
    CK(cudaMemsetAsync(p, 0, (size_t)ctx->v.n3 * (std::max(ctx->v.nsrc, 1) + 1) * sizeof(double), ctx->stream2));
    ctx->src_buf[1] = p;
    ctx->owned.push_back(p);
  }
  const uvic_b200_stepinfo &h = ctx->hint_si;
  DevView vv = ctx->v;
  vv.dtts = h.dtts;
  vv.c2dtts = h.leapfrog ? 2.0 * h.dtts : h.dtts;
  // after the rotation tau-1 of the next step is today's tau; a mixing step reads today's tau+1 instead
  vv.t_m1 = h.leapfrog ? ctx->t_slot[ctx->lev[1]] : ctx->t_slot[ctx->lev[2]];
  const int other = ctx->src_cur ^ 1;
  vv.src = ctx->src_buf[other];
  CK(cudaEventRecord(ctx->main_done_event, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->stream2, ctx->main_done_event, 0));
  cudaStream_t main_stream = ctx->stream;
  ctx->stream = ctx->stream2;
  launch_mobi(ctx, vv, &h);
  ctx->stream = main_stream;
  CK(cudaEventRecord(ctx->mobi_event, ctx->stream2));
  CK(cudaEventRecord(ctx->src_ready[other], ctx->stream2));
  ctx->ahead_valid = true;
  ctx->ahead_si = h;
  ctx->ahead_tm1 = vv.t_m1;
  ctx->ahead_buf = other;
  return 0;
}

static void set_step(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *si) {
  ctx->v.dtts = si->dtts;
  ctx->v.c2dtts = si->leapfrog ? 2.0 * si->dtts : si->dtts;  // source/mom/mom.F:111-146
  set_levels(ctx, si->leapfrog != 0);
}

int uvic_b200_set_host_window(uvic_b200_ctx *ctx, int jrow_first) {
  if (!ctx) return 1;
  if (jrow_first < 1 || jrow_first > ctx->v.jbase + 2) return fail(ctx, "set_host_window: the host arrays must start at global row 1 or jsmw = 2");
  ctx->host_jfirst = jrow_first;
  return 0;
}
int uvic_b200_lookahead_stats(uvic_b200_ctx *ctx, int64_t *hits, int64_t *misses) {
  if (!ctx) return 1;
  if (hits) *hits = ctx->la_hits;
  if (misses) *misses = ctx->la_misses;
  return 0;
}
// A driver that writes t, the forcing or the vertical b.c. through raw device pointers (uvic_b200_t_ptr,
// uvic_b200_device_ptr) after a step calls this: sources computed ahead from the old contents are dropped.
int uvic_b200_invalidate_lookahead(uvic_b200_ctx *ctx) {
  if (!ctx) return 1;
  if (ctx->ahead_valid && ctx->mobi_event) CK(cudaStreamWaitEvent(ctx->stream, ctx->mobi_event, 0));
  ctx->ahead_valid = false;
  ctx->hint_valid = false;
  return 0;
}
// Orders the caller's stream behind everything the library has queued on its side streams (the look-ahead MOBI, the GM
// velocities): an event recorded on the launch stream afterwards covers ALL the work of the steps issued so far.
int uvic_b200_join_streams(uvic_b200_ctx *ctx) {
  if (!ctx) return 1;
  if (ctx->mobi_event) CK(cudaStreamWaitEvent(ctx->stream, ctx->mobi_event, 0));
  if (ctx->ev_gm) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_gm, 0));
  return 0;
}
int uvic_b200_hint_next_step(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *next) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx) return 1;
  ctx->hint_valid = next != nullptr;
  if (next) ctx->hint_si = *next;
  return 0;
}

// The halo rows of the newest time level are first read by the advection kernels of a leapfrog step (t(tau), two rows
// either side); everything before them works on t(tau-1).  A slab driver can therefore let its exchange of t(tau+1) run
// beside the coefficient / diffusion kernels of the next step: it hands the event that marks the end of the exchange to
// the library, which waits for it right before the first advection kernel -- or at once when the step is a mixing
// step (t(tau-1) := t(tau)) or is driven call site by call site.  The event is consumed by the wait.
int uvic_b200_wait_before_advection(uvic_b200_ctx *ctx, void *cuda_event) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx) return 1;
  ctx->halo_event = (cudaEvent_t)cuda_event;
  return 0;
}
static void halo_wait_now(uvic_b200_ctx *ctx) {
  if (!ctx->halo_event) return;
  cudaStreamWaitEvent(ctx->stream, ctx->halo_event, 0);
  ctx->halo_event = nullptr;
}

int uvic_b200_isopyc(uvic_b200_ctx *ctx) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  halo_wait_now(ctx);
  launch_isopyc(ctx);
  CK(cudaGetLastError());
  return 0;
}
int uvic_b200_vmixc(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *si) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  set_step(ctx, si);
  if (!si->leapfrog) halo_wait_now(ctx);
  launch_vmixc(ctx);
  CK(cudaGetLastError());
  return 0;
}
int uvic_b200_tracer(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *si) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  set_step(ctx, si);
  if (!si->leapfrog) halo_wait_now(ctx);
  begin_mobi(ctx, si);
  launch_tracer(ctx, si);
  ctx->mobi_inflight = false;
  CK(cudaGetLastError());
  if (si->diag) {
    launch_tbar(ctx);
    launch_travar_dtabs(ctx, ctx->travar, ctx->dtabs);
    if (ctx->v.mskhr) launch_sumbk(ctx);
    CK(cudaGetLastError());
  }
  return lookahead_mobi(ctx);
}
int uvic_b200_step(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *si) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  set_step(ctx, si);
  if (!si->leapfrog || ctx->prof_on) halo_wait_now(ctx);   // a mixing step reads the newest level from its first kernel on
  begin_mobi(ctx, si);
  launch_isopyc(ctx);
  CK(cudaGetLastError());
  if (uvic_b200_vmixc(ctx, si)) return 1;
  return uvic_b200_tracer(ctx, si);   // launch_tracer waits for a pending halo event before its first advection kernel
}

int uvic_b200_tracer_step(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *si, const double *t_taum1, const double *t_tau,
                          const double *adv_vet, const double *adv_vnt, const double *adv_vbt, const double *stf,
                          const double *btf, double *t_taup1) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  DevView &v = ctx->v;
  static const bool trace = getenv("UVIC_B200_E2E_TRACE") != nullptr;   // phase times of the call on stderr (diagnostics)
  cudaEvent_t *tev = ctx->trace_ev;      // per context: events belong to one device
  if (trace && !tev[0])
    for (int q = 0; q < 5; q++) cudaEventCreate(&tev[q]);
  if (trace) cudaEventRecord(tev[0], ctx->stream);
  halo_wait_now(ctx);
  if (t_taum1 && uvic_b200_upload_t(ctx, -1, t_taum1)) return 1;
  if (t_tau && uvic_b200_upload_t(ctx, 0, t_tau)) return 1;
  // velocities and vertical b.c. travel on the input copy stream while the kernels that do not need them
  // (MOBI, Redi / GM coefficients, vmixc) already run; the copy stream first waits for everything queued so far
  // on the main stream (the previous step still reads these arrays)
  CK(cudaEventRecord(ctx->fork_event, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->copy_in, ctx->fork_event, 0));
  // the vertical b.c. first (the diffusion pass needs them), then the velocities (only the FCT needs them: launch_tracer
  // waits for them, and runs the velocity part of isopyc, right before its first advection kernel)
  if (stf) CK(cudaMemcpyAsync(v.stf, stf, (size_t)v.n2 * v.nt * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_in));
  if (btf) CK(cudaMemcpyAsync(v.btf, btf, (size_t)v.n2 * v.nt * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_in));
  CK(cudaEventRecord(ctx->h2d_vbc_event, ctx->copy_in));
  if (adv_vet) H2D_VEL(v.adv_vet, adv_vet, (long long)v.imt * v.km, v.n3, ctx->copy_in);
  if (adv_vnt) H2D_VEL(v.adv_vnt, adv_vnt, (long long)v.imt * v.km, v.n3, ctx->copy_in);
  if (adv_vbt) H2D_VEL(v.adv_vbt, adv_vbt, (long long)v.imt * (v.km + 1), v.n3z, ctx->copy_in);
  CK(cudaEventRecord(ctx->h2d_event, ctx->copy_in));
  if (trace) cudaEventRecord(tev[1], ctx->copy_in);
  set_step(ctx, si);
  begin_mobi(ctx, si);
  launch_isopyc_coef(ctx);
  launch_isopyc_vel_after(ctx, ctx->h2d_event);   // on the GM side stream, as soon as the velocities have arrived
  launch_vmixc(ctx);
  if (trace) cudaEventRecord(tev[2], ctx->stream);
  CK(cudaStreamWaitEvent(ctx->stream, ctx->h2d_vbc_event, 0));
  CK(cudaGetLastError());
  // finished tracer batches stream to the host while the next batch computes (launch_tracer)
  ctx->d2h_dst = t_taup1;
  ctx->d2h_ntr = ctx->v.nt;
  int rc = uvic_b200_tracer(ctx, si);
  ctx->d2h_dst = nullptr;
  if (rc) return rc;
  if (trace) { cudaEventRecord(tev[3], ctx->stream); cudaEventRecord(tev[4], ctx->copy_out); }
  CK(cudaStreamSynchronize(ctx->copy_out));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  if (trace) {
    float a = 0, b = 0, c2 = 0, d = 0;
    cudaEventElapsedTime(&a, tev[0], tev[1]); cudaEventElapsedTime(&b, tev[0], tev[2]);
    cudaEventElapsedTime(&c2, tev[0], tev[3]); cudaEventElapsedTime(&d, tev[0], tev[4]);
    fprintf(stderr, "[uvic_b200 e2e] h2d done %.3f ms, coefficients done %.3f, kernels done %.3f, d2h done %.3f\n", a, b, c2, d);
  }
  return 0;
}

// One ocean step with the surface boundary conditions on the device: what crosses PCIe is what the rest of the model
// exchanges with the tracer step -- the advective velocities from clinic and (when they changed) the coupler's flux
// array in; T and S of t(tau+1) for the density in clinic / loadmw every step and the surface accumulators at the end
// of an ocean segment out.  The other tracers stay resident (uvic_b200_download_tracer fetches them for output steps).
int uvic_b200_tracer_step_coupled(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *si, const double *adv_vet, const double *adv_vnt,
                                  const double *adv_vbt, const double *sbc_in, const double *bhf, int eots, int osegs, int osege,
                                  int ntspos, double *ts_taup1, double *sbc_out) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  DevView &v = ctx->v;
  if (!v.sbc) return fail(ctx, "tracer_step_coupled: call uvic_b200_sbc_setup first");
  if (ntspos < 1) return fail(ctx, "tracer_step_coupled: ntspos must be >= 1");
  halo_wait_now(ctx);
  // The coupler's array and the bottom heat flux are read by setvbc / written by set_sbc until the end of the previous
  // step: their upload (start of an ocean segment only) waits for everything queued so far.  The velocities do not:
  // with the FCT their device copies are free as soon as the previous step formed its total velocities (vel_free), so
  // this step's upload runs while the previous step's tracer kernels are still at work.
  const bool early_vel = ctx->vel_free_valid && v.fct && !sbc_in && !bhf;
  if (early_vel) {
    CK(cudaStreamWaitEvent(ctx->copy_in, ctx->vel_free, 0));
  } else {
    CK(cudaEventRecord(ctx->fork_event, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_in, ctx->fork_event, 0));
  }
  ctx->vel_free_valid = false;
  if (sbc_in) CK(cudaMemcpyAsync(v.sbc, sbc_in, (size_t)v.n2 * v.numsbc * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_in));
  if (bhf) CK(cudaMemcpyAsync(v.bhf, bhf, (size_t)v.n2 * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_in));
  CK(cudaEventRecord(ctx->h2d_vbc_event, ctx->copy_in));
  if (adv_vet) H2D_VEL(v.adv_vet, adv_vet, (long long)v.imt * v.km, v.n3, ctx->copy_in);
  if (adv_vnt) H2D_VEL(v.adv_vnt, adv_vnt, (long long)v.imt * v.km, v.n3, ctx->copy_in);
  if (adv_vbt) H2D_VEL(v.adv_vbt, adv_vbt, (long long)v.imt * (v.km + 1), v.n3z, ctx->copy_in);
  CK(cudaEventRecord(ctx->h2d_event, ctx->copy_in));
  set_step(ctx, si);
  begin_mobi(ctx, si);
  launch_isopyc_coef(ctx);
  launch_isopyc_vel_after(ctx, ctx->h2d_event);        // on the GM side stream, as soon as the velocities have arrived
  launch_vmixc(ctx);
  CK(cudaStreamWaitEvent(ctx->stream, ctx->h2d_vbc_event, 0));
  launch_setvbc(ctx);                                  // call setvbc, source/mom/mom.F:360
  CK(cudaGetLastError());
  ctx->d2h_dst = ts_taup1;
  ctx->d2h_ntr = 2;
  int rc = uvic_b200_tracer(ctx, si);
  ctx->d2h_dst = nullptr;
  ctx->d2h_ntr = v.nt;
  if (rc) return rc;
  launch_set_sbc(ctx, eots, osegs, osege, ntspos);     // 09/mom/tracer.F:1270-1288
  CK(cudaGetLastError());
  // Complete for the outputs requested on return: the host buffers of the velocities have been read, T and S
  // of t(tau+1) are on the host; the coupler's array too on the step that returns it.  The tracers that stay resident may
  // still be in flight -- every later call is ordered behind them on the context's stream -- so the host's own work between
  // two tracer steps (clinic, tropic, the coupler) and the next step's upload overlap the rest of this step.
  CK(cudaEventSynchronize(ctx->h2d_event));
  CK(cudaStreamSynchronize(ctx->copy_out));
  if (sbc_out && eots && osege) {
    CK(cudaMemcpyAsync(sbc_out, v.sbc, (size_t)v.n2 * v.numsbc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  CK(cudaGetLastError());
  return 0;
}

int uvic_b200_pin_host(void *host, size_t bytes) {
  // page-lock a host array (e.g. the COMMON block the Fortran shim passes every step) so that the copies above are
  // truly asynchronous; returns 0 when the range is already registered
  cudaError_t e = cudaHostRegister(host, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
  return e == cudaSuccess ? 0 : 1;
}
int uvic_b200_unpin_host(void *host) { return cudaHostUnregister(host) == cudaSuccess ? 0 : 1; }

int uvic_b200_inventory(uvic_b200_ctx *ctx, int level, double *out) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (level < -1 || level > 1) return fail(ctx, "inventory: bad level");
  launch_inventory(ctx, ctx->t_slot[ctx->lev[lev_index(level)]], ctx->red_out);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, ctx->red_out, (size_t)ctx->v.nt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int uvic_b200_tbar(uvic_b200_ctx *ctx, double *h) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  size_t n = (size_t)ctx->v.km * ctx->v.nt * (ctx->v.jhi - ctx->v.jlo + 1);
  CK(cudaMemcpyAsync(h, ctx->tbar, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int uvic_b200_travar_dtabs(uvic_b200_ctx *ctx, double *travar_host, double *dtabs_host) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  size_t n = (size_t)ctx->v.km * ctx->v.nt * (ctx->v.jhi - ctx->v.jlo + 1);
  if (travar_host) CK(cudaMemcpyAsync(travar_host, ctx->travar, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (dtabs_host) CK(cudaMemcpyAsync(dtabs_host, ctx->dtabs, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int uvic_b200_sumbk(uvic_b200_ctx *ctx, double *h) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  CK(cudaMemcpyAsync(h, ctx->sumbk, (size_t)3 * ctx->v.km * ctx->v.nt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

void *uvic_b200_device_ptr(uvic_b200_ctx *ctx, const char *name, size_t *nelem) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  for (auto &a : ctx->arrs)
    if (a.name == name) {
      if (nelem) *nelem = a.nelem;
      return *a.slot;
    }
  return nullptr;
}
int uvic_b200_fetch(uvic_b200_ctx *ctx, const char *name, double *host, size_t *nelem) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  size_t n = 0;
  void *p = uvic_b200_device_ptr(ctx, name, &n);
  if (!p) return fail(ctx, std::string("fetch: unknown array ") + name);
  if (nelem) *nelem = n;
  if (!host) return 0;
  gm_join(ctx);
  for (auto &a : ctx->arrs)
    if (a.name == name && a.is_int) return fail(ctx, "fetch: integer arrays are not fetchable as double");
  CK(cudaMemcpyAsync(host, p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
void *uvic_b200_t_ptr(uvic_b200_ctx *ctx, int level) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (level < -1 || level > 1) return nullptr;
  return ctx->t_slot[ctx->lev[lev_index(level)]];
}
int64_t uvic_b200_kernel_launches(const uvic_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

int uvic_b200_profile_enable(uvic_b200_ctx *ctx, int on) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx) return 1;
  ctx->prof_on = on != 0;
  return 0;
}
// drain pending event pairs into the per-kernel totals
static int prof_drain(uvic_b200_ctx *ctx) {
  CK(cudaStreamSynchronize(ctx->stream));
  for (auto &r : ctx->prof_pending) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, r.a, r.b));
    ctx->prof_ms[r.id] += ms;
    ctx->prof_count[r.id] += 1;
    ctx->prof_free.push_back(r.a);
    ctx->prof_free.push_back(r.b);
  }
  ctx->prof_pending.clear();
  return 0;
}
int uvic_b200_profile_count(uvic_b200_ctx *ctx) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx || prof_drain(ctx)) return -1;
  return (int)ctx->prof_names.size();
}
int uvic_b200_profile_get(uvic_b200_ctx *ctx, int idx, char *name, int name_len, double *total_ms, int64_t *count) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx || idx < 0 || idx >= (int)ctx->prof_names.size()) return 1;
  if (name && name_len > 0) {
    strncpy(name, ctx->prof_names[idx].c_str(), name_len - 1);
    name[name_len - 1] = 0;
  }
  if (total_ms) *total_ms = ctx->prof_ms[idx];
  if (count) *count = ctx->prof_count[idx];
  return 0;
}
int uvic_b200_profile_reset(uvic_b200_ctx *ctx) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx || prof_drain(ctx)) return 1;
  for (auto &x : ctx->prof_ms) x = 0.0;
  for (auto &x : ctx->prof_count) x = 0;
  return 0;
}
int uvic_b200_local_rows(const uvic_b200_ctx *ctx, int32_t *jbase, int32_t *jl) {
  if (ctx) cudaSetDevice(ctx->device);   // one host thread may drive several devices
  if (!ctx) return 1;
  if (jbase) *jbase = ctx->v.jbase;
  if (jl) *jl = ctx->v.jl;
  return 0;
}

}  // extern "C"
