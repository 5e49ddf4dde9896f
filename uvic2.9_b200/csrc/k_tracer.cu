// k_tracer.cu -- the per-tracer part of `call tracer` (source/mom/mom.F:389 ->
// 09/mom/tracer.F:902-1203) on the device, all nt tracers batched per launch:
//
//   k_fct_tlo     low-order (upstream) fluxes and the low-order solution t_lo
//                 09/mom/tracer_adv_flx.F:496-580
//   k_fct_rfac    raw antidiffusive fluxes and the one-dimensional Zalesak ratios
//                 R+-x, R+-y, R+-z of every cell            :582-712, 714-770, 786-958
//   k_update      delimited + low-order advective fluxes (:696-712,772-784,960-1002),
//                 explicit horizontal/vertical diffusion (09/mom/tracer.F:930-961,
//                 1025-1032), isopycnal fluxes (isoflux, 09/mom/isopyc.F:923-1108),
//                 vertical b.c. (tracer.F:1053-1067), source, explicit update
//                 (:1109-1130 with source/mom/fdift.h)
//   k_invtri      implicit vertical diffusion, per-tracer sweeps of the Thomas solve
//                 (source/mom/invtri.F:75-110) using the factors from k_vmix_column,
//                 then the cyclic boundary (setbcx, tracer.F:1153-1155)
//   k_convect     convct2 full convective adjustment (source/mom/convect.F:99-311)
//
// The reference's j loop in adv_flux looks sequential (iteration j limits anti_fn(j) with
// R+-Y(j) from the previous iteration) but is algebraically parallel; what must be kept
// are its boundary rules: R+-Y(row 1) = anti_fn(row 1) = 0 (:467-482), the clamps
// jp1=min(j+1,jmt-1), jp2=min(j+2,jmt) (:554-556) -- which only touch row jmt, whose
// ratios are zero because tmask(row jmt)=0 -- and the cyclic wrap of R+-x (:693-694).
#include "ctx.h"

struct TrPtr {
  const double *tm1;  // t(tau-1) of this tracer
  const double *t0;   // t(tau)
  double *tp1;        // t(tau+1)
};

__device__ __forceinline__ double tmk(const DevView &v, int i, int k, int j) { return (v.kmt[X2(i, j)] >= k) ? 1.0 : 0.0; }
__device__ __forceinline__ int wrap_i(const DevView &v, int i) {
  if (i < 2) return i + (v.imt - 2);
  if (i > v.imt - 1) return i - (v.imt - 2);
  return i;
}

// ---- low-order (upstream) fluxes, 09/mom/tracer_adv_flx.F:496-548 ----
__device__ __forceinline__ double lowfe(const DevView &v, const double *tm1, int i, int k, int j) {
  double totadv = v.ue[X3(i, k, j)];
  double a = tm1[X3(i, k, j)], b = tm1[X3(i + 1, k, j)];
  return totadv * (a + b) + fabs(totadv) * (a - b);
}
__device__ __forceinline__ double lowfn(const DevView &v, const double *tm1, int i, int k, int j) {
  double totadv = v.vn[X3(i, k, j)];
  double a = tm1[X3(i, k, j)], b = tm1[X3(i, k, j + 1)];
  return totadv * (a + b) + fabs(totadv) * (a - b);
}
__device__ __forceinline__ double lowfb(const DevView &v, const double *tm1, int i, int k, int j) {
  if (k == 0) return v.wb[X3Z(i, 0, j)] * 2.0 * tm1[X3(i, 1, j)];
  if (k >= v.km) return 0.0;
  double totadv = v.wb[X3Z(i, k, j)];
  double a = tm1[X3(i, k + 1, j)], b = tm1[X3(i, k, j)];
  return totadv * (a + b) + fabs(totadv) * (a - b);
}
// ---- raw antidiffusive fluxes, :582-620 ----
__device__ __forceinline__ double antife(const DevView &v, const TrPtr &p, int i, int k, int j) {
  return v.ue[X3(i, k, j)] * (p.t0[X3(i, k, j)] + p.t0[X3(i + 1, k, j)]) - lowfe(v, p.tm1, i, k, j);
}
__device__ __forceinline__ double antifn(const DevView &v, const TrPtr &p, int i, int k, int j) {
  if (j < 2) return 0.0;  // anti_fn(i,k,1,n) = c0 (:475)
  return v.vn[X3(i, k, j)] * (p.t0[X3(i, k, j)] + p.t0[X3(i, k, j + 1)]) - lowfn(v, p.tm1, i, k, j);
}
__device__ __forceinline__ double antifb(const DevView &v, const TrPtr &p, int i, int k, int j) {
  if (k == 0) return v.wb[X3Z(i, 0, j)] * 2.0 * p.tm1[X3(i, 1, j)];
  if (k >= v.km) return 0.0;
  return v.wb[X3Z(i, k, j)] * (p.t0[X3(i, k, j)] + p.t0[X3(i, k + 1, j)]) - lowfb(v, p.tm1, i, k, j) * tmk(v, i, k, j);
}

__device__ __forceinline__ bool decode_cell(const DevView &v, long long idx, int jfirst, int nrow, int &i, int &k, int &j) {
  int ni = v.imt - 2;
  long long tot = (long long)ni * v.km * nrow;
  if (idx >= tot) return false;
  i = (int)(idx % ni) + 2;
  long long r = idx / ni;
  k = (int)(r % v.km) + 1;
  j = (int)(r / v.km) + jfirst;
  return true;
}

__device__ __forceinline__ TrPtr tracer_ptr(const DevView &v, int n0) {
  TrPtr p;
  p.tm1 = v.t_m1 + (long long)n0 * v.n3;
  p.t0 = v.t_0 + (long long)n0 * v.n3;
  p.tp1 = v.t_p1 + (long long)n0 * v.n3;
  return p;
}

// rows of t_lo / R: max(2,jlo-1) .. min(jmt-1,jhi+1)
__global__ void __launch_bounds__(256) k_fct_tlo(const DevView v, int nbase, int jfirst, int nrow) {
  int i, k, j;
  if (!decode_cell(v, (long long)blockIdx.x * blockDim.x + threadIdx.x, jfirst, nrow, i, k, j)) return;
  const int g = blockIdx.y;
  const TrPtr p = tracer_ptr(v, nbase + g);
  double *t_lo = v.t_lo + (long long)g * v.n3;
  // ADV_Tx, ADV_Ty, ADV_Tz of source/mom/fdift.h:25-39 on the low-order fluxes
  double cstdxt2r = v.cstr[j - 1] * v.dxtr[i - 1] * 0.5;  // 09/mom/tracer.F:240
  double tx = (lowfe(v, p.tm1, i, k, j) - lowfe(v, p.tm1, i - 1, k, j)) * cstdxt2r;
  double ty = (lowfn(v, p.tm1, i, k, j) - lowfn(v, p.tm1, i, k, j - 1)) * v.cstdyt2r[j - 1];
  double tz = (lowfb(v, p.tm1, i, k - 1, j) - lowfb(v, p.tm1, i, k, j)) * v.dzt2r[k - 1];
  double twodt = v.c2dtts * v.dtxcel[k - 1];
  double val = p.tm1[X3(i, k, j)] - twodt * (tx + ty + tz) * tmk(v, i, k, j);
  long long line = X3(1, k, j);
  t_lo[line + i - 1] = val;
  if (i == 2) t_lo[line + v.imt - 1] = val;
  if (i == v.imt - 1) t_lo[line] = val;
}

__device__ __forceinline__ void ratio(double c2dtts, double dcf, double flxlft, double flxrgt, double fxa, double fxb, double tlo,
                                      double m, double &rpl, double &rmn) {
  double trmax = fmax(fmax(fxa, fxb), tlo);
  double trmin = fmin(fmin(fxa, fxb), tlo);
  double pplus = c2dtts * dcf * (fmax(0.0, flxlft) - fmin(0.0, flxrgt));
  double pminus = c2dtts * dcf * (fmax(0.0, flxrgt) - fmin(0.0, flxlft));
  double qplus = trmax - tlo;
  double qminus = tlo - trmin;
  rpl = fmin(1., m * qplus / (pplus + UVIC_EPSLN));
  rmn = fmin(1., m * qminus / (pminus + UVIC_EPSLN));
}

__global__ void __launch_bounds__(256) k_fct_rfac(const DevView v, int nbase, int jfirst, int nrow) {
  int i, k, j;
  if (!decode_cell(v, (long long)blockIdx.x * blockDim.x + threadIdx.x, jfirst, nrow, i, k, j)) return;
  const int g = blockIdx.y;
  const TrPtr p = tracer_ptr(v, nbase + g);
  const double *t_lo = v.t_lo + (long long)g * v.n3;
  double *R = v.Rfac + (long long)g * 6 * v.n3;
  const long long c = X3(i, k, j);
  const double m = tmk(v, i, k, j);
  const double tlo = t_lo[c];
  const double t0c = p.t0[c];
  double rpl, rmn;

  // ---- x (:635-694) ----
  {
    double mw = tmk(v, i - 1, k, j), me = tmk(v, i + 1, k, j);
    double fxa = mw * (0.5 * (p.t0[X3(i - 1, k, j)] + t0c)) + (1.0 - mw) * tlo;
    double fxb = me * (0.5 * (t0c + p.t0[X3(i + 1, k, j)])) + (1.0 - me) * tlo;
    double dcf = v.cstr[j - 1] * v.dxtr[i - 1] * 0.5;
    ratio(v.c2dtts, dcf, antife(v, p, i - 1, k, j), antife(v, p, i, k, j), fxa, fxb, tlo, m, rpl, rmn);
    R[c] = rpl;
    R[c + v.n3] = rmn;
  }
  // ---- y (:714-770) ----
  {
    int jp2 = min(j + 1, v.jmt);
    double ms = tmk(v, i, k, j - 1), mn = tmk(v, i, k, jp2);
    double fxa = 0.5 * ms * (p.t0[X3(i, k, j - 1)] + t0c) + (1.0 - ms) * tlo;
    double fxb = 0.5 * mn * (t0c + p.t0[X3(i, k, jp2)]) + (1.0 - mn) * tlo;
    ratio(v.c2dtts, v.cstdyt2r[j - 1], antifn(v, p, i, k, j - 1), antifn(v, p, i, k, j), fxa, fxb, tlo, m, rpl, rmn);
    R[c + 2 * v.n3] = rpl;
    R[c + 3 * v.n3] = rmn;
  }
  // ---- z (:786-958) ----
  {
    double fxa, fxb;
    if (k > 1) {
      double mu = tmk(v, i, k - 1, j);
      fxa = 0.5 * mu * (p.t0[X3(i, k - 1, j)] + t0c) + (1.0 - mu) * tlo;
    } else {
      fxa = tlo;
    }
    if (k < v.km) {
      double md = tmk(v, i, k + 1, j);
      fxb = 0.5 * md * (t0c + p.t0[X3(i, k + 1, j)]) + (1.0 - md) * tlo;
    } else {
      fxb = tlo;
    }
    // flxlft = anti_fb(k), flxrgt = anti_fb(k-1)  (:792-793)
    ratio(v.c2dtts, v.dzt2r[k - 1], antifb(v, p, i, k, j), antifb(v, p, i, k - 1, j), fxa, fxb, tlo, m, rpl, rmn);
    R[c + 4 * v.n3] = rpl;
    R[c + 5 * v.n3] = rmn;
  }
}

__device__ __forceinline__ double delimit(double cpos, double cneg, double a) {
  // :706-711, 777-782, 972-977
  return 0.5 * ((cpos + cneg) * a + (cpos - cneg) * fabs(a));
}

// corrected 2*advective flux through the east face of cell column f (f = 1..imt-1)
__device__ __forceinline__ double advfe(const DevView &v, const TrPtr &p, const double *R, int f, int k, int j) {
  long long cl = X3(wrap_i(v, f), k, j), cr = X3(wrap_i(v, f + 1), k, j);
  double cpos = fmin(R[cr], R[cl + v.n3]);   // Cpos(i) = min(Rpl(i+1),Rmn(i))  (:698-701)
  double cneg = fmin(R[cl], R[cr + v.n3]);   // Cneg(i) = min(Rpl(i),Rmn(i+1))
  return delimit(cpos, cneg, antife(v, p, f, k, j)) + lowfe(v, p.tm1, f, k, j);
}
__device__ __forceinline__ double advfn(const DevView &v, const TrPtr &p, const double *R, int i, int k, int g) {
  // rows 1 and jmt carry zero ratios (row 1: :477-478; row jmt: tmask = 0)
  double rpl_s = 0.0, rmn_s = 0.0, rpl_n = 0.0, rmn_n = 0.0;
  if (g >= 2) { rpl_s = R[X3(i, k, g) + 2 * v.n3]; rmn_s = R[X3(i, k, g) + 3 * v.n3]; }
  if (g + 1 <= v.jmt - 1) { rpl_n = R[X3(i, k, g + 1) + 2 * v.n3]; rmn_n = R[X3(i, k, g + 1) + 3 * v.n3]; }
  double cpos = fmin(rpl_n, rmn_s);  // min(R_plusY(j+1),R_minusY(j))  (:772-775)
  double cneg = fmin(rpl_s, rmn_n);
  return (delimit(cpos, cneg, antifn(v, p, i, k, g)) + lowfn(v, p.tm1, i, k, g)) * tmk(v, i, k, g);
}
__device__ __forceinline__ double advfb(const DevView &v, const TrPtr &p, const double *R, int i, int h, int j) {
  if (h == 0) {
    double t1 = p.t0[X3(i, 1, j)];
    return v.adv_vbt[X3Z(i, 0, j)] * (t1 + t1);  // 09/mom/tracer.F:1063-1064
  }
  if (h == v.km) return v.adv_vbt[X3Z(i, v.km, j)] * p.t0[X3(i, v.km, j)];  // :1065
  long long cu = X3(i, h, j), cd = X3(i, h + 1, j);
  double cneg = fmin(R[cd + 4 * v.n3], R[cu + 5 * v.n3]);  // min(Rpl(k+1),Rmn(k))  (:966-969)
  double cpos = fmin(R[cu + 4 * v.n3], R[cd + 5 * v.n3]);
  return (delimit(cpos, cneg, antifb(v, p, i, h, j)) + lowfb(v, p.tm1, i, h, j)) * tmk(v, i, h, j);
}

// total diffusive flux through the east face of column f (f in 2..imt-1 after wrap):
// background (tracer.F:930-940) + K11 + off-diagonal Redi terms (isopyc.F:950-1002)
__device__ __forceinline__ double difffe(const DevView &v, const double *tm1, int f, int k, int j) {
  f = wrap_i(v, f);
  long long c = X3(f, k, j);
  double d = tm1[X3(f + 1, k, j)] - tm1[c];
  double cstdxur = v.cstr[j - 1] * v.dxur[f - 1];
  double fe = v.diff_cet * v.cstr[j - 1] * v.dxur[f - 1] * d;
  if (!v.isopycmix) return fe;
  double dzt4r = 0.5 * v.dzt2r[k - 1];
  double sumz = 0.0;
#pragma unroll
  for (int kr = 0; kr <= 1; kr++) {
    int km1kr = max(k - 1 + kr, 1), kpkr = min(k + kr, v.km);
#pragma unroll
    for (int ip = 0; ip <= 1; ip++)
      sumz = sumz - v.ce[c + (ip + 2 * kr) * v.n3] * (tm1[X3(f + ip, km1kr, j)] - tm1[X3(f + ip, kpkr, j)]);
  }
  double flux_x = dzt4r * sumz;
  return fe + v.K11[c] * cstdxur * d + flux_x;
}
__device__ __forceinline__ double difffn(const DevView &v, const double *tm1, int i, int k, int g) {
  long long c = X3(i, k, g);
  double d = tm1[X3(i, k, g + 1)] - tm1[c];
  double fn = v.diff_cnt * v.csu_dyur[g - 1] * d;
  if (!v.isopycmix) return fn;
  double csu_dzt4r = v.csu[g - 1] * 0.5 * v.dzt2r[k - 1];
  double sumz = 0.0;
#pragma unroll
  for (int kr = 0; kr <= 1; kr++) {
    int km1kr = max(k - 1 + kr, 1), kpkr = min(k + kr, v.km);
#pragma unroll
    for (int jq = 0; jq <= 1; jq++)
      sumz = sumz - v.cn[c + (jq + 2 * kr) * v.n3] * (tm1[X3(i, km1kr, g + jq)] - tm1[X3(i, kpkr, g + jq)]);
  }
  double flux_y = csu_dzt4r * sumz;
  return fn + v.K22[c] * v.csu_dyur[g - 1] * d + flux_y;
}
// vertical diffusive flux with the b.c. of tracer.F:1053-1062
__device__ __forceinline__ double difffb(const DevView &v, const double *tm1, int n0, int i, int h, int j, int kb) {
  if (h == kb) return v.btf[X2(i, j) + (long long)n0 * v.n2];
  if (h == 0) return v.stf[X2(i, j) + (long long)n0 * v.n2];
  if (h >= v.km) return 0.0;
  return v.diff_cbt[X3(i, h, j)] * v.dzwr[h] * (tm1[X3(i, h, j)] - tm1[X3(i, h + 1, j)]);
}
// K31, K32 part, solved explicitly (isopyc.F:1062-1108)
__device__ __forceinline__ double difffbiso(const DevView &v, const double *tm1, int i, int h, int j) {
  if (h == 0 || h >= v.km) return 0.0;
  long long c = X3(i, h, j);
  double sumx = 0.0, sumy = 0.0;
#pragma unroll
  for (int ip = 0; ip <= 1; ip++)
#pragma unroll
    for (int kr = 0; kr <= 1; kr++)
      sumx = sumx - v.cbx[c + (ip + 2 * kr) * v.n3] * (tm1[X3(i + ip, h + kr, j)] - tm1[X3(i - 1 + ip, h + kr, j)]);
#pragma unroll
  for (int jq = 0; jq <= 1; jq++)
#pragma unroll
    for (int kr = 0; kr <= 1; kr++)
      sumy = sumy - v.cby[c + (jq + 2 * kr) * v.n3] * (tm1[X3(i, h + kr, j + jq)] - tm1[X3(i, h + kr, j - 1 + jq)]);
  return v.dxt4r[i - 1] * sumx + v.dyt4r[j - 1] * v.cstr[j - 1] * sumy;
}

__global__ void __launch_bounds__(256) k_update(const DevView v, int nbase, int jfirst, int nrow) {
  int i, k, j;
  if (!decode_cell(v, (long long)blockIdx.x * blockDim.x + threadIdx.x, jfirst, nrow, i, k, j)) return;
  const int g = blockIdx.y;
  const int n0 = nbase + g;
  const TrPtr p = tracer_ptr(v, n0);
  const double *R = v.Rfac + (long long)g * 6 * v.n3;
  const long long c = X3(i, k, j);
  const int kb = v.kmt[X2(i, j)];
  const double m = (kb >= k) ? 1.0 : 0.0;

  // advective flux divergences (source/mom/fdift.h:25-39)
  double cstdxt2r = v.cstr[j - 1] * v.dxtr[i - 1] * 0.5;
  double cstdxtr = v.cstr[j - 1] * v.dxtr[i - 1];
  double adv_tx = (advfe(v, p, R, i, k, j) - advfe(v, p, R, i - 1, k, j)) * cstdxt2r;
  double adv_ty = (advfn(v, p, R, i, k, j) - advfn(v, p, R, i, k, j - 1)) * v.cstdyt2r[j - 1];
  double adv_tz = (advfb(v, p, R, i, k - 1, j) - advfb(v, p, R, i, k, j)) * v.dzt2r[k - 1];

  // diffusive flux divergences (fdift.h:61-88)
  double diff_tx = (difffe(v, p.tm1, i, k, j) * tmk(v, i + 1, k, j) - difffe(v, p.tm1, i - 1, k, j) * tmk(v, i - 1, k, j)) * cstdxtr;
  double diff_ty = (difffn(v, p.tm1, i, k, j) * tmk(v, i, k, j + 1) - difffn(v, p.tm1, i, k, j - 1) * tmk(v, i, k, j - 1)) * v.cstdytr[j - 1];
  double diff_tz;
  {
    double fb_u = difffb(v, p.tm1, n0, i, k - 1, j, kb), fb_d = difffb(v, p.tm1, n0, i, k, j, kb);
    if (v.isopycmix)
      diff_tz = (fb_u - fb_d) * v.dztr[k - 1] * (1.0 - v.aidif) +
                (difffbiso(v, p.tm1, i, k - 1, j) - difffbiso(v, p.tm1, i, k, j)) * v.dztr[k - 1];
    else
      diff_tz = (fb_u - fb_d) * v.dztr[k - 1];
  }
  double source = 0.0;
  int is = v.itrc[n0];
  if (is != 0) source = v.src[c + (long long)(is - 1) * v.n3];

  double twodt = v.c2dtts * v.dtxcel[k - 1];
  p.tp1[c] = p.tm1[c] + twodt * (diff_tx + diff_ty + diff_tz - adv_tx - adv_ty - adv_tz + source) * m;
}

// one thread per (column, tracer): source/mom/invtri.F:75-110 with precomputed a, e, bet
__global__ void __launch_bounds__(128) k_invtri(const DevView v, int nbase) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int ni = v.imt - 2;
  int nrow = v.jhi - v.jlo + 1;
  if (idx >= (long long)ni * nrow) return;
  int i = (int)(idx % ni) + 2;
  int j = (int)(idx / ni) + v.jlo;
  const int n0 = nbase + blockIdx.y;
  double *z = v.t_p1 + (long long)n0 * v.n3;
  const int km = v.km;
  const int kb = v.kmt[X2(i, j)];
  const int kbot = max(2, kb);
  double topbc = v.stf[X2(i, j) + (long long)n0 * v.n2];
  double botbc = v.btf[X2(i, j) + (long long)n0 * v.n2];
  double zprev = 0.0;
  for (int k = 1; k <= km; k++) {
    long long c = X3(i, k, j);
    double mk = (kb >= k) ? 1.0 : 0.0;
    double tdt = v.c2dtts * v.dtxcel[k - 1];
    double f = z[c] * mk;
    if (k == 1) f = z[c] + topbc * tdt * v.dztr[0] * v.aidif * mk;
    if (k == kbot) f = z[c] - botbc * tdt * v.dztr[k - 1] * v.aidif * mk;
    double zk;
    if (k == 1)
      zk = f * v.tri_bet[c];
    else
      zk = (f - v.tri_a[c] * zprev) * v.tri_bet[c];
    z[c] = zk;
    zprev = zk;
  }
  // back substitution + cyclic boundary
  double znext = zprev;
  {
    long long line = X3(1, km, j);
    if (i == 2) z[line + v.imt - 1] = znext;
    if (i == v.imt - 1) z[line] = znext;
  }
  for (int k = km - 1; k >= 1; k--) {
    long long c = X3(i, k, j);
    double zk = z[c] - v.tri_e[X3(i, k + 1, j)] * znext;
    z[c] = zk;
    znext = zk;
    long long line = X3(1, k, j);
    if (i == 2) z[line + v.imt - 1] = zk;
    if (i == v.imt - 1) z[line] = zk;
  }
}

// source/mom/dens.h:18-19
#define ECC(k, m) v.eosc[((k)-1) + v.km * ((m)-1)]
__device__ __forceinline__ double dens_f(const DevView &v, double tq, double sq, int k) {
  return (ECC(k, 1) + (ECC(k, 4) + ECC(k, 7) * sq) * sq + (ECC(k, 3) + ECC(k, 8) * sq + ECC(k, 6) * tq) * tq) * tq +
         (ECC(k, 2) + (ECC(k, 5) + ECC(k, 9) * sq) * sq) * sq;
}

// one thread per column: convct2 (source/mom/convect.F:99-311), all nt tracers
__global__ void __launch_bounds__(128) k_convect(const DevView v) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int ni = v.imt - 2;
  int nrow = v.jhi - v.jlo + 1;
  if (idx >= (long long)ni * nrow) return;
  int i = (int)(idx % ni) + 2;
  int j = (int)(idx / ni) + v.jlo;
  const int kbo = v.kmt[X2(i, j)];
  double *T = v.t_p1, *S = v.t_p1 + v.n3;
  const double *dz = v.dztxcl;
#define TSV(a, k) a[X3(i, k, j)]
  int kt = 1, kb = 2;
  while (kt < kbo) {
    double ru = dens_f(v, TSV(T, kt) - v.to[kb - 1], TSV(S, kt) - v.so[kb - 1], kb);
    double rl = dens_f(v, TSV(T, kb) - v.to[kb - 1], TSV(S, kb) - v.so[kb - 1], kb);
    if (ru > rl) {
      bool chk_la = true, chk_lb = true;
      double zsm = dz[kt - 1] + dz[kb - 1];
      double tsm1 = TSV(T, kt) * dz[kt - 1] + TSV(T, kb) * dz[kb - 1];
      double tmx1 = tsm1 / zsm;
      double tsm2 = TSV(S, kt) * dz[kt - 1] + TSV(S, kb) * dz[kb - 1];
      double tmx2 = tsm2 / zsm;
      while (chk_lb || chk_la) {
        if (kb >= kbo) chk_lb = false;
        while (chk_lb) {
          chk_lb = false;
          int lb = kb + 1;
          ru = dens_f(v, tmx1 - v.to[lb - 1], tmx2 - v.so[lb - 1], lb);
          rl = dens_f(v, TSV(T, lb) - v.to[lb - 1], TSV(S, lb) - v.so[lb - 1], lb);
          if (ru > rl) {
            kb = lb;
            zsm = zsm + dz[kb - 1];
            tsm1 = tsm1 + TSV(T, kb) * dz[kb - 1];
            tmx1 = tsm1 / zsm;
            tsm2 = tsm2 + TSV(S, kb) * dz[kb - 1];
            tmx2 = tsm2 / zsm;
            chk_la = true;
            if (kb < kbo) chk_lb = true;
          }
        }
        chk_la = true;  // Rahmstorf variant is the active line (convect.F:237)
        if (kt <= 1) chk_la = false;
        while (chk_la) {
          chk_la = false;
          int la = kt - 1;
          ru = dens_f(v, TSV(T, la) - v.to[kt - 1], TSV(S, la) - v.so[kt - 1], kt);
          rl = dens_f(v, tmx1 - v.to[kt - 1], tmx2 - v.so[kt - 1], kt);
          if (ru > rl) {
            kt = la;
            zsm = zsm + dz[kt - 1];
            tsm1 = tsm1 + TSV(T, kt) * dz[kt - 1];
            tmx1 = tsm1 / zsm;
            tsm2 = tsm2 + TSV(S, kt) * dz[kt - 1];
            tmx2 = tsm2 / zsm;
            chk_lb = true;
          }
        }
      }
      for (int k = kt; k <= kb; k++) {
        TSV(T, k) = tmx1;
        TSV(S, k) = tmx2;
      }
      for (int n = 3; n <= v.nt; n++) {
        double *X = v.t_p1 + (long long)(n - 1) * v.n3;
        double tsm3 = 0.0;
        for (int k = kt; k <= kb; k++) tsm3 = tsm3 + TSV(X, k) * dz[k - 1];
        double tmx3 = tsm3 / zsm;
        for (int k = kt; k <= kb; k++) TSV(X, k) = tmx3;
      }
      kt = kb + 1;
    } else {
      kt = kb;
    }
    kb = kt + 1;
  }
#undef TSV
  // cyclic boundary of every tracer (09/mom/tracer.F:1199-1203)
  if (i == 2 || i == v.imt - 1) {
    int iw = (i == 2) ? v.imt : 1;
    for (int n = 0; n < v.nt; n++) {
      double *X = v.t_p1 + (long long)n * v.n3;
      for (int k = 1; k <= v.km; k++) X[X3(iw, k, j)] = X[X3(i, k, j)];
    }
  }
}

void launch_tracer(uvic_b200_ctx *c, const uvic_b200_stepinfo *si) {
  DevView &v = c->v;
  (void)si;
  const int jf_r = max(2, v.jlo - 1), jl_r = min(v.jmt - 1, v.jhi + 1);
  const int nrow_r = jl_r - jf_r + 1;
  const int nrow_c = v.jhi - v.jlo + 1;
  const long long ncell_r = (long long)(v.imt - 2) * v.km * nrow_r;
  const long long ncell_c = (long long)(v.imt - 2) * v.km * nrow_c;
  const long long ncol = (long long)(v.imt - 2) * nrow_c;
  for (int nbase = 0; nbase < v.nt; nbase += v.ngroup) {
    int ng = min(v.ngroup, v.nt - nbase);
    if (v.fct) {
      dim3 gr(cdiv(ncell_r, 256), ng);
      KLAUNCH("k_fct_tlo", k_fct_tlo, gr, 256, v, nbase, jf_r, nrow_r);
      KLAUNCH("k_fct_rfac", k_fct_rfac, gr, 256, v, nbase, jf_r, nrow_r);
    }
    dim3 gc(cdiv(ncell_c, 256), ng);
    KLAUNCH("k_update", k_update, gc, 256, v, nbase, v.jlo, nrow_c);
    dim3 gi(cdiv(ncol, 128), ng);
    KLAUNCH("k_invtri", k_invtri, gi, 128, v, nbase);
  }
  if (c->par.fullconvect) {
    KLAUNCH("k_convect", k_convect, cdiv(ncol, 128), 128, v);
  }
}
