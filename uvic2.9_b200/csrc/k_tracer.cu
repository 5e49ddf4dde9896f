// k_tracer.cu -- the per-tracer part of `call tracer` (source/mom/mom.F:389 ->
// 09/mom/tracer.F:902-1203) on the device:
//
//   k_fct_rfac    low-order (upstream) fluxes and the low-order solution t_lo of the cell
//                 (09/mom/tracer_adv_flx.F:496-580), raw antidiffusive fluxes and the one-dimensional Zalesak ratios
//                 R+-x, R+-y, R+-z of every cell            :582-712, 714-770, 786-958
//   k_update      delimited + low-order advective fluxes (:696-712,772-784,960-1002),
//                 explicit horizontal/vertical diffusion (09/mom/tracer.F:930-961,
//                 1025-1032), isopycnal fluxes (isoflux, 09/mom/isopyc.F:923-1108),
//                 vertical b.c. (tracer.F:1053-1067), source, explicit update
//                 (:1109-1130 with source/mom/fdift.h)
//   k_invtri      implicit vertical diffusion, per-tracer sweeps of the Thomas solve
//                 (source/mom/invtri.F:75-110) using the factors from k_vmix_column,
//                 then the cyclic boundary (setbcx, tracer.F:1153-1155)
//   k_convect_ts / k_convect_tr   convct2 full convective adjustment
//                 (source/mom/convect.F:99-311): region search on T,S per column, then
//                 the mixing of the other tracers in parallel over (column, tracer)
//
// Thread mapping of the three flux kernels: one thread per cell (i fastest, so a warp
// reads 32 consecutive i), and each thread loops over a chunk of tracers.  Everything
// that does not depend on the tracer -- the face velocities, the 38 Redi / vertical
// coefficients of the six faces, masks and metric factors -- is loaded once into
// registers and reused for every tracer of the chunk; per tracer a thread reads its
// 15-point t(tau-1) neighbourhood, 7 points of t(tau), 18 ratios and the source.
//
// The reference's j loop in adv_flux looks sequential (iteration j limits anti_fn(j) with
// R+-Y(j) from the previous iteration) but is algebraically parallel; what must be kept
// are its boundary rules: R+-Y(row 1) = anti_fn(row 1) = 0 (:467-482), the clamps
// jp1=min(j+1,jmt-1), jp2=min(j+2,jmt) (:554-556) -- which only touch row jmt, whose
// ratios are zero because tmask(row jmt)=0 -- and the cyclic wrap of R+-x (:693-694).
#include "ctx.h"
#include "fct_common.h"
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <utility>
#include <vector>

// decoded cell: 1-based i,k, global j, and 32-bit offsets inside one 3-D field
struct Cell {
  int i, k, j;
  int c;        // (i,k,j) in an (imt,km,jl) field
  int cz;       // (i,k,j) in an (imt,0:km,jl) field
  int sk, sj;   // strides of k and j in an (imt,km,jl) field
  int c2;       // (i,j) in an (imt,jl) field
};

__device__ __forceinline__ bool decode_cell(const DevView &v, long long idx, int jfirst, int nrow, Cell &q) {
  const int ni = v.imt - 2;
  if (idx >= (long long)ni * v.km * nrow) return false;
  unsigned u = (unsigned)idx;
  unsigned r = u / (unsigned)ni;
  q.i = (int)(u - r * (unsigned)ni) + 2;
  unsigned jj = r / (unsigned)v.km;
  q.k = (int)(r - jj * (unsigned)v.km) + 1;
  q.j = (int)jj + jfirst;
  const int jloc = q.j - v.jbase;
  q.sk = v.imt;
  q.sj = v.imt * v.km;
  q.c = (q.i - 1) + v.imt * ((q.k - 1) + v.km * jloc);
  q.cz = q.c + v.imt * (jloc + 1);
  q.c2 = (q.i - 1) + v.imt * jloc;
  return true;
}

// The same cell record for a CTA that owns a tile of 32 consecutive i x (blockDim.x / 32) consecutive levels of one row
// (warp = level): the rows k-1, k, k+1 a warp reads are the rows its neighbour warps read, so they are fetched from L2
// once per CTA instead of once per warp.  Grid: upd_tiles(v) * nrow CTAs.
__device__ __forceinline__ bool decode_cell_tiled(const DevView &v, int jfirst, int nrow, int tiled, Cell &q) {
  // tiled = 1: all warps of the CTA stacked in k; 2: pairs of warps in k, the pairs side by side in j (experiment)
  const int nwj = (tiled == 2) ? 2 : 1;
  const int ni = v.imt - 2, nit = (ni + 31) >> 5, nw = (blockDim.x >> 5) / nwj, nkt = (v.km + nw - 1) / nw;
  unsigned b = blockIdx.x;
  const unsigned it = b % (unsigned)nit;
  b /= (unsigned)nit;
  const unsigned kt = b % (unsigned)nkt;
  const int w = threadIdx.x >> 5;
  const unsigned jj = (b / (unsigned)nkt) * nwj + (w / nw);
  q.i = 2 + (int)it * 32 + (threadIdx.x & 31);
  q.k = 1 + (int)kt * nw + (w % nw);
  q.j = (int)jj + jfirst;
  if ((int)jj >= nrow || q.i > v.imt - 1 || q.k > v.km) return false;
  const int jloc = q.j - v.jbase;
  q.sk = v.imt;
  q.sj = v.imt * v.km;
  q.c = (q.i - 1) + v.imt * ((q.k - 1) + v.km * jloc);
  q.cz = q.c + v.imt * (jloc + 1);
  q.c2 = (q.i - 1) + v.imt * jloc;
  return true;
}
static long long upd_tiles(const DevView &v, int threads, int tiled) {
  const int nw = threads / 32 / ((tiled == 2) ? 2 : 1);
  return (long long)((v.imt - 2 + 31) / 32) * ((v.km + nw - 1) / nw);
}

// ------------------------------------------------------------------------------------
// t_lo and R+-x, R+-y, R+-z, rows max(2,jlo-1) .. min(jmt-1,jhi+1)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fct_rfac(const DevView v, int nbase, int ng, int tch, int jfirst, int nrow) {
  Cell q;
  if (!decode_cell(v, (long long)blockIdx.x * blockDim.x + threadIdx.x, jfirst, nrow, q)) return;
  const int km = v.km, k = q.k, c = q.c, j = q.j;
  const double ue_c = v.ue[c], ue_w = v.ue[c - 1], vn_c = v.vn[c], vn_s = v.vn[c - q.sj];
  const double wb_d = v.wb[q.cz], wb_u = v.wb[q.cz - q.sk];
  const int kmc = v.kmt[q.c2];
  const double m = (kmc >= k) ? 1.0 : 0.0, mu = (kmc >= k - 1) ? 1.0 : 0.0, md = (kmc >= k + 1) ? 1.0 : 0.0;
  const double mw = (v.kmt[q.c2 - 1] >= k) ? 1.0 : 0.0, me = (v.kmt[q.c2 + 1] >= k) ? 1.0 : 0.0;
  const double ms = (v.kmt[q.c2 - v.imt] >= k) ? 1.0 : 0.0, mn = (v.kmt[q.c2 + v.imt] >= k) ? 1.0 : 0.0;  // jp2 = j+1 <= jmt
  const double dcfx = v.cstr[j - 1] * v.dxtr[q.i - 1] * 0.5, dcfy = v.cstdyt2r[j - 1], dcfz = v.dzt2r[k - 1];
  const double twodt = v.c2dtts * v.dtxcel[k - 1];
  const int cu = (k > 1) ? c - q.sk : c, cd = (k < km) ? c + q.sk : c;
  const int g0 = blockIdx.y * tch, g1 = min(g0 + tch, ng);
  for (int g = g0; g < g1; g++) {
    const double *__restrict__ T = v.t_m1 + (long long)(nbase + g) * v.n3;
    const double *__restrict__ U = v.t_0 + (long long)(nbase + g) * v.n3;
    double *__restrict__ R = v.Rfac + (long long)g * 6 * v.n3;
    const double Tc = T[c], Te = T[c + 1], Tw = T[c - 1], Tn = T[c + q.sj], Ts = T[c - q.sj], Tu = T[cu], Td = T[cd];
    const double Uc = U[c], Ue = U[c + 1], Uw = U[c - 1], Un = U[c + q.sj], Us = U[c - q.sj], Uu = U[cu], Ud = U[cd];
    // low-order solution of this cell (:496-580); only the cell's own t_lo enters its ratios, so it is formed here
    // instead of in a separate pass (its upstream fluxes are the ones the antidiffusive fluxes below need anyway)
    double tlo;
    {
      double tx = (upw(ue_c, Tc, Te) - upw(ue_w, Tw, Tc)) * dcfx;
      double ty = (upw(vn_c, Tc, Tn) - upw(vn_s, Ts, Tc)) * dcfy;
      double fb_u = (k == 1) ? wb_u * 2.0 * Tc : upw(wb_u, Tc, Tu);   // adv_fb(i,0,j) = adv_vbt(i,0,j)*c2*t(i,1,j) (:543)
      double fb_d = (k == km) ? 0.0 : upw(wb_d, Td, Tc);              // adv_fb(i,km,j) = c0 (:544)
      double tz = (fb_u - fb_d) * dcfz;
      tlo = Tc - twodt * (tx + ty + tz) * m;
    }
    double rpl, rmn;
    // ---- x (:635-694): flxlft = anti_fe(i-1), flxrgt = anti_fe(i) ----
    {
      double fxa = mw * (0.5 * (Uw + Uc)) + (1.0 - mw) * tlo;
      double fxb = me * (0.5 * (Uc + Ue)) + (1.0 - me) * tlo;
      double a_w = ue_w * (Uw + Uc) - upw(ue_w, Tw, Tc);
      double a_e = ue_c * (Uc + Ue) - upw(ue_c, Tc, Te);
      ratio(v.c2dtts, dcfx, a_w, a_e, fxa, fxb, tlo, m, rpl, rmn);
      R[c] = rpl;
      R[c + v.n3] = rmn;
    }
    // ---- y (:714-770): flxlft = anti_fn(j-1) (= 0 for row 1, :475), flxrgt = anti_fn(j) ----
    {
      double fxa = 0.5 * ms * (Us + Uc) + (1.0 - ms) * tlo;
      double fxb = 0.5 * mn * (Uc + Un) + (1.0 - mn) * tlo;
      double a_s = (j - 1 < 2) ? 0.0 : vn_s * (Us + Uc) - upw(vn_s, Ts, Tc);
      double a_n = vn_c * (Uc + Un) - upw(vn_c, Tc, Tn);
      ratio(v.c2dtts, dcfy, a_s, a_n, fxa, fxb, tlo, m, rpl, rmn);
      R[c + 2 * v.n3] = rpl;
      R[c + 3 * v.n3] = rmn;
    }
    // ---- z (:786-958): flxlft = anti_fb(k), flxrgt = anti_fb(k-1) ----
    {
      double fxa = (k > 1) ? 0.5 * mu * (Uu + Uc) + (1.0 - mu) * tlo : tlo;
      double fxb = (k < km) ? 0.5 * md * (Uc + Ud) + (1.0 - md) * tlo : tlo;
      // anti_fb(i,0,j) = adv_vbt(i,0,j)*c2*t(i,1,j,taum1) (:617); anti_fb(i,km,j) = 0
      double a_d = (k == km) ? 0.0 : wb_d * (Uc + Ud) - upw(wb_d, Td, Tc) * m;
      double a_u = (k == 1) ? wb_u * 2.0 * Tc : wb_u * (Uu + Uc) - upw(wb_u, Tc, Tu) * mu;
      ratio(v.c2dtts, dcfz, a_d, a_u, fxa, fxb, tlo, m, rpl, rmn);
      R[c + 4 * v.n3] = rpl;
      R[c + 5 * v.n3] = rmn;
    }
  }
}

// ------------------------------------------------------------------------------------
// fluxes + explicit update, rows jlo..jhi
// ------------------------------------------------------------------------------------
#ifndef UPD_T
#define UPD_T 128   // threads per CTA of k_update
#endif
__device__ __forceinline__ void pf_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// per-thread column of the shared-memory coefficient table: element a lives at p[a*UPD_T]
struct SmCol {
  double *p;
  __device__ __forceinline__ double &operator[](int a) const { return p[a * UPD_T]; }
};

// MODE 0: 2nd-order centred advection + diffusion; 1: FCT advection (ratios of k_fct_rfac) + diffusion;
// 2: diffusive tendency only; 3: FCT advection only, subtracted from the tendency MODE 2 left in t(tau+1).
// 2 followed by 3 performs the operations of 1 in the same order with half the registers each.
template <int MODE>
#ifdef UPD_MINB   // experiment builds: resident CTAs per SM the register allocation must allow
#define UPD_BOUNDS __launch_bounds__(UPD_T, UPD_MINB)
#else
#define UPD_BOUNDS __launch_bounds__(UPD_T)
#endif
__global__ void UPD_BOUNDS k_update(const DevView v, int nbase, int ng, int tch, int jfirst, int nrow, int tiled) {
  // The 36 tracer-independent Redi coefficients of a cell's six faces are parked in shared
  // memory (thread-private slots, conflict free) instead of registers: 36 KB per 128-thread CTA
  // buys ~70 registers per thread, i.e. three CTAs per SM instead of one 256-thread CTA.
  __shared__ double sco[36][UPD_T];
  Cell q;
  if (tiled) {
    if (!decode_cell_tiled(v, jfirst, nrow, tiled, q)) return;
  } else if (!decode_cell(v, (long long)blockIdx.x * blockDim.x + threadIdx.x, jfirst, nrow, q)) {
    return;
  }
  const int tid = threadIdx.x;
  const int km = v.km, k = q.k, c = q.c, i = q.i, j = q.j, sk = q.sk, sj = q.sj;
  const long long n3 = v.n3;
  const bool iso = v.isopycmix != 0;
  // cyclic neighbours in i for quantities stored on i = 2..imt-1 only (ratios, Redi coefficients)
  const int cwr = (i == 2) ? c + (v.imt - 3) : c - 1;           // cell/face i-1 -> imt-1
  const int cer = (i == v.imt - 1) ? c - (v.imt - 3) : c + 1;   // cell i+1 -> 2
  // masks (09/mom/loadmw.F:60-77)
  const int kb = v.kmt[q.c2];
  if (MODE == 2 && __all_sync(__activemask(), kb < k)) {
    // A warp of land cells (continents, levels below the bottom): the implicit solve multiplies the tendency by
    // tmask = 0 (09/mom/tracer.F:1114-1127), so nothing here can reach t(tau+1).  Leave a clean zero and move on.
    const int g0z = blockIdx.y * tch, g1z = min(g0z + tch, ng);
    for (int g = g0z; g < g1z; g++) v.t_p1[(long long)(nbase + g) * n3 + c] = 0.0;
    return;
  }
  const double m = (kb >= k) ? 1.0 : 0.0, mu = (kb >= k - 1) ? 1.0 : 0.0;
  const double mw = (v.kmt[q.c2 - 1] >= k) ? 1.0 : 0.0, me = (v.kmt[q.c2 + 1] >= k) ? 1.0 : 0.0;
  const double ms = (v.kmt[q.c2 - v.imt] >= k) ? 1.0 : 0.0, mn = (v.kmt[q.c2 + v.imt] >= k) ? 1.0 : 0.0;
  // face velocities
  const double ue_c = v.ue[c], ue_w = v.ue[c - 1], vn_c = v.vn[c], vn_s = v.vn[c - sj];
  const double wb_d = v.wb[q.cz], wb_u = v.wb[q.cz - sk];
  // metric factors (09/mom/tracer.F:234-249, source/mom/fdift.h)
  const double cstr = v.cstr[j - 1];
  const double cstdxtr = cstr * v.dxtr[i - 1];
  const double cstdxt2r = cstr * v.dxtr[i - 1] * 0.5;
  const int iw = (i == 2) ? v.imt - 1 : i - 1;
  const double cstdxur_e = cstr * v.dxur[i - 1], cstdxur_w = cstr * v.dxur[iw - 1];
  const double ah_e = v.diff_cet * cstr * v.dxur[i - 1], ah_w = v.diff_cet * cstr * v.dxur[iw - 1];
  const double cstdyt2r = v.cstdyt2r[j - 1], cstdytr = v.cstdytr[j - 1];
  const double csu_dyur_n = v.csu_dyur[j - 1], csu_dyur_s = v.csu_dyur[j - 2];
  const double dzt2r = v.dzt2r[k - 1], dztr = v.dztr[k - 1], dzt4r = 0.5 * v.dzt2r[k - 1];
  const double csu_dzt4r_n = v.csu[j - 1] * 0.5 * v.dzt2r[k - 1], csu_dzt4r_s = v.csu[j - 2] * 0.5 * v.dzt2r[k - 1];
  const double dxt4r = v.dxt4r[i - 1], dyt4r_cstr = v.dyt4r[j - 1] * cstr;
  const double one_m_aidif = 1.0 - v.aidif;
  // tracer-independent coefficients of the six faces
  const SmCol ce_e{&sco[0][tid]}, ce_w{&sco[4][tid]}, cn_n{&sco[8][tid]}, cn_s{&sco[12][tid]};
  const SmCol cbx_d{&sco[16][tid]}, cby_d{&sco[20][tid]}, cbx_u{&sco[24][tid]}, cby_u{&sco[28][tid]};
  double &K11_e = sco[32][tid], &K11_w = sco[33][tid], &K22_n = sco[34][tid], &K22_s = sco[35][tid];
  const bool has_d = (k <= km - 1), has_u = (k >= 2);   // diff_fbiso exists on faces 1..km-1
  if (MODE != 3 && iso) {
#pragma unroll
    for (int a = 0; a < 4; a++) {
      ce_e[a] = v.ce[c + a * n3];
      ce_w[a] = v.ce[cwr + a * n3];
      cn_n[a] = v.cn[c + a * n3];
      cn_s[a] = v.cn[c - sj + a * n3];
      cbx_d[a] = has_d ? v.cbx[c + a * n3] : 0.0;
      cby_d[a] = has_d ? v.cby[c + a * n3] : 0.0;
      cbx_u[a] = has_u ? v.cbx[c - sk + a * n3] : 0.0;
      cby_u[a] = has_u ? v.cby[c - sk + a * n3] : 0.0;
    }
    K11_e = v.K11[c];
    K11_w = v.K11[cwr];
    K22_n = v.K22[c];
    K22_s = v.K22[c - sj];
  }
  // diff_cbt*dzwr on faces k and k-1 (09/mom/tracer.F:1025-1032)
  const double dcb_d = has_d ? v.diff_cbt[c] : 0.0, dcb_u = has_u ? v.diff_cbt[c - sk] : 0.0;
  const double dzwr_d = v.dzwr[k], dzwr_u = v.dzwr[k - 1];
  const int cu = has_u ? c - sk : c, cd = has_d ? c + sk : c;   // clamped k-1, k+1 (km1kr / kpkr of isoflux)
  const bool row_s_has_R = (j - 1 >= 2), row_n_has_R = (j + 1 <= v.jmt - 1);

  const int g0 = blockIdx.y * tch, g1 = min(g0 + tch, ng);
  for (int g = g0; g < g1; g++) {
    const int n0 = nbase + g;
    const double *__restrict__ T = v.t_m1 + (long long)n0 * n3;
    const double *__restrict__ U = v.t_0 + (long long)n0 * n3;
    const double *__restrict__ R = v.Rfac + (long long)g * 6 * n3;
    if (g + 1 < g1) {
      // next tracer's neighbourhood: nine cache lines per warp, requested while this tracer is computed
      const double *__restrict__ Tn1 = T + n3;
      pf_l2(Tn1 + c); pf_l2(Tn1 + c + sj); pf_l2(Tn1 + c - sj); pf_l2(Tn1 + cu); pf_l2(Tn1 + cd);
      pf_l2(Tn1 + cu + sj); pf_l2(Tn1 + cu - sj); pf_l2(Tn1 + cd + sj); pf_l2(Tn1 + cd - sj);
    }
    // t(tau-1): the 15-point neighbourhood
    const double Tc = T[c], Te = T[c + 1], Tw = T[c - 1], Tn = T[c + sj], Ts = T[c - sj], Tu = T[cu], Td = T[cd];
    const double Teu = T[cu + 1], Twu = T[cu - 1], Tnu = T[cu + sj], Tsu = T[cu - sj];
    const double Ted = T[cd + 1], Twd = T[cd - 1], Tnd = T[cd + sj], Tsd = T[cd - sj];
    // t(tau)
    const double Uc = U[c], Ue = U[c + 1], Uw = U[c - 1], Un = U[c + sj], Us = U[c - sj], Uu = U[cu], Ud = U[cd];

    // ---------------- advective fluxes ----------------
    double adv_tx = 0.0, adv_ty = 0.0, adv_tz = 0.0, adv_tz_iso[3] = {0.0, 0.0, 0.0};
    bool adv_iso = false;
    if constexpr (MODE == 1 || MODE == 3) {
      // east / west faces: Cpos(f) = min(Rpl(f+1),Rmn(f)), Cneg(f) = min(Rpl(f),Rmn(f+1)) (:698-701); no mask (:987)
      const double rplx_c = R[c], rmnx_c = R[c + n3];
      double adv_fe_e, adv_fe_w;
      {
        double lo = upw(ue_c, Tc, Te);
        double a = ue_c * (Uc + Ue) - lo;
        adv_fe_e = delimit(dmin(R[cer], rmnx_c), dmin(rplx_c, R[cer + n3]), a) + lo;
        lo = upw(ue_w, Tw, Tc);
        a = ue_w * (Uw + Uc) - lo;
        adv_fe_w = delimit(dmin(rplx_c, R[cwr + n3]), dmin(R[cwr], rmnx_c), a) + lo;
      }
      // north / south faces: Cpos(g) = min(R_plusY(g+1),R_minusY(g)), Cneg(g) = min(R_plusY(g),R_minusY(g+1)) (:772-775)
      const double rply_c = R[c + 2 * n3], rmny_c = R[c + 3 * n3];
      double adv_fn_n, adv_fn_s;
      {
        double rpl_n = 0.0, rmn_n = 0.0, rpl_s = 0.0, rmn_s = 0.0;   // rows 1 and jmt carry zero ratios
        if (row_n_has_R) { rpl_n = R[c + sj + 2 * n3]; rmn_n = R[c + sj + 3 * n3]; }
        if (row_s_has_R) { rpl_s = R[c - sj + 2 * n3]; rmn_s = R[c - sj + 3 * n3]; }
        double lo = upw(vn_c, Tc, Tn);
        double a = vn_c * (Uc + Un) - lo;
        adv_fn_n = (delimit(dmin(rpl_n, rmny_c), dmin(rply_c, rmn_n), a) + lo) * m;
        lo = upw(vn_s, Ts, Tc);
        a = row_s_has_R ? vn_s * (Us + Uc) - lo : 0.0;                // anti_fn(i,k,1,n) = c0 (:475)
        adv_fn_s = (delimit(dmin(rply_c, rmn_s), dmin(rpl_s, rmny_c), a) + lo) * ms;
      }
      // bottom / top faces: Cneg(h) = min(Rpl(h+1),Rmn(h)), Cpos(h) = min(Rpl(h),Rmn(h+1)) (:966-969);
      // adv_fb(0), adv_fb(km) are overwritten in tracer (09/mom/tracer.F:1063-1065)
      const double rplz_c = R[c + 4 * n3], rmnz_c = R[c + 5 * n3];
      double adv_fb_d, adv_fb_u;
      if (k == km) {
        adv_fb_d = wb_d * Uc;
      } else {
        double lo = upw(wb_d, Td, Tc);
        double a = wb_d * (Uc + Ud) - lo * m;
        adv_fb_d = (delimit(dmin(rplz_c, R[c + sk + 5 * n3]), dmin(R[c + sk + 4 * n3], rmnz_c), a) + lo) * m;
      }
      if (k == 1) {
        adv_fb_u = wb_u * (Uc + Uc);
      } else {
        double lo = upw(wb_u, Tc, Tu);
        double a = wb_u * (Uu + Uc) - lo * mu;
        adv_fb_u = (delimit(dmin(R[c - sk + 4 * n3], rmnz_c), dmin(rplz_c, R[c - sk + 5 * n3]), a) + lo) * mu;
      }
      adv_tx = (adv_fe_e - adv_fe_w) * cstdxt2r;
      adv_ty = (adv_fn_n - adv_fn_s) * cstdyt2r;
      adv_tz = (adv_fb_u - adv_fb_d) * dzt2r;
    } else if constexpr (MODE == 0) {
      // 2nd-order centred (09/mom/tracer_adv_flx.F:1030-1082, source/mom/fdift.h:25-39) plus the
      // Gent-McWilliams advective terms ADV_Txiso/Tyiso/Tziso (fdift.h:44-52, isoflux :1110-1134)
      const double vet_c = v.adv_vet[c], vet_w = v.adv_vet[c - 1], vnt_c = v.adv_vnt[c], vnt_s = v.adv_vnt[c - sj];
      const double vbt_d = v.adv_vbt[q.cz], vbt_u = v.adv_vbt[q.cz - sk];
      adv_tx = (vet_c * (Uc + Ue) - vet_w * (Uw + Uc)) * cstdxt2r;
      adv_ty = (vnt_c * (Uc + Un) - vnt_s * (Us + Uc)) * cstdyt2r;
      const double fb_dn = (k == km) ? vbt_d * Uc : vbt_d * (Uc + Ud);
      const double fb_up = (k == 1) ? vbt_u * (Uc + Uc) : vbt_u * (Uu + Uc);
      adv_tz = (fb_up - fb_dn) * dzt2r;
      if (iso) {
        const double ve_c = v.adv_vetiso[c], ve_w = v.adv_vetiso[c - 1], vn_c2 = v.adv_vntiso[c], vn_s2 = v.adv_vntiso[c - sj];
        const double txi = cstdxt2r * (ve_c * (Te + Tc) - ve_w * (Tc + Tw));
        const double tyi = cstdyt2r * (vn_c2 * (Tn + Tc) - vn_s2 * (Tc + Ts));
        const double fbi_d = (k == km) ? 0.0 : v.adv_vbtiso[q.cz] * (Tc + Td);
        const double fbi_u = (k == 1) ? 0.0 : v.adv_vbtiso[q.cz - sk] * (Tu + Tc);
        const double tzi = dzt2r * (fbi_u - fbi_d);
        adv_iso = true;
        // subtracted in the reference's order: ... - ADV_Tz - ADV_Txiso - ADV_Tyiso - ADV_Tziso
        adv_tz_iso[0] = txi; adv_tz_iso[1] = tyi; adv_tz_iso[2] = tzi;
      }
    }

    // ---------------- diffusive fluxes ----------------
    // east / west (09/mom/tracer.F:930-940 + isoflux 09/mom/isopyc.F:950-1002)
    double diff_fe_e, diff_fe_w;
    {
      double d = Te - Tc;
      diff_fe_e = ah_e * d;
      if (iso) {
        double sumz = 0.0;
        sumz = sumz - ce_e[0] * (Tu - Tc);
        sumz = sumz - ce_e[1] * (Teu - Te);
        sumz = sumz - ce_e[2] * (Tc - Td);
        sumz = sumz - ce_e[3] * (Te - Ted);
        diff_fe_e = diff_fe_e + K11_e * cstdxur_e * d + dzt4r * sumz;
      }
      d = Tc - Tw;
      diff_fe_w = ah_w * d;
      if (iso) {
        double sumz = 0.0;
        sumz = sumz - ce_w[0] * (Twu - Tw);
        sumz = sumz - ce_w[1] * (Tu - Tc);
        sumz = sumz - ce_w[2] * (Tw - Twd);
        sumz = sumz - ce_w[3] * (Tc - Td);
        diff_fe_w = diff_fe_w + K11_w * cstdxur_w * d + dzt4r * sumz;
      }
    }
    // north / south (09/mom/tracer.F:945-961 + isoflux :1007-1053)
    double diff_fn_n, diff_fn_s;
    {
      double d = Tn - Tc;
      diff_fn_n = v.diff_cnt * csu_dyur_n * d;
      if (iso) {
        double sumz = 0.0;
        sumz = sumz - cn_n[0] * (Tu - Tc);
        sumz = sumz - cn_n[1] * (Tnu - Tn);
        sumz = sumz - cn_n[2] * (Tc - Td);
        sumz = sumz - cn_n[3] * (Tn - Tnd);
        diff_fn_n = diff_fn_n + K22_n * csu_dyur_n * d + csu_dzt4r_n * sumz;
      }
      d = Tc - Ts;
      diff_fn_s = v.diff_cnt * csu_dyur_s * d;
      if (iso) {
        double sumz = 0.0;
        sumz = sumz - cn_s[0] * (Tsu - Ts);
        sumz = sumz - cn_s[1] * (Tu - Tc);
        sumz = sumz - cn_s[2] * (Ts - Tsd);
        sumz = sumz - cn_s[3] * (Tc - Td);
        diff_fn_s = diff_fn_s + K22_s * csu_dyur_s * d + csu_dzt4r_s * sumz;
      }
    }
    // vertical with the b.c. of tracer.F:1053-1062: diff_fb(0)=stf, then diff_fb(kmt)=btf
    double fb_d, fb_u;
    {
      const double stf = v.stf[q.c2 + (long long)n0 * v.n2], btf = v.btf[q.c2 + (long long)n0 * v.n2];
      fb_d = (k == kb) ? btf : (has_d ? dcb_d * dzwr_d * (Tc - Td) : 0.0);
      fb_u = (k - 1 == kb) ? btf : ((k == 1) ? stf : dcb_u * dzwr_u * (Tu - Tc));
    }
    double diff_tz;
    if (iso) {
      // K31, K32 part solved explicitly (09/mom/isopyc.F:1062-1108)
      double fbiso_d = 0.0, fbiso_u = 0.0;
      if (has_d) {
        double sumx = 0.0, sumy = 0.0;
        sumx = sumx - cbx_d[0] * (Tc - Tw);
        sumx = sumx - cbx_d[2] * (Td - Twd);
        sumx = sumx - cbx_d[1] * (Te - Tc);
        sumx = sumx - cbx_d[3] * (Ted - Td);
        sumy = sumy - cby_d[0] * (Tc - Ts);
        sumy = sumy - cby_d[2] * (Td - Tsd);
        sumy = sumy - cby_d[1] * (Tn - Tc);
        sumy = sumy - cby_d[3] * (Tnd - Td);
        fbiso_d = dxt4r * sumx + dyt4r_cstr * sumy;
      }
      if (has_u) {
        double sumx = 0.0, sumy = 0.0;
        sumx = sumx - cbx_u[0] * (Tu - Twu);
        sumx = sumx - cbx_u[2] * (Tc - Tw);
        sumx = sumx - cbx_u[1] * (Teu - Tu);
        sumx = sumx - cbx_u[3] * (Te - Tc);
        sumy = sumy - cby_u[0] * (Tu - Tsu);
        sumy = sumy - cby_u[2] * (Tc - Ts);
        sumy = sumy - cby_u[1] * (Tnu - Tu);
        sumy = sumy - cby_u[3] * (Tn - Tc);
        fbiso_u = dxt4r * sumx + dyt4r_cstr * sumy;
      }
      diff_tz = (fb_u - fb_d) * dztr * one_m_aidif + (fbiso_u - fbiso_d) * dztr;
    } else {
      diff_tz = (fb_u - fb_d) * dztr;
    }
    const double diff_tx = (diff_fe_e * me - diff_fe_w * mw) * cstdxtr;
    const double diff_ty = (diff_fn_n * mn - diff_fn_s * ms) * cstdytr;

    // The source term is added in k_invtri (same operation order as 09/mom/tracer.F:1114-1127:
    // t(tau-1) + twodt*(DIFF - ADV + source)*tmask), so the MOBI kernels, which run on a side
    // stream, only have to finish before the implicit solve.
    double P;
    if constexpr (MODE == 3) P = v.t_p1[(long long)n0 * n3 + c];   // the diffusive tendency left by k_update<2>
    else P = diff_tx + diff_ty + diff_tz;
    if constexpr (MODE != 2) P = P - adv_tx - adv_ty - adv_tz;
    if (MODE == 0 && adv_iso) P = P - adv_tz_iso[0] - adv_tz_iso[1] - adv_tz_iso[2];
    v.t_p1[(long long)n0 * n3 + c] = P;
  }
}

// Implicit vertical diffusion.  One thread per (column, tracer): source/mom/invtri.F:75-110 with the
// precomputed factors a, e, bet of k_vmix_factor.  A CTA is 32 consecutive columns x 4 tracers (warp =
// tracer) and the grid runs the tracer quads of one column block back to back, so the tracer-independent
// a / bet / e lines are fetched from HBM once and hit L1 / L2 for the other tracers.  The forward sweep
// parks z(k) in a thread-private shared-memory column instead of writing it to t(tau+1) and reading it
// back for the substitution: per cell and tracer the kernel moves the tendency, t(tau-1), the source
// (reads) and t(tau+1) (one write).
#define INV_T 128
// INV_KC: levels whose loads are in flight together; 8 for deep grids, 4 for km <= 24 (19 levels are 5 chunks of 4 or 3
// of 8 with 5 idle slots; measured 47 against 54 us on the 100x100x19 grid, 1.87 against 2.33 ms on 0.5 degree x 40 levels)
// resident CTAs per SM: left to the compiler (it settles on 128 registers = 4 CTAs for the deep-grid variant); forcing 5 / 6
// CTAs (scripts/build_variants.py inv5=-DUVIC_INV_MINBLOCKS=5) measured 2.95 / 3.49 ms against 1.85 on 0.5 degree x 40 levels
#ifdef UVIC_INV_MINBLOCKS
#define INV_BOUNDS __launch_bounds__(INV_T, UVIC_INV_MINBLOCKS)
#else
#define INV_BOUNDS __launch_bounds__(INV_T)
#endif
template <int INV_KC>
__global__ void INV_BOUNDS k_invtri(const DevView v, int nbase, int ng, int ntq) {
  extern __shared__ double zsm[];   // [km][INV_T]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int tq = blockIdx.x % ntq, cb = blockIdx.x / ntq;
  const int g = tq * 4 + w;
  if (g >= ng) return;
  const int ni = v.imt - 2;
  const int nrow = v.jhi - v.jlo + 1;
  const long long idx = (long long)cb * 32 + lane;
  if (idx >= (long long)ni * nrow) return;
  const int i = (int)(idx % ni) + 2;
  const int j = (int)(idx / ni) + v.jlo;
  const int n0 = nbase + g;
  double *__restrict__ z = v.t_p1 + (long long)n0 * v.n3;
  const double *__restrict__ tm1 = v.t_m1 + (long long)n0 * v.n3;
  const int isrc = v.itrc[n0];
  const double *__restrict__ srcp = (isrc != 0) ? v.src + (long long)(isrc - 1) * v.n3 : nullptr;
  const double *__restrict__ tri_a = v.tri_a, *__restrict__ tri_bet = v.tri_bet, *__restrict__ tri_e = v.tri_e;
  double *zc = zsm + threadIdx.x;
  const int km = v.km;
  const int kb = v.kmt[X2(i, j)];
  const int kbot = max(2, kb);
  const double topbc = v.stf[X2(i, j) + (long long)n0 * v.n2];
  const double botbc = v.btf[X2(i, j) + (long long)n0 * v.n2];
  const double aidif = v.aidif, c2dtts = v.c2dtts;
  double zprev = 0.0;
  const int c1 = (int)X3(i, 1, j), sk = v.imt;
  // levels are processed in chunks of INV_KC: all loads of a chunk are issued before its (serial) recurrence, so a
  // thread keeps 5*INV_KC independent 8-byte loads in flight instead of one level's worth
  const double *__restrict__ dtx = v.dtxcel;
  const double dztr_top = v.dztr[0], dztr_bot = v.dztr[kbot - 1];
  for (int k0 = 1; k0 <= km; k0 += INV_KC) {
    double pp[INV_KC], tt[INV_KC], ss[INV_KC], aa[INV_KC], bb[INV_KC], td[INV_KC];
#pragma unroll
    for (int q = 0; q < INV_KC; q++) {
      const int k = min(k0 + q, km);
      const int c = c1 + (k - 1) * sk;
      // below the bottom (and on land) bet = tmask/(...) = 0 and f = (...)*tmask = 0: z(k) = 0 whatever the inputs are,
      // so their loads are not issued
      const bool wet = k <= kb;
      pp[q] = wet ? z[c] : 0.0;
      tt[q] = wet ? tm1[c] : 0.0;
      ss[q] = (wet && srcp) ? srcp[c] : 0.0;
      aa[q] = wet ? tri_a[c] : 0.0;
      bb[q] = wet ? tri_bet[c] : 0.0;
      td[q] = dtx[k - 1];
    }
    asm volatile("" ::: "memory");   // keep the chunk's loads together, ahead of the recurrence
#pragma unroll
    for (int q = 0; q < INV_KC; q++) {
      const int k = k0 + q;
      if (k <= km) {
        const double mk = (kb >= k) ? 1.0 : 0.0;
        const double tdt = c2dtts * td[q];
        // explicit update (09/mom/tracer.F:1114-1127) from the partial tendency left by the flux kernels
        const double zc0 = tt[q] + tdt * (pp[q] + ss[q]) * mk;
        double f = zc0 * mk;
        if (k == 1) f = zc0 + topbc * tdt * dztr_top * aidif * mk;
        if (k == kbot) f = zc0 - botbc * tdt * dztr_bot * aidif * mk;
        double zk;
        if (k == 1)
          zk = f * bb[q];
        else
          zk = (f - aa[q] * zprev) * bb[q];
        zc[(k - 1) * INV_T] = zk;
        zprev = zk;
      }
    }
  }
  // back substitution + cyclic boundary (setbcx, 09/mom/tracer.F:1153-1155)
  const int wrap = (i == 2) ? (v.imt - 2) : ((i == v.imt - 1) ? -(v.imt - 2) : 0);
  double znext = zprev;
  {
    const int c = c1 + (km - 1) * sk;
    z[c] = znext;
    if (wrap) z[c + wrap] = znext;
  }
  for (int k0 = km - 1; k0 >= 1; k0 -= INV_KC) {
    double ee[INV_KC];
#pragma unroll
    for (int q = 0; q < INV_KC; q++) {
      const int k = max(k0 - q, 1);
      ee[q] = (k < kb) ? tri_e[c1 + k * sk] : 0.0;   // e(k+1) = c(k)*bet(k) = 0 from the bottom level down
    }
    asm volatile("" ::: "memory");
#pragma unroll
    for (int q = 0; q < INV_KC; q++) {
      const int k = k0 - q;
      if (k >= 1) {
        const int c = c1 + (k - 1) * sk;
        const double zk = zc[(k - 1) * INV_T] - ee[q] * znext;
        z[c] = zk;
        znext = zk;
        if (wrap) z[c + wrap] = zk;
      }
    }
  }
}

// source/mom/dens.h:18-19
#define ECC(k, m) v.eosc[((k)-1) + v.km * ((m)-1)]
__device__ __forceinline__ double dens_f(const DevView &v, double tq, double sq, int k) {
  return (ECC(k, 1) + (ECC(k, 4) + ECC(k, 7) * sq) * sq + (ECC(k, 3) + ECC(k, 8) * sq + ECC(k, 6) * tq) * tq) * tq +
         (ECC(k, 2) + (ECC(k, 5) + ECC(k, 9) * sq) * sq) * sq;
}

// convct2 phase 1 (source/mom/convect.F:193-268): one thread per column finds the unstable
// regions from T and S, mixes T and S, and records (kt, kb, zsm) of every region so that the
// other tracers can be mixed in parallel.  zsm is recorded, not recomputed, because it is
// accumulated in the order the levels were discovered.
__global__ void __launch_bounds__(128) k_convect_ts(const DevView v) {
  // The T and S columns are fetched once, all levels in flight together, into thread-private shared-memory columns; the
  // data-dependent search below then runs out of shared memory instead of paying a memory round trip per level it
  // touches (one column per thread: the kernel is latency bound).  Columns that convected are written back.
  extern __shared__ double cts[];   // [2][km][128]
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int ni = v.imt - 2;
  int nrow = v.jhi - v.jlo + 1;
  if (idx >= (long long)ni * nrow) return;
  int i = (int)(idx % ni) + 2;
  int j = (int)(idx / ni) + v.jlo;
  const int kbo = v.kmt[X2(i, j)];
  double *gT = v.t_p1, *gS = v.t_p1 + v.n3;
  double *T = cts + threadIdx.x, *S = cts + (size_t)v.km * 128 + threadIdx.x;
  const double *dz = v.dztxcl;
  const int c1 = (int)X3(i, 1, j), sk = v.imt;
  const long long col = X2(i, j);
  const int maxreg = v.km / 2 + 1;
  for (int k0 = 1; k0 <= kbo; k0 += 8) {
    double a_[8], b_[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int k = min(k0 + q, kbo);
      a_[q] = gT[c1 + (k - 1) * sk];
      b_[q] = gS[c1 + (k - 1) * sk];
    }
#pragma unroll
    for (int q = 0; q < 8; q++)
      if (k0 + q <= kbo) {
        T[(k0 + q - 1) * 128] = a_[q];
        S[(k0 + q - 1) * 128] = b_[q];
      }
  }
#define TSV(a, k) a[((k)-1) * 128]
  int nreg = 0;
  int kt = 1, kb = 2;
  while (kt < kbo) {
    double ru = dens_f(v, TSV(T, kt) - v.to[kb - 1], TSV(S, kt) - v.so[kb - 1], kb);
    double rl = dens_f(v, TSV(T, kb) - v.to[kb - 1], TSV(S, kb) - v.so[kb - 1], kb);
    if (ru > rl) {
      bool chk_la = true, chk_lb = true;
      double zsm = dz[kt - 1] + dz[kb - 1];
      double tsm1 = TSV(T, kt) * dz[kt - 1] + TSV(T, kb) * dz[kb - 1];
      double tmx1 = qdiv(tsm1, zsm);
      double tsm2 = TSV(S, kt) * dz[kt - 1] + TSV(S, kb) * dz[kb - 1];
      double tmx2 = qdiv(tsm2, zsm);
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
            tmx1 = qdiv(tsm1, zsm);
            tsm2 = tsm2 + TSV(S, kb) * dz[kb - 1];
            tmx2 = qdiv(tsm2, zsm);
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
            tmx1 = qdiv(tsm1, zsm);
            tsm2 = tsm2 + TSV(S, kt) * dz[kt - 1];
            tmx2 = qdiv(tsm2, zsm);
            chk_lb = true;
          }
        }
      }
      for (int k = kt; k <= kb; k++) {
        TSV(T, k) = tmx1;
        TSV(S, k) = tmx2;
      }
      if (nreg < maxreg) {
        v.conv_kt[col + (long long)nreg * v.n2] = kt | (kb << 16);
        v.conv_zsm[col + (long long)nreg * v.n2] = zsm;
      }
      nreg++;
      kt = kb + 1;
    } else {
      kt = kb;
    }
    kb = kt + 1;
  }
  v.conv_n[col] = nreg;
  if (nreg > 0) {
    // write the adjusted column back, with its cyclic copy (09/mom/tracer.F:1199-1203; untouched columns keep the
    // copy k_invtri made)
    const int off = (i == 2) ? (v.imt - 2) : ((i == v.imt - 1) ? -(v.imt - 2) : 0);
    for (int k = 1; k <= kbo; k++) {
      const double tq = TSV(T, k), sq = TSV(S, k);
      gT[c1 + (k - 1) * sk] = tq;
      gS[c1 + (k - 1) * sk] = sq;
      if (off) {
        gT[c1 + (k - 1) * sk + off] = tq;
        gS[c1 + (k - 1) * sk + off] = sq;
      }
    }
  }
#undef TSV
}

// convct2 phase 2 (source/mom/convect.F:269-277): one thread per (column, tracer n >= 3)
__global__ void __launch_bounds__(128) k_convect_tr(const DevView v, int nfirst) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int ni = v.imt - 2;
  int nrow = v.jhi - v.jlo + 1;
  if (idx >= (long long)ni * nrow) return;
  int i = (int)(idx % ni) + 2;
  int j = (int)(idx / ni) + v.jlo;
  const int n0 = nfirst + blockIdx.y;
  const long long col = X2(i, j);
  const int nreg = v.conv_n[col];
  const bool edge = (i == 2 || i == v.imt - 1);
  if (nreg == 0) return;   // untouched column: k_invtri already set its cyclic copy
  double *X = v.t_p1 + (long long)n0 * v.n3;
  const double *dz = v.dztxcl;
  const int c1 = (int)X3(i, 1, j), sk = v.imt;
  for (int r = 0; r < nreg; r++) {
    int pk = v.conv_kt[col + (long long)r * v.n2];
    int kt = pk & 0xffff, kb = pk >> 16;
    double zsm = v.conv_zsm[col + (long long)r * v.n2];
    double tsm3 = 0.0;
    for (int k = kt; k <= kb; k++) tsm3 = tsm3 + X[c1 + (k - 1) * sk] * dz[k - 1];
    double tmx3 = qdiv(tsm3, zsm);
    for (int k = kt; k <= kb; k++) X[c1 + (k - 1) * sk] = tmx3;
  }
  if (edge) {
    int off = (i == 2) ? (v.imt - 2) : -(v.imt - 2);
    for (int k = 1; k <= v.km; k++) X[c1 + (k - 1) * sk + off] = X[c1 + (k - 1) * sk];
  }
}

// FCT variants: the default is the marching kernel of k_fct.cu behind the diffusion pass; UVIC_B200_FCT=split keeps
// the two-pass version (ratios through HBM: k_fct_rfac, then k_update<3>), =merged runs diffusion and the FCT fluxes
// of the two-pass version in one kernel (k_update<1>).  All three agree bit for bit (tests/test_gpu_parity.py).
int fct_variant() {
  const char *e = getenv("UVIC_B200_FCT");
  if (e && !strcmp(e, "merged")) return 2;
  if (e && !strcmp(e, "split")) return 1;
  return 0;
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
  // Tracer batches.  Device resident: as many tracers per batch as the FCT scratch holds.  With a host destination
  // (uvic_b200_tracer_step) T and S go first -- they need no MOBI source and clinic wants them -- and the rest is cut into
  // a few batches, so that every finished batch streams to the host on the copy stream while the next one computes.
  std::vector<std::pair<int, int>> batches;
  if (c->d2h_dst && v.nt > 2) {
    batches.push_back({0, 2});
    if (c->d2h_ntr <= 2) {
      // the coupled entry point only returns T and S: the other tracers go in as few batches as possible
      for (int q = 2; q < v.nt; q += v.ngroup) batches.push_back({q, std::min(v.ngroup, v.nt - q)});
    } else {
      // Every finished batch streams to the host while the next one computes, and the link (not the kernels) is what
      // the call waits for.
      // A batch costs a fixed ~0.09 ms of short kernels plus ~13 us per tracer on the 100x100 grid, its copy ~29 us per
      // tracer: batches of six are the smallest that keep the copy engine fed, and the last one leaves a short tail.
      int q = 2;
      while (q < v.nt) {
        const int n = std::min(std::min(6, v.ngroup), v.nt - q);   // 5 and 8 measured slower
        batches.push_back({q, n});
        q += n;
      }
    }
  } else {
    for (int nbase = 0; nbase < v.nt; nbase += v.ngroup) batches.push_back({nbase, std::min(v.ngroup, v.nt - nbase)});
  }
  bool mobi_waited = false;
  size_t nb_done = 0;
  for (auto &bt : batches) {
    const int nbase = bt.first, ng = bt.second;
    // tracers per thread: keep at least ~4 CTAs-worth of threads per SM in flight, then amortise the
    // tracer-independent loads over as many tracers as possible
    int nchunk = (int)min((long long)ng, max(1LL, (148LL * 2048 * 2 + ncell_c - 1) / ncell_c));
    int tch = (ng + nchunk - 1) / nchunk;
    nchunk = (ng + tch - 1) / tch;
    // CTA = 32 i x 4 levels where that tiling wastes few threads (imt - 2 close to a multiple of 32, km to one of 4): the
    // kernel is bound by L2 -> L1 traffic and the tile halves it.  UVIC_B200_UPD_TILE=0/1 overrides (experiments).
    const int tile_env = getenv("UVIC_B200_UPD_TILE") ? atoi(getenv("UVIC_B200_UPD_TILE")) : -1;   // read per call: the tests toggle it
    const int tiled = (tile_env >= 0) ? tile_env : ((double)upd_tiles(v, UPD_T, 1) * UPD_T <= 1.08 * (double)(v.imt - 2) * v.km ? 1 : 0);
    const long long ntile = upd_tiles(v, UPD_T, tiled);
    dim3 gc(tiled ? (unsigned)(ntile * (tiled == 2 ? (nrow_c + 1) / 2 : nrow_c)) : cdiv(ncell_c, UPD_T), nchunk);
    // the total velocities come from the GM chain on its side stream: the advection kernels below are its first readers
    auto velocities = [&]() {
      gm_join(c);
      if (c->halo_event) {   // uvic_b200_wait_before_advection: the halo rows of t(tau) are about to be read
        cudaStreamWaitEvent(c->stream, c->halo_event, 0);
        c->halo_event = nullptr;
      }
    };
    if (v.fct) {
      const int variant = fct_variant();
      if (variant == 0) {
        KLAUNCH("k_diffuse", k_update<2>, gc, UPD_T, v, nbase, ng, tch, v.jlo, nrow_c, tiled);
        velocities();
        launch_fct_march(c, nbase, ng);
      } else {
        velocities();
        dim3 gr(cdiv(ncell_r, 256), nchunk);
        KLAUNCH("k_fct_rfac", k_fct_rfac, gr, 256, v, nbase, ng, tch, jf_r, nrow_r);
        if (variant == 1) {
          KLAUNCH("k_diffuse", k_update<2>, gc, UPD_T, v, nbase, ng, tch, v.jlo, nrow_c, tiled);
          KLAUNCH("k_fct_apply", k_update<3>, gc, UPD_T, v, nbase, ng, tch, v.jlo, nrow_c, tiled);
        } else {
          KLAUNCH("k_update", k_update<1>, gc, UPD_T, v, nbase, ng, tch, v.jlo, nrow_c, tiled);
        }
      }
    } else {
      velocities();
      KLAUNCH("k_update", k_update<0>, gc, UPD_T, v, nbase, ng, tch, v.jlo, nrow_c, tiled);
    }
    // the source term enters in k_invtri: the first batch with a sourced tracer waits for MOBI
    bool sourced = false;
    for (int n = nbase; n < nbase + ng; n++) sourced = sourced || c->itrc_h[n] != 0;
    if (sourced && !mobi_waited && c->mobi_event && c->mobi_inflight) {
      cudaStreamWaitEvent(c->stream, c->src_ready[c->src_cur], 0);   // this step's sources, not a look-ahead queued behind them
      mobi_waited = true;
    }
    {
      const int ntq = (ng + 3) / 4;
      const size_t shm = (size_t)v.km * INV_T * sizeof(double);
      ensure_dyn_smem(c, (const void *)k_invtri<8>, shm);
      ensure_dyn_smem(c, (const void *)k_invtri<4>, shm);
      ProfScope ps_(c, "k_invtri");
      if (v.km <= 24)
        k_invtri<4><<<cdiv(ncol, 32) * ntq, INV_T, shm, c->stream>>>(v, nbase, ng, ntq);
      else
        k_invtri<8><<<cdiv(ncol, 32) * ntq, INV_T, shm, c->stream>>>(v, nbase, ng, ntq);
    }
    if (c->par.fullconvect) {
      if (nbase == 0) {
        const size_t shm = (size_t)2 * v.km * 128 * sizeof(double);
        ensure_dyn_smem(c, (const void *)k_convect_ts, shm);
        ProfScope ps_(c, "k_convect_ts");
        k_convect_ts<<<cdiv(ncol, 128), 128, shm, c->stream>>>(v);
      }
      const int nfirst = max(2, nbase), ntr = nbase + ng - nfirst;
      if (ntr > 0) {
        dim3 gt(cdiv(ncol, 128), ntr);
        KLAUNCH("k_convect_tr", k_convect_tr, gt, 128, v, nfirst);
      }
    }
    // Fourier filter of the polar rows + cyclic boundary (09/mom/tracer.F:1245-1262)
    launch_filter(c, nbase, ng);
    if (c->d2h_dst && nbase < c->d2h_ntr) {
      // this batch of t(tau+1) is final: hand it to the copy stream
      if (nb_done >= c->ev_batch.size()) {
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        c->ev_batch.push_back(e);
      }
      cudaEventRecord(c->ev_batch[nb_done], c->stream);
      cudaStreamWaitEvent(c->copy_out, c->ev_batch[nb_done], 0);
      cudaMemcpyAsync(c->d2h_dst + (size_t)nbase * v.n3, v.t_p1 + (size_t)nbase * v.n3, (size_t)ng * v.n3 * sizeof(double),
                      cudaMemcpyDeviceToHost, c->copy_out);
      nb_done++;
    }
  }
}
