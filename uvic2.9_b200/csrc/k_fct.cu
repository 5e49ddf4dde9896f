// k_fct.cu -- flux-corrected transport (adv_flux with O_fct, 09/mom/tracer_adv_flx.F:381-1029)
// as ONE kernel that marches through the latitude rows with the 3-D stencil staged in shared
// memory.
//
//   k_fct_march   low-order (upstream) fluxes and t_lo (:496-580), raw antidiffusive fluxes and
//                 the one-dimensional Zalesak ratios R+-x, R+-y, R+-z (:582-712, 714-770,
//                 786-958), delimited + low-order fluxes (:696-712, 772-784, 960-1002) and their
//                 divergence ADV_Tx + ADV_Ty + ADV_Tz (source/mom/fdift.h:25-39), subtracted
//                 from the diffusive tendency k_diffuse left in t(tau+1).
//
// Work decomposition.  A CTA owns FM_TI = 30 consecutive i (lane = i, lanes 0 and 31 are the
// west / east halo cells), a tile of TK <= 20 levels (warp = k, one halo warp above and below
// when the tile does not touch the surface / the bottom), one tracer, and a chunk of latitude
// rows that it walks south to north.  Per row r it
//   1. issues the cp.async copies that stage row r+3 of t(tau-1), t(tau) and the three face
//      velocities (34 x (TK+4) doubles each, cyclic wrap in i applied per element) into a
//      four-slot shared-memory ring; rows r and r+1 landed before the previous barrier;
//   2. forms t_lo and the six ratios of its cell (i,k,r): the i and k neighbours come from shared
//      memory, the j neighbours are the thread's own values of the rows before and after;
//   3. publishes R+-x / R+-z of row r in shared memory (double buffered) -- one barrier per row;
//   4. delimits the x and z face fluxes of row r-1 with the neighbours' ratios published one
//      barrier earlier, and its north face, whose limiter needed R+-y of row r; writes row r-1.
// Nothing but t(tau-1), t(tau), the velocities, kmt and the tendency crosses HBM: the six ratio
// fields the two-pass version (k_fct_rfac / k_update<3>, kept as the UVIC_B200_FCT=split
// reference path) wrote and re-read are gone, and the low-order fluxes are formed once per cell.  Every expression is evaluated with the operands and the order of the two-pass
// kernels, so both paths agree bit for bit.
//
// The reference's j loop looks sequential (iteration j limits anti_fn(j) with R+-Y(j) from the
// previous iteration); what is kept are its boundary rules: R+-Y(row 1) = anti_fn(row 1) = 0
// (:467-482), the clamps jp1/jp2 (:554-556, which only reach row jmt where tmask = 0, so
// R+-Y(row jmt) = 0) and the cyclic wrap of R+-x (:693-694).
#include "ctx.h"
#include "fct_common.h"
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <array>
#include <cuda.h>   // CUtensorMap (the encode function itself is fetched through cudaGetDriverEntryPoint: no -lcuda)

#define FM_TI 30      // interior cells in i per CTA
#define FM_W 36       // staged row width: element E <-> i = i0 - 3 + E; E = 1..34 (i0-2 .. i0+31) are used, E = 0 and 35 only make
                      // the row start 16-byte aligned in global memory (i0 - 4 is even, 0-based) and its size a multiple of 16: TMA
#define FM_E 2        // element of lane 0 is FM_E - 1: a lane's own element is lane + FM_E
#define FM_MAXW 21    // most warps per CTA (shared-memory layout of the ratio planes)
#define FM_NSLOT 4    // rows r-1, r, r+1 in use, r+2 in flight

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMA staging (interior i tiles of grids with an even imt): one elected thread issues five tensor copies per row
// (t(tau-1), t(tau), the three face velocities: boxes of 34 x nrows doubles, out-of-range levels zero filled) that complete
// on the mbarrier of the ring slot; the 512 threads no longer compute 6-8 global addresses each per row ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned mb, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mb, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mb, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "FM_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra FM_DONE;\n"
      "bra FM_WAIT;\n"
      "FM_DONE:\n"
      "}\n" ::"r"(mb), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned mb) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(mb), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, unsigned mb) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(mb), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

struct FmMaps {
  CUtensorMap tm1, t0, ue, vn, wb;   // t(tau-1), t(tau) as (imt,km,jl,nt); ue, vn (imt,km,jl); wb (imt,km+1,jl)
};

// staged plane: MAXW + 4 rows of FM_W doubles, padded to a multiple of 128 bytes (TMA destination alignment); a
// compile-time constant, so that every shared-memory access of the march is base register + immediate
template <int MAXW>
struct FmCfg {
  static constexpr int PLANE = (((MAXW + 4) * FM_W + 15) / 16) * 16;
};

struct FmGeom {
  int nit, nkt, TK, nchunk, chunk;   // i tiles, k tiles, levels per k tile, row chunks, rows per chunk
  int nrows;                         // staged levels per row: TK + 4
  int tma;                           // 1: interior i tiles stage their rows with TMA
};

template <int MAXW>
__global__ void __launch_bounds__(32 * MAXW, (MAXW <= 8) ? 2 : 1) k_fct_march(const DevView v, int nbase, int ng, FmGeom gm, const __grid_constant__ FmMaps maps) {
  extern __shared__ __align__(128) unsigned char fm_raw[];
  constexpr int plane = FmCfg<MAXW>::PLANE;         // plane row p <-> level ka_lo - 1 + p
  // skipping the face-flux phase of land warps and the write of land cells pays where registers are not the limit
  // (128-register variants); the 96-register variant (columns of <= 20 levels in one tile) is faster without (measured)
  constexpr bool LAND_SKIP = (MAXW <= 16);
  // ---- which piece of the domain ----
  int bid = blockIdx.x;
  const int g = bid % ng; bid /= ng;
  const int it = bid % gm.nit; bid /= gm.nit;
  const int kt = bid % gm.nkt;
  const int ch = bid / gm.nkt;
  const int imt = v.imt, km = v.km, jmt = v.jmt;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int k0 = 1 + kt * gm.TK, k1 = min(km, k0 + gm.TK - 1);     // levels this CTA updates
  const int ka_lo = max(1, k0 - 1), ka_hi = min(km, k1 + 1);       // levels whose ratios it needs
  const int nA = ka_hi - ka_lo + 1;
  const int ja = v.jlo + ch * gm.chunk, jb = min(v.jhi, ja + gm.chunk - 1);   // rows this CTA updates
  if (ja > jb) return;
  const int rA0 = max(2, ja - 1);
  if (w >= nA) {
    // a warp without a level in this k tile only keeps the barrier count
    if (gm.tma && it > 0 && it < gm.nit - 1) __syncthreads();
    __syncthreads();
    __syncthreads();
    for (int r = rA0; r <= jb + 1; r++) __syncthreads();
    return;
  }
  const int k = ka_lo + w;
  const int jtop = v.jbase + v.jl - 1;
  // own cell in i: lane 0 / 31 are halo cells; wrapped into 2..imt-1 (cyclic, 09/mom/tracer_adv_flx.F:693-694)
  const int i0 = 2 + it * FM_TI;
  auto wrap_i = [&](int i) {
    if (i < 2) i += imt - 2;
    else if (i > imt - 1) i -= imt - 2;
    return min(max(i, 1), imt);
  };
  const int ig = i0 - 1 + lane;
  const int iw = wrap_i(ig);
  const bool out_cell = lane >= 1 && lane <= FM_TI && ig <= imt - 1 && k >= k0 && k <= k1;
  const bool halo_k = LAND_SKIP && (k < k0 || k > k1);   // halo level of the k tile: only its R+-z are needed (by the level next to it)

  // ---- shared memory ----
  double *sT = reinterpret_cast<double *>(fm_raw);  // [FM_NSLOT][nrows][FM_W]
  double *sU = sT + FM_NSLOT * plane;
  double *sUe = sU + FM_NSLOT * plane;              // east-face velocity
  double *sVn = sUe + FM_NSLOT * plane;             // north-face velocity
  double *sWb = sVn + FM_NSLOT * plane;             // bottom-face velocity, plane row p <-> face ka_lo - 1 + p
  double *sR = sWb + FM_NSLOT * plane;              // [2][4][MAXW][32]: R+x, R-x, R+z, R-z, double buffered over rows
  unsigned long long *sMb = reinterpret_cast<unsigned long long *>(sR + 2 * 4 * MAXW * 32);   // [FM_NSLOT] mbarriers (TMA staging)
  int *sK = reinterpret_cast<int *>(sMb + FM_NSLOT);                                         // [FM_NSLOT][FM_W] kmt
  // TMA staging for the i tiles whose 34-wide window lies inside 1..imt (no cyclic wrap); the two edge tiles keep cp.async
  const bool tma = gm.tma && it > 0 && it < gm.nit - 1;
  const unsigned mb0 = smem_u32(sMb);
  const unsigned tx_bytes = 5u * FM_W * (unsigned)gm.nrows * 8u;
  if (tma && threadIdx.x == 0) {
    for (int q = 0; q < FM_NSLOT; q++) mbar_init(mb0 + 8 * q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const double *__restrict__ T = v.t_m1 + (long long)(nbase + g) * v.n3;
  const double *__restrict__ U = v.t_0 + (long long)(nbase + g) * v.n3;
  double *__restrict__ P = v.t_p1 + (long long)(nbase + g) * v.n3;
  const int sj = imt * km, sjz = imt * (km + 1);

  // ---- staging: per-thread element offsets, fixed for the whole march ----
  // element E of a staged row <-> i = i0 - 3 + E; this thread copies E = lane + 2, lanes 0 / 1 also E = 1 / 34;
  // the first / last warp also copies the level above / below the tile's ratio levels
  const int e2 = (lane == 0) ? 1 : FM_W - 2;
  const int ie2 = wrap_i(i0 - 3 + e2);
  const int g_own = (iw - 1) + imt * (k - 1), gz_own = (iw - 1) + imt * k;
  const int g_h = (ie2 - 1) + imt * (k - 1);
  const int s_own = (w + 1) * FM_W + lane + FM_E, s_h = (w + 1) * FM_W + e2;
  const int kx = (w == 0) ? ka_lo - 1 : ((w == nA - 1) ? ka_hi + 1 : -1);
  const bool x_w = kx >= 0 && kx <= km, x_t = kx >= 1 && kx <= km;
  const int g_x = (iw - 1) + imt * (kx - 1), gz_x = (iw - 1) + imt * kx;
  const int s_x = ((w == 0) ? 0 : nA + 1) * FM_W + lane + FM_E;
  const int g_k = iw - 1, g_kh = ie2 - 1;
  const double *__restrict__ gUe = v.ue, *__restrict__ gVn = v.vn, *__restrict__ gWb = v.wb;
  const int *__restrict__ gK = v.kmt;
  const int first_row = max(2, ja - 1) - 1;   // the first staged row: use q of a ring slot holds row first_row + (slot offset) + 4 q
  auto stage = [&](int rr) {
    const int so = (rr & (FM_NSLOT - 1)) * plane;
    const int jloc = min(max(rr, v.jbase), jtop) - v.jbase;
    if (tma) {
      if (threadIdx.x == 0) {
        const unsigned mb = mb0 + 8 * (rr & (FM_NSLOT - 1));
        mbar_expect_tx(mb, tx_bytes);
        // element E of a staged row <-> i = i0 - 3 + E (0-based i0 - 4 + E: an even start, 16-byte aligned -- a box whose
        // first byte is not 16-byte aligned raises `illegal instruction`); plane row p <-> level ka_lo - 1 + p
        // (0-based ka_lo - 2 + p; -1 and km are out of range: zero filled), face ka_lo - 1 + p for the bottom-face velocity
        tma_load_4d(&sT[so], &maps.tm1, i0 - 4, ka_lo - 2, jloc, nbase + g, mb);
        tma_load_4d(&sU[so], &maps.t0, i0 - 4, ka_lo - 2, jloc, nbase + g, mb);
        tma_load_3d(&sUe[so], &maps.ue, i0 - 4, ka_lo - 2, jloc, mb);
        tma_load_3d(&sVn[so], &maps.vn, i0 - 4, ka_lo - 2, jloc, mb);
        tma_load_3d(&sWb[so], &maps.wb, i0 - 4, ka_lo - 1, jloc, mb);
      }
      if (w == 0) {
        const int sko = (rr & (FM_NSLOT - 1)) * FM_W;
        cp_async4(&sK[sko + lane + FM_E], gK + (jloc * imt + g_k));
        if (lane < 2) cp_async4(&sK[sko + e2], gK + (jloc * imt + g_kh));
      }
      cp_async_commit();
      return;
    }
    const int o3 = jloc * sj, o3z = jloc * sjz;
    cp_async8(&sT[so + s_own], T + (o3 + g_own));
    cp_async8(&sU[so + s_own], U + (o3 + g_own));
    cp_async8(&sUe[so + s_own], gUe + (o3 + g_own));
    cp_async8(&sVn[so + s_own], gVn + (o3 + g_own));
    cp_async8(&sWb[so + s_own], gWb + (o3z + gz_own));
    if (lane < 2) {
      cp_async8(&sT[so + s_h], T + (o3 + g_h));
      cp_async8(&sU[so + s_h], U + (o3 + g_h));
      cp_async8(&sUe[so + s_h], gUe + (o3 + g_h));
    }
    if (x_w) cp_async8(&sWb[so + s_x], gWb + (o3z + gz_x));
    if (x_t) {
      cp_async8(&sT[so + s_x], T + (o3 + g_x));
      cp_async8(&sU[so + s_x], U + (o3 + g_x));
    }
    if (w == 0) {
      const int sko = (rr & (FM_NSLOT - 1)) * FM_W;
      cp_async4(&sK[sko + lane + FM_E], gK + (jloc * imt + g_k));
      if (lane < 2) cp_async4(&sK[sko + e2], gK + (jloc * imt + g_kh));
    }
    cp_async_commit();
  };

  // loop-invariant factors of this thread
  const double dxtr_i = v.dxtr[iw - 1];
  const double dcfz = v.dzt2r[k - 1];
  const double c2dtts = v.c2dtts;
  const double twodt = c2dtts * v.dtxcel[k - 1];
  const double *__restrict__ cstr = v.cstr, *__restrict__ cstdyt2r = v.cstdyt2r;
  const int o_c = (w + 1) * FM_W + lane + FM_E;              // own element in a staged plane
  const int o_u = (k > 1) ? o_c - FM_W : o_c;                // clamped k-1 (Tu = Tc at k = 1)
  const int o_d = (k < km) ? o_c + FM_W : o_c;               // clamped k+1
  const int ro = w * 32 + lane;
  const bool has_dn = (w + 1 < nA);                          // the level below is in this CTA (else k = km or lower halo)
  const bool has_up = (w >= 1);

  // ---- prologue: rows rA0-1 .. rA0+2 ----
  // row rr landed?  (slot rr & 3, its use number (rr - first_row) >> 2 gives the mbarrier phase)
  auto wait_row = [&](int rr) {
    if (tma) mbar_wait(mb0 + 8 * (rr & (FM_NSLOT - 1)), ((rr - first_row) >> 2) & 1);
  };
  if (tma) __syncthreads();   // the mbarriers are initialised before anybody waits on them
  stage(rA0 - 1);
  stage(rA0);
  stage(rA0 + 1);
  stage(rA0 + 2);
  cp_async_wait<0>();
  wait_row(rA0 - 1);
  wait_row(rA0);
  wait_row(rA0 + 1);
  wait_row(rA0 + 2);
  __syncthreads();
  double Tc, Uc, Um, lo_n_p, a_n_p;
  int kmc_p;
  double ryp_p = 0.0, rym_p = 0.0, Fn_pp = 0.0;
  {
    const int sm1 = ((rA0 - 1) & 3) * plane + o_c, s0 = (rA0 & 3) * plane + o_c;
    const double Tm = sT[sm1];
    Um = sU[sm1];
    Tc = sT[s0];
    Uc = sU[s0];
    const double vn_m = sVn[sm1];
    // north face of row rA0-1: low-order flux and antidiffusive flux (anti_fn(row 1) = 0, :475)
    lo_n_p = upw(vn_m, Tm, Tc);
    a_n_p = (rA0 - 1 < 2) ? 0.0 : vn_m * (Um + Uc) - lo_n_p;
    kmc_p = sK[((rA0 - 1) & 3) * FM_W + lane + FM_E];
  }
  __syncthreads();   // row rA0-1 has been read: its slot may be refilled
  // state of the previous row's x / z faces, carried across the barrier to where their neighbours' ratios are visible
  double c_ae = 0.0, c_loe = 0.0, c_ad = 0.0, c_lod = 0.0, c_au = 0.0, c_lou = 0.0;
  double c_rxp = 0.0, c_rxm = 0.0, c_rzp = 0.0, c_rzm = 0.0, c_dcfx = 0.0;

#pragma unroll 1
  for (int r = rA0; r <= jb + 1; r++) {
    const bool doA = r <= jmt - 1;
    stage(r + 3);
    // land cells are not written: the implicit solve multiplies the tendency by tmask (09/mom/tracer.F:1114-1127)
    const bool will_out = (r - 1 >= ja) && (r - 1 <= jb) && out_cell && (!LAND_SKIP || kmc_p >= k);
    const int cj = (r - 1 - v.jbase) * sj + g_own;
    double pd = 0.0;
    if (will_out) pd = P[cj];   // diffusive tendency of row r-1
    // ---- x and z faces of row r-1 (ratios of the neighbours became visible at the last barrier) ----
    double tx_p = 0.0, tz_p = 0.0;
    const double m_p = (kmc_p >= k) ? 1.0 : 0.0;   // tmask(i,k,r-1)
    if (!halo_k && (!LAND_SKIP || __any_sync(0xffffffffu, kmc_p >= k))) {   // a warp of land cells, or a halo level, has no face flux anybody reads
      const double *R = sR + ((r - 1) & 1) * (4 * MAXW * 32);
      // east / west faces: Cpos(f) = min(Rpl(f+1),Rmn(f)), Cneg(f) = min(Rpl(f),Rmn(f+1)) (:698-701); no mask (:987)
      // the west face of a cell is the east face of its western neighbour: one lane over (lane 0 is a halo cell)
      const int le = (lane < 31) ? ro + 1 : ro;
      const double Fe = delimit(dmin(R[le], c_rxm), dmin(c_rxp, R[MAXW * 32 + le]), c_ae) + c_loe;
      const double Fw = __shfl_up_sync(0xffffffffu, Fe, 1);
      // bottom / top faces: Cpos(h) = min(Rpl(h),Rmn(h+1)), Cneg(h) = min(Rpl(h+1),Rmn(h)) (:966-969);
      // adv_fb(0), adv_fb(km) are overwritten in tracer (09/mom/tracer.F:1063-1065): level 1 / km carry those in c_lou / c_lod
      const int ld = has_dn ? ro + 32 : ro, lu = has_up ? ro - 32 : ro;
      const double mu_p = (kmc_p >= k - 1) ? 1.0 : 0.0;
      double Fb = (delimit(dmin(c_rzp, R[3 * MAXW * 32 + ld]), dmin(R[2 * MAXW * 32 + ld], c_rzm), c_ad) + c_lod) * m_p;
      double Fu = (delimit(dmin(R[2 * MAXW * 32 + lu], c_rzm), dmin(c_rzp, R[3 * MAXW * 32 + lu]), c_au) + c_lou) * mu_p;
      if (k == km) Fb = c_lod;
      if (k == 1) Fu = c_lou;
      tx_p = (Fe - Fw) * c_dcfx;
      tz_p = (Fu - Fb) * dcfz;
    }
    double ryp = 0.0, rym = 0.0, lo_n = 0.0, a_n = 0.0, Un = 0.0, Tn = 0.0;
    int kmc = 0;
    if (doA) {
      const int so = (r & 3) * plane, sn = ((r + 1) & 3) * plane;
      const int sko = (r & 3) * FM_W + lane + FM_E;
      const double Te = sT[so + o_c + 1], Tw = sT[so + o_c - 1], Tu = sT[so + o_u], Td = sT[so + o_d];
      const double Ue = sU[so + o_c + 1], Uw = sU[so + o_c - 1], Uu = sU[so + o_u], Ud = sU[so + o_d];
      Tn = sT[sn + o_c];
      Un = sU[sn + o_c];
      const double ue_c = sUe[so + o_c], ue_w = sUe[so + o_c - 1], vn_c = sVn[so + o_c];
      const double wb_d = sWb[so + o_c], wb_u = sWb[so + o_c - FM_W];
      kmc = sK[sko];
      const double m = (kmc >= k) ? 1.0 : 0.0;
      const double mu = (kmc >= k - 1) ? 1.0 : 0.0;
      const bool mw_b = sK[sko - 1] >= k, me_b = sK[sko + 1] >= k;
      const bool ms_b = kmc_p >= k, mn_b = sK[((r + 1) & 3) * FM_W + lane + FM_E] >= k;
      const double dcfx = cstr[r - 1] * dxtr_i * 0.5;
      const double dcfy = cstdyt2r[r - 1];
      // low-order fluxes of the six faces and the low-order solution (:496-580)
      const double lo_e = upw(ue_c, Tc, Te);
      const double lo_w = upw(ue_w, Tw, Tc);
      lo_n = upw(vn_c, Tc, Tn);
      const double lo_s = lo_n_p;
      const double lo_d = upw(wb_d, Td, Tc);
      const double lo_u = upw(wb_u, Tc, Tu);
      double tlo;
      {
        const double tx = (lo_e - lo_w) * dcfx;
        const double ty = (lo_n - lo_s) * dcfy;
        const double fb_u = (k == 1) ? wb_u * 2.0 * Tc : lo_u;   // adv_fb(i,0,j) = adv_vbt(i,0,j)*c2*t(i,1,j) (:543)
        const double fb_d = (k == km) ? 0.0 : lo_d;              // adv_fb(i,km,j) = c0 (:544)
        const double tz = (fb_u - fb_d) * dcfz;
        tlo = Tc - twodt * (tx + ty + tz) * m;
      }
      double rxp = 0.0, rxm = 0.0, rzp = 0.0, rzm = 0.0;
      // A land cell has tmask = 0 in front of Q+-, so its six ratios are exactly zero whatever the fluxes are
      // (:688-691, 764-767, 952-955).  Warps whose 32 cells are all land (continents, levels below the bottom) skip the
      // three limiters -- two thirds of the work of a row -- without diverging; everything else runs as usual.
      const bool any_ocean = __any_sync(0xffffffffu, kmc >= k);
      // ---- x (:635-694): flxlft = anti_fe(i-1), flxrgt = anti_fe(i) ----
      const double a_w = ue_w * (Uw + Uc) - lo_w;
      const double a_e = ue_c * (Uc + Ue) - lo_e;
      if (any_ocean && !halo_k) {
        // mask*(average) + (1-mask)*t_lo with a 0/1 mask is one of the two terms (up to the sign of a zero): select it
        const double fxa = mw_b ? 0.5 * (Uw + Uc) : tlo;
        const double fxb = me_b ? 0.5 * (Uc + Ue) : tlo;
        ratio(c2dtts, dcfx, a_w, a_e, fxa, fxb, tlo, m, rxp, rxm);
      }
      // ---- y (:714-770): flxlft = anti_fn(j-1) (= 0 for row 1, :475), flxrgt = anti_fn(j) ----
      a_n = vn_c * (Uc + Un) - lo_n;
      if (any_ocean && !halo_k) {
        const double fxa = ms_b ? 0.5 * (Um + Uc) : tlo;
        const double fxb = mn_b ? 0.5 * (Uc + Un) : tlo;
        ratio(c2dtts, dcfy, a_n_p, a_n, fxa, fxb, tlo, m, ryp, rym);
      }
      // ---- z (:786-958): flxlft = anti_fb(k), flxrgt = anti_fb(k-1) ----
      // anti_fb(i,0,j) = adv_vbt(i,0,j)*c2*t(i,1,j,taum1) (:617); anti_fb(i,km,j) = 0
      const double a_d = wb_d * (Uc + Ud) - lo_d * m;
      const double a_u = wb_u * (Uu + Uc) - lo_u * mu;
      if (any_ocean) {
        const double fxa = (k > 1 && kmc >= k - 1) ? 0.5 * (Uu + Uc) : tlo;
        const double fxb = (k < km && kmc >= k + 1) ? 0.5 * (Uc + Ud) : tlo;
        ratio(c2dtts, dcfz, (k == km) ? 0.0 : a_d, (k == 1) ? wb_u * 2.0 * Tc : a_u, fxa, fxb, tlo, m, rzp, rzm);
      }
      double *R = sR + (r & 1) * (4 * MAXW * 32);
      R[ro] = rxp;
      R[MAXW * 32 + ro] = rxm;
      R[2 * MAXW * 32 + ro] = rzp;
      R[3 * MAXW * 32 + ro] = rzm;
      c_ae = a_e; c_loe = lo_e; c_ad = a_d; c_au = a_u;
      c_lod = (k == km) ? wb_d * Uc : lo_d;           // adv_fb(i,km,j) = adv_vbt(i,km,j)*t(i,km,j,tau)
      c_lou = (k == 1) ? wb_u * (Uc + Uc) : lo_u;     // adv_fb(i,0,j) = adv_vbt(i,0,j)*2*t(i,1,j,tau)
      c_rxp = rxp; c_rxm = rxm; c_rzp = rzp; c_rzm = rzm; c_dcfx = dcfx;
    }
    // ---- finish row j = r-1: its north face needed R+-y of row r (:772-784, 985-1002) ----
    {
      const double Fn = (delimit(dmin(ryp, rym_p), dmin(ryp_p, rym), a_n_p) + lo_n_p) * m_p;
      if (will_out) {
        const double adv_ty = (Fn - Fn_pp) * cstdyt2r[r - 2];
        P[cj] = pd - tx_p - adv_ty - tz_p;
      }
      Fn_pp = Fn;
    }
    lo_n_p = lo_n; a_n_p = a_n; ryp_p = ryp; rym_p = rym; kmc_p = kmc;
    Um = Uc; Uc = Un; Tc = Tn;
    cp_async_wait<1>();   // everything but the copies issued in this iteration has landed: rows <= r+2
    wait_row(r + 2);
    __syncthreads();      // ... and, like the ratios of row r, is visible to every thread
  }
  cp_async_wait<0>();
  wait_row(jb + 4);       // the last row staged ahead: no copy may be in flight when the CTA retires
}

// launch geometry: a CTA has at most maxw warps (one per level whose ratios it needs); latitude chunks sized to give
// every SM several CTAs while keeping the two warm-up rows of a chunk a small fraction of its work
static FmGeom fct_geometry(const DevView &v, int ng, int maxw) {
  FmGeom g;
  g.nit = (v.imt - 2 + FM_TI - 1) / FM_TI;
  // one tile if the column fits the CTA; two tiles need TK + 1 warps each, more need TK + 2 in the middle
  g.nkt = (v.km <= maxw) ? 1 : ((v.km <= 2 * (maxw - 1)) ? 2 : (v.km + maxw - 3) / (maxw - 2));
  g.TK = (v.km + g.nkt - 1) / g.nkt;
  g.nrows = g.TK + 4;
  const int rows = v.jhi - v.jlo + 1;
  // rows per chunk: every chunk recomputes the ratios of the two rows before its first (warm-up), so long chunks win as
  // long as the grid keeps many CTAs per SM -- measured on B200, 0.5 degree x 40 tracers (358 rows): 64 rows 11.06 ms,
  // 96: 10.86, 128: 10.76, 180: 10.68.  UVIC_B200_FCT_CHUNK overrides (experiments).
  const int chunk_rows = getenv("UVIC_B200_FCT_CHUNK") ? std::max(8, atoi(getenv("UVIC_B200_FCT_CHUNK"))) : 192;
  int nchunk = std::max(1, (rows + chunk_rows - 1) / chunk_rows);
  const long long per_chunk = (long long)ng * g.nit * g.nkt;
  while (per_chunk * nchunk < 148 * 4 && rows / (nchunk + 1) >= 8) nchunk++;   // small tracer batches: shorter marches, more CTAs
  g.chunk = (rows + nchunk - 1) / nchunk;
  g.nchunk = (rows + g.chunk - 1) / g.chunk;
  return g;
}

// ---- tensor maps (host): built once per device pointer and box height, cached in the context ----
typedef CUresult (*FmEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                               CUtensorMapFloatOOBfill);
static FmEncodeFn fm_encode_fn() {
  static FmEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (FmEncodeFn)p;
  }
  return fn;
}
// (imt, nlev, jl[, nt]) doubles, box 34 x nrows x 1 [x 1], no swizzle, zero fill outside
static bool fm_map(uvic_b200_ctx *c, const double *ptr, int imt, int nlev, int jl, int nt4, int nrows, CUtensorMap *out) {
  auto key = std::make_pair((const void *)ptr, nrows * 8 + (nt4 > 0 ? 1 : 0));
  auto it = c->tma_maps.find(key);
  if (it != c->tma_maps.end()) {
    memcpy(out, it->second.data(), sizeof(CUtensorMap));
    return true;
  }
  FmEncodeFn enc = fm_encode_fn();
  if (!enc) return false;
  const int rank = nt4 > 0 ? 4 : 3;
  cuuint64_t dims[4] = {(cuuint64_t)imt, (cuuint64_t)nlev, (cuuint64_t)jl, (cuuint64_t)std::max(nt4, 1)};
  cuuint64_t strides[3] = {(cuuint64_t)imt * 8, (cuuint64_t)imt * nlev * 8, (cuuint64_t)imt * nlev * jl * 8};
  cuuint32_t box[4] = {FM_W, (cuuint32_t)nrows, 1, 1}, es[4] = {1, 1, 1, 1};
  CUtensorMap m;
  if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, (void *)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  std::array<unsigned char, 128> blob;
  memcpy(blob.data(), &m, sizeof(CUtensorMap));
  c->tma_maps[key] = blob;
  *out = m;
  return true;
}

template <int MAXW>
static void fct_launch_t(uvic_b200_ctx *c, int nbase, int ng) {
  DevView &v = c->v;
  FmGeom g = fct_geometry(v, ng, MAXW);
  constexpr size_t plane = FmCfg<MAXW>::PLANE;
  const size_t shm = (5 * FM_NSLOT * plane + 8 * MAXW * 32) * sizeof(double) + FM_NSLOT * sizeof(unsigned long long) + FM_NSLOT * FM_W * sizeof(int);
  ensure_dyn_smem(c, (const void *)k_fct_march<MAXW>, shm);
  // TMA staging needs 16-byte multiples for the row strides (imt even) and at least one interior i tile
  static const bool tma_env = !(getenv("UVIC_B200_FCT_TMA") && atoi(getenv("UVIC_B200_FCT_TMA")) == 0);
  FmMaps maps;
  memset(&maps, 0, sizeof maps);
  g.tma = 0;
  if (tma_env && (v.imt % 2) == 0 && g.nit >= 3 && g.nrows <= 256) {
    const bool ok = fm_map(c, v.t_m1, v.imt, v.km, v.jl, v.nt, g.nrows, &maps.tm1) && fm_map(c, v.t_0, v.imt, v.km, v.jl, v.nt, g.nrows, &maps.t0) &&
                    fm_map(c, v.ue, v.imt, v.km, v.jl, 0, g.nrows, &maps.ue) && fm_map(c, v.vn, v.imt, v.km, v.jl, 0, g.nrows, &maps.vn) &&
                    fm_map(c, v.wb, v.imt, v.km + 1, v.jl, 0, g.nrows, &maps.wb);
    g.tma = ok ? 1 : 0;
  }
  const int nwarp = (g.nkt == 1) ? v.km : std::min(MAXW, g.TK + (g.nkt == 2 ? 1 : 2));   // most levels any k tile needs ratios for
  const long long nblk = (long long)ng * g.nit * g.nkt * g.nchunk;
  ProfScope ps_(c, "k_fct_march");
  k_fct_march<MAXW><<<(unsigned)nblk, 32 * nwarp, shm, c->stream>>>(v, nbase, ng, g, maps);
}

void launch_fct_march(uvic_b200_ctx *c, int nbase, int ng) {
  // Warps per CTA decide the registers a thread may hold (a sub-partition owns 16 K registers): 16 warps -> 128,
  // 20 -> 96, 21 -> 80.  The kernel wants ~160, so fewer, fatter CTAs win once the column does not fit 20 warps
  // (measured on B200, profiles/).  UVIC_B200_FCT_MAXW overrides the choice (experiments).
  int maxw = (c->v.km <= 20) ? 20 : 16;
  if (const char *e = getenv("UVIC_B200_FCT_MAXW")) maxw = atoi(e);
  if (maxw <= 8) fct_launch_t<8>(c, nbase, ng);        // two CTAs per SM (experiment: UVIC_B200_FCT_MAXW=8)
  else if (maxw <= 12) fct_launch_t<12>(c, nbase, ng);
  else if (maxw <= 16) fct_launch_t<16>(c, nbase, ng);
  else if (maxw <= 20) fct_launch_t<20>(c, nbase, ng);
  else fct_launch_t<21>(c, nbase, ng);
}
