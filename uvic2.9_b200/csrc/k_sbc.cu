// k_sbc.cu -- the surface boundary conditions either side of the tracer step, on the device
// (SURVEY.md 8f rank 2): the host then moves the coupler's 2-D flux array per step instead of
// stf/btf for every tracer, and reads the surface accumulators back at the end of an ocean
// segment instead of full 3-D fields.
//
//   k_setvbc    09/mom/setvbc.F:60-140: stf(i,j,n) = sbc(i,jrow,flux slot of n)*tmask(i,1,j) for
//               every tracer that owns a flux slot, zero otherwise; btf = 0 except
//               btf(i,j,itemp) = -bhf(i,jrow)*tmask(i,1,j)
//   k_set_sbc   09/mom/set_sbc.F:36-83 as called at the end of `tracer` (09/mom/tracer.F:1270-1288)
//               with doAccum = .true.: zero at the start of an ocean segment (ocean cells only),
//               accumulate t(i,1,j,n,tau+1) every step, average at the end of the segment
#include "ctx.h"

__global__ void __launch_bounds__(128) k_setvbc(const DevView v) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ni = v.imt - 2;
  if (idx >= (long long)ni * v.jl) return;
  const int i = (int)(idx % ni) + 2;                 // istrt = 2 .. iend = imt-1
  const int j = (int)(idx / ni) + v.jbase;
  const long long c2 = X2(i, j);
  const double tm = (v.kmt[c2] >= 1) ? 1.0 : 0.0;    // tmask(i,1,j), 09/mom/loadmw.F:60-77
  for (int n = 0; n < v.nt; n++) {
    const int m = v.sbc_flx[n];
    v.stf[c2 + (long long)n * v.n2] = (m > 0) ? v.sbc[c2 + (long long)(m - 1) * v.n2] * tm : 0.0;
    v.btf[c2 + (long long)n * v.n2] = 0.0;
  }
  v.btf[c2] = -v.bhf[c2] * tm;                       // itemp = 1
}

__global__ void __launch_bounds__(128) k_set_sbc(const DevView v, int eots, int osegs, int osege, double rts) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ni = v.imt - 2;
  const int nrow = v.jhi - v.jlo + 1;
  if (idx >= (long long)ni * nrow) return;
  const int i = (int)(idx % ni) + 2;
  const int j = (int)(idx / ni) + v.jlo;
  const long long c2 = X2(i, j);
  const bool ocean = v.kmt[c2] != 0;
  const long long c1 = X3(i, 1, j);
  for (int n = 0; n < v.nt; n++) {
    const int m = v.sbc_acc[n];
    if (m <= 0) continue;
    double s = v.sbc[c2 + (long long)(m - 1) * v.n2];
    if (eots && osegs && ocean) s = 0.0;
    if (eots) s = s + v.t_p1[c1 + (long long)n * v.n3];
    if (eots && osege && ocean) s = rts * s;
    v.sbc[c2 + (long long)(m - 1) * v.n2] = s;
  }
}

void launch_setvbc(uvic_b200_ctx *c) {
  DevView &v = c->v;
  const long long n = (long long)(v.imt - 2) * v.jl;
  KLAUNCH("k_setvbc", k_setvbc, cdiv(n, 128), 128, v);
}

void launch_set_sbc(uvic_b200_ctx *c, int eots, int osegs, int osege, int ntspos) {
  DevView &v = c->v;
  const long long n = (long long)(v.imt - 2) * (v.jhi - v.jlo + 1);
  KLAUNCH("k_set_sbc", k_set_sbc, cdiv(n, 128), 128, v, eots, osegs, osege, 1.0 / (double)ntspos);
}

// ---------------------------------------------------------------------------------------
// Time averages of the tracers: the tracer part of avgvar / avgout
// (09/mom/timeavgs.F:206-375, 398-420; called from 09/mom/diag.F:138-146 with t(tau)) on the
// default averaging grid (the whole model grid).  With the sums on the device the host fetches
// full 3-D fields once per averaging period instead of once per step.
//   k_tavg_accumulate   spbuf(i,k,j,n) += t(i,k,j,n,tau);
//                       spbuf2(i,j,n) += stf(i,j,n) [- vflux(i,j)*gaost(n) for all but T and S]
//   k_tavg_mean         avg = rnavgt * spbuf (rnavgt = 1/navgts)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tavg_accumulate(const DevView v, double *__restrict__ sum_t, double *__restrict__ sum_stf,
                                                         const double *__restrict__ vflux, const double *__restrict__ gaost) {
  const long long per = (long long)v.imt * v.km * (v.jhi - v.jlo + 1);   // cells of the owned rows, all i
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per) return;
  const long long row0 = (long long)v.imt * v.km * (v.jlo - v.jbase);
  const long long c = row0 + idx;
  for (int n = 0; n < v.nt; n++) sum_t[c + (long long)n * v.n3] = sum_t[c + (long long)n * v.n3] + v.t_0[c + (long long)n * v.n3];
  // the 2-D part: one thread per (i, j)
  const long long per2 = (long long)v.imt * (v.jhi - v.jlo + 1);
  if (idx < per2) {
    const long long c2 = (long long)v.imt * (v.jlo - v.jbase) + idx;
    for (int n = 0; n < v.nt; n++) {
      double s = sum_stf[c2 + (long long)n * v.n2] + v.stf[c2 + (long long)n * v.n2];
      if (n >= 2) s = s - (vflux ? vflux[c2] : 0.0) * (gaost ? gaost[n] : 0.0);
      sum_stf[c2 + (long long)n * v.n2] = s;
    }
  }
}

__global__ void __launch_bounds__(256) k_tavg_mean(const double *__restrict__ sum, double *__restrict__ avg, long long n, double rnavgt) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) avg[idx] = rnavgt * sum[idx];
}

void launch_tavg_accumulate(uvic_b200_ctx *c, const double *vflux_dev, const double *gaost_dev) {
  DevView &v = c->v;
  const long long per = (long long)v.imt * v.km * (v.jhi - v.jlo + 1);
  KLAUNCH("k_tavg_accumulate", k_tavg_accumulate, cdiv(per, 256), 256, v, c->tavg_t, c->tavg_stf, vflux_dev, gaost_dev);
}
void launch_tavg_mean(uvic_b200_ctx *c, const double *sum, double *avg, long long n, double rnavgt) {
  KLAUNCH("k_tavg_mean", k_tavg_mean, cdiv(n, 256), 256, sum, avg, n, rnavgt);
}
