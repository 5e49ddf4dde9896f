// fct_common.h -- scalar building blocks of the FCT scheme shared by k_tracer.cu and k_fct.cu
#pragma once
#include "ctx.h"

// upstream flux 2*(v*T)_face, 09/mom/tracer_adv_flx.F:500-503: totadv*(a+b) + |totadv|*(a-b)
__device__ __forceinline__ double upw(double totadv, double a, double b) { return totadv * (a + b) + fabs(totadv) * (a - b); }

// Fortran max/min on finite operands: one DSETP + two selects (CUDA's fmax/fmin add NaN handling)
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }

// the limiter ratios divide Q >= 0 by P + 1e-20 > 0: qdiv (ctx.h) applies
__device__ __forceinline__ double fdiv_pos(double x, double y) { return qdiv(x, y); }

__device__ __forceinline__ void ratio(double c2dtts, double dcf, double flxlft, double flxrgt, double fxa, double fxb, double tlo,
                                      double m, double &rpl, double &rmn) {
  double trmax = dmax(dmax(fxa, fxb), tlo);
  double trmin = dmin(dmin(fxa, fxb), tlo);
  // max(0,f) - min(0,g) with min(0,g) = g - max(0,g) (exact: one of the two terms is g, the other 0)
  const double lp = dmax(0.0, flxlft), rp = dmax(0.0, flxrgt);
  double pplus = c2dtts * dcf * (lp - (flxrgt - rp));
  double pminus = c2dtts * dcf * (rp - (flxlft - lp));
  double qplus = trmax - tlo;
  double qminus = tlo - trmin;
  rpl = dmin(1., fdiv_pos(m * qplus, pplus + UVIC_EPSLN));
  rmn = dmin(1., fdiv_pos(m * qminus, pminus + UVIC_EPSLN));
}


__device__ __forceinline__ double delimit(double cpos, double cneg, double a) {
  // :706-711, 777-782, 972-977
  return 0.5 * ((cpos + cneg) * a + (cpos - cneg) * fabs(a));
}

