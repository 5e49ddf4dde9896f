// fct_common.h -- scalar building blocks of the FCT scheme shared by k_tracer.cu and k_fct.cu
#pragma once
#include "ctx.h"

// upstream flux 2*(v*T)_face, 09/mom/tracer_adv_flx.F:500-503: totadv*(a+b) + |totadv|*(a-b)
__device__ __forceinline__ double upw(double totadv, double a, double b) { return totadv * (a + b) + fabs(totadv) * (a - b); }

// Fortran max/min on finite operands: one DSETP + two selects (CUDA's fmax/fmin add NaN handling)
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }

__device__ __forceinline__ void ratio(double c2dtts, double dcf, double flxlft, double flxrgt, double fxa, double fxb, double tlo,
                                      double m, double &rpl, double &rmn) {
  double trmax = dmax(dmax(fxa, fxb), tlo);
  double trmin = dmin(dmin(fxa, fxb), tlo);
  double pplus = c2dtts * dcf * (dmax(0.0, flxlft) - dmin(0.0, flxrgt));
  double pminus = c2dtts * dcf * (dmax(0.0, flxrgt) - dmin(0.0, flxlft));
  double qplus = trmax - tlo;
  double qminus = tlo - trmin;
  rpl = dmin(1., div0(m * qplus, pplus + UVIC_EPSLN));
  rmn = dmin(1., div0(m * qminus, pminus + UVIC_EPSLN));
}


__device__ __forceinline__ double delimit(double cpos, double cneg, double a) {
  // :706-711, 777-782, 972-977
  return 0.5 * ((cpos + cneg) * a + (cpos - cneg) * fabs(a));
}

