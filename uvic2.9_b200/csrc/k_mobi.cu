// k_mobi.cu -- MOBI biogeochemical source terms on the device (options of run/mk.in:
// O_mobi, O_mobi_alk/_caco3/_o2/_nitrogen/_nitrogen_15/_silicon/_iron, O_carbon,
// O_carbon_13, O_carbon_14).
//
//   k_mobi_light   day fraction, incoming solar and the light at the top of every level
//                  (09/mom/tracer.F:370-390, 09/mom/mobi.F:795-812), one thread per column.
//   k_mobi_cell    everything per cell that does not depend on the sinking chain, one thread
//                  per ocean cell: co2calc_SWS + drtsafe + ta_iter_SWS (09/common/co2calc.F),
//                  AOU, rate factors, the Evans-Parslow integrals, denitrification switches.
//   k_mobi_column  mobi_driver (09/mom/mobi.F:519-1483) and the Euler sub-steps of mobi_src
//                  (:2148-3313), one thread per water column (columns sorted by depth).
//   k_mobi_ws      the same, warp specialised: 8 warps share 32 columns (small grids).
//                  The reference's three k-loops are fused: loops 2 and 3 of mobi_driver
//                  (:1301-1400) only touch level-k quantities, so running them directly
//                  after level k of loop 1 performs the same operations in the same order
//                  on every src(k,slot).  The c14 source (tracer.F:848-867) and the dust /
//                  hydrothermal iron (:536-545) are applied in the same pass.
//
// FP64 throughout; the sinking chain (expo -> impo) is inherently serial down the column
// and the nbio Euler sub-steps are serial in time, so the parallel axis is the column.
#include "ctx.h"
#include "mobi_par.h"
#include <stdlib.h>
#include <algorithm>

// k_mobi_cell is long straight-line code (~120 KB of SASS when everything is inlined, most of it copies of exp / log /
// pow / tanh and of ta_iter's divides) and was stalled on instruction fetch; the pre-pass kernels call one out-of-line
// body per function.  Same library code, same results.  The sub-step loops of the column kernels keep the inlined
// versions (measured: out-of-line calls cost them 2-4 %).
__device__ __noinline__ double m_exp(double x) { return exp(x); }
__device__ __noinline__ double m_log(double x) { return log(x); }
__device__ __noinline__ double m_log10(double x) { return log10(x); }
__device__ __noinline__ double m_pow(double x, double y) { return pow(x, y); }
__device__ __noinline__ double m_tanh(double x) { return tanh(x); }

#define TRCMIN 5e-12        // 09/mom/mobi.h:199
#define RN15STD 0.0036765   // 09/mom/mobi.h
#define RC13STD 0.0112372
#define RC14STD 1.176e-12

__device__ __forceinline__ double fsign(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }
__device__ __forceinline__ double tflag(double x) { return 0.5 + fsign(0.5, x - TRCMIN); }
__device__ __forceinline__ double sq(double x) { return x * x; }

// ------------------------------------------------------------------------------------
// carbonate chemistry
// ------------------------------------------------------------------------------------
struct Carb {
  double k1, k2, kw, kb, ks, kf, k1p, k2p, k3p, ksi;  // COMMON /const/
  double bt, st, ft, sit, pt, dic, ta;                // COMMON /species/
};


// 09/common/co2calc.F:455-526
__device__ __forceinline__ void ta_iter(const Carb &q, double x, double &fn, double &df) {
  double x2 = x * x;
  double x3 = x2 * x;
  double k12 = q.k1 * q.k2;
  double k12p = q.k1p * q.k2p;
  double k123p = k12p * q.k3p;
  double c = 1.0 + qdiv(q.st, q.ks) + qdiv(q.ft, q.kf);
  double a = x3 + q.k1p * x2 + k12p * x + k123p;
  double a2 = a * a;
  double da = 3.0 * x2 + 2.0 * q.k1p * x + k12p;
  double b = x2 + q.k1 * x + k12;
  double b2 = b * b;
  double db = 2.0 * x + q.k1;
  double xc = qdiv(x, c);
  double hs = 1.0 + qdiv(q.ks, xc), hf = 1.0 + qdiv(q.kf, xc);
  double bb = 1.0 + qdiv(x, q.kb), ss = 1.0 + qdiv(x, q.ksi);
  fn = qdiv(q.k1 * x * q.dic, b) + qdiv(2.0 * q.dic * k12, b) + qdiv(q.bt, bb) + qdiv(q.kw, x) + qdiv(q.pt * k12p * x, a) +
       qdiv(2.0 * q.pt * k123p, a) + qdiv(q.sit, ss) - qdiv(x, c) - qdiv(q.st, hs) - qdiv(q.ft, hf) - qdiv(q.pt * x3, a) - q.ta;
  df = qdiv((q.k1 * q.dic * b) - q.k1 * x * q.dic * db, b2) - qdiv(2.0 * q.dic * k12 * db, b2) - qdiv(qdiv(q.bt, q.kb), bb * bb) -
       qdiv(q.kw, x2) + qdiv(q.pt * k12p * (a - x * da), a2) - qdiv(2.0 * q.pt * k123p * da, a2) -
       qdiv(qdiv(q.sit, q.ksi), ss * ss) - qdiv(1.0, c) - q.st * qdiv(1.0, hs * hs) * qdiv(q.ks * c, x2) -
       q.ft * qdiv(1.0, hf * hf) * qdiv(q.kf * c, x2) - qdiv(q.pt * x2 * (3.0 * a - x * da), a2);
}

// 09/common/co2calc.F:401-454 (Numerical Recipes rtsafe, error trapping removed).  The reference evaluates ta_iter at
// x1, at x2, at the midpoint and then once per iteration; here all evaluations go through ONE inlined copy of ta_iter
// (it holds ~30 divides: four copies were a third of k_mobi_cell's code) driven by a small state machine that performs
// the same evaluations in the same order.
__device__ double drtsafe(const Carb &q, double x1, double x2, double xacc) {
  double fl = 0.0, f = 0.0, df = 0.0, xl = 0.0, xh = 0.0, r = 0.0, dxold = 0.0, dx = 0.0;
  double x = x1;
  int stage = 0, it = 0;
  while (true) {
    double fv, dfv;
    ta_iter(q, x, fv, dfv);
    if (stage == 0) {          // f(x1)
      fl = fv;
      x = x2;
      stage = 1;
      continue;
    }
    if (stage == 1) {          // f(x2): orient the bracket, start from the midpoint
      if (fl < 0.0) {
        xl = x1;
        xh = x2;
      } else {
        xh = x1;
        xl = x2;
      }
      r = 0.5 * (x1 + x2);
      dxold = fabs(x2 - x1);
      dx = dxold;
      x = r;
      stage = 2;
      continue;
    }
    f = fv;
    df = dfv;
    if (stage == 3) {          // evaluation at the end of an iteration: shrink the bracket
      if (f < 0.0)
        xl = r;
      else
        xh = r;
    }
    stage = 3;
    if (++it > 100) return r;
    if (((r - xh) * df - f) * ((r - xl) * df - f) >= 0. || fabs(2.0 * f) > fabs(dxold * df)) {
      dxold = dx;
      dx = 0.5 * (xh - xl);
      r = xl + dx;
      if (xl == r) return r;
    } else {
      dxold = dx;
      dx = f / df;
      double temp = r;
      r = r - dx;
      if (temp == r) return r;
    }
    if (fabs(dx) < xacc) return r;
    x = r;
  }
}

// 09/common/co2calc.F:1-400; returns CO2* (mol m-3) and Omega_calcite, the two outputs MOBI uses
// AIR: also the air-sea difference dco2star = co2*ff*atmpres - co2star (:320-331; only gasbc asks for it)
template <bool AIR>
__device__ void co2calc_sws(double t, double s, double dic_in, double ta_in, double depth, double &co2star_out, double &omega_c,
                            double co2_in = 0.0, double atmpres = 1.0, double *dco2star_out = nullptr) {
  Carb q;
  const double permil = 1.0 / 1024.5;
  q.pt = 0.5125e-3 * permil;   // hard-wired phosphate and silicate (:113-114)
  q.sit = 7.6875e-03 * permil;
  q.ta = ta_in * permil;
  q.dic = dic_in * permil;
  const double pres = depth * 0.1;
  double tk = 273.15 + t;
  double tk100 = QDIV(tk, 100.0);
  double tk1002 = tk100 * tk100;
  double invtk = QDIV(1.0, tk);
  double dlogtk = m_log(tk);
  double is = QDIV(19.924 * s, (1000. - 1.005 * s));
  double is2 = is * is;
  double sqrtis = sqrt(is);
  double s2 = s * s;
  double t2 = t * t;
  double sqrts = sqrt(s);
  double s15 = m_pow(s, 1.5);
  double scl = QDIV(s, 1.80655);
  double pitkR = QDIV(QDIV(pres, tk), 83.15);
  double p2itkR = pres * pitkR;
  q.bt = QDIV(0.000232 * scl, 10.811);
  q.st = QDIV(0.14 * scl, 96.062);
  q.ft = QDIV(0.000067 * scl, 18.9984);
  (void)tk1002;

  q.k1 = m_pow(10., (-1. * (3670.7 * invtk - 62.008 + 9.7944 * dlogtk - 0.0118 * s + 0.000116 * s2))) *
         m_exp((25.5 - 0.1271 * t) * pitkR + 0.5 * (-3.08e-3 + 8.77e-5 * t) * p2itkR);
  q.k2 = m_pow(10., (-1 * (1394.7 * invtk + 4.777 - 0.0184 * s + 0.000118 * s2))) *
         m_exp((15.82 + 0.0219 * t) * pitkR + 0.5 * (1.13e-3 - 1.475e-4 * t) * p2itkR);
  q.k1p = m_exp(-4576.752 * invtk + 115.540 - 18.453 * dlogtk + (-106.736 * invtk + 0.69171) * sqrts + (-0.65643 * invtk - 0.01844) * s) *
          m_exp((14.51 - 0.1211 * t + 3.21e-4 * t2) * pitkR + 0.5 * (-2.67e-3 + 4.27e-5 * t) * p2itkR);
  q.k2p = m_exp(-8814.715 * invtk + 172.1033 - 27.927 * dlogtk + (-160.340 * invtk + 1.3566) * sqrts + (0.37335 * invtk - 0.05778) * s) *
          m_exp((23.12 - 0.1758 * t + 2.647e-3 * t2) * pitkR + 0.5 * (-5.15e-3 + 9.0e-5 * t) * p2itkR);
  q.k3p = m_exp(-3070.75 * invtk - 18.126 + (17.27039 * invtk + 2.81197) * sqrts + (-44.99486 * invtk - 0.09984) * s) *
          m_exp((26.57 - 0.202 * t + 3.042e-3 * t2) * pitkR + 0.5 * (-4.08e-3 + 7.14e-5 * t) * p2itkR);
  q.ksi = m_exp(-8904.2 * invtk + 117.400 - 19.334 * dlogtk + (-458.79 * invtk + 3.5913) * sqrtis + (188.74 * invtk - 1.5998) * is +
              (-12.1652 * invtk + 0.07871) * is2 + m_log(1.0 - 0.001005 * s)) *
          m_exp((29.48 - 0.1622 * t - 2.608e-3 * t2) * pitkR + 0.5 * (-2.84e-3) * p2itkR);
  q.kw = m_exp(-13847.26 * invtk + 148.9802 - 23.6521 * dlogtk + (118.67 * invtk - 5.977 + 1.0495 * dlogtk) * sqrts - 0.01615 * s) *
         m_exp((20.02 - 0.1119 * t + 1.409e-3 * t2) * pitkR + 0.5 * (-5.13e-3 + 7.94e-5 * t) * p2itkR);
  q.ks = m_exp(-4276.1 * invtk + 141.328 - 23.093 * dlogtk + (-13856 * invtk + 324.57 - 47.986 * dlogtk) * sqrtis +
             (35474 * invtk - 771.54 + 114.723 * dlogtk) * is - 2698 * invtk * m_pow(is, 1.5) + 1776 * invtk * is2 +
             m_log(1.0 - 0.001005 * s)) *
         m_exp((18.03 - .0466 * t - 3.16e-4 * t2) * pitkR + 0.5 * (-4.53e-3 + 9.0e-5 * t) * p2itkR);
  q.kf = m_exp(1590.2 * invtk - 12.641 + 1.525 * sqrtis + m_log(1.0 - 0.001005 * s)) *
         m_exp((9.78 + 9.0e-3 * t + 9.42e-4 * t2) * pitkR + 0.5 * (-3.91e-3 + 5.4e-5 * t) * p2itkR);
  q.kb = m_exp((-8966.90 - 2890.53 * sqrts - 77.942 * s + 1.728 * s15 - 0.0996 * s2) * invtk +
             (148.0248 + 137.1942 * sqrts + 1.62142 * s) + (-24.4344 - 25.085 * sqrts - 0.2474 * s) * dlogtk + 0.053105 * sqrts * tk +
             m_log(QDIV((1 + (QDIV(q.st, q.ks)) + (QDIV(q.ft, q.kf))), (1 + (QDIV(q.st, q.ks)))))) *
         m_exp((29.48 - 0.1622 * t - 2.608e-3 * t2) * pitkR + 0.5 * (-2.84e-3) * p2itkR);

  // [H+] on the seawater scale in [1e-10, 1e-6], xacc = 1e-10 (:343-346)
  double x1 = m_pow(10.0, -6.), x2 = m_pow(10.0, -10.);
  double hSWS = drtsafe(q, x1, x2, 1.e-10);
  double hSWS2 = hSWS * hSWS;
  double co2star = QDIV(q.dic * hSWS2, (hSWS2 + q.k1 * hSWS + q.k1 * q.k2));
  double CO3 = QDIV(q.k1 * q.k2 * co2star, hSWS2);
  // calcite solubility product with pressure dependence (:360-388)
  double Kspc = m_exp(-395.8293 + (QDIV(6537.773, tk)) + 71.595 * m_log(tk) - 0.17959 * tk +
                    (-1.78938 + (QDIV(410.64, tk)) + 0.0065453 * tk) * sqrt(s) - 0.17755 * s + 0.0094979 * s15);
  double DVc = -65.28 + 0.397 * t - 0.005155 * (t * t) + (19.816 - 0.0441 * t - 0.00017 * (t * t)) * sqrt(QDIV(s, 35.));
  double DK = 0.01847 + 0.0001956 * t - 0.000002212 * (t * t) + (-0.03217 - 0.0000711 * t + 0.000002212) * sqrt(QDIV(s, 35.));
  Kspc = Kspc * m_exp(-DVc * pitkR + 0.5 * DK * p2itkR);
  const double Ca = 10.28E-3;
  omega_c = QDIV(Ca * CO3, Kspc);
  co2star_out = QDIV(co2star, permil);
  if constexpr (AIR) {
    // solubility ff of Weiss & Price (1980) (:150-153)
    const double tk100a = tk / 100.0, tk1002a = tk100a * tk100a;
    const double ff = m_exp(-162.8301 + 218.2968 / tk100a + 90.9241 * m_log(tk100a) - 1.47696 * tk1002a +
                            s * (.025695 - .025225 * tk100a + 0.0049867 * tk1002a));
    const double co2starair = (co2_in * 1.e-6) * ff * atmpres;
    *dco2star_out = (co2starair - co2star) / permil;
  }
}

// ------------------------------------------------------------------------------------
// Pre-pass.  Everything mobi_driver / mobi_src evaluate per level that does NOT depend
// on the sinking chain (the export of the level above) is a function of the tau-1 inputs
// alone and is computed here at full parallelism, leaving only the nbio Euler sub-steps
// and the export chain to the column kernels.
// ------------------------------------------------------------------------------------
enum MobiPre {
  PR_AC13B = 0, PR_DISSK1, PR_CAPR, PR_BCT, PR_BCTZ, PR_NUD, PR_AOU8, PR_O2FLAG, PR_AVEJ, PR_AVEJ_D, PR_AVEJ_DIAT,
  PR_FO2, PR_P099, PR_LNO3A, PR_LNO3B, PR_GL, PR_N
};
static_assert(PR_N == MOBI_NPRE, "mobi_pre field count");

// k_mobi_light: one thread per water column.  Day fraction and incoming solar
// (09/mom/tracer.F:370-390), then the light at the top of every level: the attenuation by the
// plankton and calcite of the levels above uses their tau-1 inputs only (09/mom/mobi.F:795-812).
__global__ void __launch_bounds__(128) k_mobi_light(const DevView v, double declin) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int ni = v.imt - 2, nrow = v.jhi - v.jlo + 1;
  if (idx >= (long long)ni * nrow) return;
  int i = (int)(idx % ni) + 2;
  int j = (int)(idx / ni) + v.jlo;
  const int kmx = v.kmt[X2(i, j)];
  if (kmx <= 0) return;
  const MobiPar *__restrict__ P = v.mobi_par;
  const int *__restrict__ ix = v.mobi_idx;
  const double pi = 3.14159265358979323846;  // atan(1.0)*4.0 in FP64
  const double radian = QDIV(360., (2. * pi));
  double lat = v.tlat[X2(i, j)];
  double rctheta = fmax(-1.5, fmin(1.5, QDIV(lat, radian) - declin));
  double cr = cos(rctheta);
  rctheta = QDIV(P->kw, sqrt(1. - QDIV((1. - cr * cr), (1.33 * 1.33))));
  double dayfrac = fmin(1., -tan(QDIV(lat, radian)) * tan(declin));
  dayfrac = fmax(1e-12, QDIV(acos(fmax(-1., dayfrac)), pi));
  double swr = P->tap * v.dnswr[X2(i, j)] * 1e-3 * (1. + v.aice[X2(i, j)] * (m_exp(-P->ki * (v.hice[X2(i, j)] + v.hsno[X2(i, j)])) - 1.));
  v.mobi_day[X2(i, j)] = dayfrac;
  const long long n3 = v.n3;
  const double *__restrict__ phyt = v.t_m1 + (long long)(ix[IX_TR + V_PHYT] - 1) * n3;
  const double *__restrict__ diaz = v.t_m1 + (long long)(ix[IX_TR + V_DIAZ] - 1) * n3;
  const double *__restrict__ diat = v.t_m1 + (long long)(ix[IX_TR + V_DIAT] - 1) * n3;
  const double *__restrict__ caco3 = v.t_m1 + (long long)(ix[IX_TR + V_CACO3] - 1) * n3;
  double phin = 0.0, caco3in = 0.0;
  // the tracer loads of 8 levels are issued together, ahead of the exp chain (one column per thread: latency bound)
  const long long c1 = X3(i, 1, j);
  const int sk = v.imt;
  double *__restrict__ gl_out = v.mobi_pre + (long long)PR_GL * n3;
  for (int k0 = 1; k0 <= kmx; k0 += 8) {
    double p_[8], z_[8], d_[8], c_[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const long long c = c1 + (long long)(min(k0 + q, kmx) - 1) * sk;
      p_[q] = phyt[c]; z_[q] = diaz[c]; d_[q] = diat[c]; c_[q] = caco3[c];
    }
    asm volatile("" ::: "memory");
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int k = k0 + q;
      if (k <= kmx) {
        const double dztk = v.dzt[k - 1];
        swr = swr * m_exp(-P->kc * phin - P->kc_c * caco3in);
        phin = fmax(p_[q], TRCMIN) * dztk + fmax(z_[q], TRCMIN) * dztk + fmax(d_[q], TRCMIN) * dztk;
        caco3in = caco3in + c_[q] * dztk;
        gl_out[c1 + (long long)(k - 1) * sk] = swr * m_exp(P->ztt[k - 1] * rctheta);
      }
    }
  }
}

// Evans & Parslow daily-mean light-limited growth (09/mom/mobi.F:1984-2003), one species
__device__ __forceinline__ double evans_parslow(double gl_x, double gd, double f1, double kirr, double dzt) {
  double u1 = fmax(QDIV(gl_x, gd), 1.e-6), u2 = u1 * f1;
  double phi1 = m_log(u1 + sqrt(1. + u1 * u1)) - QDIV((sqrt(1. + u1 * u1) - 1.), u1);
  double phi2 = m_log(u2 + sqrt(1. + u2 * u2)) - QDIV((sqrt(1. + u2 * u2) - 1.), u2);
  return QDIV(gd * (phi1 - phi2), (-kirr * dzt));
}

// k_mobi_cell: one thread per ocean cell.  co2calc_SWS (09/common/co2calc.F, called from
// mobi_driver :768-772), O2 saturation -> AOU (09/mom/tracer.F:457-476), the temperature and
// oxygen dependent rate factors (09/mom/mobi.F:775-835), the pre-loop light harvesting and
// Evans-Parslow integrals of mobi_src (:1928-2003), and the oxygen / nitrate switches of
// the denitrification terms (:1035-1046, 1301-1322).
// Two register allocations of the same body.  On a large grid (many waves of CTAs) twelve resident CTAs per SM at 40
// registers, with ~380 bytes of spills that stay in L1, beat four at 128 registers: the kernel is FP64-issue bound and
// short of warps to cover the pipe latency (measured on B200, 0.5 degree x 40 levels: 160 registers 2.72 ms, 128: 2.49,
// 96: 2.48, 80: 2.44, 64: 2.35, 48: 2.32, 40: 2.29, 32: 2.30).  On a small grid, where every CTA is resident at once
// anyway, the spills only lengthen each thread: 100x100x19 takes 83 us at 128 registers and 101 us at 40.
#ifndef MOBI_CELL_MINB
#define MOBI_CELL_MINB 12
#endif
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_mobi_cell(const DevView v) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int ni = v.imt - 2, nrow = v.jhi - v.jlo + 1;
  if (idx >= (long long)ni * v.km * nrow) return;
  int i = (int)(idx % ni) + 2;
  long long r = idx / ni;
  int k = (int)(r % v.km) + 1;
  int j = (int)(r / v.km) + v.jlo;
  if (k > v.kmt[X2(i, j)]) return;
  const MobiPar *__restrict__ P = v.mobi_par;
  const int *ix = v.mobi_idx;
  const long long n3 = v.n3;
  const long long c = X3(i, k, j);
#define TIN(slot) v.t_m1[c + (long long)(ix[slot] - 1) * n3]
  const double t_in = TIN(IX_ITEMP);
  const double s_in = 1.e3 * TIN(IX_ISALT) + 35.0;
  const double dic_in = TIN(IX_TR + V_DIC);
  const double alk_in = TIN(IX_IALK);
  const double o2_in = TIN(IX_IO2) * 1000.;
  double *__restrict__ pre = v.mobi_pre + c;
  double co2star, Omega_c;
  co2calc_sws<false>(t_in, s_in, dic_in, alk_in, QDIV(v.zt[k - 1], 100.), co2star, Omega_c);
  {
    double ac13_DIC_aq = -1.0512994e-4 * t_in + 1.011765;
    double ac13_aq_POC = -0.017 * m_log10(fmin(fmax(co2star * 1000., 2.), 74.)) + 1.0034;
    pre[(long long)PR_AC13B * n3] = QDIV(ac13_aq_POC, ac13_DIC_aq);
    pre[(long long)PR_DISSK1 * n3] = P->dissk0 * fmax(0., (1. - Omega_c));
    pre[(long long)PR_CAPR * n3] = P->caprmax * fmax(0., (Omega_c - 1.) / (P->kcapr + Omega_c - 1.));   // singular at Omega_c = 1 - kcapr: the full IEEE divide
  }
  // oxygen saturation -> AOU for the ligand parameterisation
  double aou_in;
  {
    double f1 = m_log(QDIV((298.15 - t_in), (273.15 + t_in)));
    double f2 = f1 * f1, f3 = f2 * f1, f4 = f3 * f1, f5 = f4 * f1;
    double o2sat = m_exp(2.00907 + 3.22014 * f1 + 4.05010 * f2 + 4.94457 * f3 - 2.56847E-1 * f4 + 3.88767 * f5 +
                       s_in * (-6.24523e-3 - 7.37614e-3 * f1 - 1.03410e-2 * f2 - 8.17083E-3 * f3) - 4.88682E-7 * s_in * s_in);
    o2sat = QDIV(o2sat, 22391.6) * 1000.0 * 1000.;
    aou_in = o2sat - o2_in;
  }
  const double bct = m_pow(P->bbio, (P->cbio * t_in));
  const double fo2 = m_tanh(0.22 * fmax(o2_in, 0.));
  pre[(long long)PR_BCT * n3] = bct;
  pre[(long long)PR_BCTZ * n3] = (0.5 * (m_tanh(o2_in - 8.) + 1)) * bct;   // same bbio**(cbio*t) value (:830-835)
  pre[(long long)PR_NUD * n3] = P->nud0 * (0.6 + 0.4 * fo2);
  pre[(long long)PR_FO2 * n3] = fo2;
  pre[(long long)PR_AOU8 * n3] = m_pow(fmax(aou_in, 40.), 0.8);
  pre[(long long)PR_O2FLAG * n3] = m_tanh(fmax(o2_in, 0.));
  {
    const double tno3 = fmax(TIN(IX_TR + V_NO3), TRCMIN);   // tnpzd(k,ino3) after the clip of mobi_src
    pre[(long long)PR_P099 * n3] = m_pow(0.99, (fmax(o2_in, TRCMIN) - fmax(tno3, TRCMIN)));
    pre[(long long)PR_LNO3A * n3] = 0.5 * m_tanh(tno3 * 10 - 5.0);
    pre[(long long)PR_LNO3B * n3] = 0.5 * m_tanh(tno3 - 2.5);
  }
  // light harvesting and daily growth integrals on the clipped inputs
  {
    const double bphyt = fmax(TIN(IX_TR + V_PHYT), TRCMIN), bdiaz = fmax(TIN(IX_TR + V_DIAZ), TRCMIN);
    const double bdiat = fmax(TIN(IX_TR + V_DIAT), TRCMIN), bcaco3 = fmax(TIN(IX_TR + V_CACO3), TRCMIN);
    const double bdfe = fmax(TIN(IX_TR + V_DFE), TRCMIN);
    const double gl = pre[(long long)PR_GL * n3], dayfrac = v.mobi_day[X2(i, j)], dzt = v.dzt[k - 1];
    double p1 = fmin(bphyt, P->pmax), p2 = fmax(0.0, bphyt - P->pmax);
    double kfevar = QDIV((P->kfemin * p1 + P->kfemax * p2), (p1 + p2));
    double deffe = QDIV(bdfe, (kfevar + bdfe));
    double thetamax = P->thetamaxlo + (P->thetamaxhi - P->thetamaxlo) * deffe;
    double alpha_O = P->alphamin + (P->alphamax - P->alphamin) * deffe;
    double gl_O = gl * thetamax * alpha_O;
    p1 = fmin(bdiat, P->pmax_Diat);
    p2 = fmax(0.0, bdiat - P->pmax_Diat);
    double kfevar_Diat = QDIV((P->kfemin_Diat * p1 + P->kfemax_Diat * p2), (p1 + p2));
    double deffe_Diat = QDIV(bdfe, (kfevar_Diat + bdfe));
    double gl_Diat = gl * (P->thetamaxlo + (P->thetamaxhi - P->thetamaxlo) * deffe_Diat) * (P->alphamin + (P->alphamax - P->alphamin) * deffe_Diat);
    double deffe_D = QDIV(bdfe, (P->kfe_D + bdfe));
    double gl_D = gl * (P->thetamaxlo + (P->thetamaxhi - P->thetamaxlo) * deffe_D) * (P->alphamin + (P->alphamax - P->alphamin) * deffe_D);
    double kirr = -P->kw - P->kc * (bphyt + bdiaz + bdiat) - P->kc_c * bcaco3;
    double f1 = m_exp(kirr * dzt);
    double jmax = P->abio_P * bct * deffe;
    pre[(long long)PR_AVEJ * n3] = evans_parslow(gl_O, jmax * dayfrac, f1, kirr, dzt);
    double jmax_D = fmax(0., P->abio_P * (bct - P->dbct_D) * deffe_D) * P->jdiar;
    pre[(long long)PR_AVEJ_D * n3] = evans_parslow(gl_D, fmax(1.e-14, jmax_D * dayfrac), f1, kirr, dzt);
    double jmax_Diat = P->abiodiat * bct * deffe_Diat;
    pre[(long long)PR_AVEJ_DIAT * n3] = evans_parslow(gl_Diat, jmax_Diat * dayfrac, f1, kirr, dzt);
  }
#undef TIN
}

// ------------------------------------------------------------------------------------
// ecosystem ODE, 09/mom/mobi.F:1485-3313.  b[] holds the 32 state variables (clipped in
// place like bioin, :1894); on return b[] holds the increments bioout.
// ------------------------------------------------------------------------------------
struct SrcIO {
  // in
  double bct, impo, impo_phos, wwd, nud, impocaco3, wwc, dissk1, impoopl, wwo, opl_disk1, nudop, nudon, bctz;
  double rn15impo, rc13impo, ac13b, rcaco3c13impo, impofe, capr, avej, avej_D, avej_Diat, o2flag, aou8;
  // out
  double nfix, expo, expo_phos, calpro, dissl, expocaco3, expoopl, rn15expo, rc13expo, rcaco3c13expo, expofe;
};

#define CL15(x) fmax(fmin((x), 2. * RN15STD / (1 + RN15STD)), RN15STD / (1 + RN15STD) / 2.)
#define CL13(x) fmax(fmin((x), 2. * RC13STD / (1 + RC13STD)), 0.5 * RC13STD / (1 + RC13STD))
__device__ __forceinline__ void mobi_src(const MobiPar *__restrict__ P, int nbio, double dtbio, double (&b)[MOBI_NVAR], double (&clip)[MOBI_NVAR], SrcIO &io) {
  const double gamma1 = P->gamma1, redptn = P->redptn, redctn = P->redctn, redntp = P->redntp, diazntp = P->diazntp;
  const double diazptn = P->diazptn, dfr = P->dfr, pfr = P->pfr, dfrt = P->dfrt, geZ = P->geZ, rfeton = P->rfeton;
  const double bct = io.bct;
  // ratios from the raw inputs (:1781-1784)
  double ptn_P = QDIV(b[V_PHYT_PHOS], b[V_PHYT]);
  double ptn_detr = QDIV(b[V_DETR_PHOS], b[V_DETR]);
  // flags from the raw inputs (:1814-1890), kept as a bit mask: bit m set <=> flag of state m is 1
  unsigned flm = 0u;
#pragma unroll
  for (int m = 0; m < MOBI_NVAR; m++)
    if (b[m] - TRCMIN >= 0.0) flm |= (1u << m);
#define fl(m) (((flm >> (m)) & 1u) ? 1.0 : 0.0)
  const double sf_P_phosflag = 0.5 + fsign(0.5, ptn_P - gamma1 * redptn);
  const double sf_detr_phosflag = 0.5 + fsign(0.5, ptn_detr - gamma1 * redptn);
  // limit tracers to positive values; the clipped inputs are what the caller sees afterwards (:1893-1926)
#pragma unroll
  for (int m = 0; m < MOBI_NVAR; m++) {
    b[m] = fmax(b[m], TRCMIN);
    clip[m] = b[m];
  }
  // the pre-loop light harvesting and Evans-Parslow integrals (:1928-2003) come from k_mobi_cell
  const double avej = io.avej, avej_D = io.avej_D, avej_Diat = io.avej_Diat;
  double p1, p2, kfevar, deffe, kfevar_Diat, deffe_Diat, deffe_D;
  const double gmax = P->gbio * io.bctz;
  const double nupt = P->nupt0 * bct, nupt_D = P->nupt0_D * bct, nudt = P->nudt0 * bct;
  double nfixout = 0.0, expoout = 0.0, expo_phosout = 0.0, rn15expoout = 0.0, rc13expoout = 0.0, rcaco3c13expoout = 0.0;
  double calproout = 0.0, disslout = 0.0, expocaco3out = 0.0, expooplout = 0.0, expofeout = 0.0;
  const double o2flag = io.o2flag;  // tanh(max(o2,0)) (:2267) and max(aou,40)**0.8 (:2263): loop invariant, from k_mobi_cell
  const double aou8 = io.aou8;

  for (int n = 1; n <= nbio; n++) {
    const double biopo4 = b[V_PO4], biophyt = b[V_PHYT], biophyt_phos = b[V_PHYT_PHOS], biozoop = b[V_ZOOP], biodetr = b[V_DETR];
    const double biodetr_phos = b[V_DETR_PHOS], biodic = b[V_DIC], biodop = b[V_DOP], biono3 = b[V_NO3], biodon = b[V_DON];
    const double biodiaz = b[V_DIAZ], biocaco3 = b[V_CACO3], biodiat = b[V_DIAT], biosil = b[V_SIL], bioopl = b[V_OPL];
    const double biodfe = b[V_DFE], biodetrfe = b[V_DETRFE];
    // half-saturation constants and maximum rates (:2150-2166)
    p1 = fmin(biophyt, P->pmax);
    p2 = fmax(0.0, biophyt - P->pmax);
    double k1n = QDIV((P->knmin * p1 + P->knmax * p2), (p1 + p2));
    double k1p_P = k1n * ptn_P;
    kfevar = QDIV((P->kfemin * p1 + P->kfemax * p2), (p1 + p2));
    deffe = QDIV(biodfe, (kfevar + biodfe));
    double jmax = P->abio_P * bct * deffe;
    p1 = fmin(biodiat, P->pmax_Diat);
    p2 = fmax(0.0, biodiat - P->pmax_Diat);
    kfevar_Diat = QDIV((P->kfemin_Diat * p1 + P->kfemax_Diat * p2), (p1 + p2));
    double k1n_Diat = QDIV((P->knmin_Diat * p1 + P->knmax_Diat * p2), (p1 + p2));
    double k1p_Diat = k1n_Diat * redptn;
    deffe_Diat = QDIV(biodfe, (kfevar_Diat + biodfe));
    double jmax_Diat = P->abiodiat * bct * deffe_Diat;
    deffe_D = QDIV(biodfe, (P->kfe_D + biodfe));
    double jmax_D = fmax(0., P->abio_P * (bct - P->dbct_D) * deffe_D) * P->jdiar;
    // growth rates (:2168-2206)
    double limP_dop = QDIV(P->hdop * biodop, (k1p_P + biodop));
    double limP_po4 = QDIV(biopo4, (k1p_P + biopo4));
    double dopupt_flag = 0.5 + fsign(0.5, limP_dop - limP_po4);
    double limP = limP_dop * dopupt_flag + limP_po4 * (1. - dopupt_flag);
    double u_P = fmin(avej, jmax * limP);
    double limSi = QDIV(biosil, (5.e-3 + biosil));  // k1si = 5.e-3 (:2181)
    limP_dop = QDIV(P->hdop * biodop, (k1p_Diat + biodop));
    limP_po4 = QDIV(biopo4, (k1p_Diat + biopo4));
    double dopupt_Diat_flag = 0.5 + fsign(0.5, limP_dop - limP_po4);
    double limP_Diat = limP_dop * dopupt_Diat_flag + limP_po4 * (1. - dopupt_Diat_flag);
    double u_Diat = fmin(avej_Diat, jmax_Diat * limSi);
    u_Diat = fmin(u_Diat, jmax_Diat * limP_Diat);
    u_P = fmin(u_P, QDIV(jmax * biono3, (k1n + biono3)));
    u_Diat = fmin(u_Diat, QDIV(jmax_Diat * biono3, (k1n_Diat + biono3)));
    double u_D = fmin(avej_D, jmax_D * limP);
    double dopupt_D_flag = dopupt_flag;
    // grazing (:2208-2245)
    double thetaZ = P->zprefP * biophyt + P->zprefDet * biodetr + P->zprefZ * biozoop + P->zprefDiaz * biodiaz + P->kzoo +
                    P->zprefDiat * biodiat;
    double ing_P = QDIV(P->zprefP, thetaZ), ing_Det = QDIV(P->zprefDet, thetaZ), ing_Z = QDIV(P->zprefZ, thetaZ);
    double ing_D = QDIV(P->zprefDiaz, thetaZ), ing_Diat = QDIV(P->zprefDiat, thetaZ);
    double npp = u_P * biophyt;
    double npp_Diat = u_Diat * biodiat;
    double dopupt = npp * dopupt_flag;
    double dopupt_Diat = npp_Diat * dopupt_Diat_flag;
    double npp_D = fmax(0., u_D * biodiaz);
    double graz_D = gmax * ing_D * biodiaz * biozoop;
    double morpt_D = nupt_D * biodiaz;
    double morp_D = P->nup_D * biodiaz * biodiaz;
    double no3upt_D = (0.5 + 0.5 * tanh(biono3 - 5.)) * npp_D;
    double dopupt_D = npp_D * dopupt_D_flag;
    double graz = gmax * ing_P * biophyt * biozoop;
    double graz_Z = gmax * ing_Z * biozoop * biozoop;
    double graz_Det = gmax * ing_Det * biodetr * biozoop;
    double morp = P->nup * biophyt;
    double morpt = nupt * biophyt;
    double recy_don = io.nudon * bct * biodon;
    double recy_dop = io.nudop * bct * biodop;
    double morz = P->nuz * biozoop * biozoop;
    double remi = io.nud * bct * biodetr;
    double expo = io.wwd * biodetr;
    double expo_phos = io.wwd * biodetr_phos;
    double dissl = biocaco3 * io.dissk1;
    double expocaco3 = io.wwc * biocaco3;
    double graz_Diat = gmax * ing_Diat * biodiat * biozoop;
    double morp_Diat = P->nu_diat * biodiat;
    double morpt_Diat = nudt * biodiat;
    double opldis = bioopl * io.opl_disk1;
    double expoopl = io.wwo * bioopl;
    double remife = io.nud * bct * biodetrfe;
    // iron scavenging (:2262-2283)
    double ligand = QDIV(fmax(QDIV(aou8, 66.) + QDIV(pow(biodon, 0.8), 4.8), 0.5), 1000.);
    double fepa = (1.0 + P->kfeleq * (ligand - biodfe)) * o2flag;
    double feprime = (QDIV((-fepa + sqrt(fepa * fepa + 4.0 * P->kfeleq * biodfe)), (2.0 * P->kfeleq))) * o2flag;
    double feorgads = (P->kfeorg * (pow(((biodetr * fl(V_DETR)) * P->mc * redctn), 0.58)) * feprime) * o2flag;
    double fecol = P->kfecol * (feprime * feprime) * o2flag;
    double expofe = io.wwd * biodetrfe;
    // flags switch outgoing fluxes off when a pool is exhausted (:2284-2334)
    graz = graz * fl(V_PHYT) * fl(V_PHYT_PHOS) * sf_P_phosflag * fl(V_PHYTN15);
    graz_Z = graz_Z * fl(V_ZOOP) * fl(V_ZOOPN15);
    graz_Det = graz_Det * fl(V_DETR) * fl(V_DETR_PHOS) * sf_detr_phosflag * fl(V_DETRN15);
    morp = morp * fl(V_PHYT) * fl(V_PHYT_PHOS) * fl(V_PHYTN15);
    morpt = morpt * fl(V_PHYT) * fl(V_PHYT_PHOS) * fl(V_PHYTN15);
    morz = morz * fl(V_ZOOP) * fl(V_ZOOPN15);
    remi = remi * fl(V_DETR) * fl(V_DETR_PHOS) * fl(V_DETRN15);
    expo = expo * fl(V_DETR) * fl(V_DETRN15);
    expo_phos = expo_phos * fl(V_DETR_PHOS);
    recy_dop = recy_dop * fl(V_DOP);
    npp = npp * fl(V_NO3) * (dopupt_flag * fl(V_DOP) + (1. - dopupt_flag) * fl(V_PO4)) * fl(V_DIN15);
    npp_Diat = npp_Diat * fl(V_NO3) * (dopupt_Diat_flag * fl(V_DOP) + (1. - dopupt_Diat_flag) * fl(V_PO4)) * fl(V_DIN15);
    npp_D = npp_D * (dopupt_D_flag * fl(V_DOP) + (1. - dopupt_D_flag) * fl(V_PO4)) * fl(V_DIN15);
    graz_D = graz_D * fl(V_DIAZ) * fl(V_DIAZN15);
    morpt_D = morpt_D * fl(V_DIAZ) * fl(V_DIAZN15);
    morp_D = morp_D * fl(V_DIAZ) * fl(V_DIAZN15);
    no3upt_D = no3upt_D * fl(V_NO3) * fl(V_DIN15);
    recy_don = recy_don * fl(V_DON) * fl(V_DON15);
    dissl = dissl * fl(V_CACO3);
    expocaco3 = expocaco3 * fl(V_CACO3);
    graz_Diat = graz_Diat * fl(V_DIAT);
    morp_Diat = morp_Diat * fl(V_DIAT);
    morpt_Diat = morpt_Diat * fl(V_DIAT);
    remife = remife * fl(V_DETRFE);
    feorgads = feorgads * fl(V_DFE);
    expofe = expofe * fl(V_DETRFE);
    fecol = fecol * fl(V_DFE);
    // digestion, excretion, sloppy feeding (:2335-2440)
    double dig_P = gamma1 * graz, dig_Z = gamma1 * graz_Z, dig_Det = gamma1 * graz_Det, dig_Diat = gamma1 * graz_Diat;
    double dig = dig_Z + dig_P + dig_Det + dig_Diat;
    double excr = gamma1 * (1 - geZ) * graz_Z + gamma1 * (1 - geZ) * graz + gamma1 * (1 - geZ) * graz_Det + gamma1 * (1 - geZ) * graz_Diat;
    double sf_P = (1. - gamma1) * graz, sf_Z = (1. - gamma1) * graz_Z, sf_Det = (1. - gamma1) * graz_Det;
    double sf_Diat = (1. - gamma1) * graz_Diat;
    double sf = sf_P + sf_Z + sf_Det + sf_Diat;
    double sf_P_phos = (graz * ptn_P - dig_P * redptn);
    double sf_Det_phos = (graz_Det * ptn_detr - dig_Det * redptn);
    const double nr_excr_P = 0.0, nr_excr_detr = 0.0;
    double sf_phos = sf_P_phos + sf_Z * redptn + sf_Det_phos + sf_Diat * redptn;
    double dig_D = gamma1 * graz_D * (QDIV(redntp, diazntp));
    dig = dig + dig_D;
    excr = excr + gamma1 * (1 - geZ) * graz_D * (QDIV(redntp, diazntp));
    double nr_excr_D = gamma1 * graz_D * (1 - (QDIV(redntp, diazntp))) + (1 - gamma1) * graz_D * (1 - (QDIV(redntp, diazntp)));
    double sf_D = (1 - gamma1) * graz_D * (QDIV(redntp, diazntp));
    sf = sf + sf_D;
    sf_phos = sf_phos + sf_D * redptn;
    // isotope fractionation factors (:2441-2530)
    double uno3 = fmax(fmin(QDIV(npp * dtbio, biono3), 0.999), TRCMIN);
    double rno3 = fmax(fmin(QDIV(b[V_DIN15], (biono3 - b[V_DIN15])), 2 * RN15STD), RN15STD / 2.);
    double bassim = rno3 + QDIV(QDIV(P->eps_assim * (1 - uno3), uno3) * log(1 - uno3) * rno3, 1000.);
    double fcassim = QDIV(bassim, (1 + bassim));
    double udon = fmax(fmin(QDIV(recy_don * dtbio, biodon), 0.999), TRCMIN);
    double rdon = fmax(fmin(QDIV(b[V_DON15], (biodon - b[V_DON15])), 2 * RN15STD), RN15STD / 2.);
    double brecy = rdon + QDIV(QDIV(P->eps_recy * (1 - udon), udon) * log(1 - udon) * rdon, 1000.);
    double fcrecy = QDIV(brecy, (1 + brecy));
    double rzoop = fmax(fmin(QDIV(b[V_ZOOPN15], (biozoop - b[V_ZOOPN15])), 2. * RN15STD), RN15STD / 2.);
    double bexcr = rzoop - QDIV(P->eps_excr * rzoop, 1000.);
    double fcexcr = QDIV(bexcr, (1 + bexcr));
    double bnfix = RN15STD - QDIV(P->eps_nfix * RN15STD, 1000.);
    double fcnfix = QDIV(bnfix, (1 + bnfix));
    double rtphytn15 = CL15(QDIV(b[V_PHYTN15], biophyt));
    double rtdiatn15 = CL15(QDIV(b[V_DIATN15], biodiat));
    double rtzoopn15 = CL15(QDIV(b[V_ZOOPN15], biozoop));
    double rtdetrn15 = CL15(QDIV(b[V_DETRN15], biodetr));
    double rtdiazn15 = CL15(QDIV(b[V_DIAZN15], biodiaz));
    double rdic13 = fmax(fmin(QDIV(b[V_DIC13], (biodic - b[V_DIC13])), 2. * RC13STD), 0.5 * RC13STD);
    double bc13npp = io.ac13b * rdic13;
    double fcnpp = QDIV(bc13npp, (1 + bc13npp));
    double rtdic13 = CL13(QDIV(b[V_DIC13], biodic));
    double rtphytc13 = CL13(QDIV(b[V_PHYTC13], (biophyt * redctn)));
    double rtdiatc13 = CL13(QDIV(b[V_DIATC13], (biodiat * redctn)));
    double rtcaco3c13 = CL13(QDIV(b[V_CACO3C13], biocaco3));
    double rtzoopc13 = CL13(QDIV(b[V_ZOOPC13], (biozoop * redctn)));
    double rtdetrc13 = CL13(QDIV(b[V_DETRC13], (biodetr * redctn)));
    double rtdoc13 = CL13(QDIV(b[V_DOC13], (biodon * redctn)));
    double rtdiazc13 = CL13(QDIV(b[V_DIAZC13], (biodiaz * redctn)));
    // CaCO3 and opal production (:2532-2548)
    double calpro = ((sf_Z + morz) * io.capr + (sf_P + morp) * io.capr) * redctn * 1.e3;
    double sipr0 = (-0.46204044117647 * tanh(6.9 * biodfe * 1.e3 + -3.673092) + 1.60266544117647);
    double oplpro = (morp_Diat + sf_Diat) * sipr0 * fl(V_SIL) * (1.e-3);
    opldis = opldis * fl(V_OPL);
    expoopl = expoopl * fl(V_OPL);
    double GM15ptc = 0.0060 + 0.0069 * biopo4;
    double GM15ptn = GM15ptc * redctn * 1.e3;
    const double rnd = QDIV(redntp, diazntp);

    // prognostic equations, forward Euler (:2552-2760)
    b[V_PO4] = biopo4 + dtbio * (dopupt * ptn_P - GM15ptn * npp + (1. - dfrt) * morpt * ptn_P + (1. - pfr) * remi * ptn_detr +
                                 diazptn * (morpt_D - (npp_D - dopupt_D)) + recy_dop +
                                 redptn * (excr + (1. - dfrt) * morpt_Diat - (npp_Diat - dopupt_Diat)));
    b[V_DOP] = biodop + dtbio * (dfr * morp * ptn_P + redptn * (dfr * morp_Diat + dfrt * morpt_Diat - dopupt_Diat) +
                                 dfrt * morpt * ptn_P + pfr * remi * ptn_detr - ptn_P * dopupt - diazptn * dopupt_D - recy_dop);
    b[V_PHYT] = biophyt + dtbio * (npp - morp - graz - morpt);
    b[V_PHYT_PHOS] = biophyt_phos + dtbio * (npp * GM15ptn - morp * ptn_P - graz * ptn_P - morpt * ptn_P);
    b[V_ZOOP] = biozoop + dtbio * (dig - morz - graz_Z - excr);
    b[V_DETR] = biodetr + dtbio * ((1. - dfr) * morp + sf + morz - remi - graz_Det - expo + io.impo + morp_D * rnd + (1. - dfr) * morp_Diat);
    b[V_DETR_PHOS] = biodetr_phos + dtbio * ((1. - dfr) * morp * ptn_P + sf_phos + morz * redptn - remi * ptn_detr - graz_Det * ptn_detr -
                                             expo_phos + io.impo_phos + morp_D * rnd * redptn + (1. - dfr) * morp_Diat * redptn);
    b[V_DIC] = biodic + dtbio * redctn *
                            (excr + (1. - pfr) * remi + (1. - dfrt) * morpt - npp + (1. - dfrt) * morpt_Diat - npp_Diat + morpt_D - npp_D +
                             recy_don + nr_excr_D + nr_excr_P + nr_excr_detr + morp_D * (1. - rnd));
    b[V_NO3] = biono3 + dtbio * (excr + (1. - pfr) * remi + (1. - dfrt) * morpt - npp + (1. - dfrt) * morpt_Diat - npp_Diat + morpt_D -
                                 no3upt_D + recy_don + nr_excr_D + nr_excr_P + nr_excr_detr + morp_D * (1. - rnd));
    b[V_DON] = biodon + dtbio * (dfr * morp + dfrt * morpt + pfr * remi - recy_don + dfr * morp_Diat + dfrt * morpt_Diat);
    b[V_DIAZ] = biodiaz + dtbio * (npp_D - morp_D - morpt_D - graz_D);
    const double ptn_P_old = ptn_P, ptn_detr_old = ptn_detr;
    (void)ptn_P_old; (void)ptn_detr_old;
    ptn_P = QDIV(b[V_PHYT_PHOS], b[V_PHYT]);
    ptn_detr = QDIV(b[V_DETR_PHOS], b[V_DETR]);
    b[V_CACO3] = biocaco3 + dtbio * (calpro - dissl - expocaco3 + io.impocaco3);
    b[V_DIAT] = biodiat + dtbio * (npp_Diat - morp_Diat - graz_Diat - morpt_Diat);
    b[V_SIL] = biosil + dtbio * (opldis - oplpro);
    b[V_OPL] = bioopl + dtbio * (oplpro - opldis - expoopl + io.impoopl);
    b[V_DFE] = biodfe + dtbio * (rfeton * (excr + (1. - dfrt) * morpt - npp + morpt_D - npp_D + recy_don + nr_excr_D + nr_excr_P +
                                           nr_excr_detr + morp_D * (1. - rnd)) -
                                 feorgads + remife - fecol + rfeton * ((1. - dfrt) * morpt_Diat - npp_Diat));
    b[V_DETRFE] = biodetrfe + dtbio * (rfeton * (sf + (1. - dfr) * morp + morp_D * rnd + morz - graz_Det) + feorgads + P->iscr * fecol -
                                       remife - expofe + io.impofe + rfeton * (1. - dfr) * morp_Diat);
    b[V_DIN15] = b[V_DIN15] + dtbio * (rtphytn15 * (1. - dfrt) * morpt + rtphytn15 * nr_excr_P + rtdiatn15 * (1. - dfrt) * morpt_Diat -
                                       fcassim * npp_Diat + fcexcr * excr + rtdiazn15 * morpt_D + rtdiazn15 * nr_excr_D +
                                       rtdiazn15 * morp_D * (1. - rnd) + rtdetrn15 * (1. - pfr) * remi + rtdetrn15 * nr_excr_detr +
                                       fcrecy * recy_don - fcassim * npp - fcassim * no3upt_D);
    b[V_DON15] = b[V_DON15] + dtbio * (dfr * rtphytn15 * morp + dfr * rtdiatn15 * morp_Diat + dfrt * rtdiatn15 * morpt_Diat +
                                       dfrt * rtphytn15 * morpt + rtdetrn15 * pfr * remi - fcrecy * recy_don);
    b[V_PHYTN15] = b[V_PHYTN15] + dtbio * (fcassim * npp - rtphytn15 * morp - rtphytn15 * graz - rtphytn15 * morpt);
    b[V_DIATN15] = b[V_DIATN15] + dtbio * (fcassim * npp_Diat - rtdiatn15 * morp_Diat - rtdiatn15 * graz_Diat - rtdiatn15 * morpt_Diat);
    b[V_ZOOPN15] = b[V_ZOOPN15] + dtbio * (rtphytn15 * dig_P + rtdiatn15 * dig_Diat + rtzoopn15 * dig_Z + rtdetrn15 * dig_Det +
                                           rtdiazn15 * dig_D - rtzoopn15 * morz - rtzoopn15 * graz_Z - fcexcr * excr);
    b[V_DETRN15] = b[V_DETRN15] + dtbio * (rtphytn15 * (1. - dfr) * morp + rtdiatn15 * (1. - dfr) * morp_Diat + rtdiatn15 * sf_Diat +
                                           rtphytn15 * sf_P + rtzoopn15 * sf_Z + rtdetrn15 * sf_Det + rtdiazn15 * sf_D + rtzoopn15 * morz -
                                           rtdetrn15 * remi - rtdetrn15 * graz_Det - rtdetrn15 * expo + io.rn15impo * io.impo +
                                           rtdiazn15 * morp_D * rnd);
    b[V_DIAZN15] = b[V_DIAZN15] + dtbio * (fcnfix * (npp_D - no3upt_D) + fcassim * no3upt_D - rtdiazn15 * morp_D - rtdiazn15 * graz_D -
                                           rtdiazn15 * morpt_D);
    b[V_DIC13] = b[V_DIC13] + dtbio * redctn *
                                  (rtphytc13 * (1. - dfrt) * morpt + rtphytc13 * nr_excr_P + rtzoopc13 * excr + rtdiazc13 * morpt_D +
                                   rtdiazc13 * nr_excr_D + rtdiazc13 * morp_D * (1 - rnd) + rtdetrc13 * (1. - pfr) * remi +
                                   rtdetrc13 * nr_excr_detr + rtdiatc13 * (1. - dfrt) * morpt_Diat - fcnpp * npp_Diat + rtdoc13 * recy_don -
                                   fcnpp * npp - fcnpp * npp_D);
    b[V_DOC13] = b[V_DOC13] + dtbio * redctn *
                                  (dfr * rtphytc13 * morp + rtdiatc13 * (dfr * morp_Diat + dfrt * morpt_Diat) + rtphytc13 * dfrt * morpt +
                                   rtdetrc13 * pfr * remi - rtdoc13 * recy_don);
    b[V_PHYTC13] = b[V_PHYTC13] + dtbio * redctn * (fcnpp * npp - rtphytc13 * morp - rtphytc13 * graz - rtphytc13 * morpt);
    b[V_ZOOPC13] = b[V_ZOOPC13] + dtbio * redctn *
                                      (rtphytc13 * dig_P + rtdiatc13 * dig_Diat + rtzoopc13 * dig_Z + rtdetrc13 * dig_Det +
                                       rtdiazc13 * dig_D - rtzoopc13 * morz - rtzoopc13 * graz_Z - rtzoopc13 * excr);
    b[V_DETRC13] = b[V_DETRC13] + dtbio * redctn *
                                      (rtphytc13 * (1. - dfr) * morp + rtdiatc13 * (1. - dfr) * morp_Diat + rtdiatc13 * sf_Diat +
                                       rtphytc13 * sf_P + rtzoopc13 * sf_Z + rtdetrc13 * sf_Det + rtdiazc13 * sf_D + rtzoopc13 * morz -
                                       rtdetrc13 * remi - rtdetrc13 * graz_Det - rtdetrc13 * expo + io.rc13impo + rtdiazc13 * morp_D * rnd);
    b[V_DIAZC13] = b[V_DIAZC13] + dtbio * redctn * (fcnpp * npp_D - rtdiazc13 * (morp_D + graz_D + morpt_D));
    b[V_CACO3C13] = b[V_CACO3C13] + dtbio * (rtdic13 * calpro - rtcaco3c13 * dissl - rtcaco3c13 * expocaco3 + io.rcaco3c13impo);
    b[V_DIATC13] = b[V_DIATC13] + dtbio * redctn * (fcnpp * npp_Diat - rtdiatc13 * (morp_Diat + graz_Diat + morpt_Diat));
    // accumulate the outputs (:2762-2777)
    expoout = expoout + expo;
    expo_phosout = expo_phosout + expo_phos;
    rn15expoout = rn15expoout + rtdetrn15;
    rc13expoout = rc13expoout + rtdetrc13 * expo;
    rcaco3c13expoout = rcaco3c13expoout + rtcaco3c13 * expocaco3;
    calproout = calproout + calpro;
    disslout = disslout + dissl;
    expocaco3out = expocaco3out + expocaco3;
    expooplout = expooplout + expoopl;
    nfixout = nfixout + npp_D - no3upt_D;
    expofeout = expofeout + expofe;
    // a flag that is still 1 is re-evaluated on the updated pool; once 0 it stays 0 (:3175-3251)
#pragma unroll
    for (int m = 0; m < MOBI_NVAR; m++)
      if (b[m] - TRCMIN < 0.0) flm &= ~(1u << m);
  }
  // increments relative to the clipped inputs (:3254-3311)
#pragma unroll
  for (int m = 0; m < MOBI_NVAR; m++) b[m] = b[m] - clip[m];
  io.nfix = nfixout; io.expo = expoout; io.expo_phos = expo_phosout; io.calpro = calproout; io.dissl = disslout;
  io.expocaco3 = expocaco3out; io.expoopl = expooplout; io.rn15expo = rn15expoout; io.rc13expo = rc13expoout;
  io.rcaco3c13expo = rcaco3c13expoout; io.expofe = expofeout;
#undef fl
}

#ifdef MOBI_COL_MINB   // experiment builds: resident CTAs (= warps) per SM the register allocation must allow
#define MOBI_COL_BOUNDS __launch_bounds__(32, MOBI_COL_MINB)
#else
#define MOBI_COL_BOUNDS __launch_bounds__(32)
#endif
__global__ void MOBI_COL_BOUNDS k_mobi_column(const DevView v, int mi, int nbio, double dtbio, double rdtts, double rnbio) {
  // ocean columns of the owned rows, sorted by depth so that the 32 columns of a warp run
  // the same number of levels
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= v.mobi_ncols) return;
  const int col = v.mobi_cols[idx];
  const int i = col % v.imt + 1, j = col / v.imt + v.jbase;
  const int kmx = v.kmt[col];
  const MobiPar *__restrict__ P = v.mobi_par;
  const int *__restrict__ ix = v.mobi_idx;
  const double redctn = P->redctn;

  double expo = 0.0, expo_phos = 0.0, rn15expo = 0.0, rc13expo = 0.0, rcaco3c13expo = 0.0;
  double expofe = 0.0, expocaco3 = 0.0, expoopl = 0.0;
  const long long n3 = v.n3;
  double b[MOBI_NVAR], clip[MOBI_NVAR];
  const int s_alk = ix[IX_ISALK], s_o2 = ix[IX_ISO2], s_c14 = ix[IX_ISC14];

  for (int k = 1; k <= kmx; k++) {
    const long long c = X3(i, k, j);
    const double dztk = v.dzt[k - 1], dztrk = v.dztr[k - 1];
    // gather (tracer.F:393-503)
#pragma unroll
    for (int m = 0; m < MOBI_NVAR; m++) b[m] = v.t_m1[c + (long long)(ix[IX_TR + m] - 1) * n3];
    const double o2_in = v.t_m1[c + (long long)(ix[IX_IO2] - 1) * n3] * 1000.;
    const double dic_in = b[V_DIC];
    const double c14_in = v.t_m1[c + (long long)(ix[IX_IC14] - 1) * n3];
    const double sgb = v.sg_bathy[XIJK(i, j, k)];
    const double *__restrict__ pre = v.mobi_pre + c;
    // ---- mobi_driver, level k of loop 1 (mobi.F:763-1289) ----
    SrcIO io;
    io.rn15impo = rn15expo;
    io.ac13b = pre[(long long)PR_AC13B * n3];
    io.rc13impo = rc13expo * dztrk;
    io.rcaco3c13impo = rcaco3c13expo * dztrk;
    io.dissk1 = pre[(long long)PR_DISSK1 * n3];
    io.capr = pre[(long long)PR_CAPR * n3];
    io.opl_disk1 = P->opl_disk0;
    io.impocaco3 = expocaco3 * dztrk;
    io.impo = expo * dztrk;
    io.impo_phos = expo_phos * dztrk;
    io.impofe = expofe * dztrk;
    io.bct = pre[(long long)PR_BCT * n3];
    io.impoopl = expoopl * dztrk;
    io.bctz = pre[(long long)PR_BCTZ * n3];
    io.nud = pre[(long long)PR_NUD * n3];
    io.nudon = P->nudon0;
    io.nudop = P->nudop0;
    io.wwd = P->wd[k - 1];
    io.wwc = P->wc[k - 1];
    io.wwo = P->wo[k - 1];
    io.avej = pre[(long long)PR_AVEJ * n3];
    io.avej_D = pre[(long long)PR_AVEJ_D * n3];
    io.avej_Diat = pre[(long long)PR_AVEJ_DIAT * n3];
    io.o2flag = pre[(long long)PR_O2FLAG * n3];
    io.aou8 = pre[(long long)PR_AOU8 * n3];
    mobi_src(P, nbio, dtbio, b, clip, io);   // b: increments; clip: the clipped tnpzd(k,:)
    // rates (mobi.F:880-895)
#pragma unroll
    for (int m = 0; m < MOBI_NVAR; m++) b[m] = b[m] * rdtts;
    expofe = io.expofe * rnbio;
    expocaco3 = io.expocaco3 * rnbio;
    expoopl = io.expoopl * rnbio;
    const double rexpoopl = expoopl;
    expo = io.expo * rnbio;
    expo_phos = io.expo_phos * rnbio;
    rn15expo = io.rn15expo * rnbio;
    rc13expo = io.rc13expo * rnbio;
    rcaco3c13expo = io.rcaco3c13expo * rnbio;
    const double rcalpro = io.calpro * rnbio;
    const double rdissl = io.dissl * rnbio;
    const double rexpocaco3 = expocaco3;
    const double nfix = io.nfix;
    // benthic denitrification, Bohlen et al. 2012 (:1035-1075); the flags see the clipped tnpzd
    const double tno3 = clip[V_NO3], tdin15 = clip[V_DIN15];
    double no3flag = 0.5 + fsign(0.5, tno3 - TRCMIN);
    double din15flag = 0.5 + fsign(0.5, tdin15 - TRCMIN);
    double lno3 = pre[(long long)PR_LNO3A * n3];
    double sg_bdeni = (0.06 + 0.19 * pre[(long long)PR_P099 * n3]) * fmax(expo * sgb, TRCMIN) * redctn * 1.e3;
    sg_bdeni = fmin(sg_bdeni, sgb * expo);
    sg_bdeni = fmax(sg_bdeni, 0.);
    sg_bdeni = sg_bdeni * (0.5 + lno3) * no3flag * din15flag;
    const double bdeni = sg_bdeni;
    b[V_NO3] = b[V_NO3] + sgb * expo - sg_bdeni;
    double rno3 = QDIV(fmax(tdin15, TRCMIN * RN15STD / (1 + RN15STD)), fmax(tno3 - tdin15, TRCMIN * RN15STD / (1 + RN15STD)));
    rno3 = fmin(rno3, 2. * RN15STD);
    rno3 = fmax(rno3, RN15STD / 2.);
    double eps_bdeni = v.mobi_epsbd[k - 1];
    double bbdeni = rno3 - QDIV(eps_bdeni * rno3, 1000.);
    b[V_DIN15] = b[V_DIN15] + rn15expo * sgb * expo - QDIV(bbdeni, (1 + bbdeni)) * sg_bdeni;
    // sediment carbon oxidation (Flogel 2011 / Somes 2021) and iron release (Dale 2015) (:1076-1110)
    double coxdepth = fmin(fmax(v.zt[k - 1], 50000.), 150000.);
    double oblinc = -1.26e-6 * coxdepth + 0.203;
    double obexpc = -6.e-7 * coxdepth + 1.14;
    double nburial = QDIV((oblinc * pow((QDIV(expo * sgb * dztk, 100) * 86400. * 365. * redctn * 1000.), obexpc)), (QDIV(86400. * 365. * dztk, 100) * redctn * 1000.));
    double coxsed = expo * sgb - nburial;
    double fesed = QDIV(85. * tanh(QDIV(QDIV(coxsed * redctn * 1000 * dztk, 100) * 86400., o2_in)), (QDIV(dztk, 100) * 86400 * 1000));
    b[V_DFE] = b[V_DFE] + fesed;
    b[V_PO4] = b[V_PO4] + sgb * expo_phos;
    b[V_DIC] = b[V_DIC] + sgb * expo * redctn;
    b[V_DIC13] = b[V_DIC13] + rc13expo * sgb * redctn;
    rc13expo = rc13expo - sgb * rc13expo;
    expo = expo - sgb * expo;
    expo_phos = expo_phos - sgb * expo_phos;
    const double dic_npzd_sms = b[V_DIC];
    // isotope ratios of DIC and CaCO3 for the calcite terms (:1258-1276)
    double rtdic13 = QDIV(fmax(clip[V_DIC13], TRCMIN * RC13STD / (1 + RC13STD)), fmax(dic_in, TRCMIN));
    rtdic13 = fmin(rtdic13, 2. * RC13STD / (1 + RC13STD));
    rtdic13 = fmax(rtdic13, 0.5 * RC13STD / (1 + RC13STD));
    double rtcaco3c13 = QDIV(fmax(clip[V_CACO3C13], TRCMIN * RC13STD / (1 + RC13STD)), fmax(clip[V_CACO3], TRCMIN));
    rtcaco3c13 = fmin(rtcaco3c13, 2. * RC13STD / (1 + RC13STD));
    rtcaco3c13 = fmax(rtcaco3c13, 0.5 * RC13STD / (1 + RC13STD));
    double src_alk = -b[V_DIC] * P->redntc * 1.e-3;
    // total export -> import for the next layer (:1280-1288)
    expo = expo * dztk;
    expo_phos = expo_phos * dztk;
    rc13expo = rc13expo * dztk;
    rcaco3c13expo = rcaco3c13expo * dztk;
    expofe = expofe * dztk;
    expocaco3 = expocaco3 * dztk;
    expoopl = expoopl * dztk;

    // ---- level k of loop 2: O2, water-column denitrification, ALK (:1301-1366) ----
    double fo2 = pre[(long long)PR_FO2 * n3];
    double so2 = dic_npzd_sms * P->redotc + nfix * rnbio * 1.25e-3;
    lno3 = pre[(long long)PR_LNO3B * n3];
    double wcdeni = 800. * no3flag * so2 * (1.0 - fo2) * (0.5 + lno3) * din15flag;
    wcdeni = fmax(wcdeni, 0.);
    b[V_NO3] = b[V_NO3] - wcdeni;
    double uno3 = QDIV(wcdeni * v.c2dtts, tno3);
    uno3 = fmin(uno3, 0.999);
    uno3 = fmax(uno3, TRCMIN);
    double bwcdeni = rno3 + QDIV(QDIV(P->eps_wcdeni * (1 - uno3), uno3) * log(1 - uno3) * rno3, 1000.);
    b[V_DIN15] = b[V_DIN15] - (QDIV(bwcdeni, (1 + bwcdeni))) * wcdeni;
    src_alk = src_alk + wcdeni * 1.e-3;
    src_alk = src_alk + bdeni * 1.e-3;
    src_alk = src_alk - nfix * rnbio * 1.e-3;
    const double src_o2 = -so2 * fo2;

    // ---- level k of loop 3: calcite dissolution / production (:1372-1400), opal leftovers (:1478) ----
    if (k < kmx) {
      b[V_DIC] = b[V_DIC] + rdissl * 1.e-3 - rcalpro * 1.e-3;
      b[V_DIC13] = b[V_DIC13] + rdissl * 1.e-3 * rtcaco3c13 - rcalpro * 1.e-3 * rtdic13;
      src_alk = src_alk + 2. * rdissl * 1.e-3 - 2. * rcalpro * 1.e-3;
    } else {
      b[V_DIC] = b[V_DIC] + rdissl * 1.e-3 - rcalpro * 1.e-3 + rexpocaco3 * 1.e-3;
      b[V_DIC13] = b[V_DIC13] + rdissl * 1.e-3 * rtcaco3c13 - rcalpro * 1.e-3 * rtdic13 + rexpocaco3 * 1.e-3 * rtcaco3c13;
      src_alk = src_alk + 2. * rdissl * 1.e-3 - 2. * rcalpro * 1.e-3 + 2. * rexpocaco3 * 1.e-3;
      b[V_SIL] = b[V_SIL] + rexpoopl;
    }
    // dust (surface) and hydrothermal iron (tracer.F:536-545)
    if (k == 1) b[V_DFE] = b[V_DFE] + QDIV(v.fe_atmdep[X2(i, j) + (long long)(mi - 1) * v.n2] * 1000, (QDIV(v.dzt[0], 100.)));
    b[V_DFE] = b[V_DFE] + v.fe_hydr[XIJK(i, j, k)];

    // scatter into src (mobi.F:1149-1204)
#pragma unroll
    for (int m = 0; m < MOBI_NVAR; m++) v.src[c + (long long)(ix[IX_SRC + m] - 1) * n3] = b[m];
    v.src[c + (long long)(s_alk - 1) * n3] = src_alk;
    v.src[c + (long long)(s_o2 - 1) * n3] = src_o2;
    // c14 source (tracer.F:848-867)
    v.src[c + (long long)(s_c14 - 1) * n3] = b[V_DIC] * RC14STD - 3.836e-12 * c14_in;
  }
}


// ------------------------------------------------------------------------------------
// k_mobi_ws: the same column physics, warp specialised.  With few water columns (the
// 100x100 grid has ~7000) one thread per column leaves the SMs almost empty and the
// kernel time is the latency of one column's serial instruction chain.  Here a CTA of
// 8 warps owns 32 columns (lane = column) and every warp executes a different slice of
// the Euler sub-step for those columns -- no divergence inside a warp -- exchanging the
// rates through shared memory:
//   stage 1  every warp derives its share of the process rates / isotope factors from the
//            state at the start of the sub-step                 (09/mom/mobi.F:2150-2548)
//   stage 2  every warp advances the state variables it owns and re-evaluates their
//            flags                                             (:2552-2777, 3175-3251)
// Each expression is evaluated exactly as in mobi_src above (same operands, same order), so
// the two kernels agree bit for bit.  The level epilogue of mobi_driver (:880-1400) is split
// the same way.  The export chain (expo -> impo of the next level) stays serial in k.
// ------------------------------------------------------------------------------------
enum WsRate {
  R_npp = 0, R_dopupt, R_npp_D, R_no3upt_D, R_dopupt_D, R_fcassim, R_npp_Diat, R_dopupt_Diat, R_sipr0, R_recy_don, R_fcrecy,
  R_graz, R_graz_Z, R_graz_Det, R_graz_D, R_graz_Diat, R_morp, R_morpt, R_morp_D, R_morpt_D, R_morp_Diat, R_morpt_Diat, R_morz,
  R_remi, R_expo, R_expo_phos, R_recy_dop, R_dissl, R_expocaco3, R_opldis, R_expoopl, R_remife, R_expofe, R_dig, R_dig_P, R_dig_Z,
  R_dig_Det, R_dig_Diat, R_dig_D, R_excr, R_sf, R_sf_P, R_sf_Z, R_sf_Det, R_sf_Diat, R_sf_D, R_sf_phos, R_nr_excr_D, R_calpro,
  R_feprime, R_fecol, R_pw58, R_fcexcr, R_fcnfix, R_rtdic13, R_rtcaco3c13, R_GM15ptn, R_rtphytn15, R_rtdiatn15, R_rtzoopn15,
  R_rtdetrn15, R_rtdiazn15, R_fcnpp, R_rtphytc13, R_rtdiatc13, R_rtzoopc13, R_rtdetrc13, R_rtdoc13, R_rtdiazc13, R_N
};
enum WsLev {
  // export chain carried from level to level (09/mom/mobi.F:1280-1288)
  X_expo = 0, X_expo_phos, X_rn15expo, X_rc13expo, X_rcaco3c13expo, X_expofe, X_expocaco3, X_expoopl,
  // per-level sums over the sub-steps (the *out arguments of mobi_src, :2762-2777)
  S_expo, S_expo_phos, S_rn15expo, S_rc13expo, S_rcaco3c13expo, S_expofe, S_expocaco3, S_expoopl, S_calpro, S_dissl, S_nfix,
  // phytoplankton / detritus P:N ratios, double buffered over the sub-steps, and their flags (:1781-1784, 1814-1890)
  L_ptn_P0, L_ptn_P1, L_ptn_detr0, L_ptn_detr1, L_sf_P_phosflag, L_sf_detr_phosflag,
  // epilogue exchange
  L_wcdeni, L_bwfrac, L_N
};
struct WsSm {
  double B[MOBI_NVAR][32];   // state at the start of the sub-step
  double F[MOBI_NVAR][32];   // flags (1.0 / 0.0)
  double C[MOBI_NVAR][32];   // clipped inputs tnpzd(k,:)
  double R[R_N][32];
  double L[L_N][32];
};
#define WS_WARPS 4
// barrier of one column group: named barrier 1 + group, WS_WARPS warps
#define WS_BAR() asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(32 * WS_WARPS) : "memory")
#define WS_ROLES 8   // every warp runs roles w and w + WS_WARPS one after the other

// G column groups per CTA: the groups of a CTA are independent (own shared-memory block, own named barrier) but run
// the same instruction stream, so an SM fetches the (instruction-cache sized) sub-step code once for G groups.
template <int G>
__global__ void __launch_bounds__(32 * WS_WARPS * G) k_mobi_ws(const DevView v, int mi, int nbio, double dtbio, double rdtts, double rnbio) {
  extern __shared__ __align__(16) unsigned char ws_raw[];
  const int grp = threadIdx.x / (32 * WS_WARPS);
  const int gidx = blockIdx.x * G + grp;                    // column group of these WS_WARPS warps
  if (gidx * 32 >= v.mobi_ncols) return;                    // whole group idle (its barrier has no other users)
  WsSm &sm = *reinterpret_cast<WsSm *>(ws_raw + (size_t)grp * sizeof(WsSm));
  const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) % WS_WARPS;
  const int cidx = gidx * 32 + lane;
  const bool valid = cidx < v.mobi_ncols;
  const int col = valid ? v.mobi_cols[cidx] : 0;
  const int i = col % v.imt + 1, j = col / v.imt + v.jbase;
  const int kmx = valid ? v.kmt[col] : 0;
  const int kmax = v.kmt[v.mobi_cols[gidx * 32]];   // columns are sorted by depth, deepest first
  const MobiPar *__restrict__ P = v.mobi_par;
  const int *__restrict__ ix = v.mobi_idx;
  const long long n3 = v.n3;
  const double gamma1 = P->gamma1, redptn = P->redptn, redctn = P->redctn, redntp = P->redntp, diazntp = P->diazntp;
  const double diazptn = P->diazptn, dfr = P->dfr, pfr = P->pfr, dfrt = P->dfrt, geZ = P->geZ, rfeton = P->rfeton;
  const double rnd = QDIV(redntp, diazntp);
#define SB(m) sm.B[m][lane]
#define SF(m) sm.F[m][lane]
#define SC(m) sm.C[m][lane]
#define SR(x) sm.R[R_##x][lane]
#define SL(x) sm.L[x][lane]
  if (w == 0) {
#pragma unroll
    for (int q = X_expo; q <= X_expoopl; q++) SL(q) = 0.0;
  }
  // per-warp accumulators of the *out sums
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0, acc4 = 0.0, acc5 = 0.0, acc6 = 0.0;
  // epilogue carry of warp 1 (finished after the barrier, see E3)
  double e3_no3 = 0.0, e3_din15 = 0.0;
  bool e3_pending = false;
  long long e3_c = 0;

  for (int k = 1; k <= kmax; k++) {
    const bool act = k <= kmx;
    const long long c = X3(i, k, j);
    const double dztk = v.dzt[k - 1], dztrk = v.dztr[k - 1];
    const double *__restrict__ pre = v.mobi_pre + c;
#define PRE(f) (act ? pre[(long long)(f) * n3] : 1.0)
    // ---- E3 of the previous level (warp 1): water-column denitrification closes NO3, DIN15 and ALK ----
    if (w == 1 && e3_pending) {
      const double wcdeni = SL(L_wcdeni), bwfrac = SL(L_bwfrac);
      v.src[e3_c + (long long)(ix[IX_SRC + V_NO3] - 1) * n3] = e3_no3 - wcdeni;
      v.src[e3_c + (long long)(ix[IX_SRC + V_DIN15] - 1) * n3] = e3_din15 - bwfrac * wcdeni;
      e3_pending = false;
    }
    // ---- P1: gather, flags from the raw inputs, clip (09/mom/tracer.F:393-503, mobi.F:1781-1926) ----
    {
      double raw[8];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int m = 8 * w + q;
        raw[q] = act ? v.t_m1[c + (long long)(ix[IX_TR + m] - 1) * n3] : 1.0;
        SF(m) = (raw[q] - TRCMIN >= 0.0) ? 1.0 : 0.0;
        const double cl = fmax(raw[q], TRCMIN);
        SB(m) = cl;
        SC(m) = cl;
      }
      if (w == 0) {   // states 0..7 include PHYT, PHYT_PHOS, DETR, DETR_PHOS
        const double ptn_P = QDIV(raw[V_PHYT_PHOS], raw[V_PHYT]);
        SL(L_ptn_P0) = ptn_P;
        SL(L_sf_P_phosflag) = 0.5 + fsign(0.5, ptn_P - gamma1 * redptn);
        const double ptn_detr = QDIV(raw[V_DETR_PHOS], raw[V_DETR]);
        SL(L_ptn_detr0) = ptn_detr;
        SL(L_sf_detr_phosflag) = 0.5 + fsign(0.5, ptn_detr - gamma1 * redptn);
      }
    }
    // level constants of this warp's role
    const double bct = PRE(PR_BCT);
    double lc0 = 0.0, lc1 = 0.0, lc2 = 0.0, lc3 = 0.0, lc4 = 0.0, lc5 = 0.0, lc6 = 0.0, lc7 = 0.0;
    switch (w) {
      case 0: lc0 = PRE(PR_AVEJ); break;
      case 1: lc0 = PRE(PR_AVEJ_DIAT); break;
      case 2: lc0 = P->gbio * PRE(PR_BCTZ); lc1 = PRE(PR_NUD); lc2 = PRE(PR_DISSK1); lc3 = PRE(PR_CAPR);
              lc4 = P->wd[k - 1]; lc5 = P->wc[k - 1]; lc6 = P->wo[k - 1];
              lc7 = PRE(PR_AC13B); break;                                             // role 6
      case 3: lc0 = PRE(PR_AOU8); lc1 = PRE(PR_O2FLAG); lc2 = PRE(PR_AVEJ_D); break;   // roles 3 and 7
      default: break;
    }
    acc0 = acc1 = acc2 = acc3 = acc4 = acc5 = acc6 = 0.0;
    WS_BAR();

#define UPD(m, expr)                                   \
  do {                                                 \
    const double nv_ = (expr);                         \
    SB(m) = nv_;                                       \
    if (nv_ - TRCMIN < 0.0) SF(m) = 0.0;               \
  } while (0)
    // Role-major loops: every warp runs its own compact sub-step loop (stage 1, barrier, stage 2, barrier), so its
    // instruction stream is a short predictable loop instead of two indirect branches per sub-step into a kernel-sized
    // switch.  All warps execute the same number of barriers.

#define WS_LOOP_HEAD                          \
    for (int n = 1; n <= nbio; n++) {         \
      const int par = (n - 1) & 1;            \
      const double ptn_P = SL(L_ptn_P0 + par), ptn_detr = SL(L_ptn_detr0 + par); \
      (void)ptn_P; (void)ptn_detr;
    switch (w) {
      case 0: {   // roles 0 and 4
        WS_LOOP_HEAD
        {
          // phytoplankton growth (:2150-2206), NO3 assimilation fractionation (:2441-2452)
          const double biophyt = SB(V_PHYT), biodfe = SB(V_DFE), biodop = SB(V_DOP), biopo4 = SB(V_PO4), biono3 = SB(V_NO3);
          double p1 = fmin(biophyt, P->pmax), p2 = fmax(0.0, biophyt - P->pmax);
          double k1n = QDIV((P->knmin * p1 + P->knmax * p2), (p1 + p2));
          double k1p_P = k1n * ptn_P;
          double kfevar = QDIV((P->kfemin * p1 + P->kfemax * p2), (p1 + p2));
          double deffe = QDIV(biodfe, (kfevar + biodfe));
          double jmax = P->abio_P * bct * deffe;
          double limP_dop = QDIV(P->hdop * biodop, (k1p_P + biodop));
          double limP_po4 = QDIV(biopo4, (k1p_P + biopo4));
          double dopupt_flag = 0.5 + fsign(0.5, limP_dop - limP_po4);
          double limP = limP_dop * dopupt_flag + limP_po4 * (1. - dopupt_flag);
          double u_P = fmin(lc0, jmax * limP);
          u_P = fmin(u_P, QDIV(jmax * biono3, (k1n + biono3)));
          double npp = u_P * biophyt;
          SR(dopupt) = npp * dopupt_flag;
          npp = npp * SF(V_NO3) * (dopupt_flag * SF(V_DOP) + (1. - dopupt_flag) * SF(V_PO4)) * SF(V_DIN15);
          SR(npp) = npp;
          double uno3 = fmax(fmin(QDIV(npp * dtbio, biono3), 0.999), TRCMIN);
          double rno3 = fmax(fmin(QDIV(SB(V_DIN15), (biono3 - SB(V_DIN15))), 2 * RN15STD), RN15STD / 2.);
          double bassim = rno3 + QDIV(QDIV(P->eps_assim * (1 - uno3), uno3) * log(1 - uno3) * rno3, 1000.);
          SR(fcassim) = QDIV(bassim, (1 + bassim));
        }
        {
          // organic iron adsorption exponent (:2278), excretion / N2-fixation fractionation (:2466-2478), calcite 13C ratios
          const double biozoop = SB(V_ZOOP);
          SR(pw58) = pow(((SB(V_DETR) * SF(V_DETR)) * P->mc * redctn), 0.58);
          double rzoop = fmax(fmin(QDIV(SB(V_ZOOPN15), (biozoop - SB(V_ZOOPN15))), 2. * RN15STD), RN15STD / 2.);
          double bexcr = rzoop - QDIV(P->eps_excr * rzoop, 1000.);
          SR(fcexcr) = QDIV(bexcr, (1 + bexcr));
          double bnfix = RN15STD - QDIV(P->eps_nfix * RN15STD, 1000.);
          SR(fcnfix) = QDIV(bnfix, (1 + bnfix));
          SR(rtdic13) = CL13(QDIV(SB(V_DIC13), SB(V_DIC)));
          SR(rtcaco3c13) = CL13(QDIV(SB(V_CACO3C13), SB(V_CACO3)));
          double GM15ptc = 0.0060 + 0.0069 * SB(V_PO4);
          SR(GM15ptn) = GM15ptc * redctn * 1.e3;
        }
        WS_BAR();
        {
          const double npp = SR(npp), dopupt = SR(dopupt), morp = SR(morp), morpt = SR(morpt), graz = SR(graz), remi = SR(remi);
          const double morpt_D = SR(morpt_D), npp_D = SR(npp_D), dopupt_D = SR(dopupt_D), recy_dop = SR(recy_dop), excr = SR(excr);
          const double morpt_Diat = SR(morpt_Diat), npp_Diat = SR(npp_Diat), dopupt_Diat = SR(dopupt_Diat), morp_Diat = SR(morp_Diat);
          const double GM15ptn = SR(GM15ptn);
          const double biophyt = SB(V_PHYT), biophyt_phos = SB(V_PHYT_PHOS);
          UPD(V_PO4, SB(V_PO4) + dtbio * (dopupt * ptn_P - GM15ptn * npp + (1. - dfrt) * morpt * ptn_P + (1. - pfr) * remi * ptn_detr +
                                         diazptn * (morpt_D - (npp_D - dopupt_D)) + recy_dop +
                                         redptn * (excr + (1. - dfrt) * morpt_Diat - (npp_Diat - dopupt_Diat))));
          UPD(V_DOP, SB(V_DOP) + dtbio * (dfr * morp * ptn_P + redptn * (dfr * morp_Diat + dfrt * morpt_Diat - dopupt_Diat) +
                                         dfrt * morpt * ptn_P + pfr * remi * ptn_detr - ptn_P * dopupt - diazptn * dopupt_D - recy_dop));
          const double nphyt = biophyt + dtbio * (npp - morp - graz - morpt);
          const double nphos = biophyt_phos + dtbio * (npp * GM15ptn - morp * ptn_P - graz * ptn_P - morpt * ptn_P);
          UPD(V_PHYT, nphyt);
          UPD(V_PHYT_PHOS, nphos);
          SL(L_ptn_P0 + (par ^ 1)) = QDIV(nphos, nphyt);
        }
        {
          const double rtphytn15 = SR(rtphytn15), rtdiatn15 = SR(rtdiatn15), rtdiazn15 = SR(rtdiazn15), rtdetrn15 = SR(rtdetrn15);
          const double fcassim = SR(fcassim), fcexcr = SR(fcexcr), fcrecy = SR(fcrecy);
          const double morpt = SR(morpt), morpt_Diat = SR(morpt_Diat), npp_Diat = SR(npp_Diat), excr = SR(excr), morpt_D = SR(morpt_D);
          const double nr_excr_D = SR(nr_excr_D), morp_D = SR(morp_D), remi = SR(remi), recy_don = SR(recy_don), npp = SR(npp);
          const double no3upt_D = SR(no3upt_D), morp = SR(morp), morp_Diat = SR(morp_Diat);
          const double nr_excr_P = 0.0, nr_excr_detr = 0.0;
          UPD(V_DIN15, SB(V_DIN15) + dtbio * (rtphytn15 * (1. - dfrt) * morpt + rtphytn15 * nr_excr_P + rtdiatn15 * (1. - dfrt) * morpt_Diat -
                                             fcassim * npp_Diat + fcexcr * excr + rtdiazn15 * morpt_D + rtdiazn15 * nr_excr_D +
                                             rtdiazn15 * morp_D * (1. - rnd) + rtdetrn15 * (1. - pfr) * remi + rtdetrn15 * nr_excr_detr +
                                             fcrecy * recy_don - fcassim * npp - fcassim * no3upt_D));
          UPD(V_DON15, SB(V_DON15) + dtbio * (dfr * rtphytn15 * morp + dfr * rtdiatn15 * morp_Diat + dfrt * rtdiatn15 * morpt_Diat +
                                             dfrt * rtphytn15 * morpt + rtdetrn15 * pfr * remi - fcrecy * recy_don));
        }
        {
          // 13C of the diazotrophs (taken over from warp 3)
          const double rtdiazc13 = SR(rtdiazc13);
          UPD(V_DIAZC13, SB(V_DIAZC13) + dtbio * redctn * (SR(fcnpp) * SR(npp_D) - rtdiazc13 * (SR(morp_D) + SR(graz_D) + SR(morpt_D))));
          // ... and of the diatoms
          const double rtdiatc13 = SR(rtdiatc13);
          UPD(V_DIATC13, SB(V_DIATC13) + dtbio * redctn * (SR(fcnpp) * SR(npp_Diat) - rtdiatc13 * (SR(morp_Diat) + SR(graz_Diat) + SR(morpt_Diat))));
        }
        WS_BAR();
        }
      } break;
      case 1: {   // roles 1 and 5
        WS_LOOP_HEAD
        {
          // diatom growth (:2160-2200), opal production ratio (:2540)
          const double biodiat = SB(V_DIAT), biodfe = SB(V_DFE), biosil = SB(V_SIL), biodop = SB(V_DOP), biopo4 = SB(V_PO4);
          const double biono3 = SB(V_NO3);
          double p1 = fmin(biodiat, P->pmax_Diat), p2 = fmax(0.0, biodiat - P->pmax_Diat);
          double kfevar_Diat = QDIV((P->kfemin_Diat * p1 + P->kfemax_Diat * p2), (p1 + p2));
          double k1n_Diat = QDIV((P->knmin_Diat * p1 + P->knmax_Diat * p2), (p1 + p2));
          double k1p_Diat = k1n_Diat * redptn;
          double deffe_Diat = QDIV(biodfe, (kfevar_Diat + biodfe));
          double jmax_Diat = P->abiodiat * bct * deffe_Diat;
          double limSi = QDIV(biosil, (5.e-3 + biosil));  // k1si = 5.e-3 (:2181)
          double limP_dop = QDIV(P->hdop * biodop, (k1p_Diat + biodop));
          double limP_po4 = QDIV(biopo4, (k1p_Diat + biopo4));
          double dopupt_Diat_flag = 0.5 + fsign(0.5, limP_dop - limP_po4);
          double limP_Diat = limP_dop * dopupt_Diat_flag + limP_po4 * (1. - dopupt_Diat_flag);
          double u_Diat = fmin(lc0, jmax_Diat * limSi);
          u_Diat = fmin(u_Diat, jmax_Diat * limP_Diat);
          u_Diat = fmin(u_Diat, QDIV(jmax_Diat * biono3, (k1n_Diat + biono3)));
          double npp_Diat = u_Diat * biodiat;
          SR(dopupt_Diat) = npp_Diat * dopupt_Diat_flag;
          SR(npp_Diat) = npp_Diat * SF(V_NO3) * (dopupt_Diat_flag * SF(V_DOP) + (1. - dopupt_Diat_flag) * SF(V_PO4)) * SF(V_DIN15);
          SR(sipr0) = (-0.46204044117647 * tanh(6.9 * biodfe * 1.e3 + -3.673092) + 1.60266544117647);
        }
        {
          // DON recycling and its fractionation (:2453-2465), 15N ratios of the organic pools (:2479-2500)
          const double biodon = SB(V_DON);
          double recy_don = P->nudon0 * bct * biodon;
          recy_don = recy_don * SF(V_DON) * SF(V_DON15);
          SR(recy_don) = recy_don;
          double udon = fmax(fmin(QDIV(recy_don * dtbio, biodon), 0.999), TRCMIN);
          double rdon = fmax(fmin(QDIV(SB(V_DON15), (biodon - SB(V_DON15))), 2 * RN15STD), RN15STD / 2.);
          double brecy = rdon + QDIV(QDIV(P->eps_recy * (1 - udon), udon) * log(1 - udon) * rdon, 1000.);
          SR(fcrecy) = QDIV(brecy, (1 + brecy));
          // (the 15N ratios of phytoplankton, diatoms, zooplankton and diazotrophs are formed by warp 3: load balance)
          const double rtdetrn15 = CL15(QDIV(SB(V_DETRN15), SB(V_DETR)));
          SR(rtdetrn15) = rtdetrn15;
          acc0 = acc0 + rtdetrn15;   // rn15expoout
        }
        WS_BAR();
        {
          const double morp = SR(morp), sf = SR(sf), morz = SR(morz), remi = SR(remi), graz_Det = SR(graz_Det), expo = SR(expo);
          const double morp_D = SR(morp_D), morp_Diat = SR(morp_Diat), sf_phos = SR(sf_phos), expo_phos = SR(expo_phos);
          const double impo = SL(X_expo) * dztrk, impo_phos = SL(X_expo_phos) * dztrk;
          const double ndetr = SB(V_DETR) + dtbio * ((1. - dfr) * morp + sf + morz - remi - graz_Det - expo + impo + morp_D * rnd +
                                                     (1. - dfr) * morp_Diat);
          const double ndphos = SB(V_DETR_PHOS) + dtbio * ((1. - dfr) * morp * ptn_P + sf_phos + morz * redptn - remi * ptn_detr -
                                                           graz_Det * ptn_detr - expo_phos + impo_phos + morp_D * rnd * redptn +
                                                           (1. - dfr) * morp_Diat * redptn);
          UPD(V_DETR, ndetr);
          UPD(V_DETR_PHOS, ndphos);
          SL(L_ptn_detr0 + (par ^ 1)) = QDIV(ndphos, ndetr);
          UPD(V_ZOOP, SB(V_ZOOP) + dtbio * (SR(dig) - morz - SR(graz_Z) - SR(excr)));
          UPD(V_DIAZ, SB(V_DIAZ) + dtbio * (SR(npp_D) - morp_D - SR(morpt_D) - SR(graz_D)));
          UPD(V_DIAT, SB(V_DIAT) + dtbio * (SR(npp_Diat) - morp_Diat - SR(graz_Diat) - SR(morpt_Diat)));
        }
        {
          const double rtphytn15 = SR(rtphytn15), rtdiatn15 = SR(rtdiatn15), rtdiazn15 = SR(rtdiazn15), rtdetrn15 = SR(rtdetrn15);
          const double rtzoopn15 = SR(rtzoopn15), fcassim = SR(fcassim), fcexcr = SR(fcexcr), fcnfix = SR(fcnfix);
          const double npp = SR(npp), morp = SR(morp), graz = SR(graz), morpt = SR(morpt), npp_Diat = SR(npp_Diat);
          const double morp_Diat = SR(morp_Diat), graz_Diat = SR(graz_Diat), morpt_Diat = SR(morpt_Diat), morz = SR(morz);
          const double graz_Z = SR(graz_Z), excr = SR(excr), npp_D = SR(npp_D), no3upt_D = SR(no3upt_D), morp_D = SR(morp_D);
          const double graz_D = SR(graz_D), morpt_D = SR(morpt_D);
          UPD(V_PHYTN15, SB(V_PHYTN15) + dtbio * (fcassim * npp - rtphytn15 * morp - rtphytn15 * graz - rtphytn15 * morpt));
          UPD(V_DIATN15, SB(V_DIATN15) + dtbio * (fcassim * npp_Diat - rtdiatn15 * morp_Diat - rtdiatn15 * graz_Diat - rtdiatn15 * morpt_Diat));
          UPD(V_ZOOPN15, SB(V_ZOOPN15) + dtbio * (rtphytn15 * SR(dig_P) + rtdiatn15 * SR(dig_Diat) + rtzoopn15 * SR(dig_Z) +
                                                 rtdetrn15 * SR(dig_Det) + rtdiazn15 * SR(dig_D) - rtzoopn15 * morz - rtzoopn15 * graz_Z -
                                                 fcexcr * excr));
          UPD(V_DIAZN15, SB(V_DIAZN15) + dtbio * (fcnfix * (npp_D - no3upt_D) + fcassim * no3upt_D - rtdiazn15 * morp_D - rtdiazn15 * graz_D -
                                                 rtdiazn15 * morpt_D));
        }
        WS_BAR();
        }
      } break;
      case 2: {   // roles 2 and 6
        WS_LOOP_HEAD
        {
          // grazing, mortality, remineralisation, sinking (:2208-2260), flags (:2284-2334), digestion / sloppy feeding (:2335-2440)
          const double biophyt = SB(V_PHYT), biodetr = SB(V_DETR), biozoop = SB(V_ZOOP), biodiaz = SB(V_DIAZ), biodiat = SB(V_DIAT);
          const double gmax = lc0, nud = lc1;
          const double nupt = P->nupt0 * bct, nupt_D = P->nupt0_D * bct, nudt = P->nudt0 * bct;
          double thetaZ = P->zprefP * biophyt + P->zprefDet * biodetr + P->zprefZ * biozoop + P->zprefDiaz * biodiaz + P->kzoo +
                          P->zprefDiat * biodiat;
          double ing_P = QDIV(P->zprefP, thetaZ), ing_Det = QDIV(P->zprefDet, thetaZ), ing_Z = QDIV(P->zprefZ, thetaZ);
          double ing_D = QDIV(P->zprefDiaz, thetaZ), ing_Diat = QDIV(P->zprefDiat, thetaZ);
          double graz_D = gmax * ing_D * biodiaz * biozoop;
          double morpt_D = nupt_D * biodiaz;
          double morp_D = P->nup_D * biodiaz * biodiaz;
          double graz = gmax * ing_P * biophyt * biozoop;
          double graz_Z = gmax * ing_Z * biozoop * biozoop;
          double graz_Det = gmax * ing_Det * biodetr * biozoop;
          double morp = P->nup * biophyt;
          double morpt = nupt * biophyt;
          double recy_dop = P->nudop0 * bct * SB(V_DOP);
          double morz = P->nuz * biozoop * biozoop;
          double remi = nud * bct * biodetr;
          double expo = lc4 * biodetr;
          double expo_phos = lc4 * SB(V_DETR_PHOS);
          double dissl = SB(V_CACO3) * lc2;
          double expocaco3 = lc5 * SB(V_CACO3);
          double graz_Diat = gmax * ing_Diat * biodiat * biozoop;
          double morp_Diat = P->nu_diat * biodiat;
          double morpt_Diat = nudt * biodiat;
          double opldis = SB(V_OPL) * P->opl_disk0;
          double expoopl = lc6 * SB(V_OPL);
          double remife = nud * bct * SB(V_DETRFE);
          double expofe = lc4 * SB(V_DETRFE);
          graz = graz * SF(V_PHYT) * SF(V_PHYT_PHOS) * SL(L_sf_P_phosflag) * SF(V_PHYTN15);
          graz_Z = graz_Z * SF(V_ZOOP) * SF(V_ZOOPN15);
          graz_Det = graz_Det * SF(V_DETR) * SF(V_DETR_PHOS) * SL(L_sf_detr_phosflag) * SF(V_DETRN15);
          morp = morp * SF(V_PHYT) * SF(V_PHYT_PHOS) * SF(V_PHYTN15);
          morpt = morpt * SF(V_PHYT) * SF(V_PHYT_PHOS) * SF(V_PHYTN15);
          morz = morz * SF(V_ZOOP) * SF(V_ZOOPN15);
          remi = remi * SF(V_DETR) * SF(V_DETR_PHOS) * SF(V_DETRN15);
          expo = expo * SF(V_DETR) * SF(V_DETRN15);
          expo_phos = expo_phos * SF(V_DETR_PHOS);
          recy_dop = recy_dop * SF(V_DOP);
          graz_D = graz_D * SF(V_DIAZ) * SF(V_DIAZN15);
          morpt_D = morpt_D * SF(V_DIAZ) * SF(V_DIAZN15);
          morp_D = morp_D * SF(V_DIAZ) * SF(V_DIAZN15);
          dissl = dissl * SF(V_CACO3);
          expocaco3 = expocaco3 * SF(V_CACO3);
          graz_Diat = graz_Diat * SF(V_DIAT);
          morp_Diat = morp_Diat * SF(V_DIAT);
          morpt_Diat = morpt_Diat * SF(V_DIAT);
          remife = remife * SF(V_DETRFE);
          expofe = expofe * SF(V_DETRFE);
          double dig_P = gamma1 * graz, dig_Z = gamma1 * graz_Z, dig_Det = gamma1 * graz_Det, dig_Diat = gamma1 * graz_Diat;
          double dig = dig_Z + dig_P + dig_Det + dig_Diat;
          double excr = gamma1 * (1 - geZ) * graz_Z + gamma1 * (1 - geZ) * graz + gamma1 * (1 - geZ) * graz_Det + gamma1 * (1 - geZ) * graz_Diat;
          double sf_P = (1. - gamma1) * graz, sf_Z = (1. - gamma1) * graz_Z, sf_Det = (1. - gamma1) * graz_Det;
          double sf_Diat = (1. - gamma1) * graz_Diat;
          double sf = sf_P + sf_Z + sf_Det + sf_Diat;
          double sf_P_phos = (graz * ptn_P - dig_P * redptn);
          double sf_Det_phos = (graz_Det * ptn_detr - dig_Det * redptn);
          double sf_phos = sf_P_phos + sf_Z * redptn + sf_Det_phos + sf_Diat * redptn;
          double dig_D = gamma1 * graz_D * (QDIV(redntp, diazntp));
          dig = dig + dig_D;
          excr = excr + gamma1 * (1 - geZ) * graz_D * (QDIV(redntp, diazntp));
          double nr_excr_D = gamma1 * graz_D * (1 - (QDIV(redntp, diazntp))) + (1 - gamma1) * graz_D * (1 - (QDIV(redntp, diazntp)));
          double sf_D = (1 - gamma1) * graz_D * (QDIV(redntp, diazntp));
          sf = sf + sf_D;
          sf_phos = sf_phos + sf_D * redptn;
          double calpro = ((sf_Z + morz) * lc3 + (sf_P + morp) * lc3) * redctn * 1.e3;
          opldis = opldis * SF(V_OPL);
          expoopl = expoopl * SF(V_OPL);
          SR(graz) = graz; SR(graz_Z) = graz_Z; SR(graz_Det) = graz_Det; SR(graz_D) = graz_D; SR(graz_Diat) = graz_Diat;
          SR(morp) = morp; SR(morpt) = morpt; SR(morp_D) = morp_D; SR(morpt_D) = morpt_D; SR(morp_Diat) = morp_Diat;
          SR(morpt_Diat) = morpt_Diat; SR(morz) = morz; SR(remi) = remi; SR(expo) = expo; SR(expo_phos) = expo_phos;
          SR(recy_dop) = recy_dop; SR(dissl) = dissl; SR(expocaco3) = expocaco3; SR(opldis) = opldis; SR(expoopl) = expoopl;
          SR(remife) = remife; SR(expofe) = expofe; SR(dig) = dig; SR(dig_P) = dig_P; SR(dig_Z) = dig_Z; SR(dig_Det) = dig_Det;
          SR(dig_Diat) = dig_Diat; SR(dig_D) = dig_D; SR(excr) = excr; SR(sf) = sf; SR(sf_P) = sf_P; SR(sf_Z) = sf_Z;
          SR(sf_Det) = sf_Det; SR(sf_Diat) = sf_Diat; SR(sf_D) = sf_D; SR(sf_phos) = sf_phos; SR(nr_excr_D) = nr_excr_D;
          SR(calpro) = calpro;
          // *out sums (:2762-2777)
          acc0 = acc0 + expo; acc1 = acc1 + expo_phos; acc2 = acc2 + calpro; acc3 = acc3 + dissl; acc4 = acc4 + expocaco3;
          acc5 = acc5 + expoopl; acc6 = acc6 + expofe;
        }
        {
          // 13C fractionation of primary production and ratios of the living pools (:2501-2530)
          const double biodic = SB(V_DIC);
          double rdic13 = fmax(fmin(QDIV(SB(V_DIC13), (biodic - SB(V_DIC13))), 2. * RC13STD), 0.5 * RC13STD);
          double bc13npp = lc7 * rdic13;
          SR(fcnpp) = QDIV(bc13npp, (1 + bc13npp));
          SR(rtphytc13) = CL13(QDIV(SB(V_PHYTC13), (SB(V_PHYT) * redctn)));
          SR(rtdiatc13) = CL13(QDIV(SB(V_DIATC13), (SB(V_DIAT) * redctn)));
          SR(rtzoopc13) = CL13(QDIV(SB(V_ZOOPC13), (SB(V_ZOOP) * redctn)));
          SR(rtdetrc13) = CL13(QDIV(SB(V_DETRC13), (SB(V_DETR) * redctn)));
        }
        WS_BAR();
        {
          const double excr = SR(excr), remi = SR(remi), morpt = SR(morpt), npp = SR(npp), morpt_Diat = SR(morpt_Diat);
          const double npp_Diat = SR(npp_Diat), morpt_D = SR(morpt_D), npp_D = SR(npp_D), recy_don = SR(recy_don);
          const double nr_excr_D = SR(nr_excr_D), morp_D = SR(morp_D), no3upt_D = SR(no3upt_D), morp = SR(morp);
          const double morp_Diat = SR(morp_Diat);
          const double nr_excr_P = 0.0, nr_excr_detr = 0.0;
          UPD(V_DIC, SB(V_DIC) + dtbio * redctn *
                                     (excr + (1. - pfr) * remi + (1. - dfrt) * morpt - npp + (1. - dfrt) * morpt_Diat - npp_Diat + morpt_D -
                                      npp_D + recy_don + nr_excr_D + nr_excr_P + nr_excr_detr + morp_D * (1. - rnd)));
          UPD(V_NO3, SB(V_NO3) + dtbio * (excr + (1. - pfr) * remi + (1. - dfrt) * morpt - npp + (1. - dfrt) * morpt_Diat - npp_Diat +
                                         morpt_D - no3upt_D + recy_don + nr_excr_D + nr_excr_P + nr_excr_detr + morp_D * (1. - rnd)));
          UPD(V_DON, SB(V_DON) + dtbio * (dfr * morp + dfrt * morpt + pfr * remi - recy_don + dfr * morp_Diat + dfrt * morpt_Diat));
          UPD(V_CACO3, SB(V_CACO3) + dtbio * (SR(calpro) - SR(dissl) - SR(expocaco3) + SL(X_expocaco3) * dztrk));
          {
            // 13C of calcite (taken over from warp 3)
            const double rtcaco3c13 = SR(rtcaco3c13), rcaco3c13impo = SL(X_rcaco3c13expo) * dztrk;
            UPD(V_CACO3C13, SB(V_CACO3C13) + dtbio * (SR(rtdic13) * SR(calpro) - rtcaco3c13 * SR(dissl) - rtcaco3c13 * SR(expocaco3) + rcaco3c13impo));
          }
        }
        {
          const double rtphytn15 = SR(rtphytn15), rtdiatn15 = SR(rtdiatn15), rtdiazn15 = SR(rtdiazn15), rtdetrn15 = SR(rtdetrn15);
          const double rtzoopn15 = SR(rtzoopn15);
          const double morp = SR(morp), morp_Diat = SR(morp_Diat), sf_Diat = SR(sf_Diat), sf_P = SR(sf_P), sf_Z = SR(sf_Z);
          const double sf_Det = SR(sf_Det), sf_D = SR(sf_D), morz = SR(morz), remi = SR(remi), graz_Det = SR(graz_Det), expo = SR(expo);
          const double morp_D = SR(morp_D);
          const double impo = SL(X_expo) * dztrk, rn15impo = SL(X_rn15expo);
          UPD(V_DETRN15, SB(V_DETRN15) + dtbio * (rtphytn15 * (1. - dfr) * morp + rtdiatn15 * (1. - dfr) * morp_Diat + rtdiatn15 * sf_Diat +
                                                 rtphytn15 * sf_P + rtzoopn15 * sf_Z + rtdetrn15 * sf_Det + rtdiazn15 * sf_D +
                                                 rtzoopn15 * morz - rtdetrn15 * remi - rtdetrn15 * graz_Det - rtdetrn15 * expo +
                                                 rn15impo * impo + rtdiazn15 * morp_D * rnd));
          const double rtphytc13 = SR(rtphytc13), rtzoopc13 = SR(rtzoopc13), rtdiazc13 = SR(rtdiazc13), rtdetrc13 = SR(rtdetrc13);
          const double rtdiatc13 = SR(rtdiatc13), rtdoc13 = SR(rtdoc13), fcnpp = SR(fcnpp);
          const double morpt = SR(morpt), excr = SR(excr), morpt_D = SR(morpt_D), nr_excr_D = SR(nr_excr_D);
          const double morpt_Diat = SR(morpt_Diat), npp_Diat = SR(npp_Diat), recy_don = SR(recy_don), npp = SR(npp), npp_D = SR(npp_D);
          const double nr_excr_P = 0.0, nr_excr_detr = 0.0;
          UPD(V_DIC13, SB(V_DIC13) + dtbio * redctn *
                                         (rtphytc13 * (1. - dfrt) * morpt + rtphytc13 * nr_excr_P + rtzoopc13 * excr + rtdiazc13 * morpt_D +
                                          rtdiazc13 * nr_excr_D + rtdiazc13 * morp_D * (1 - rnd) + rtdetrc13 * (1. - pfr) * remi +
                                          rtdetrc13 * nr_excr_detr + rtdiatc13 * (1. - dfrt) * morpt_Diat - fcnpp * npp_Diat +
                                          rtdoc13 * recy_don - fcnpp * npp - fcnpp * npp_D));
        }
        WS_BAR();
        }
      } break;
      case 3: {   // roles 3 and 7
        WS_LOOP_HEAD
        {
          // iron speciation and scavenging (:2262-2283)
          const double biodon = SB(V_DON), biodfe = SB(V_DFE), aou8 = lc0, o2flag = lc1;
          double ligand = QDIV(fmax(QDIV(aou8, 66.) + QDIV(pow(biodon, 0.8), 4.8), 0.5), 1000.);
          double fepa = (1.0 + P->kfeleq * (ligand - biodfe)) * o2flag;
          double feprime = (QDIV((-fepa + sqrt(fepa * fepa + 4.0 * P->kfeleq * biodfe)), (2.0 * P->kfeleq))) * o2flag;
          double fecol = P->kfecol * (feprime * feprime) * o2flag;
          SR(feprime) = feprime;
          SR(fecol) = fecol * SF(V_DFE);
        }
        {
          // diazotroph growth and N2 fixation (:2164-2166, 2200-2206, 2226-2232); the phosphorus limitation of the
          // ordinary phytoplankton is re-derived here (same expressions as warp 0)
          const double biophyt = SB(V_PHYT), biodfe = SB(V_DFE), biodop = SB(V_DOP), biopo4 = SB(V_PO4), biono3 = SB(V_NO3);
          const double biodiaz = SB(V_DIAZ);
          double p1 = fmin(biophyt, P->pmax), p2 = fmax(0.0, biophyt - P->pmax);
          double k1n = QDIV((P->knmin * p1 + P->knmax * p2), (p1 + p2));
          double k1p_P = k1n * ptn_P;
          double deffe_D = QDIV(biodfe, (P->kfe_D + biodfe));
          double jmax_D = fmax(0., P->abio_P * (bct - P->dbct_D) * deffe_D) * P->jdiar;
          double limP_dop = QDIV(P->hdop * biodop, (k1p_P + biodop));
          double limP_po4 = QDIV(biopo4, (k1p_P + biopo4));
          double dopupt_flag = 0.5 + fsign(0.5, limP_dop - limP_po4);
          double limP = limP_dop * dopupt_flag + limP_po4 * (1. - dopupt_flag);
          double u_D = fmin(lc2, jmax_D * limP);
          double npp_D = fmax(0., u_D * biodiaz);
          double no3upt_D = (0.5 + 0.5 * tanh(biono3 - 5.)) * npp_D;
          SR(dopupt_D) = npp_D * dopupt_flag;
          npp_D = npp_D * (dopupt_flag * SF(V_DOP) + (1. - dopupt_flag) * SF(V_PO4)) * SF(V_DIN15);
          no3upt_D = no3upt_D * SF(V_NO3) * SF(V_DIN15);
          SR(npp_D) = npp_D;
          SR(no3upt_D) = no3upt_D;
          acc0 = acc0 + npp_D - no3upt_D;   // nfixout
          SR(rtdoc13) = CL13(QDIV(SB(V_DOC13), (SB(V_DON) * redctn)));
          SR(rtdiazc13) = CL13(QDIV(SB(V_DIAZC13), (biodiaz * redctn)));
          // 15N ratios of four organic pools (:2479-2500), taken over from warp 1
          SR(rtphytn15) = CL15(QDIV(SB(V_PHYTN15), SB(V_PHYT)));
          SR(rtdiatn15) = CL15(QDIV(SB(V_DIATN15), SB(V_DIAT)));
          SR(rtzoopn15) = CL15(QDIV(SB(V_ZOOPN15), SB(V_ZOOP)));
          SR(rtdiazn15) = CL15(QDIV(SB(V_DIAZN15), SB(V_DIAZ)));
        }
        WS_BAR();
        {
          const double excr = SR(excr), morpt = SR(morpt), npp = SR(npp), morpt_D = SR(morpt_D), npp_D = SR(npp_D);
          const double recy_don = SR(recy_don), nr_excr_D = SR(nr_excr_D), morp_D = SR(morp_D), remife = SR(remife), fecol = SR(fecol);
          const double morpt_Diat = SR(morpt_Diat), npp_Diat = SR(npp_Diat), sf = SR(sf), morp = SR(morp), morz = SR(morz);
          const double graz_Det = SR(graz_Det), expofe = SR(expofe), morp_Diat = SR(morp_Diat), sf_Diat = SR(sf_Diat);
          const double nr_excr_P = 0.0, nr_excr_detr = 0.0;
          const double o2flag = lc1;
          double feorgads = (P->kfeorg * SR(pw58) * SR(feprime)) * o2flag;
          feorgads = feorgads * SF(V_DFE);
          const double opldis = SR(opldis);
          const double oplpro = (morp_Diat + sf_Diat) * SR(sipr0) * SF(V_SIL) * (1.e-3);
          UPD(V_DFE, SB(V_DFE) + dtbio * (rfeton * (excr + (1. - dfrt) * morpt - npp + morpt_D - npp_D + recy_don + nr_excr_D + nr_excr_P +
                                                    nr_excr_detr + morp_D * (1. - rnd)) -
                                          feorgads + remife - fecol + rfeton * ((1. - dfrt) * morpt_Diat - npp_Diat)));
          UPD(V_DETRFE, SB(V_DETRFE) + dtbio * (rfeton * (sf + (1. - dfr) * morp + morp_D * rnd + morz - graz_Det) + feorgads +
                                                P->iscr * fecol - remife - expofe + SL(X_expofe) * dztrk +
                                                rfeton * (1. - dfr) * morp_Diat));
          UPD(V_SIL, SB(V_SIL) + dtbio * (opldis - oplpro));
          UPD(V_OPL, SB(V_OPL) + dtbio * (oplpro - opldis - SR(expoopl) + SL(X_expoopl) * dztrk));
        }
        {
          const double rtphytc13 = SR(rtphytc13), rtzoopc13 = SR(rtzoopc13), rtdiazc13 = SR(rtdiazc13), rtdetrc13 = SR(rtdetrc13);
          const double rtdiatc13 = SR(rtdiatc13), rtdoc13 = SR(rtdoc13), fcnpp = SR(fcnpp);
          const double rtcaco3c13 = SR(rtcaco3c13);
          const double morp = SR(morp), morp_Diat = SR(morp_Diat), morpt_Diat = SR(morpt_Diat), morpt = SR(morpt), remi = SR(remi);
          const double recy_don = SR(recy_don), npp = SR(npp), graz = SR(graz), morz = SR(morz), graz_Z = SR(graz_Z), excr = SR(excr);
          const double sf_Diat = SR(sf_Diat), sf_P = SR(sf_P), sf_Z = SR(sf_Z), sf_Det = SR(sf_Det), sf_D = SR(sf_D);
          const double graz_Det = SR(graz_Det), expo = SR(expo), morp_D = SR(morp_D);
          const double expocaco3 = SR(expocaco3);
          const double rc13impo = SL(X_rc13expo) * dztrk;
          UPD(V_DOC13, SB(V_DOC13) + dtbio * redctn *
                                         (dfr * rtphytc13 * morp + rtdiatc13 * (dfr * morp_Diat + dfrt * morpt_Diat) + rtphytc13 * dfrt * morpt +
                                          rtdetrc13 * pfr * remi - rtdoc13 * recy_don));
          UPD(V_PHYTC13, SB(V_PHYTC13) + dtbio * redctn * (fcnpp * npp - rtphytc13 * morp - rtphytc13 * graz - rtphytc13 * morpt));
          UPD(V_ZOOPC13, SB(V_ZOOPC13) + dtbio * redctn *
                                             (rtphytc13 * SR(dig_P) + rtdiatc13 * SR(dig_Diat) + rtzoopc13 * SR(dig_Z) + rtdetrc13 * SR(dig_Det) +
                                              rtdiazc13 * SR(dig_D) - rtzoopc13 * morz - rtzoopc13 * graz_Z - rtzoopc13 * excr));
          UPD(V_DETRC13, SB(V_DETRC13) + dtbio * redctn *
                                             (rtphytc13 * (1. - dfr) * morp + rtdiatc13 * (1. - dfr) * morp_Diat + rtdiatc13 * sf_Diat +
                                              rtphytc13 * sf_P + rtzoopc13 * sf_Z + rtdetrc13 * sf_Det + rtdiazc13 * sf_D + rtzoopc13 * morz -
                                              rtdetrc13 * remi - rtdetrc13 * graz_Det - rtdetrc13 * expo + rc13impo + rtdiazc13 * morp_D * rnd));
          // (DIAZC13, CACO3C13 and DIATC13 are advanced by warps 0, 1 and 2: load balance)
          acc1 = acc1 + rtdetrc13 * expo;            // rc13expoout
          acc2 = acc2 + rtcaco3c13 * expocaco3;      // rcaco3c13expoout
        }
        WS_BAR();
        }
      } break;
    }
#undef WS_LOOP_HEAD

    // ================= E1: increments as rates (:880-895); publish the sums =================
    // States the driver does not touch afterwards go straight to src; the others stay in B.
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int m = 8 * w + q;
      const double inc = (SB(m) - SC(m)) * rdtts;
      SB(m) = inc;
      const bool later = (m == V_NO3 || m == V_DIN15 || m == V_DFE || m == V_PO4 || m == V_DIC || m == V_DIC13 || m == V_SIL);
      if (!later && act) v.src[c + (long long)(ix[IX_SRC + m] - 1) * n3] = inc;
    }
    switch (w) {
      case 2: SL(S_expo) = acc0; SL(S_expo_phos) = acc1; SL(S_calpro) = acc2; SL(S_dissl) = acc3; SL(S_expocaco3) = acc4;
              SL(S_expoopl) = acc5; SL(S_expofe) = acc6; break;
      case 1: SL(S_rn15expo) = acc0; break;                                                   // role 5
      case 3: SL(S_nfix) = acc0; SL(S_rc13expo) = acc1; SL(S_rcaco3c13expo) = acc2; break;   // role 7
      default: break;
    }
    WS_BAR();

    // ================= E2: mobi_driver after mobi_src (:896-1400) =================
    {
      const double sgb = act ? v.sg_bathy[XIJK(i, j, k)] : 0.0;
      const double expo = SL(S_expo) * rnbio;   // export of this level before the sea-floor share is removed
      for (int rr = 0; rr < 2; rr++)
      switch (w + WS_WARPS * rr) {
        case 0: {
          // export chain -> import of the next level (:1000-1010, 1112-1120, 1280-1288); sea-floor phosphorus
          double expofe = SL(S_expofe) * rnbio, expocaco3 = SL(S_expocaco3) * rnbio, expoopl = SL(S_expoopl) * rnbio;
          double expo_phos = SL(S_expo_phos) * rnbio, rn15expo = SL(S_rn15expo) * rnbio, rc13expo = SL(S_rc13expo) * rnbio;
          double rcaco3c13expo = SL(S_rcaco3c13expo) * rnbio;
          double ex = expo;
          const double po4 = SB(V_PO4) + sgb * expo_phos;
          if (act) v.src[c + (long long)(ix[IX_SRC + V_PO4] - 1) * n3] = po4;
          rc13expo = rc13expo - sgb * rc13expo;
          ex = ex - sgb * ex;
          expo_phos = expo_phos - sgb * expo_phos;
          SL(X_expo) = ex * dztk;
          SL(X_expo_phos) = expo_phos * dztk;
          SL(X_rn15expo) = rn15expo;
          SL(X_rc13expo) = rc13expo * dztk;
          SL(X_rcaco3c13expo) = rcaco3c13expo * dztk;
          SL(X_expofe) = expofe * dztk;
          SL(X_expocaco3) = expocaco3 * dztk;
          SL(X_expoopl) = expoopl * dztk;
        } break;
        case 1: {
          // benthic denitrification, Bohlen et al. 2012 (:1035-1075); NO3 / DIN15 / ALK are closed in E3
          const double tno3 = SC(V_NO3), tdin15 = SC(V_DIN15);
          const double rn15expo = SL(S_rn15expo) * rnbio;
          double no3flag = 0.5 + fsign(0.5, tno3 - TRCMIN);
          double din15flag = 0.5 + fsign(0.5, tdin15 - TRCMIN);
          double lno3 = PRE(PR_LNO3A);
          double sg_bdeni = (0.06 + 0.19 * PRE(PR_P099)) * fmax(expo * sgb, TRCMIN) * redctn * 1.e3;
          sg_bdeni = fmin(sg_bdeni, sgb * expo);
          sg_bdeni = fmax(sg_bdeni, 0.);
          sg_bdeni = sg_bdeni * (0.5 + lno3) * no3flag * din15flag;
          double rno3 = QDIV(fmax(tdin15, TRCMIN * RN15STD / (1 + RN15STD)), fmax(tno3 - tdin15, TRCMIN * RN15STD / (1 + RN15STD)));
          rno3 = fmin(rno3, 2. * RN15STD);
          rno3 = fmax(rno3, RN15STD / 2.);
          double eps_bdeni = v.mobi_epsbd[k - 1];
          double bbdeni = rno3 - QDIV(eps_bdeni * rno3, 1000.);
          e3_no3 = SB(V_NO3) + sgb * expo - sg_bdeni;
          e3_din15 = SB(V_DIN15) + rn15expo * sgb * expo - QDIV(bbdeni, (1 + bbdeni)) * sg_bdeni;
          e3_c = c;
          e3_pending = act;
        } break;
        case 2: {
          // sea-floor carbon, O2, water-column denitrification, ALK (:1100-1110, 1301-1366), calcite (:1372-1400), c14 (tracer.F:848-867)
          const double tno3 = SC(V_NO3), tdin15 = SC(V_DIN15);
          const double nfix = SL(S_nfix), rdissl = SL(S_dissl) * rnbio, rcalpro = SL(S_calpro) * rnbio;
          const double rexpocaco3 = SL(S_expocaco3) * rnbio;
          double no3flag = 0.5 + fsign(0.5, tno3 - TRCMIN);
          double din15flag = 0.5 + fsign(0.5, tdin15 - TRCMIN);
          double dic = SB(V_DIC) + sgb * expo * redctn;
          const double dic_npzd_sms = dic;
          double src_alk = -dic * P->redntc * 1.e-3;
          // the benthic term of warp 1 enters ALK: same expressions as there
          double bdeni;
          {
            double lno3a = PRE(PR_LNO3A);
            double sg_bdeni = (0.06 + 0.19 * PRE(PR_P099)) * fmax(expo * sgb, TRCMIN) * redctn * 1.e3;
            sg_bdeni = fmin(sg_bdeni, sgb * expo);
            sg_bdeni = fmax(sg_bdeni, 0.);
            bdeni = sg_bdeni * (0.5 + lno3a) * no3flag * din15flag;
          }
          double rno3 = QDIV(fmax(tdin15, TRCMIN * RN15STD / (1 + RN15STD)), fmax(tno3 - tdin15, TRCMIN * RN15STD / (1 + RN15STD)));
          rno3 = fmin(rno3, 2. * RN15STD);
          rno3 = fmax(rno3, RN15STD / 2.);
          const double fo2 = PRE(PR_FO2);
          double so2 = dic_npzd_sms * P->redotc + nfix * rnbio * 1.25e-3;
          double lno3 = PRE(PR_LNO3B);
          double wcdeni = 800. * no3flag * so2 * (1.0 - fo2) * (0.5 + lno3) * din15flag;
          wcdeni = fmax(wcdeni, 0.);
          double uno3 = QDIV(wcdeni * v.c2dtts, tno3);
          uno3 = fmin(uno3, 0.999);
          uno3 = fmax(uno3, TRCMIN);
          double bwcdeni = rno3 + QDIV(QDIV(P->eps_wcdeni * (1 - uno3), uno3) * log(1 - uno3) * rno3, 1000.);
          SL(L_wcdeni) = wcdeni;
          SL(L_bwfrac) = (QDIV(bwcdeni, (1 + bwcdeni)));
          src_alk = src_alk + wcdeni * 1.e-3;
          src_alk = src_alk + bdeni * 1.e-3;
          src_alk = src_alk - nfix * rnbio * 1.e-3;
          const double src_o2 = -so2 * fo2;
          if (k < kmx) {
            dic = dic + rdissl * 1.e-3 - rcalpro * 1.e-3;
            src_alk = src_alk + 2. * rdissl * 1.e-3 - 2. * rcalpro * 1.e-3;
          } else {
            dic = dic + rdissl * 1.e-3 - rcalpro * 1.e-3 + rexpocaco3 * 1.e-3;
            src_alk = src_alk + 2. * rdissl * 1.e-3 - 2. * rcalpro * 1.e-3 + 2. * rexpocaco3 * 1.e-3;
          }
          if (act) {
            const double c14_in = v.t_m1[c + (long long)(ix[IX_IC14] - 1) * n3];
            v.src[c + (long long)(ix[IX_SRC + V_DIC] - 1) * n3] = dic;
            v.src[c + (long long)(ix[IX_ISALK] - 1) * n3] = src_alk;
            v.src[c + (long long)(ix[IX_ISO2] - 1) * n3] = src_o2;
            v.src[c + (long long)(ix[IX_ISC14] - 1) * n3] = dic * RC14STD - 3.836e-12 * c14_in;
          }
        } break;
        case 3: {
          // sediment carbon oxidation (Flogel 2011 / Somes 2021), iron release (Dale 2015) (:1076-1099); dust and
          // hydrothermal iron (tracer.F:536-545)
          const double o2_in = act ? v.t_m1[c + (long long)(ix[IX_IO2] - 1) * n3] * 1000. : 1.0;
          double coxdepth = fmin(fmax(v.zt[k - 1], 50000.), 150000.);
          double oblinc = -1.26e-6 * coxdepth + 0.203;
          double obexpc = -6.e-7 * coxdepth + 1.14;
          double nburial = QDIV((oblinc * pow((QDIV(expo * sgb * dztk, 100) * 86400. * 365. * redctn * 1000.), obexpc)), (QDIV(86400. * 365. * dztk, 100) * redctn * 1000.));
          double coxsed = expo * sgb - nburial;
          double fesed = QDIV(85. * tanh(QDIV(QDIV(coxsed * redctn * 1000 * dztk, 100) * 86400., o2_in)), (QDIV(dztk, 100) * 86400 * 1000));
          double dfe = SB(V_DFE) + fesed;
          if (act) {
            if (k == 1) dfe = dfe + QDIV(v.fe_atmdep[X2(i, j) + (long long)(mi - 1) * v.n2] * 1000, (QDIV(v.dzt[0], 100.)));
            dfe = dfe + v.fe_hydr[XIJK(i, j, k)];
            v.src[c + (long long)(ix[IX_SRC + V_DFE] - 1) * n3] = dfe;
          }
        } break;
        case 4: {
          // 13C of the sea-floor remineralisation and of calcite dissolution / production (:1104, 1258-1276, 1372-1400)
          const double rc13expo = SL(S_rc13expo) * rnbio, rdissl = SL(S_dissl) * rnbio, rcalpro = SL(S_calpro) * rnbio;
          const double rexpocaco3 = SL(S_expocaco3) * rnbio;
          const double dic_in = act ? v.t_m1[c + (long long)(ix[IX_TR + V_DIC] - 1) * n3] : 1.0;
          double dic13 = SB(V_DIC13) + rc13expo * sgb * redctn;
          double rtdic13 = QDIV(fmax(SC(V_DIC13), TRCMIN * RC13STD / (1 + RC13STD)), fmax(dic_in, TRCMIN));
          rtdic13 = fmin(rtdic13, 2. * RC13STD / (1 + RC13STD));
          rtdic13 = fmax(rtdic13, 0.5 * RC13STD / (1 + RC13STD));
          double rtcaco3c13 = QDIV(fmax(SC(V_CACO3C13), TRCMIN * RC13STD / (1 + RC13STD)), fmax(SC(V_CACO3), TRCMIN));
          rtcaco3c13 = fmin(rtcaco3c13, 2. * RC13STD / (1 + RC13STD));
          rtcaco3c13 = fmax(rtcaco3c13, 0.5 * RC13STD / (1 + RC13STD));
          if (k < kmx)
            dic13 = dic13 + rdissl * 1.e-3 * rtcaco3c13 - rcalpro * 1.e-3 * rtdic13;
          else
            dic13 = dic13 + rdissl * 1.e-3 * rtcaco3c13 - rcalpro * 1.e-3 * rtdic13 + rexpocaco3 * 1.e-3 * rtcaco3c13;
          if (act) v.src[c + (long long)(ix[IX_SRC + V_DIC13] - 1) * n3] = dic13;
        } break;
        case 5: {
          // opal reaching the sea floor dissolves there (:1478)
          double sil = SB(V_SIL);
          if (k >= kmx) sil = sil + SL(S_expoopl) * rnbio;
          if (act) v.src[c + (long long)(ix[IX_SRC + V_SIL] - 1) * n3] = sil;
        } break;
        default: break;
      }
    }
    WS_BAR();
#undef PRE
  }  // levels
  if (w == 1 && e3_pending) {
    const double wcdeni = SL(L_wcdeni), bwfrac = SL(L_bwfrac);
    v.src[e3_c + (long long)(ix[IX_SRC + V_NO3] - 1) * n3] = e3_no3 - wcdeni;
    v.src[e3_c + (long long)(ix[IX_SRC + V_DIN15] - 1) * n3] = e3_din15 - bwfrac * wcdeni;
  }
#undef SB
#undef SF
#undef SC
#undef SR
#undef SL
#undef UPD
}

// ------------------------------------------------------------------------------------
// Air-sea gas exchange: the flux loop of gasbc (09/common/gasbc.F:62-69, 76-80, 148-266), one thread per surface point.
// Carbon (DIC, DI13C, 14C) and oxygen fluxes from the segment-mean surface state set_sbc left in sbc, with co2calc_SWS
// at the surface (depth 0, atmpres 1); land points take the land carbon fluxes; then the cyclic boundary of the four
// flux slots.  (the only other caller of co2calc_SWS.)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_gasbc(const DevView v, const uvic_b200_gasbc_par gp) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ni = v.imt - 2;
  const int jfirst = max(2, v.jbase), jlast = min(v.jmt - 1, v.jbase + v.jl - 1);
  if (idx >= (long long)ni * (jlast - jfirst + 1)) return;
  const int i = (int)(idx % ni) + 2;
  const int j = (int)(idx / ni) + jfirst;
  const long long c2 = X2(i, j);
  const long long n2 = v.n2;
#define GS(slot) v.sbc[c2 + (long long)((slot)-1) * n2]
  const double rc13std = RC13STD, rc14std = RC14STD, C2K = 273.15;
  const double ak = 0.99915, aaqg = 0.998764;
  const double r13a = (gp.dc13ccn * 0.001 + 1.) * rc13std;
  const double batmc13 = ak * aaqg * r13a;
  double xconv = 33.7 / 3.6e+05;
  xconv = xconv * 0.75;
  double fdic = 0.0, fdic13 = 0.0, fc14 = 0.0, fo2 = 0.0;
  bool wr_c = false, wr_o = false;
  if (v.kmt[c2] > 0) {
    double sss = 1000.0 * GS(gp.isss) + 35.0;
    double sst = GS(gp.isst);
    sst = fmin(35., fmax(sst, -2.));
    sss = fmin(45., fmax(sss, 0.));
    const double ao = 1. - v.aice[c2];
    const double dic = GS(gp.issdic);
    double co2star, Omega_c, dco2star;
    co2calc_sws<true>(sst, sss, dic, GS(gp.issalk), 0.0, co2star, Omega_c, gp.co2ccn, 1.0, &dco2star);
    const double scco2 = 2073.1 - 125.62 * sst + 3.6276 * (sst * sst) - 0.043219 * (sst * sst * sst);
    const double ws = GS(gp.iws) * 0.01;
    const double ws2 = ws * ws;
    const double piston_vel = ao * xconv * ws2 * pow(scco2 / 660., -0.5);
    fdic = piston_vel * dco2star;
    const double adicg = 1.01051 - 1.05e-4 * sst;
    const double dic13 = GS(gp.issdic13);
    double r13dic = dic13 / (dic - dic13);
    r13dic = fmin(r13dic, 2. * rc13std);
    r13dic = fmax(r13dic, 0.5 * rc13std);
    const double bdic13 = ak * aaqg * r13dic / adicg;
    fdic13 = piston_vel * ((batmc13 / (1 + batmc13)) * (dco2star + co2star) - (bdic13 / (1 + bdic13)) * co2star);
    fc14 = piston_vel * ((dco2star + co2star) * (1 + gp.dc14ccn * 0.001) * rc14std - co2star * GS(gp.issc14) / dic);
    const double sco2 = 1638.0 - 81.83 * sst + 1.483 * (sst * sst) - 0.008004 * (sst * sst * sst);
    const double piston_o2 = ao * xconv * ws2 * pow(sco2 / 660.0, -0.5);
    const double f1 = log((298.15 - sst) / (C2K + sst));
    const double f2 = f1 * f1, f3 = f2 * f1, f4 = f3 * f1, f5 = f4 * f1;
    double o2sat = exp(2.00907 + 3.22014 * f1 + 4.05010 * f2 + 4.94457 * f3 - 2.56847E-1 * f4 + 3.88767 * f5 +
                       sss * (-6.24523e-3 - 7.37614e-3 * f1 - 1.03410e-2 * f2 - 8.17083E-3 * f3) - 4.88682E-7 * sss * sss);
    o2sat = o2sat / 22391.6 * 1000.0;
    fo2 = piston_o2 * (o2sat - GS(gp.isso2));
    wr_c = wr_o = true;
  } else if (gp.inpp > 0) {
    const double f = GS(gp.inpp) - GS(gp.isr) - GS(gp.iburn);
    fdic = f * 0.1 / 12.e-6;
    fdic13 = f * 0.1 / 12.e-6 * rc13std / (1 + rc13std);
    fc14 = f * rc14std * 0.1 / 12.e-6;
    wr_c = true;
  }
  // write, with the cyclic copies (setbcx, :248-266): a(1) = a(imt-1), a(imt) = a(2)
  const int wrap = (i == 2) ? (v.imt - 2) : ((i == v.imt - 1) ? -(v.imt - 2) : 0);
  if (wr_c) {
    GS(gp.idicflx) = fdic; GS(gp.idic13flx) = fdic13; GS(gp.ic14flx) = fc14;
  }
  if (wr_o) GS(gp.io2flx) = fo2;
  if (wrap) {
    // the boundary copy takes whatever the interior point holds now (written above or left untouched)
    v.sbc[c2 + wrap + (long long)(gp.idicflx - 1) * n2] = GS(gp.idicflx);
    v.sbc[c2 + wrap + (long long)(gp.idic13flx - 1) * n2] = GS(gp.idic13flx);
    v.sbc[c2 + wrap + (long long)(gp.ic14flx - 1) * n2] = GS(gp.ic14flx);
    v.sbc[c2 + wrap + (long long)(gp.io2flx - 1) * n2] = GS(gp.io2flx);
  }
#undef GS
}

void launch_gasbc(uvic_b200_ctx *c, const uvic_b200_gasbc_par *gp) {
  DevView &v = c->v;
  const int jfirst = std::max(2, v.jbase), jlast = std::min(v.jmt - 1, v.jbase + v.jl - 1);
  const long long n = (long long)(v.imt - 2) * (jlast - jfirst + 1);
  KLAUNCH("k_gasbc", k_gasbc, cdiv(n, 128), 128, v, *gp);
}

static int mobi_ws_mode() {
  // UVIC_B200_MOBI_WS = 0 (one thread per column), 1 (warp specialised), unset = by column count
  const char *e = getenv("UVIC_B200_MOBI_WS");
  return e ? atoi(e) : -1;
}

void launch_mobi(uvic_b200_ctx *c, const DevView &v, const uvic_b200_stepinfo *si) {
  // month index and declination (09/mom/tracer.F:310-343)
  double yrtime = fmod(si->relyr, 1.);
  int mi = 12;
  for (int m = 1; m <= 12; m++)
    if (yrtime <= m / 12.) { mi = m; break; }
  const double pi = atan(1.0) * 4.0;
  double declin = sin((fmod(si->relyr, 1.) - 0.22) * 2. * pi) * 0.4;
  int nbio = (int)(v.c2dtts / c->mobi_dtnpzd);
  double dtbio = v.c2dtts / nbio;
  double rdtts = 1. / v.c2dtts;
  double rnbio = 1. / nbio;
  long long ncell = (long long)(v.imt - 2) * v.km * (v.jhi - v.jlo + 1);
  long long ncol = (long long)(v.imt - 2) * (v.jhi - v.jlo + 1);
  if (v.mobi_ncols <= 0) return;
  KLAUNCH("k_mobi_light", k_mobi_light, cdiv(ncol, 128), 128, v, declin);
  if (cdiv(ncell, 128) >= 148LL * MOBI_CELL_MINB * 3) {   // at least three full waves at the high occupancy
    KLAUNCH("k_mobi_cell", k_mobi_cell<MOBI_CELL_MINB>, cdiv(ncell, 128), 128, v);
  } else {
    KLAUNCH("k_mobi_cell", k_mobi_cell<4>, cdiv(ncell, 128), 128, v);
  }
  const int ngroups = (v.mobi_ncols + 31) / 32;
  int mode = mobi_ws_mode();
  // Few column groups: the chain latency of a column dominates and the warp-specialised kernel wins; many groups
  // fill the machine either way and the one-thread-per-column kernel issues fewer instructions in total.
  bool ws = (mode < 0) ? (ngroups <= 148 * 8) : (mode != 0);
  if (ws) {
    ensure_dyn_smem(c, (const void *)k_mobi_ws<1>, sizeof(WsSm));
    ensure_dyn_smem(c, (const void *)k_mobi_ws<2>, 2 * sizeof(WsSm));
    ensure_dyn_smem(c, (const void *)k_mobi_ws<4>, 4 * sizeof(WsSm));
    // G = 4 confines MOBI to a quarter of the SMs it would otherwise touch (48 of 148 on the 100x100 grid), leaving the
    // rest to the kernels of the main stream it overlaps with; measured best for the whole step.
    int G = 4;
    if (const char *e = getenv("UVIC_B200_MOBI_WS_G")) G = atoi(e);
    ProfScope ps_(c, "k_mobi_ws");
    if (G >= 4)
      k_mobi_ws<4><<<(ngroups + 3) / 4, 32 * WS_WARPS * 4, 4 * sizeof(WsSm), c->stream>>>(v, mi, nbio, dtbio, rdtts, rnbio);
    else if (G == 2)
      k_mobi_ws<2><<<(ngroups + 1) / 2, 32 * WS_WARPS * 2, 2 * sizeof(WsSm), c->stream>>>(v, mi, nbio, dtbio, rdtts, rnbio);
    else
      k_mobi_ws<1><<<ngroups, 32 * WS_WARPS, sizeof(WsSm), c->stream>>>(v, mi, nbio, dtbio, rdtts, rnbio);
  } else {
    KLAUNCH("k_mobi_column", k_mobi_column, ngroups, 32, v, mi, nbio, dtbio, rdtts, rnbio);
  }
}
