// k_mobi.cu -- MOBI biogeochemical source terms on the device (options of run/mk.in:
// O_mobi, O_mobi_alk/_caco3/_o2/_nitrogen/_nitrogen_15/_silicon/_iron, O_carbon,
// O_carbon_13, O_carbon_14).
//
//   k_mobi_co2     co2calc_SWS + drtsafe + ta_iter_SWS (09/common/co2calc.F), one thread
//                  per ocean cell.  The carbonate solve depends only on T, S, DIC, ALK of
//                  the cell, so it is lifted out of the column loop of mobi_driver
//                  (09/mom/mobi.F:768-772) and run at full cell parallelism; it hands
//                  CO2* and Omega_calcite to the column kernel.
//   k_mobi_column  the column prologue of tracer (09/mom/tracer.F:310-545), mobi_driver
//                  (09/mom/mobi.F:519-1483) and mobi_src (:1485-3313), one thread per
//                  water column, i fastest so the 37 tracer reads of a level coalesce.
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
  double c = 1.0 + q.st / q.ks + q.ft / q.kf;
  double a = x3 + q.k1p * x2 + k12p * x + k123p;
  double a2 = a * a;
  double da = 3.0 * x2 + 2.0 * q.k1p * x + k12p;
  double b = x2 + q.k1 * x + k12;
  double b2 = b * b;
  double db = 2.0 * x + q.k1;
  double xc = x / c;
  double hs = 1.0 + q.ks / xc, hf = 1.0 + q.kf / xc;
  double bb = 1.0 + x / q.kb, ss = 1.0 + x / q.ksi;
  fn = q.k1 * x * q.dic / b + 2.0 * q.dic * k12 / b + q.bt / bb + q.kw / x + q.pt * k12p * x / a + 2.0 * q.pt * k123p / a + q.sit / ss -
       x / c - q.st / hs - q.ft / hf - q.pt * x3 / a - q.ta;
  df = ((q.k1 * q.dic * b) - q.k1 * x * q.dic * db) / b2 - 2.0 * q.dic * k12 * db / b2 - q.bt / q.kb / (bb * bb) - q.kw / x2 +
       (q.pt * k12p * (a - x * da)) / a2 - 2.0 * q.pt * k123p * da / a2 - q.sit / q.ksi / (ss * ss) - 1.0 / c -
       q.st * (1.0 / (hs * hs)) * (q.ks * c / x2) - q.ft * (1.0 / (hf * hf)) * (q.kf * c / x2) - q.pt * x2 * (3.0 * a - x * da) / a2;
}

// 09/common/co2calc.F:401-454 (Numerical Recipes rtsafe, error trapping removed)
__device__ double drtsafe(const Carb &q, double x1, double x2, double xacc) {
  double fl, fh, df, f, xl, xh;
  ta_iter(q, x1, fl, df);
  ta_iter(q, x2, fh, df);
  if (fl < 0.0) {
    xl = x1;
    xh = x2;
  } else {
    xh = x1;
    xl = x2;
  }
  double r = 0.5 * (x1 + x2);
  double dxold = fabs(x2 - x1);
  double dx = dxold;
  ta_iter(q, r, f, df);
  for (int it = 1; it <= 100; it++) {
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
    ta_iter(q, r, f, df);
    if (f < 0.0)
      xl = r;
    else
      xh = r;
  }
  return r;
}

// 09/common/co2calc.F:1-400; returns CO2* (mol m-3) and Omega_calcite, the two outputs MOBI uses
__device__ void co2calc_sws(double t, double s, double dic_in, double ta_in, double depth, double &co2star_out, double &omega_c) {
  Carb q;
  const double permil = 1.0 / 1024.5;
  q.pt = 0.5125e-3 * permil;   // hard-wired phosphate and silicate (:113-114)
  q.sit = 7.6875e-03 * permil;
  q.ta = ta_in * permil;
  q.dic = dic_in * permil;
  const double pres = depth * 0.1;
  double tk = 273.15 + t;
  double tk100 = tk / 100.0;
  double tk1002 = tk100 * tk100;
  double invtk = 1.0 / tk;
  double dlogtk = log(tk);
  double is = 19.924 * s / (1000. - 1.005 * s);
  double is2 = is * is;
  double sqrtis = sqrt(is);
  double s2 = s * s;
  double t2 = t * t;
  double sqrts = sqrt(s);
  double s15 = pow(s, 1.5);
  double scl = s / 1.80655;
  double pitkR = pres / tk / 83.15;
  double p2itkR = pres * pitkR;
  q.bt = 0.000232 * scl / 10.811;
  q.st = 0.14 * scl / 96.062;
  q.ft = 0.000067 * scl / 18.9984;
  (void)tk1002;

  q.k1 = pow(10., (-1. * (3670.7 * invtk - 62.008 + 9.7944 * dlogtk - 0.0118 * s + 0.000116 * s2))) *
         exp((25.5 - 0.1271 * t) * pitkR + 0.5 * (-3.08e-3 + 8.77e-5 * t) * p2itkR);
  q.k2 = pow(10., (-1 * (1394.7 * invtk + 4.777 - 0.0184 * s + 0.000118 * s2))) *
         exp((15.82 + 0.0219 * t) * pitkR + 0.5 * (1.13e-3 - 1.475e-4 * t) * p2itkR);
  q.k1p = exp(-4576.752 * invtk + 115.540 - 18.453 * dlogtk + (-106.736 * invtk + 0.69171) * sqrts + (-0.65643 * invtk - 0.01844) * s) *
          exp((14.51 - 0.1211 * t + 3.21e-4 * t2) * pitkR + 0.5 * (-2.67e-3 + 4.27e-5 * t) * p2itkR);
  q.k2p = exp(-8814.715 * invtk + 172.1033 - 27.927 * dlogtk + (-160.340 * invtk + 1.3566) * sqrts + (0.37335 * invtk - 0.05778) * s) *
          exp((23.12 - 0.1758 * t + 2.647e-3 * t2) * pitkR + 0.5 * (-5.15e-3 + 9.0e-5 * t) * p2itkR);
  q.k3p = exp(-3070.75 * invtk - 18.126 + (17.27039 * invtk + 2.81197) * sqrts + (-44.99486 * invtk - 0.09984) * s) *
          exp((26.57 - 0.202 * t + 3.042e-3 * t2) * pitkR + 0.5 * (-4.08e-3 + 7.14e-5 * t) * p2itkR);
  q.ksi = exp(-8904.2 * invtk + 117.400 - 19.334 * dlogtk + (-458.79 * invtk + 3.5913) * sqrtis + (188.74 * invtk - 1.5998) * is +
              (-12.1652 * invtk + 0.07871) * is2 + log(1.0 - 0.001005 * s)) *
          exp((29.48 - 0.1622 * t - 2.608e-3 * t2) * pitkR + 0.5 * (-2.84e-3) * p2itkR);
  q.kw = exp(-13847.26 * invtk + 148.9802 - 23.6521 * dlogtk + (118.67 * invtk - 5.977 + 1.0495 * dlogtk) * sqrts - 0.01615 * s) *
         exp((20.02 - 0.1119 * t + 1.409e-3 * t2) * pitkR + 0.5 * (-5.13e-3 + 7.94e-5 * t) * p2itkR);
  q.ks = exp(-4276.1 * invtk + 141.328 - 23.093 * dlogtk + (-13856 * invtk + 324.57 - 47.986 * dlogtk) * sqrtis +
             (35474 * invtk - 771.54 + 114.723 * dlogtk) * is - 2698 * invtk * pow(is, 1.5) + 1776 * invtk * is2 +
             log(1.0 - 0.001005 * s)) *
         exp((18.03 - .0466 * t - 3.16e-4 * t2) * pitkR + 0.5 * (-4.53e-3 + 9.0e-5 * t) * p2itkR);
  q.kf = exp(1590.2 * invtk - 12.641 + 1.525 * sqrtis + log(1.0 - 0.001005 * s)) *
         exp((9.78 + 9.0e-3 * t + 9.42e-4 * t2) * pitkR + 0.5 * (-3.91e-3 + 5.4e-5 * t) * p2itkR);
  q.kb = exp((-8966.90 - 2890.53 * sqrts - 77.942 * s + 1.728 * s15 - 0.0996 * s2) * invtk +
             (148.0248 + 137.1942 * sqrts + 1.62142 * s) + (-24.4344 - 25.085 * sqrts - 0.2474 * s) * dlogtk + 0.053105 * sqrts * tk +
             log((1 + (q.st / q.ks) + (q.ft / q.kf)) / (1 + (q.st / q.ks)))) *
         exp((29.48 - 0.1622 * t - 2.608e-3 * t2) * pitkR + 0.5 * (-2.84e-3) * p2itkR);

  // [H+] on the seawater scale in [1e-10, 1e-6], xacc = 1e-10 (:343-346)
  double x1 = pow(10.0, -6.), x2 = pow(10.0, -10.);
  double hSWS = drtsafe(q, x1, x2, 1.e-10);
  double hSWS2 = hSWS * hSWS;
  double co2star = q.dic * hSWS2 / (hSWS2 + q.k1 * hSWS + q.k1 * q.k2);
  double CO3 = q.k1 * q.k2 * co2star / hSWS2;
  // calcite solubility product with pressure dependence (:360-388)
  double Kspc = exp(-395.8293 + (6537.773 / tk) + 71.595 * log(tk) - 0.17959 * tk +
                    (-1.78938 + (410.64 / tk) + 0.0065453 * tk) * sqrt(s) - 0.17755 * s + 0.0094979 * s15);
  double DVc = -65.28 + 0.397 * t - 0.005155 * (t * t) + (19.816 - 0.0441 * t - 0.00017 * (t * t)) * sqrt(s / 35.);
  double DK = 0.01847 + 0.0001956 * t - 0.000002212 * (t * t) + (-0.03217 - 0.0000711 * t + 0.000002212) * sqrt(s / 35.);
  Kspc = Kspc * exp(-DVc * pitkR + 0.5 * DK * p2itkR);
  const double Ca = 10.28E-3;
  omega_c = Ca * CO3 / Kspc;
  co2star_out = co2star / permil;
}

__global__ void __launch_bounds__(128) k_mobi_co2(const DevView v) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int ni = v.imt - 2, nrow = v.jhi - v.jlo + 1;
  if (idx >= (long long)ni * v.km * nrow) return;
  int i = (int)(idx % ni) + 2;
  long long r = idx / ni;
  int k = (int)(r % v.km) + 1;
  int j = (int)(r / v.km) + v.jlo;
  if (k > v.kmt[X2(i, j)]) return;
  const int *ix = v.mobi_idx;
  long long c = X3(i, k, j);
  double t_in = v.t_m1[c + (long long)(ix[IX_ITEMP] - 1) * v.n3];
  double s_in = 1.e3 * v.t_m1[c + (long long)(ix[IX_ISALT] - 1) * v.n3] + 35.0;
  double dic_in = v.t_m1[c + (long long)(ix[IX_TR + V_DIC] - 1) * v.n3];
  double alk_in = v.t_m1[c + (long long)(ix[IX_IALK] - 1) * v.n3];
  double depth = v.zt[k - 1] / 100.;
  double co2star, omega_c;
  co2calc_sws(t_in, s_in, dic_in, alk_in, depth, co2star, omega_c);
  v.co2_star[c] = co2star;
  v.co2_omega[c] = omega_c;
}

// ------------------------------------------------------------------------------------
// ecosystem ODE, 09/mom/mobi.F:1485-3313.  b[] holds the 32 state variables (clipped in
// place like bioin, :1894); on return b[] holds the increments bioout.
// ------------------------------------------------------------------------------------
struct SrcIO {
  // in
  double gl, bct, impo, dzt, impo_phos, dayfrac, wwd, nud, impocaco3, wwc, dissk1, impoopl, wwo, opl_disk1, nudop, nudon, bctz;
  double rn15impo, rc13impo, ac13b, rcaco3c13impo, impofe, o2, aou, capr;
  // out
  double nfix, expo, expo_phos, calpro, dissl, expocaco3, expoopl, rn15expo, rc13expo, rcaco3c13expo, expofe;
};

#define CL15(x) fmax(fmin((x), 2. * RN15STD / (1 + RN15STD)), RN15STD / (1 + RN15STD) / 2.)
#define CL13(x) fmax(fmin((x), 2. * RC13STD / (1 + RC13STD)), 0.5 * RC13STD / (1 + RC13STD))

__device__ __forceinline__ void mobi_src(const MobiPar *__restrict__ P, int nbio, double dtbio, double (&b)[MOBI_NVAR], double (&clip)[MOBI_NVAR], SrcIO &io) {
  const double gamma1 = P->gamma1, redptn = P->redptn, redctn = P->redctn, redntp = P->redntp, diazntp = P->diazntp;
  const double diazptn = P->diazptn, dfr = P->dfr, pfr = P->pfr, dfrt = P->dfrt, geZ = P->geZ, rfeton = P->rfeton;
  const double bct = io.bct, dzt = io.dzt, gl = io.gl;
  // ratios from the raw inputs (:1781-1784)
  double ptn_P = b[V_PHYT_PHOS] / b[V_PHYT];
  double ptn_detr = b[V_DETR_PHOS] / b[V_DETR];
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
  // iron-dependent light harvesting, pre-loop values (:1928-1952)
  double p1 = fmin(b[V_PHYT], P->pmax), p2 = fmax(0.0, b[V_PHYT] - P->pmax);
  double kfevar = (P->kfemin * p1 + P->kfemax * p2) / (p1 + p2);
  double deffe = b[V_DFE] / (kfevar + b[V_DFE]);
  double thetamax = P->thetamaxlo + (P->thetamaxhi - P->thetamaxlo) * deffe;
  double alpha_O = P->alphamin + (P->alphamax - P->alphamin) * deffe;
  double gl_O = gl * thetamax * alpha_O;
  p1 = fmin(b[V_DIAT], P->pmax_Diat);
  p2 = fmax(0.0, b[V_DIAT] - P->pmax_Diat);
  double kfevar_Diat = (P->kfemin_Diat * p1 + P->kfemax_Diat * p2) / (p1 + p2);
  double deffe_Diat = b[V_DFE] / (kfevar_Diat + b[V_DFE]);
  double gl_Diat = gl * (P->thetamaxlo + (P->thetamaxhi - P->thetamaxlo) * deffe_Diat) * (P->alphamin + (P->alphamax - P->alphamin) * deffe_Diat);
  double deffe_D = b[V_DFE] / (P->kfe_D + b[V_DFE]);
  double gl_D = gl * (P->thetamaxlo + (P->thetamaxhi - P->thetamaxlo) * deffe_D) * (P->alphamin + (P->alphamax - P->alphamin) * deffe_D);
  // photosynthesis after Evans & Parslow (:1954-2003)
  double kirr = -P->kw - P->kc * (b[V_PHYT] + b[V_DIAZ] + b[V_DIAT]) - P->kc_c * b[V_CACO3];
  double f1 = exp(kirr * dzt);
  double avej, avej_D, avej_Diat;
  {
    double jmax = P->abio_P * bct * deffe;
    double gd = jmax * io.dayfrac;
    double u1 = fmax(gl_O / gd, 1.e-6), u2 = u1 * f1;
    double phi1 = log(u1 + sqrt(1. + u1 * u1)) - (sqrt(1. + u1 * u1) - 1.) / u1;
    double phi2 = log(u2 + sqrt(1. + u2 * u2)) - (sqrt(1. + u2 * u2) - 1.) / u2;
    avej = gd * (phi1 - phi2) / (-kirr * dzt);
    double jmax_D = fmax(0., P->abio_P * (bct - P->dbct_D) * deffe_D) * P->jdiar;
    double gd_D = fmax(1.e-14, jmax_D * io.dayfrac);
    u1 = fmax(gl_D / gd_D, 1.e-6);
    u2 = u1 * f1;
    phi1 = log(u1 + sqrt(1. + u1 * u1)) - (sqrt(1. + u1 * u1) - 1.) / u1;
    phi2 = log(u2 + sqrt(1. + u2 * u2)) - (sqrt(1. + u2 * u2) - 1.) / u2;
    avej_D = gd_D * (phi1 - phi2) / (-kirr * dzt);
    double jmax_Diat = P->abiodiat * bct * deffe_Diat;
    double gd_Diat = jmax_Diat * io.dayfrac;
    u1 = fmax(gl_Diat / gd_Diat, 1.e-6);
    u2 = u1 * f1;
    phi1 = log(u1 + sqrt(1. + u1 * u1)) - (sqrt(1. + u1 * u1) - 1.) / u1;
    phi2 = log(u2 + sqrt(1. + u2 * u2)) - (sqrt(1. + u2 * u2) - 1.) / u2;
    avej_Diat = gd_Diat * (phi1 - phi2) / (-kirr * dzt);
  }
  const double gmax = P->gbio * io.bctz;
  const double nupt = P->nupt0 * bct, nupt_D = P->nupt0_D * bct, nudt = P->nudt0 * bct;
  double nfixout = 0.0, expoout = 0.0, expo_phosout = 0.0, rn15expoout = 0.0, rc13expoout = 0.0, rcaco3c13expoout = 0.0;
  double calproout = 0.0, disslout = 0.0, expocaco3out = 0.0, expooplout = 0.0, expofeout = 0.0;
  const double o2flag = tanh(fmax(io.o2, 0.));  // (:2267) loop invariant
  const double aou8 = pow(fmax(io.aou, 40.), 0.8);

  for (int n = 1; n <= nbio; n++) {
    const double biopo4 = b[V_PO4], biophyt = b[V_PHYT], biophyt_phos = b[V_PHYT_PHOS], biozoop = b[V_ZOOP], biodetr = b[V_DETR];
    const double biodetr_phos = b[V_DETR_PHOS], biodic = b[V_DIC], biodop = b[V_DOP], biono3 = b[V_NO3], biodon = b[V_DON];
    const double biodiaz = b[V_DIAZ], biocaco3 = b[V_CACO3], biodiat = b[V_DIAT], biosil = b[V_SIL], bioopl = b[V_OPL];
    const double biodfe = b[V_DFE], biodetrfe = b[V_DETRFE];
    // half-saturation constants and maximum rates (:2150-2166)
    p1 = fmin(biophyt, P->pmax);
    p2 = fmax(0.0, biophyt - P->pmax);
    double k1n = (P->knmin * p1 + P->knmax * p2) / (p1 + p2);
    double k1p_P = k1n * ptn_P;
    kfevar = (P->kfemin * p1 + P->kfemax * p2) / (p1 + p2);
    deffe = biodfe / (kfevar + biodfe);
    double jmax = P->abio_P * bct * deffe;
    p1 = fmin(biodiat, P->pmax_Diat);
    p2 = fmax(0.0, biodiat - P->pmax_Diat);
    kfevar_Diat = (P->kfemin_Diat * p1 + P->kfemax_Diat * p2) / (p1 + p2);
    double k1n_Diat = (P->knmin_Diat * p1 + P->knmax_Diat * p2) / (p1 + p2);
    double k1p_Diat = k1n_Diat * redptn;
    deffe_Diat = biodfe / (kfevar_Diat + biodfe);
    double jmax_Diat = P->abiodiat * bct * deffe_Diat;
    deffe_D = biodfe / (P->kfe_D + biodfe);
    double jmax_D = fmax(0., P->abio_P * (bct - P->dbct_D) * deffe_D) * P->jdiar;
    // growth rates (:2168-2206)
    double limP_dop = P->hdop * biodop / (k1p_P + biodop);
    double limP_po4 = biopo4 / (k1p_P + biopo4);
    double dopupt_flag = 0.5 + fsign(0.5, limP_dop - limP_po4);
    double limP = limP_dop * dopupt_flag + limP_po4 * (1. - dopupt_flag);
    double u_P = fmin(avej, jmax * limP);
    double limSi = biosil / (5.e-3 + biosil);  // k1si = 5.e-3 (:2181)
    limP_dop = P->hdop * biodop / (k1p_Diat + biodop);
    limP_po4 = biopo4 / (k1p_Diat + biopo4);
    double dopupt_Diat_flag = 0.5 + fsign(0.5, limP_dop - limP_po4);
    double limP_Diat = limP_dop * dopupt_Diat_flag + limP_po4 * (1. - dopupt_Diat_flag);
    double u_Diat = fmin(avej_Diat, jmax_Diat * limSi);
    u_Diat = fmin(u_Diat, jmax_Diat * limP_Diat);
    u_P = fmin(u_P, jmax * biono3 / (k1n + biono3));
    u_Diat = fmin(u_Diat, jmax_Diat * biono3 / (k1n_Diat + biono3));
    double u_D = fmin(avej_D, jmax_D * limP);
    double dopupt_D_flag = dopupt_flag;
    // grazing (:2208-2245)
    double thetaZ = P->zprefP * biophyt + P->zprefDet * biodetr + P->zprefZ * biozoop + P->zprefDiaz * biodiaz + P->kzoo +
                    P->zprefDiat * biodiat;
    double ing_P = P->zprefP / thetaZ, ing_Det = P->zprefDet / thetaZ, ing_Z = P->zprefZ / thetaZ;
    double ing_D = P->zprefDiaz / thetaZ, ing_Diat = P->zprefDiat / thetaZ;
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
    double ligand = fmax(aou8 / 66. + pow(biodon, 0.8) / 4.8, 0.5) / 1000.;
    double fepa = (1.0 + P->kfeleq * (ligand - biodfe)) * o2flag;
    double feprime = ((-fepa + sqrt(fepa * fepa + 4.0 * P->kfeleq * biodfe)) / (2.0 * P->kfeleq)) * o2flag;
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
    double dig_D = gamma1 * graz_D * (redntp / diazntp);
    dig = dig + dig_D;
    excr = excr + gamma1 * (1 - geZ) * graz_D * (redntp / diazntp);
    double nr_excr_D = gamma1 * graz_D * (1 - (redntp / diazntp)) + (1 - gamma1) * graz_D * (1 - (redntp / diazntp));
    double sf_D = (1 - gamma1) * graz_D * (redntp / diazntp);
    sf = sf + sf_D;
    sf_phos = sf_phos + sf_D * redptn;
    // isotope fractionation factors (:2441-2530)
    double uno3 = fmax(fmin(npp * dtbio / biono3, 0.999), TRCMIN);
    double rno3 = fmax(fmin(b[V_DIN15] / (biono3 - b[V_DIN15]), 2 * RN15STD), RN15STD / 2.);
    double bassim = rno3 + P->eps_assim * (1 - uno3) / uno3 * log(1 - uno3) * rno3 / 1000.;
    double fcassim = bassim / (1 + bassim);
    double udon = fmax(fmin(recy_don * dtbio / biodon, 0.999), TRCMIN);
    double rdon = fmax(fmin(b[V_DON15] / (biodon - b[V_DON15]), 2 * RN15STD), RN15STD / 2.);
    double brecy = rdon + P->eps_recy * (1 - udon) / udon * log(1 - udon) * rdon / 1000.;
    double fcrecy = brecy / (1 + brecy);
    double rzoop = fmax(fmin(b[V_ZOOPN15] / (biozoop - b[V_ZOOPN15]), 2. * RN15STD), RN15STD / 2.);
    double bexcr = rzoop - P->eps_excr * rzoop / 1000.;
    double fcexcr = bexcr / (1 + bexcr);
    double bnfix = RN15STD - P->eps_nfix * RN15STD / 1000.;
    double fcnfix = bnfix / (1 + bnfix);
    double rtphytn15 = CL15(b[V_PHYTN15] / biophyt);
    double rtdiatn15 = CL15(b[V_DIATN15] / biodiat);
    double rtzoopn15 = CL15(b[V_ZOOPN15] / biozoop);
    double rtdetrn15 = CL15(b[V_DETRN15] / biodetr);
    double rtdiazn15 = CL15(b[V_DIAZN15] / biodiaz);
    double rdic13 = fmax(fmin(b[V_DIC13] / (biodic - b[V_DIC13]), 2. * RC13STD), 0.5 * RC13STD);
    double bc13npp = io.ac13b * rdic13;
    double fcnpp = bc13npp / (1 + bc13npp);
    double rtdic13 = CL13(b[V_DIC13] / biodic);
    double rtphytc13 = CL13(b[V_PHYTC13] / (biophyt * redctn));
    double rtdiatc13 = CL13(b[V_DIATC13] / (biodiat * redctn));
    double rtcaco3c13 = CL13(b[V_CACO3C13] / biocaco3);
    double rtzoopc13 = CL13(b[V_ZOOPC13] / (biozoop * redctn));
    double rtdetrc13 = CL13(b[V_DETRC13] / (biodetr * redctn));
    double rtdoc13 = CL13(b[V_DOC13] / (biodon * redctn));
    double rtdiazc13 = CL13(b[V_DIAZC13] / (biodiaz * redctn));
    // CaCO3 and opal production (:2532-2548)
    double calpro = ((sf_Z + morz) * io.capr + (sf_P + morp) * io.capr) * redctn * 1.e3;
    double sipr0 = (-0.46204044117647 * tanh(6.9 * biodfe * 1.e3 + -3.673092) + 1.60266544117647);
    double oplpro = (morp_Diat + sf_Diat) * sipr0 * fl(V_SIL) * (1.e-3);
    opldis = opldis * fl(V_OPL);
    expoopl = expoopl * fl(V_OPL);
    double GM15ptc = 0.0060 + 0.0069 * biopo4;
    double GM15ptn = GM15ptc * redctn * 1.e3;
    const double rnd = redntp / diazntp;

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
    ptn_P = b[V_PHYT_PHOS] / b[V_PHYT];
    ptn_detr = b[V_DETR_PHOS] / b[V_DETR];
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

__global__ void __launch_bounds__(32) k_mobi_column(const DevView v, int mi, double declin, int nbio, double dtbio, double rdtts,
                                                     double rnbio) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int ni = v.imt - 2, nrow = v.jhi - v.jlo + 1;
  if (idx >= (long long)ni * nrow) return;
  int i = (int)(idx % ni) + 2;
  int j = (int)(idx / ni) + v.jlo;
  const int kmx = v.kmt[X2(i, j)];
  if (kmx <= 0) return;
  const MobiPar *__restrict__ P = v.mobi_par;
  const int *__restrict__ ix = v.mobi_idx;
  const double redctn = P->redctn;
  const double pi = 3.14159265358979323846;  // atan(1.0)*4.0 in FP64
  const double radian = 360. / (2. * pi);

  // day fraction and incoming solar (09/mom/tracer.F:370-390)
  double lat = v.tlat[X2(i, j)];
  double rctheta = fmax(-1.5, fmin(1.5, lat / radian - declin));
  double cr = cos(rctheta);
  rctheta = P->kw / sqrt(1. - (1. - cr * cr) / (1.33 * 1.33));
  double dayfrac = fmin(1., -tan(lat / radian) * tan(declin));
  dayfrac = fmax(1e-12, acos(fmax(-1., dayfrac)) / pi);
  double swr = P->tap * v.dnswr[X2(i, j)] * 1e-3 * (1. + v.aice[X2(i, j)] * (exp(-P->ki * (v.hice[X2(i, j)] + v.hsno[X2(i, j)])) - 1.));

  double expo = 0.0, expo_phos = 0.0, phin = 0.0, rn15expo = 0.0, rc13expo = 0.0, rcaco3c13expo = 0.0;
  double expofe = 0.0, caco3in = 0.0, expocaco3 = 0.0, expoopl = 0.0;
  const long long n3 = v.n3;
  double b[MOBI_NVAR], clip[MOBI_NVAR];
  const int s_alk = ix[IX_ISALK], s_o2 = ix[IX_ISO2], s_c14 = ix[IX_ISC14];

  for (int k = 1; k <= kmx; k++) {
    const long long c = X3(i, k, j);
    const double dztk = v.dzt[k - 1], dztrk = v.dztr[k - 1];
    // gather (tracer.F:393-503)
#pragma unroll
    for (int m = 0; m < MOBI_NVAR; m++) b[m] = v.t_m1[c + (long long)(ix[IX_TR + m] - 1) * n3];
    const double t_in = v.t_m1[c + (long long)(ix[IX_ITEMP] - 1) * n3];
    const double o2_in = v.t_m1[c + (long long)(ix[IX_IO2] - 1) * n3] * 1000.;
    const double s_in = 1.e3 * v.t_m1[c + (long long)(ix[IX_ISALT] - 1) * n3] + 35.0;
    const double dic_in = b[V_DIC];
    const double c14_in = v.t_m1[c + (long long)(ix[IX_IC14] - 1) * n3];
    const double sgb = v.sg_bathy[XIJK(i, j, k)];
    // oxygen saturation -> AOU for the ligand parameterisation (tracer.F:457-476)
    double aou_in;
    {
      double f1 = log((298.15 - t_in) / (273.15 + t_in));
      double f2 = f1 * f1, f3 = f2 * f1, f4 = f3 * f1, f5 = f4 * f1;
      double o2sat = exp(2.00907 + 3.22014 * f1 + 4.05010 * f2 + 4.94457 * f3 - 2.56847E-1 * f4 + 3.88767 * f5 +
                         s_in * (-6.24523e-3 - 7.37614e-3 * f1 - 1.03410e-2 * f2 - 8.17083E-3 * f3) - 4.88682E-7 * s_in * s_in);
      o2sat = o2sat / 22391.6 * 1000.0 * 1000.;
      aou_in = o2sat - o2_in;
    }
    // ---- mobi_driver, level k of loop 1 (mobi.F:763-1289) ----
    SrcIO io;
    io.rn15impo = rn15expo;
    const double co2star = v.co2_star[c], Omega_c = v.co2_omega[c];
    double ac13_DIC_aq = -1.0512994e-4 * t_in + 1.011765;
    double ac13_aq_POC = -0.017 * log10(fmin(fmax(co2star * 1000., 2.), 74.)) + 1.0034;
    io.ac13b = ac13_aq_POC / ac13_DIC_aq;
    io.rc13impo = rc13expo * dztrk;
    io.rcaco3c13impo = rcaco3c13expo * dztrk;
    io.dissk1 = P->dissk0 * fmax(0., (1. - Omega_c));
    io.capr = P->caprmax * fmax(0., (Omega_c - 1.) / (P->kcapr + Omega_c - 1.));
    io.opl_disk1 = P->opl_disk0;
    swr = swr * exp(-P->kc * phin - P->kc_c * caco3in);
    phin = fmax(b[V_PHYT], TRCMIN) * dztk + fmax(b[V_DIAZ], TRCMIN) * dztk + fmax(b[V_DIAT], TRCMIN) * dztk;
    caco3in = caco3in + b[V_CACO3] * dztk;
    io.impocaco3 = expocaco3 * dztrk;
    io.gl = swr * exp(P->ztt[k - 1] * rctheta);
    io.impo = expo * dztrk;
    io.impo_phos = expo_phos * dztrk;
    io.impofe = expofe * dztrk;
    io.bct = pow(P->bbio, (P->cbio * t_in));
    io.impoopl = expoopl * dztrk;
    io.bctz = (0.5 * (tanh(o2_in - 8.) + 1)) * io.bct;   // same bbio**(cbio*t) value (:830-835)
    io.nud = P->nud0 * (0.6 + 0.4 * tanh(0.22 * fmax(o2_in, 0.)));
    io.nudon = P->nudon0;
    io.nudop = P->nudop0;
    io.dzt = dztk;
    io.dayfrac = dayfrac;
    io.wwd = P->wd[k - 1];
    io.wwc = P->wc[k - 1];
    io.wwo = P->wo[k - 1];
    io.o2 = o2_in;
    io.aou = aou_in;
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
    double lno3 = 0.5 * tanh(tno3 * 10 - 5.0);
    double sg_bdeni = (0.06 + 0.19 * pow(0.99, (fmax(o2_in, TRCMIN) - fmax(tno3, TRCMIN)))) * fmax(expo * sgb, TRCMIN) * redctn * 1.e3;
    sg_bdeni = fmin(sg_bdeni, sgb * expo);
    sg_bdeni = fmax(sg_bdeni, 0.);
    sg_bdeni = sg_bdeni * (0.5 + lno3) * no3flag * din15flag;
    const double bdeni = sg_bdeni;
    b[V_NO3] = b[V_NO3] + sgb * expo - sg_bdeni;
    double rno3 = fmax(tdin15, TRCMIN * RN15STD / (1 + RN15STD)) / fmax(tno3 - tdin15, TRCMIN * RN15STD / (1 + RN15STD));
    rno3 = fmin(rno3, 2. * RN15STD);
    rno3 = fmax(rno3, RN15STD / 2.);
    double eps_bdeni = P->eps_bdeni0 * exp(-2.5e-6 * (v.zt[k - 1]));
    double bbdeni = rno3 - eps_bdeni * rno3 / 1000.;
    b[V_DIN15] = b[V_DIN15] + rn15expo * sgb * expo - bbdeni / (1 + bbdeni) * sg_bdeni;
    // sediment carbon oxidation (Flogel 2011 / Somes 2021) and iron release (Dale 2015) (:1076-1110)
    double coxdepth = fmin(fmax(v.zt[k - 1], 50000.), 150000.);
    double oblinc = -1.26e-6 * coxdepth + 0.203;
    double obexpc = -6.e-7 * coxdepth + 1.14;
    double nburial = (oblinc * pow((expo * sgb * dztk / 100 * 86400. * 365. * redctn * 1000.), obexpc)) /
                     (86400. * 365. * dztk / 100 * redctn * 1000.);
    double coxsed = expo * sgb - nburial;
    double fesed = 85. * tanh(coxsed * redctn * 1000 * dztk / 100 * 86400. / o2_in) / (dztk / 100 * 86400 * 1000);
    b[V_DFE] = b[V_DFE] + fesed;
    b[V_PO4] = b[V_PO4] + sgb * expo_phos;
    b[V_DIC] = b[V_DIC] + sgb * expo * redctn;
    b[V_DIC13] = b[V_DIC13] + rc13expo * sgb * redctn;
    rc13expo = rc13expo - sgb * rc13expo;
    expo = expo - sgb * expo;
    expo_phos = expo_phos - sgb * expo_phos;
    const double dic_npzd_sms = b[V_DIC];
    // isotope ratios of DIC and CaCO3 for the calcite terms (:1258-1276)
    double rtdic13 = fmax(clip[V_DIC13], TRCMIN * RC13STD / (1 + RC13STD)) / fmax(dic_in, TRCMIN);
    rtdic13 = fmin(rtdic13, 2. * RC13STD / (1 + RC13STD));
    rtdic13 = fmax(rtdic13, 0.5 * RC13STD / (1 + RC13STD));
    double rtcaco3c13 = fmax(clip[V_CACO3C13], TRCMIN * RC13STD / (1 + RC13STD)) / fmax(clip[V_CACO3], TRCMIN);
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
    double fo2 = tanh(0.22 * fmax(o2_in, 0.));
    double so2 = dic_npzd_sms * P->redotc + nfix * rnbio * 1.25e-3;
    lno3 = 0.5 * tanh(tno3 - 2.5);
    double wcdeni = 800. * no3flag * so2 * (1.0 - fo2) * (0.5 + lno3) * din15flag;
    wcdeni = fmax(wcdeni, 0.);
    b[V_NO3] = b[V_NO3] - wcdeni;
    double uno3 = wcdeni * v.c2dtts / tno3;
    uno3 = fmin(uno3, 0.999);
    uno3 = fmax(uno3, TRCMIN);
    double bwcdeni = rno3 + P->eps_wcdeni * (1 - uno3) / uno3 * log(1 - uno3) * rno3 / 1000.;
    b[V_DIN15] = b[V_DIN15] - (bwcdeni / (1 + bwcdeni)) * wcdeni;
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
    if (k == 1) b[V_DFE] = b[V_DFE] + v.fe_atmdep[X2(i, j) + (long long)(mi - 1) * v.n2] * 1000 / (v.dzt[0] / 100.);
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

void launch_mobi(uvic_b200_ctx *c, const uvic_b200_stepinfo *si) {
  DevView &v = c->v;
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
  KLAUNCH("k_mobi_co2", k_mobi_co2, cdiv(ncell, 128), 128, v);
  KLAUNCH("k_mobi_column", k_mobi_column, cdiv(ncol, 32), 32, v, mi, declin, nbio, dtbio, rdtts, rnbio);
}
