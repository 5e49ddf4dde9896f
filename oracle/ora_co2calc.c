/*
 * ora_co2calc.c -- restatement of co2calc_SWS, drtsafe and ta_iter_SWS
 * (09/common/co2calc.F:1-400, 401-454, 455-526): carbonate chemistry on the seawater
 * hydrogen scale with pressure correction (Millero 1995), safe Newton for [H+].
 * The reference passes the equilibrium constants through COMMON /const/ and /species/
 * (:91-93, 463-465); here they travel in a struct.  TEST INFRASTRUCTURE ONLY (oracle.h).
 */
#include <math.h>
#include "oracle.h"

typedef struct {
  double k0, k1, k12, k2, kw, kb, ks, kf, k1p, k2p, k3p, ksi, ff; /* /const/ */
  double bt, st, ft, sit, pt, dic, ta;                            /* /species/ */
} co2c;

/* 09/common/co2calc.F:455-526 */
static void ta_iter_SWS(co2c *q, double x, double *fn, double *df) {
  double x2 = x * x;
  double x3 = x2 * x;
  q->k12 = q->k1 * q->k2;
  double k12p = q->k1p * q->k2p;
  double k123p = k12p * q->k3p;
  double c = 1.0 + q->st / q->ks + q->ft / q->kf;
  double a = x3 + q->k1p * x2 + k12p * x + k123p;
  double a2 = a * a;
  double da = 3.0 * x2 + 2.0 * q->k1p * x + k12p;
  double b = x2 + q->k1 * x + q->k12;
  double b2 = b * b;
  double db = 2.0 * x + q->k1;
  double k1 = q->k1, k12 = q->k12, dic = q->dic, bt = q->bt, kb = q->kb, kw = q->kw, pt = q->pt;
  double sit = q->sit, ksi = q->ksi, st = q->st, ks = q->ks, ft = q->ft, kf = q->kf, ta = q->ta;
  double t1, t2;
  *fn = k1 * x * dic / b + 2.0 * dic * k12 / b + bt / (1.0 + x / kb) + kw / x + pt * k12p * x / a + 2.0 * pt * k123p / a +
        sit / (1.0 + x / ksi) - x / c - st / (1.0 + ks / (x / c)) - ft / (1.0 + kf / (x / c)) - pt * x3 / a - ta;
  t1 = 1.0 + ks / (x / c);
  t2 = 1.0 + kf / (x / c);
  *df = ((k1 * dic * b) - k1 * x * dic * db) / b2 - 2.0 * dic * k12 * db / b2 - bt / kb / ((1.0 + x / kb) * (1.0 + x / kb)) - kw / x2 +
        (pt * k12p * (a - x * da)) / a2 - 2.0 * pt * k123p * da / a2 - sit / ksi / ((1.0 + x / ksi) * (1.0 + x / ksi)) - 1.0 / c -
        st * (1.0 / (t1 * t1)) * (ks * c / x2) - ft * (1.0 / (t2 * t2)) * (kf * c / x2) - pt * x2 * (3.0 * a - x * da) / a2;
}

/* 09/common/co2calc.F:401-454 */
static double drtsafe(co2c *q, double x1, double x2, double xacc) {
  const int maxit = 100;
  double fl, df, fh, xl, xh, swap, dxold, dx, f, temp, r;
  ta_iter_SWS(q, x1, &fl, &df);
  ta_iter_SWS(q, x2, &fh, &df);
  if (fl < 0.0) {
    xl = x1;
    xh = x2;
  } else {
    xh = x1;
    xl = x2;
    swap = fl;
    fl = fh;
    fh = swap;
  }
  r = 0.5 * (x1 + x2);
  dxold = fabs(x2 - x1);
  dx = dxold;
  ta_iter_SWS(q, r, &f, &df);
  for (int j = 1; j <= maxit; j++) {
    if (((r - xh) * df - f) * ((r - xl) * df - f) >= 0. || fabs(2.0 * f) > fabs(dxold * df)) {
      dxold = dx;
      dx = 0.5 * (xh - xl);
      r = xl + dx;
      if (xl == r) return r;
    } else {
      dxold = dx;
      dx = f / df;
      temp = r;
      r = r - dx;
      if (temp == r) return r;
    }
    if (fabs(dx) < xacc) return r;
    ta_iter_SWS(q, r, &f, &df);
    if (f < 0.0) {
      xl = r;
      fl = f;
    } else {
      xh = r;
      fh = f;
    }
  }
  (void)fl; (void)fh;
  return r;
}

/* 09/common/co2calc.F:1-400 */
void ora_co2calc_SWS(double t, double s, double dic_in, double ta_in, double co2_in, double atmpres, double depth, double *ph,
                     double *co2star_o, double *dco2star_o, double *pCO2_o, double *dpco2_o, double *CO3_o, double *Omega_c,
                     double *Omega_a) {
  co2c q;
  const double phhi = 6., phlo = 10.;
  const double sit_in = 7.6875e-03, pt_in = 0.5125e-3;
  double permil = 1.0 / 1024.5;
  q.pt = pt_in * permil;
  q.sit = sit_in * permil;
  q.ta = ta_in * permil;
  q.dic = dic_in * permil;
  double C2K = 273.15;
  double pres = depth * 0.1;
  double permeg = 1.e-6;
  double co2 = co2_in * permeg;

  double tk = C2K + t;
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

  q.ff = exp(-162.8301 + 218.2968 / tk100 + 90.9241 * log(tk100) - 1.47696 * tk1002 +
             s * (.025695 - .025225 * tk100 + 0.0049867 * tk1002));
  q.k0 = exp(93.4517 / tk100 - 60.2409 + 23.3585 * log(tk100) + s * (.023517 - 0.023656 * tk100 + 0.0047036 * tk1002));
  double rt_x = 83.1451 * tk;
  double delta_x = (57.7 - 0.118 * tk);
  double b_x = -1636.75 + 12.0408 * tk - 0.0327957 * tk * tk;
  b_x = b_x + 3.16528 * 1e-5 * tk * tk * tk;
  double FugFac = exp((b_x + 2 * delta_x) * 1 / rt_x);

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

  double x1 = pow(10.0, (-phhi));
  double x2 = pow(10.0, (-phlo));
  double xacc = 1.e-10;
  double hSWS = drtsafe(&q, x1, x2, xacc);

  double hSWS2 = hSWS * hSWS;
  double co2star = q.dic * hSWS2 / (hSWS2 + q.k1 * hSWS + q.k1 * q.k2);
  double co2starair = co2 * q.ff * atmpres;
  double dco2star = co2starair - co2star;
  *ph = -log10(hSWS);
  double pCO2 = co2star / (q.k0 * FugFac);
  double dpCO2 = pCO2 - co2starair;
  double CO3 = q.k1 * q.k2 * co2star / hSWS2;

  /* :343-365: the five x**0.5 stay pow(x, 0.5), as written and as in oracle/_ref (see ora_mobi.c header) */
  double Kspc = exp(-395.8293 + (6537.773 / tk) + 71.595 * log(tk) - 0.17959 * tk +
                    (-1.78938 + (410.64 / tk) + 0.0065453 * tk) * pow(s, 0.5) - 0.17755 * s + 0.0094979 * s15);
  double Kspa = exp(-395.9180 + (6685.079 / tk) + 71.595 * log(tk) - 0.17959 * tk +
                    (-0.157481 + (202.938 / tk) + 0.0039780 * tk) * pow(s, 0.5) - 0.23067 * s + 0.0136808 * s15);
  double DVc = -65.28 + 0.397 * t - 0.005155 * (t * t) + (19.816 - 0.0441 * t - 0.00017 * (t * t)) * pow(s / 35., 0.5);
  double DVa = -65.50 + 0.397 * t - 0.005155 * (t * t) + (19.82 - 0.0441 * t - 0.00017 * (t * t)) * pow(s / 35., 0.5);
  double DK = 0.01847 + 0.0001956 * t - 0.000002212 * (t * t) + (-0.03217 - 0.0000711 * t + 0.000002212) * pow(s / 35., 0.5);
  Kspc = Kspc * exp(-DVc * pitkR + 0.5 * DK * p2itkR);
  Kspa = Kspa * exp(-DVa * pitkR + 0.5 * DK * p2itkR);
  double Ca = 10.28E-3;
  *Omega_c = Ca * CO3 / Kspc;
  *Omega_a = Ca * CO3 / Kspa;

  *co2star_o = co2star / permil;
  *dco2star_o = dco2star / permil;
  *CO3_o = CO3 / permil;
  *pCO2_o = pCO2 / permeg;
  *dpco2_o = dpCO2 / permeg;
}
