/*
 * ora_sbc.c -- surface boundary conditions either side of the tracer step:
 *   ora_setvbc   09/mom/setvbc.F:60-140   vertical boundary conditions of the tracers from the coupler's flux array
 *   ora_set_sbc  09/mom/set_sbc.F:36-83 as called from 09/mom/tracer.F:1270-1288: surface tracers accumulated /
 *                averaged over an ocean segment for the atmosphere
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include "oracle.h"
#include "ora_index.h"

#define STF(i, j, n) c->stf[I2N(i, j, n)]
#define BTF(i, j, n) c->btf[I2N(i, j, n)]
#define TMASK(i, k, j) c->tmask[I3(i, k, j)]
#define KMT(i, j) c->kmt[I2(i, j)]
#define T(i, k, j, n, l) c->t[IT(i, k, j, n, l)]
/* sbc(imt,jmt,numsbc), 09/common/csbc.h */
#define SBC(i, j, m) c->sbc[((size_t)(i)-1) + (size_t)c->imt * (((size_t)(j)-1) + (size_t)c->jmt * ((size_t)(m)-1))]
#define BHF(i, j) c->bhf[((size_t)(i)-1) + (size_t)c->imt * ((size_t)(j)-1)]

void ora_setvbc(ora_ctx *c) {
  const int imt = c->imt, jmt = c->jmt, km = c->km, nt = c->nt;
  const int istrt = 2, iend = imt - 1;   /* max(2,is), min(imt-1,ie) with is=1, ie=imt (source/mom/mom.F:360) */
  /* no flux condition for all tracers at surface & bottom (09/mom/setvbc.F:68-76) */
  for (int n = 1; n <= nt; n++)
    for (int j = 1; j <= jmt; j++)
      for (int i = istrt; i <= iend; i++) {
        STF(i, j, n) = 0.0;
        BTF(i, j, n) = 0.0;
      }
  /* surface tracer fluxes from the atmosphere (:83-126): one assignment per tracer that owns a flux slot;
     the bottom heat flux enters through btf(itemp) */
  for (int j = 1; j <= jmt; j++)
    for (int i = istrt; i <= iend; i++) {
      for (int n = 1; n <= nt; n++) {
        const int m = c->sbc_flx_index[n - 1];
        if (m > 0) STF(i, j, n) = SBC(i, j, m) * TMASK(i, 1, j);
      }
      BTF(i, j, 1) = -BHF(i, j) * TMASK(i, 1, j);   /* itemp = 1 */
    }
}

/* one tracer: 09/mom/set_sbc.F:36-83 with doAccum = .true. */
static void set_sbc_one(ora_ctx *c, int isbc, int itr) {
  const int imt = c->imt, jmt = c->jmt, km = c->km;
  const int is = 2, ie = imt - 1;
  if (isbc <= 0 || itr <= 0) return;
  if (c->eots && c->osegs)
    for (int j = 1; j <= jmt; j++)
      for (int i = is; i <= ie; i++)
        if (KMT(i, j) != 0) SBC(i, j, isbc) = 0.0;
  if (c->eots)
    for (int j = 1; j <= jmt; j++)
      for (int i = is; i <= ie; i++) SBC(i, j, isbc) = SBC(i, j, isbc) + T(i, 1, j, itr, 1);   /* taup1 */
  if (c->eots && c->osege) {
    const double rts = 1.0 / c->ntspos;
    for (int j = 1; j <= jmt; j++)
      for (int i = is; i <= ie; i++)
        if (KMT(i, j) != 0) SBC(i, j, isbc) = rts * SBC(i, j, isbc);
  }
}

void ora_set_sbc(ora_ctx *c) {
  /* 09/mom/tracer.F:1270-1288: temperature and salinity first, then the other tracers */
  for (int n = 1; n <= 2 && n <= c->nt; n++)
    if (c->trsbcindex[n - 1] != 0) set_sbc_one(c, c->trsbcindex[n - 1], n);
  for (int n = 3; n <= c->nt; n++)
    if (c->trsbcindex[n - 1] != 0) set_sbc_one(c, c->trsbcindex[n - 1], n);
}

/* ---- air-sea gas exchange: the flux loop of gasbc (09/common/gasbc.F:62-69, 76-80, 148-266) ----
 * carbon (DIC, DI13C, 14C) and oxygen fluxes into the ocean from the segment-mean surface state the ocean left in sbc
 * (set_sbc), with co2calc_SWS at the surface; land points take the land carbon fluxes (O_mtlm); cyclic boundary.
 * gas_idx (1-based sbc slots): 0 isst, 1 isss, 2 issdic, 3 issalk, 4 issdic13, 5 issc14, 6 isso2, 7 iws, 8 inpp, 9 isr,
 * 10 iburn, 11 idicflx, 12 idic13flx, 13 ic14flx, 14 io2flx. */
void ora_co2calc_SWS(double t, double s, double dic_in, double ta_in, double co2_in, double atmpres, double depth, double *ph,
                     double *co2star_o, double *dco2star_o, double *pCO2_o, double *dpco2_o, double *CO3_o, double *Omega_c,
                     double *Omega_a);

void ora_gasbc(ora_ctx *c) {
  const int imt = c->imt, jmt = c->jmt, km = c->km;
  const int32_t *gx = c->gas_idx;
  const double rc13std = 0.0112372, rc14std = 1.176e-12, C2K = 273.15;   /* 09/mom/mobi.h:207,212 */
  const double atmpres = 1.0, zero = 0.;
  const double ak = 0.99915, aaqg = 0.998764;                            /* Zhang et al. 1995 (:64-66) */
  const double r13a = (c->dc13ccn * 0.001 + 1.) * rc13std;
  const double batmc13 = ak * aaqg * r13a;
  double xconv = 33.7 / 3.6e+05;                                         /* :79-80 */
  xconv = xconv * 0.75;
  (void)km;
  for (int j = 2; j <= jmt - 1; j++)
    for (int i = 2; i <= imt - 1; i++) {
      if (KMT(i, j) > 0) {   /* tmsk(i,j) >= 0.5 */
        double sss = 1000.0 * SBC(i, j, gx[1]) + 35.0;
        double sst = SBC(i, j, gx[0]);
        sst = dmin(35., dmax(sst, -2.));
        sss = dmin(45., dmax(sss, 0.));
        const double ao = 1. - c->aice[I2(i, j)];
        double pH, co2star, dco2star, pCO2, dpco2, CO3, Omega_c, Omega_a;
        ora_co2calc_SWS(sst, sss, SBC(i, j, gx[2]), SBC(i, j, gx[3]), c->co2ccn, atmpres, zero, &pH, &co2star, &dco2star, &pCO2,
                        &dpco2, &CO3, &Omega_c, &Omega_a);
        /* Schmidt number and piston velocity for CO2 */
        const double scco2 = 2073.1 - 125.62 * sst + 3.6276 * (sst * sst) - 0.043219 * (sst * sst * sst);
        const double ws2 = (SBC(i, j, gx[7]) * 0.01) * (SBC(i, j, gx[7]) * 0.01);
        const double piston_vel = ao * xconv * ws2 * pow(scco2 / 660., -0.5);
        SBC(i, j, gx[11]) = piston_vel * dco2star;
        const double adicg = 1.01051 - 1.05e-4 * sst;
        double r13dic = SBC(i, j, gx[4]) / (SBC(i, j, gx[2]) - SBC(i, j, gx[4]));
        r13dic = dmin(r13dic, 2. * rc13std);
        r13dic = dmax(r13dic, 0.5 * rc13std);
        const double bdic13 = ak * aaqg * r13dic / adicg;
        SBC(i, j, gx[12]) = piston_vel * ((batmc13 / (1 + batmc13)) * (dco2star + co2star) - (bdic13 / (1 + bdic13)) * co2star);
        SBC(i, j, gx[13]) = piston_vel * ((dco2star + co2star) * (1 + c->dc14ccn * 0.001) * rc14std -
                                          co2star * SBC(i, j, gx[5]) / SBC(i, j, gx[2]));
        /* oxygen */
        const double sco2 = 1638.0 - 81.83 * sst + 1.483 * (sst * sst) - 0.008004 * (sst * sst * sst);
        const double piston_o2 = ao * xconv * ws2 * pow(sco2 / 660.0, -0.5);
        const double f1 = log((298.15 - sst) / (C2K + sst));
        const double f2 = f1 * f1, f3 = f2 * f1, f4 = f3 * f1, f5 = f4 * f1;
        double o2sat = exp(2.00907 + 3.22014 * f1 + 4.05010 * f2 + 4.94457 * f3 - 2.56847E-1 * f4 + 3.88767 * f5 +
                           sss * (-6.24523e-3 - 7.37614e-3 * f1 - 1.03410e-2 * f2 - 8.17083E-3 * f3) - 4.88682E-7 * sss * sss);
        o2sat = o2sat / 22391.6 * 1000.0;
        SBC(i, j, gx[14]) = piston_o2 * (o2sat - SBC(i, j, gx[6]));
      } else if (gx[8] > 0) {
        /* land carbon fluxes, kg m-2 s-1 => umol cm-2 s-1 (:233-243) */
        const double f = SBC(i, j, gx[8]) - SBC(i, j, gx[9]) - SBC(i, j, gx[10]);
        SBC(i, j, gx[11]) = f * 0.1 / 12.e-6;
        SBC(i, j, gx[12]) = f * 0.1 / 12.e-6 * rc13std / (1 + rc13std);
        SBC(i, j, gx[13]) = f * rc14std * 0.1 / 12.e-6;
      }
    }
  for (int m = 11; m <= 14; m++) {   /* setbcx of the four flux slots (:248-266) */
    double *a = &SBC(1, 1, gx[m]);
    for (int j = 0; j < jmt; j++) {
      a[(size_t)j * imt] = a[(size_t)j * imt + imt - 2];
      a[(size_t)j * imt + imt - 1] = a[(size_t)j * imt + 1];
    }
  }
}
