/*
 * ora_advflux.c -- restatement of 09/mom/tracer_adv_flx.F: the O_fct branch (:381-1029,
 * one-dimensional delimiters O_fct_dlm1, the default per 09/mom/checks.F:899-903) and
 * the 2nd-order centred branch (:1030-1082).  Called as tracer does with one fully
 * open memory window: adv_flux(joff=0, js=2, je=jmt-1, is=2, ie=imt-1, n).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"
#include "ora_index.h"

#define T(i, k, j, n, l) c->t[IT(i, k, j, n, l)]
#define TM(i, k, j) c->tmask[I3(i, k, j)]
#define ADV_FE(i, k, j) c->adv_fe[I3(i, k, j)]
#define ADV_FN(i, k, j) c->adv_fn[I3(i, k, j)]
#define ADV_FB(i, k, j) c->adv_fb[I3Z(i, k, j)]
#define ANTI_FE(i, k, j) c->anti_fe[I3(i, k, j)]
#define ANTI_FN(i, k, j) c->anti_fn[I3(i, k, j)]
#define ANTI_FB(i, k, j) c->anti_fb[I3Z(i, k, j)]
#define RPY(i, k, j) c->R_plusY[I3(i, k, j)]
#define RMY(i, k, j) c->R_minusY[I3(i, k, j)]

/* source/mom/fdift.h:25-39 (O_fct form of ADV_Ty) */
#define ADV_Tx(i, k, j) ((ADV_FE(i, k, j) - ADV_FE((i)-1, k, j)) * c->cstdxt2r[I2(i, j)])
#define ADV_Ty(i, k, j, jrow) ((ADV_FN(i, k, j) - ADV_FN(i, k, (j)-1)) * c->cstdyt2r[(jrow)-1])
#define ADV_Tz(i, k, j) ((ADV_FB(i, (k)-1, j) - ADV_FB(i, k, j)) * c->dzt2r[(k)-1])

static void adv_flux_fct(ora_ctx *c, int n) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  const int joff = 0, js = 2, je = jmt - 1, is = 2, ie = imt - 1;
  const int kmm1 = km - 1;
  const double c2dtts = c->c2dtts;

  /* local data (:440-444); Fortran arrays are 1-based: allocate +2 and index directly */
  double *twodt = (double *)calloc(km + 2, sizeof(double));
  double *dcf = (double *)calloc(imt + 2, sizeof(double));
  double *Trmin = (double *)calloc(imt + 2, sizeof(double));
  double *Trmax = (double *)calloc(imt + 2, sizeof(double));
  double *Cpos = (double *)calloc(imt + 2, sizeof(double));
  double *Cneg = (double *)calloc(imt + 2, sizeof(double));
  double *flxlft = (double *)calloc(imt + 2, sizeof(double));
  double *flxrgt = (double *)calloc(imt + 2, sizeof(double));
  double *Rpl = (double *)calloc((size_t)imt * km, sizeof(double));
  double *Rmn = (double *)calloc((size_t)imt * km, sizeof(double));
  double *t_lo = (double *)calloc((size_t)imt * km, sizeof(double));
#define RPL(i, k) Rpl[((i)-1) + (size_t)imt * ((k)-1)]
#define RMN(i, k) Rmn[((i)-1) + (size_t)imt * ((k)-1)]
#define T_LO(i, k) t_lo[((i)-1) + (size_t)imt * ((k)-1)]
  /* tmaski (:484-490) is c1 - tmask; evaluated inline */
#define TMI(i, k, j) (1.0 - TM(i, k, j))

  /* limit the indices (:452-457) */
  int istrt = imax(2, is);
  int iend = imin(imt - 1, ie);
  int istrtm1 = istrt - 1;
  int iendp1 = iend + 1;
  int jstrt = js;
  int jend = imin(je, jmt - 1 - joff);

  /* initialization when calculating jrow 2 (:463-482) */
  if (joff + js == 2) {
    jstrt = js - 1;
    for (int k = 1; k <= km; k++)
      for (int i = istrt - 1; i <= iend; i++) {
        ADV_FN(i, k, 1) = 0.0;
        ANTI_FN(i, k, 1) = 0.0;
        RPY(i, k, 1) = 0.0;
        RMY(i, k, 1) = 0.0;
      }
  }

  /* 2*advective low order (upstream) flux across northern, eastern and bottom faces (:496-548) */
  int jlast = imin(jend + 1 + joff, jmt - 1) - joff;
  for (int j = js - 1; j <= jlast; j++)
    for (int k = 1; k <= km; k++)
      for (int i = istrt; i <= iend; i++) {
        double totadv = c->adv_vnt[I3(i, k, j)] + c->adv_vntiso[I3(i, k, j)];
        ADV_FN(i, k, j) = totadv * (T(i, k, j, n, TAUM1) + T(i, k, j + 1, n, TAUM1)) +
                          fabs(totadv) * (T(i, k, j, n, TAUM1) - T(i, k, j + 1, n, TAUM1));
      }
  for (int j = js; j <= jlast; j++) {
    for (int k = 1; k <= km; k++)
      for (int i = istrtm1; i <= iend; i++) {
        double totadv = c->adv_vet[I3(i, k, j)] + c->adv_vetiso[I3(i, k, j)];
        ADV_FE(i, k, j) = totadv * (T(i, k, j, n, TAUM1) + T(i + 1, k, j, n, TAUM1)) +
                          fabs(totadv) * (T(i, k, j, n, TAUM1) - T(i + 1, k, j, n, TAUM1));
      }
    for (int k = 1; k <= kmm1; k++)
      for (int i = istrt; i <= iend; i++) {
        double totadv = c->adv_vbt[I3Z(i, k, j)] + c->adv_vbtiso[I3Z(i, k, j)];
        ADV_FB(i, k, j) = totadv * (T(i, k + 1, j, n, TAUM1) + T(i, k, j, n, TAUM1)) +
                          fabs(totadv) * (T(i, k + 1, j, n, TAUM1) - T(i, k, j, n, TAUM1));
      }
    for (int i = istrt; i <= iend; i++) {
      ADV_FB(i, 0, j) = c->adv_vbt[I3Z(i, 0, j)] * 2.0 * T(i, 1, j, n, TAUM1);
      ADV_FB(i, km, j) = 0.0;
    }
  }

  /* main j loop (:553-1002): iteration j works on row j+1 */
  for (int j = jstrt; j <= jend; j++) {
    int jrow = (j + 1) + joff;
    int jp2 = imin(j + 2 + joff, jmt) - joff;
    int jp1 = imin(j + 1 + joff, jmt - 1) - joff;

    /* low order solution at row j+1 (:560-580) */
    for (int k = 1; k <= km; k++) {
      twodt[k] = c2dtts * c->dtxcel[k - 1];
      for (int i = istrt; i <= iend; i++)
        T_LO(i, k) = (T(i, k, j + 1, n, TAUM1) -
                      twodt[k] * (ADV_Tx(i, k, jp1) + ADV_Ty(i, k, jp1, jrow) + ADV_Tz(i, k, jp1)) * TM(i, k, j + 1));
    }
    ora_setbcx(t_lo, imt, km);
    memcpy(&c->t_lo_dump[I3(1, 1, j + 1)], t_lo, sizeof(double) * (size_t)imt * km);

    /* raw antidiffusive fluxes: high order (leapfrog) minus low order (:582-620) */
    for (int k = 1; k <= km; k++) {
      for (int i = istrtm1; i <= iend; i++) {
        double totadv = c->adv_vet[I3(i, k, jp1)] + c->adv_vetiso[I3(i, k, jp1)];
        ANTI_FE(i, k, j + 1) = totadv * (T(i, k, j + 1, n, TAU) + T(i + 1, k, j + 1, n, TAU)) - ADV_FE(i, k, jp1);
      }
      for (int i = istrt; i <= iend; i++) {
        double totadv = c->adv_vnt[I3(i, k, jp1)] + c->adv_vntiso[I3(i, k, jp1)];
        ANTI_FN(i, k, j + 1) = totadv * (T(i, k, j + 1, n, TAU) + T(i, k, jp2, n, TAU)) - ADV_FN(i, k, jp1);
      }
    }
    for (int k = 1; k <= kmm1; k++)
      for (int i = istrt; i <= iend; i++) {
        double totadv = c->adv_vbt[I3Z(i, k, jp1)] + c->adv_vbtiso[I3Z(i, k, jp1)];
        ANTI_FB(i, k, j + 1) = totadv * (T(i, k, j + 1, n, TAU) + T(i, k + 1, j + 1, n, TAU)) - ADV_FB(i, k, jp1) * TM(i, k, j + 1);
      }
    for (int i = istrt; i <= iend; i++) {
      ANTI_FB(i, 0, j + 1) = c->adv_vbt[I3Z(i, 0, j + 1)] * 2.0 * T(i, 1, j + 1, n, TAUM1);
      ANTI_FB(i, km, j + 1) = 0.0;
    }

    /* ---- delimit x-direction (:635-712) ---- */
    for (int k = 1; k <= km; k++) {
      for (int i = istrt; i <= iendp1; i++) Trmax[i] = 0.5 * (T(i - 1, k, j + 1, n, TAU) + T(i, k, j + 1, n, TAU));
      for (int i = istrt; i <= iend; i++) {
        double fxa = TM(i - 1, k, j + 1) * Trmax[i] + TMI(i - 1, k, j + 1) * T_LO(i, k);
        double fxb = TM(i + 1, k, j + 1) * Trmax[i + 1] + TMI(i + 1, k, j + 1) * T_LO(i, k);
        Trmax[i] = dmax(dmax(fxa, fxb), T_LO(i, k));
        Trmin[i] = dmin(dmin(fxa, fxb), T_LO(i, k));
        dcf[i] = c->cstdxt2r[I2(i, j + 1)];
        flxlft[i] = ANTI_FE(i - 1, k, j + 1);
        flxrgt[i] = ANTI_FE(i, k, j + 1);
      }
      for (int i = istrt; i <= iend; i++) {
        double Pplus = c2dtts * dcf[i] * (dmax(0.0, flxlft[i]) - dmin(0.0, flxrgt[i]));
        double Pminus = c2dtts * dcf[i] * (dmax(0.0, flxrgt[i]) - dmin(0.0, flxlft[i]));
        double Qplus = Trmax[i] - T_LO(i, k);
        double Qminus = T_LO(i, k) - Trmin[i];
        RPL(i, k) = dmin(1., TM(i, k, j + 1) * Qplus / (Pplus + EPSLN));
        RMN(i, k) = dmin(1., TM(i, k, j + 1) * Qminus / (Pminus + EPSLN));
      }
      ora_setbcx(Rpl, imt, km);
      ora_setbcx(Rmn, imt, km);
      for (int i = istrt; i <= iendp1; i++) {
        Cpos[i - 1] = dmin(RPL(i, k), RMN(i - 1, k));
        Cneg[i - 1] = dmin(RPL(i - 1, k), RMN(i, k));
      }
      for (int i = istrtm1; i <= iend; i++)
        ANTI_FE(i, k, j + 1) = 0.5 * ((Cpos[i] + Cneg[i]) * ANTI_FE(i, k, j + 1) + (Cpos[i] - Cneg[i]) * fabs(ANTI_FE(i, k, j + 1)));
    }

    /* ---- delimit y-direction (:714-784) ---- */
    for (int k = 1; k <= km; k++) {
      for (int i = istrt; i <= iend; i++) {
        double fxa = 0.5 * TM(i, k, j) * (T(i, k, j, n, TAU) + T(i, k, j + 1, n, TAU)) + TMI(i, k, j) * T_LO(i, k);
        double fxb = 0.5 * TM(i, k, jp2) * (T(i, k, j + 1, n, TAU) + T(i, k, jp2, n, TAU)) + TMI(i, k, jp2) * T_LO(i, k);
        Trmax[i] = dmax(dmax(fxa, fxb), T_LO(i, k));
        Trmin[i] = dmin(dmin(fxa, fxb), T_LO(i, k));
        dcf[i] = c->cstdyt2r[jrow - 1];
        flxlft[i] = ANTI_FN(i, k, j);
        flxrgt[i] = ANTI_FN(i, k, j + 1);
      }
      for (int i = istrt; i <= iend; i++) {
        double Pplus = c2dtts * dcf[i] * (dmax(0.0, flxlft[i]) - dmin(0.0, flxrgt[i]));
        double Pminus = c2dtts * dcf[i] * (dmax(0.0, flxrgt[i]) - dmin(0.0, flxlft[i]));
        double Qplus = Trmax[i] - T_LO(i, k);
        double Qminus = T_LO(i, k) - Trmin[i];
        RPY(i, k, j + 1) = dmin(1., TM(i, k, j + 1) * Qplus / (Pplus + EPSLN));
        RMY(i, k, j + 1) = dmin(1., TM(i, k, j + 1) * Qminus / (Pminus + EPSLN));
      }
      for (int i = istrt; i <= iend; i++) {
        Cpos[i] = dmin(RPY(i, k, j + 1), RMY(i, k, j));
        Cneg[i] = dmin(RPY(i, k, j), RMY(i, k, j + 1));
      }
      for (int i = istrt; i <= iend; i++)
        ANTI_FN(i, k, j) = 0.5 * ((Cpos[i] + Cneg[i]) * ANTI_FN(i, k, j) + (Cpos[i] - Cneg[i]) * fabs(ANTI_FN(i, k, j)));
    }

    /* ---- delimit z-direction (:786-983) ---- */
    for (int k = 1; k <= km; k++) {
      for (int i = istrt; i <= iend; i++) {
        double fxa, fxb;
        dcf[i] = c->dzt2r[k - 1];
        flxlft[i] = ANTI_FB(i, k, j + 1);
        flxrgt[i] = ANTI_FB(i, k - 1, j + 1);
        if (k > 1)
          fxa = 0.5 * TM(i, k - 1, j + 1) * (T(i, k - 1, j + 1, n, TAU) + T(i, k, j + 1, n, TAU)) + TMI(i, k - 1, j + 1) * T_LO(i, k);
        else
          fxa = T_LO(i, k);
        if (k < km)
          fxb = 0.5 * TM(i, k + 1, j + 1) * (T(i, k, j + 1, n, TAU) + T(i, k + 1, j + 1, n, TAU)) + TMI(i, k + 1, j + 1) * T_LO(i, k);
        else
          fxb = T_LO(i, k);
        Trmax[i] = dmax(dmax(fxa, fxb), T_LO(i, k));
        Trmin[i] = dmin(dmin(fxa, fxb), T_LO(i, k));
      }
      for (int i = istrt; i <= iend; i++) {
        double Pplus = c2dtts * dcf[i] * (dmax(0.0, flxlft[i]) - dmin(0.0, flxrgt[i]));
        double Pminus = c2dtts * dcf[i] * (dmax(0.0, flxrgt[i]) - dmin(0.0, flxlft[i]));
        double Qplus = Trmax[i] - T_LO(i, k);
        double Qminus = T_LO(i, k) - Trmin[i];
        RPL(i, k) = dmin(1., TM(i, k, j + 1) * Qplus / (Pplus + EPSLN));
        RMN(i, k) = dmin(1., TM(i, k, j + 1) * Qminus / (Pminus + EPSLN));
      }
    }
    for (int k = 1; k <= kmm1; k++) {
      for (int i = istrt; i <= iend; i++) {
        Cneg[i] = dmin(RPL(i, k + 1), RMN(i, k));
        Cpos[i] = dmin(RPL(i, k), RMN(i, k + 1));
      }
      for (int i = istrt; i <= iend; i++)
        ANTI_FB(i, k, j + 1) = 0.5 * ((Cpos[i] + Cneg[i]) * ANTI_FB(i, k, j + 1) + (Cpos[i] - Cneg[i]) * fabs(ANTI_FB(i, k, j + 1)));
    }
    for (int i = istrt; i <= iend; i++) {
      ANTI_FB(i, 0, j + 1) = 0.0;
      ANTI_FB(i, km, j + 1) = 0.0;
    }

    /* complete advective fluxes: add low order to delimited antidiffusive (:985-1002) */
    for (int k = 1; k <= km; k++) {
      for (int i = istrtm1; i <= iend; i++) ANTI_FE(i, k, j + 1) = ANTI_FE(i, k, j + 1) + ADV_FE(i, k, jp1);
      for (int i = istrt; i <= iend; i++) {
        ANTI_FN(i, k, j) = (ANTI_FN(i, k, j) + ADV_FN(i, k, j)) * TM(i, k, j);
        ANTI_FB(i, k, j + 1) = (ANTI_FB(i, k, j + 1) + ADV_FB(i, k, jp1)) * TM(i, k, j + 1);
      }
    }
  }

  /* set 2*corrected advective fluxes (:1004-1028) */
  for (int j = js - 1; j <= jend; j++)
    for (int k = 1; k <= km; k++)
      for (int i = istrt; i <= iend; i++) ADV_FN(i, k, j) = ANTI_FN(i, k, j);
  for (int j = js; j <= jend; j++) {
    for (int k = 1; k <= km; k++)
      for (int i = istrtm1; i <= iend; i++) ADV_FE(i, k, j) = ANTI_FE(i, k, j);
    for (int k = 1; k <= kmm1; k++)
      for (int i = istrt; i <= iend; i++) ADV_FB(i, k, j) = ANTI_FB(i, k, j);
  }

  free(twodt); free(dcf); free(Trmin); free(Trmax); free(Cpos); free(Cneg);
  free(flxlft); free(flxrgt); free(Rpl); free(Rmn); free(t_lo);
}

/* 09/mom/tracer_adv_flx.F:1030-1082  2nd order centred */
static void adv_flux_2nd(ora_ctx *c, int n) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  const int js = 2, je = jmt - 1, istrt = 2, iend = imt - 1;
  for (int j = js; j <= je; j++)
    for (int k = 1; k <= km; k++)
      for (int i = istrt - 1; i <= iend; i++)
        ADV_FE(i, k, j) = c->adv_vet[I3(i, k, j)] * (T(i, k, j, n, TAU) + T(i + 1, k, j, n, TAU));
  for (int j = js; j <= je; j++)
    for (int k = 1; k <= km - 1; k++)
      for (int i = istrt; i <= iend; i++)
        ADV_FB(i, k, j) = c->adv_vbt[I3Z(i, k, j)] * (T(i, k, j, n, TAU) + T(i, k + 1, j, n, TAU));
}

void ora_adv_flux(ora_ctx *c, int n) {
  if (c->fct)
    adv_flux_fct(c, n);
  else
    adv_flux_2nd(c, n);
}
