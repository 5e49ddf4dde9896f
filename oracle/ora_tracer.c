/*
 * ora_tracer.c -- restatement of subroutine tracer (09/mom/tracer.F:214-1364) and its
 * helpers ivdift (:1938-2032) / invtri (source/mom/invtri.F), diagt1 (:1516-1565),
 * with the flux-divergence statement functions of source/mom/fdift.h.
 * Called as mom does with one fully open window: tracer(joff=0, js=2, je=jmt-1,
 * is=2, ie=imt-1) (source/mom/mom.F:373-389).
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
#define DIFF_FE(i, k, j) c->diff_fe[I3(i, k, j)]
#define DIFF_FN(i, k, j) c->diff_fn[I3(i, k, j)]
#define DIFF_FB(i, k, j) c->diff_fb[I3Z(i, k, j)]
#define DIFF_FBISO(i, k, j) c->diff_fbiso[I3Z(i, k, j)]

/* source/mom/fdift.h:25-88 */
#define ADV_Tx(i, k, j) ((ADV_FE(i, k, j) - ADV_FE((i)-1, k, j)) * c->cstdxt2r[I2(i, j)])
#define ADV_Ty_fct(i, k, j, jrow) ((ADV_FN(i, k, j) - ADV_FN(i, k, (j)-1)) * c->cstdyt2r[(jrow)-1])
#define ADV_Ty_2nd(i, k, j, jrow, n)                                                           \
  ((c->adv_vnt[I3(i, k, j)] * (T(i, k, j, n, TAU) + T(i, k, (j) + 1, n, TAU)) -                \
    c->adv_vnt[I3(i, k, (j)-1)] * (T(i, k, (j)-1, n, TAU) + T(i, k, j, n, TAU))) *             \
   c->cstdyt2r[(jrow)-1])
#define ADV_Tz(i, k, j) ((ADV_FB(i, (k)-1, j) - ADV_FB(i, k, j)) * c->dzt2r[(k)-1])
#define ADV_Txiso(i, k, j, n)                                                                              \
  (c->cstdxt2r[I2(i, j)] * (c->adv_vetiso[I3(i, k, j)] * (T((i) + 1, k, j, n, TAUM1) + T(i, k, j, n, TAUM1)) - \
                            c->adv_vetiso[I3((i)-1, k, j)] * (T(i, k, j, n, TAUM1) + T((i)-1, k, j, n, TAUM1))))
#define ADV_Tyiso(i, k, j, jrow, n)                                                                            \
  (c->cstdyt2r[(jrow)-1] * (c->adv_vntiso[I3(i, k, j)] * (T(i, k, (j) + 1, n, TAUM1) + T(i, k, j, n, TAUM1)) - \
                            c->adv_vntiso[I3(i, k, (j)-1)] * (T(i, k, j, n, TAUM1) + T(i, k, (j)-1, n, TAUM1))))
#define ADV_Tziso(i, k, j) (c->dzt2r[(k)-1] * (c->adv_fbiso[I3Z(i, (k)-1, j)] - c->adv_fbiso[I3Z(i, k, j)]))
#define DIFF_Tx(i, k, j) ((DIFF_FE(i, k, j) * TM((i) + 1, k, j) - DIFF_FE((i)-1, k, j) * TM((i)-1, k, j)) * c->cstdxtr[I2(i, j)])
#define DIFF_Ty(i, k, j, jrow) ((DIFF_FN(i, k, j) * TM(i, k, (j) + 1) - DIFF_FN(i, k, (j)-1) * TM(i, k, (j)-1)) * c->cstdytr[(jrow)-1])
#define DIFF_Tz_iso(i, k, j)                                                               \
  ((DIFF_FB(i, (k)-1, j) - DIFF_FB(i, k, j)) * c->dztr[(k)-1] * (1.0 - c->aidif) +         \
   (DIFF_FBISO(i, (k)-1, j) - DIFF_FBISO(i, k, j)) * c->dztr[(k)-1])
#define DIFF_Tz_plain(i, k, j) ((DIFF_FB(i, (k)-1, j) - DIFF_FB(i, k, j)) * c->dztr[(k)-1])

/* 09/mom/loadmw.F:60-77  land/sea masks from kmt */
void ora_make_masks(ora_ctx *c) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  for (int j = 1; j <= jmt; j++)
    for (int k = 1; k <= km; k++)
      for (int i = 1; i <= imt; i++) TM(i, k, j) = (c->kmt[I2(i, j)] >= k) ? 1.0 : 0.0;
}

/* source/mom/invtri.F:1-115.  z(imt,km,jmt); topbc/botbc(imt,jmt); dcb(imt,km,jmt); rows 2..jmt-1 */
void ora_invtri(ora_ctx *c, double *z, const double *topbc, const double *botbc, const double *dcb, const double *tdt) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  const int js = 2, je = jmt - 1, is = 2, ie = imt - 1;
  const double aidif = c->aidif;
  size_t n3 = (size_t)imt * km * jmt, n3z = (size_t)imt * (km + 1) * jmt;
  double *a = (double *)calloc(n3, sizeof(double)), *b = (double *)calloc(n3, sizeof(double));
  double *cc = (double *)calloc(n3z, sizeof(double)), *f = (double *)calloc(n3z, sizeof(double));
  double *e = (double *)calloc(n3, sizeof(double)), *bet = (double *)calloc((size_t)imt * jmt, sizeof(double));
#define Z(i, k, j) z[I3(i, k, j)]
#define A(i, k, j) a[I3(i, k, j)]
#define B(i, k, j) b[I3(i, k, j)]
#define C(i, k, j) cc[I3Z(i, k, j)]
#define F(i, k, j) f[I3Z(i, k, j)]
#define E(i, k, j) e[I3(i, k, j)]
#define BET(i, j) bet[I2(i, j)]
  for (int j = js; j <= je; j++) {
    int jrow = j;
    for (int k = 1; k <= km; k++) {
      int km1 = imax(1, k - 1);
      int kp1 = imin(k + 1, km);
      double factu = c->dztur[k - 1] * tdt[k - 1] * aidif;
      double factl = c->dztlr[k - 1] * tdt[k - 1] * aidif;
      for (int i = is; i <= ie; i++) {
        A(i, k, j) = -dcb[I3(i, km1, j)] * factu * TM(i, k, j);
        C(i, k, j) = -dcb[I3(i, k, j)] * factl * TM(i, kp1, j);
        F(i, k, j) = Z(i, k, j) * TM(i, k, j);
        B(i, k, j) = 1.0 - A(i, k, j) - C(i, k, j);
      }
    }
    for (int i = is; i <= ie; i++) {
      A(i, 1, j) = 0.0;
      C(i, km, j) = 0.0;
      B(i, 1, j) = 1.0 - A(i, 1, j) - C(i, 1, j);
      B(i, km, j) = 1.0 - A(i, km, j) - C(i, km, j);
      /* top and bottom b.c. (:77-81) */
      F(i, 1, j) = Z(i, 1, j) + topbc[I2(i, j)] * tdt[0] * c->dztr[0] * aidif * TM(i, 1, j);
      int k = imax(2, c->kmt[I2(i, jrow)]);
      F(i, k, j) = Z(i, k, j) - botbc[I2(i, j)] * tdt[k - 1] * c->dztr[k - 1] * aidif * TM(i, k, j);
    }
  }
  /* decomposition and forward substitution (:85-100) */
  const double eps = 1.e-30;
  for (int j = js; j <= je; j++) {
    for (int i = is; i <= ie; i++) {
      BET(i, j) = TM(i, 1, j) / (B(i, 1, j) + eps);
      Z(i, 1, j) = F(i, 1, j) * BET(i, j);
    }
    for (int k = 2; k <= km; k++)
      for (int i = is; i <= ie; i++) {
        E(i, k, j) = C(i, k - 1, j) * BET(i, j);
        BET(i, j) = TM(i, k, j) / (B(i, k, j) - A(i, k, j) * E(i, k, j) + eps);
        Z(i, k, j) = (F(i, k, j) - A(i, k, j) * Z(i, k - 1, j)) * BET(i, j);
      }
  }
  /* back substitution (:102-110) */
  for (int j = js; j <= je; j++)
    for (int k = km - 1; k >= 1; k--)
      for (int i = is; i <= ie; i++) Z(i, k, j) = Z(i, k, j) - E(i, k + 1, j) * Z(i, k + 1, j);
  free(a); free(b); free(cc); free(f); free(e); free(bet);
#undef Z
#undef A
#undef B
#undef C
#undef F
#undef E
#undef BET
}

/* 09/mom/tracer.F:1516-1565  diagt1: tbar/travar/dtabs and region sums (tsiperts/tavgts steps) */
void ora_diag_tbar(ora_ctx *c, int n) {
  const int imt = c->imt, km = c->km, jmt = c->jmt, nt = c->nt;
  const int js = 2, je = jmt - 1, is = 2, ie = imt - 1;
  double *temp1 = (double *)calloc((size_t)imt * km, sizeof(double));
  double *temp2 = (double *)calloc((size_t)imt * km, sizeof(double));
  double *temp3 = (double *)calloc((size_t)imt * km, sizeof(double));
#define TB(a, k, n, jrow) c->a[((k)-1) + (size_t)km * (((n)-1) + (size_t)nt * ((jrow)-1))]
#define TMP(a, i, k) a[((i)-1) + (size_t)imt * ((k)-1)]
  for (int j = js; j <= je; j++) {
    int jrow = j;
    double r2dt = 1.0 / c->c2dtts;
    double cosdyt = c->cst[jrow - 1] * c->dyt[jrow - 1];
    for (int k = 1; k <= km; k++) {
      double fx = r2dt / c->dtxcel[k - 1];
      for (int i = is; i <= ie; i++) {
        double darea = c->dzt[k - 1] * c->dxt[i - 1] * cosdyt * TM(i, k, j);
        TMP(temp3, i, k) = T(i, k, j, n, TAU) * darea;
        TMP(temp1, i, k) = T(i, k, j, n, TAU) * T(i, k, j, n, TAU) * darea;
        TMP(temp2, i, k) = fabs(T(i, k, j, n, TAUP1) - T(i, k, j, n, TAUM1)) * darea * fx;
      }
      for (int i = is; i <= ie; i++) {
        TB(tbar, k, n, jrow) = TB(tbar, k, n, jrow) + TMP(temp3, i, k);
        TB(travar, k, n, jrow) = TB(travar, k, n, jrow) + TMP(temp1, i, k);
        TB(dtabs, k, n, jrow) = TB(dtabs, k, n, jrow) + TMP(temp2, i, k);
      }
    }
  }
  /* region sums (:1548-1565), nhreg = 3 (09/common/param.h:28) */
  for (int j = js; j <= je; j++) {
    int jrow = j;
    for (int i = is; i <= ie; i++) {
      int mask = c->mskhr[I2(i, jrow)];
      if (mask != 0) {
        double boxar = c->cst[jrow - 1] * c->dxt[i - 1] * c->dyt[jrow - 1] * TM(i, 1, j) * 0.0001;
        for (int k = 1; k <= km; k++)
          c->sumbk[(mask - 1) + 3 * ((k - 1) + (size_t)km * (n - 1))] += T(i, k, j, n, TAU) * boxar * c->dzt[k - 1] * TM(i, k, j) * 0.01;
      }
    }
  }
  free(temp1); free(temp2); free(temp3);
#undef TB
#undef TMP
}

/* subroutine tracer, 09/mom/tracer.F */
void ora_tracer(ora_ctx *c) {
  const int imt = c->imt, km = c->km, jmt = c->jmt, nt = c->nt;
  const int joff = 0, js = 2, je = jmt - 1;
  const int istrt = 2, iend = imt - 1;
  double twodt[km];

  /* metric prologue (:234-249) */
  int limit = imin(je + 1 + joff, jmt) - joff;
  for (int j = js; j <= limit; j++) {
    int jrow = j + joff;
    for (int i = istrt - 1; i <= iend; i++) {
      c->cstdxtr[I2(i, j)] = c->cstr[jrow - 1] * c->dxtr[i - 1];
      c->cstdxt2r[I2(i, j)] = c->cstr[jrow - 1] * c->dxtr[i - 1] * 0.5;
      c->cstdxur[I2(i, j)] = c->cstr[jrow - 1] * c->dxur[i - 1];
      c->ah_cstdxur[I2(i, j)] = c->diff_cet * c->cstr[jrow - 1] * c->dxur[i - 1];
    }
  }

  /* ocean biogeochemistry source terms (:306-545, 848-867) */
  if (c->do_mobi) ora_mobi_columns(c);

  for (int n = 1; n <= nt; n++) {
    /* advective flux (:918) */
    ora_adv_flux(c, n);

    /* diffusive flux on eastern / northern faces (:930-961) */
    for (int j = js; j <= je; j++)
      for (int k = 1; k <= km; k++)
        for (int i = istrt - 1; i <= iend; i++)
          DIFF_FE(i, k, j) = c->ah_cstdxur[I2(i, j)] * (T(i + 1, k, j, n, TAUM1) - T(i, k, j, n, TAUM1));
    for (int j = js - 1; j <= je; j++) {
      int jrow = j + joff;
      for (int k = 1; k <= km; k++)
        for (int i = istrt; i <= iend; i++)
          DIFF_FN(i, k, j) = c->diff_cnt * c->csu_dyur[jrow - 1] * (T(i, k, j + 1, n, TAUM1) - T(i, k, j, n, TAUM1));
    }
    /* diffusive flux across bottom face (:1025-1032) */
    for (int j = js; j <= je; j++)
      for (int k = 1; k <= km - 1; k++)
        for (int i = istrt; i <= iend; i++)
          DIFF_FB(i, k, j) = c->diff_cbt[I3(i, k, j)] * c->dzwr[k] * (T(i, k, j, n, TAUM1) - T(i, k + 1, j, n, TAUM1));

    /* isopycnal fluxes (:1041) */
    if (c->isopycmix) ora_isoflux(c, n);

    /* vertical b.c. (:1053-1067) */
    for (int j = js; j <= je; j++) {
      int jrow = j + joff;
      for (int i = istrt; i <= iend; i++) {
        int kb = c->kmt[I2(i, jrow)];
        DIFF_FB(i, 0, j) = c->stf[I2N(i, j, n)];
        DIFF_FB(i, kb, j) = c->btf[I2N(i, j, n)];
        ADV_FB(i, 0, j) = c->adv_vbt[I3Z(i, 0, j)] * (T(i, 1, j, n, TAU) + T(i, 1, j, n, TAU));
        ADV_FB(i, km, j) = c->adv_vbt[I3Z(i, km, j)] * T(i, km, j, n, TAU);
      }
    }

    /* source term (:1075-1086) */
    memset(c->source, 0, sizeof(double) * (size_t)imt * km * jmt);
    if (c->itrc[n - 1] != 0)
      for (int j = js; j <= je; j++)
        for (int k = 1; k <= km; k++)
          for (int i = istrt; i <= iend; i++) c->source[I3(i, k, j)] = c->src[IS(i, k, j, c->itrc[n - 1])];

    /* explicit update (:1109-1130) */
    for (int j = js; j <= je; j++) {
      int jrow = j + joff;
      for (int k = 1; k <= km; k++) {
        twodt[k - 1] = c->c2dtts * c->dtxcel[k - 1];
        for (int i = istrt; i <= iend; i++) {
          double rhs;
          if (c->fct) {
            rhs = DIFF_Tx(i, k, j) + DIFF_Ty(i, k, j, jrow) +
                  (c->isopycmix ? DIFF_Tz_iso(i, k, j) : DIFF_Tz_plain(i, k, j)) - ADV_Tx(i, k, j) - ADV_Ty_fct(i, k, j, jrow) -
                  ADV_Tz(i, k, j) + c->source[I3(i, k, j)];
          } else if (c->isopycmix) {
            rhs = DIFF_Tx(i, k, j) + DIFF_Ty(i, k, j, jrow) + DIFF_Tz_iso(i, k, j) - ADV_Tx(i, k, j) - ADV_Ty_2nd(i, k, j, jrow, n) -
                  ADV_Tz(i, k, j) - ADV_Txiso(i, k, j, n) - ADV_Tyiso(i, k, j, jrow, n) - ADV_Tziso(i, k, j) + c->source[I3(i, k, j)];
          } else {
            rhs = DIFF_Tx(i, k, j) + DIFF_Ty(i, k, j, jrow) + DIFF_Tz_plain(i, k, j) - ADV_Tx(i, k, j) - ADV_Ty_2nd(i, k, j, jrow, n) -
                  ADV_Tz(i, k, j) + c->source[I3(i, k, j)];
          }
          T(i, k, j, n, TAUP1) = T(i, k, j, n, TAUM1) + twodt[k - 1] * (rhs)*TM(i, k, j);
        }
      }
    }
    memcpy(&c->texp_dump[(size_t)imt * km * jmt * (n - 1)], &T(1, 1, 1, n, TAUP1), sizeof(double) * (size_t)imt * km * jmt);

    /* implicit vertical diffusion (:1148 -> ivdift :1998 -> invtri) */
    if (c->isopycmix || c->aidif != 0.0)
      ora_invtri(c, &T(1, 1, 1, n, TAUP1), &c->stf[I2N(1, 1, n)], &c->btf[I2N(1, 1, n)], c->diff_cbt, twodt);

    for (int j = js; j <= je; j++) ora_setbcx(&T(1, 1, j, n, TAUP1), imt, km);

    /* diagt1 (:1161) on tsiperts steps */
    if (c->timavgperts) ora_diag_tbar(c, n);
  }

  /* explicit convection (:1198-1203) */
  if (c->do_convect) {
    ora_convct2(c, &T(1, 1, 1, 1, TAUP1));
    for (int j = js; j <= je; j++)
      for (int n = 1; n <= nt; n++) ora_setbcx(&T(1, 1, j, n, TAUP1), imt, km);
  }

  /* Fourier filter of the polar rows (:1245-1257) */
  if (c->do_filter) {
    ora_filt(c);
    for (int n = 1; n <= nt; n++)
      for (int j = js; j <= je; j++) ora_setbcx(&T(1, 1, j, n, TAUP1), imt, km);
  }
}

/* source/mom/mom.F:340-389: isopyc -> vmixc -> tracer */
void ora_step(ora_ctx *c) {
  if (c->isopycmix) ora_isopyc(c);
  ora_vmixc(c);
  ora_tracer(c);
}
