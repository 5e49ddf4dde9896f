/*
 * ora_isopyc.c -- restatement of 09/mom/isopyc.F (Redi/GM isopycnal mixing), with the
 * cpp options of run/mk.in: O_isopycmix O_gent_mcwilliams O_anisotropic_zonal_mixing,
 * small-angle tensor with the Gerdes et al. taper (O_full_tensor, O_dm_taper off).
 * Called as mom does: isopyc(joff=0, js=1, je=jmt, is=2, ie=imt-1) (source/mom/mom.F:340).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include "oracle.h"
#include "ora_index.h"

/* statement functions of 09/common/isopyc.h:121-136 */
#define ALPHAI(i, k, j) c->alphai[I3(i, k, j)]
#define BETAI(i, k, j) c->betai[I3(i, k, j)]
#define DDXT(i, k, j, n) c->ddxt[I4(i, k, j, n)]
#define DDYT(i, k, j, n) c->ddyt[I4(i, k, j, n)]
#define DDZT(i, k, j, n) c->ddzt[I4Z(i, k, j, n)]
#define DRODXE(i, k, j, ip) (ALPHAI((i) + (ip), k, j) * DDXT(i, k, j, 1) + BETAI((i) + (ip), k, j) * DDXT(i, k, j, 2))
#define DRODZE(i, k, j, ip, kr) \
  (ALPHAI((i) + (ip), k, j) * DDZT((i) + (ip), (k)-1 + (kr), j, 1) + BETAI((i) + (ip), k, j) * DDZT((i) + (ip), (k)-1 + (kr), j, 2))
#define DRODYN(i, k, j, jq) (ALPHAI(i, k, (j) + (jq)) * DDYT(i, k, j, 1) + BETAI(i, k, (j) + (jq)) * DDYT(i, k, j, 2))
#define DRODZN(i, k, j, jq, kr) \
  (ALPHAI(i, k, (j) + (jq)) * DDZT(i, (k)-1 + (kr), (j) + (jq), 1) + BETAI(i, k, (j) + (jq)) * DDZT(i, (k)-1 + (kr), (j) + (jq), 2))
#define DRODXB(i, k, j, ip, kr) \
  (ALPHAI(i, (k) + (kr), j) * DDXT((i)-1 + (ip), (k) + (kr), j, 1) + BETAI(i, (k) + (kr), j) * DDXT((i)-1 + (ip), (k) + (kr), j, 2))
#define DRODYB(i, k, j, jq, kr) \
  (ALPHAI(i, (k) + (kr), j) * DDYT(i, (k) + (kr), (j)-1 + (jq), 1) + BETAI(i, (k) + (kr), j) * DDYT(i, (k) + (kr), (j)-1 + (jq), 2))
#define DRODZB(i, k, j, kr) (ALPHAI(i, (k) + (kr), j) * DDZT(i, k, j, 1) + BETAI(i, (k) + (kr), j) * DDZT(i, k, j, 2))

#define TM(i, k, j) c->tmask[I3(i, k, j)]
#define T(i, k, j, n, l) c->t[IT(i, k, j, n, l)]

/* 09/mom/isopyc.F:363-464  subroutine elements */
static void elements(ora_ctx *c, int js, int je, int is, int ie) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  /* alpha and beta at centers of T cells  (:391-402) */
  for (int j = js; j <= je; j++) {
    for (int k = 1; k <= km; k++)
      for (int i = is; i <= ie; i++) {
        double tprime = T(i, k, j, 1, TAUM1) - c->to[k - 1];
        double sprime = T(i, k, j, 2, TAUM1) - c->so[k - 1];
        ALPHAI(i, k, j) = ora_drodt(c->eosc, km, tprime, sprime, k);
        BETAI(i, k, j) = ora_drods(c->eosc, km, tprime, sprime, k);
      }
    ora_setbcx(&ALPHAI(1, 1, j), imt, km);
    ora_setbcx(&BETAI(1, 1, j), imt, km);
  }
  /* gradients at bottom face of T cells (:408-422) */
  for (int j = js; j <= je; j++)
    for (int n = 1; n <= 2; n++) {
      for (int k = 1; k <= km; k++) {
        int kp1 = imin(k + 1, km);
        for (int i = is; i <= ie; i++)
          DDZT(i, k, j, n) = TM(i, kp1, j) * c->dzwr[k] * (T(i, k, j, n, TAUM1) - T(i, kp1, j, n, TAUM1));
      }
      for (int i = is; i <= ie; i++) DDZT(i, 0, j, n) = 0.0;
      ora_setbcx(&DDZT(1, 0, j, n), imt, km + 1);
    }
  /* gradients at eastern face of T cells (:428-440) */
  for (int j = imax(js - 1, 2); j <= je - 1; j++) {
    int jrow = j;
    for (int n = 1; n <= 2; n++) {
      for (int k = 1; k <= km; k++)
        for (int i = is; i <= ie; i++)
          DDXT(i, k, j, n) = TM(i, k, j) * TM(i + 1, k, j) * c->cstr[jrow - 1] * c->dxur[i - 1] *
                             (T(i + 1, k, j, n, TAUM1) - T(i, k, j, n, TAUM1));
      ora_setbcx(&DDXT(1, 1, j, n), imt, km);
    }
  }
  /* gradients at northern face of T cells (:446-460) */
  for (int j = imax(js - 1, 1); j <= je - 1; j++) {
    int jrow = j;
    for (int n = 1; n <= 2; n++) {
      for (int k = 1; k <= km; k++)
        for (int i = is; i <= ie; i++)
          DDYT(i, k, j, n) = TM(i, k, j) * TM(i, k, j + 1) * c->dyur[jrow - 1] *
                             (T(i, k, j + 1, n, TAUM1) - T(i, k, j, n, TAUM1));
      ora_setbcx(&DDYT(1, 1, j, n), imt, km);
    }
  }
  (void)jmt;
}

#define FISOP(i, j, k) c->fisop[IJK(i, j, k)]

/* 09/mom/isopyc.F:559-665  subroutine ai_east */
static void ai_east(ora_ctx *c, int js, int je) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  for (int j = js; j <= je; j++) {
    int jrow = j;
    for (int k = 1; k <= km; k++) {
      double sc = 1.0 / (c->slmxr * c->dtxsqr[k - 1]);
      double dzt4r = 0.5 * c->dzt2r[k - 1];
      for (int i = 2; i <= imt - 1; i++) {
        /* O_anisotropic_zonal_mixing (:595-601) */
        double Ai0 = .5 * (FISOP(i, jrow, k) + FISOP(i + 1, jrow, k)) * c->ahisop + c->addisop[I3(i, k, jrow)];
        double sumz = 0.0;
        for (int kr = 0; kr <= 1; kr++)
          for (int ip = 0; ip <= 1; ip++) {
            double sxe = fabs(DRODXE(i, k, j, ip) / (DRODZE(i, k, j, ip, kr) + EPSLN));
            double a;
            if (sxe > sc) {
              double r = sc / (sxe + EPSLN);
              a = Ai0 * TM(i, k, j) * TM(i + 1, k, j) * (r * r);
            } else {
              a = Ai0 * TM(i, k, j) * TM(i + 1, k, j);
            }
            c->Ai_ez[IA(i, k, j, ip, kr)] = a;
            sumz = sumz + c->dzw[k - 1 + kr] * a;
          }
        c->K11[I3(i, k, j)] = dzt4r * sumz;
      }
    }
    for (int b = 0; b < 4; b++) ora_setbcx(&c->Ai_ez[IA(1, 1, j, b & 1, b >> 1)], imt, km);
    ora_setbcx(&c->K11[I3(1, 1, j)], imt, km);
  }
  (void)jmt;
}

/* 09/mom/isopyc.F:667-771  subroutine ai_north */
static void ai_north(ora_ctx *c, int js, int je) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  for (int j = js; j <= je; j++) {
    int jrow = j;
    for (int k = 1; k <= km; k++) {
      double sc = 1.0 / (c->slmxr * c->dtxsqr[k - 1]);
      double dzt4r = 0.5 * c->dzt2r[k - 1];
      for (int i = 2; i <= imt - 1; i++) {
        double Ai0 = 0.5 * (FISOP(i, jrow, k) + FISOP(i, jrow + 1, k)) * c->ahisop;
        double sumz = 0.0;
        for (int kr = 0; kr <= 1; kr++)
          for (int jq = 0; jq <= 1; jq++) {
            double syn = fabs(DRODYN(i, k, j, jq) / (DRODZN(i, k, j, jq, kr) + EPSLN));
            double a;
            if (syn > sc) {
              double r = sc / (syn + EPSLN);
              a = Ai0 * TM(i, k, j) * TM(i, k, j + 1) * (r * r);
            } else {
              a = Ai0 * TM(i, k, j) * TM(i, k, j + 1);
            }
            c->Ai_nz[IA(i, k, j, jq, kr)] = a;
            sumz = sumz + c->dzw[k - 1 + kr] * a;
          }
        c->K22[I3(i, k, j)] = dzt4r * sumz;
      }
    }
    for (int b = 0; b < 4; b++) ora_setbcx(&c->Ai_nz[IA(1, 1, j, b & 1, b >> 1)], imt, km);
    ora_setbcx(&c->K22[I3(1, 1, j)], imt, km);
  }
  (void)jmt;
}

/* 09/mom/isopyc.F:773-921  subroutine ai_bottom */
static void ai_bottom(ora_ctx *c, int js, int je) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  for (int j = js; j <= je; j++) {
    int jrow = j;
    for (int k = 1; k <= km - 1; k++) {
      double sc = 1.0 / (c->slmxr * c->dtxsqr[k - 1]);
      for (int i = 2; i <= imt - 1; i++) {
        double Ai0 = 0.5 * (FISOP(i, jrow, k + 1) + FISOP(i, jrow, k)) * c->ahisop;
        /* eastward slopes at the base of T cells */
        double sumx = 0.0;
        for (int ip = 0; ip <= 1; ip++)
          for (int kr = 0; kr <= 1; kr++) {
            double sxb = fabs(DRODXB(i, k, j, ip, kr) / (DRODZB(i, k, j, kr) + EPSLN));
            double a;
            if (sxb > sc) {
              double r = sc / (sxb + EPSLN);
              a = Ai0 * TM(i, k + 1, j) * (r * r);
            } else {
              a = Ai0 * TM(i, k + 1, j);
            }
            c->Ai_bx[IA(i, k, j, ip, kr)] = a;
            sumx = sumx + c->dxu[i - 1 + ip - 1] * a * (sxb * sxb);
          }
        /* northward slopes at the base of T cells */
        double sumy = 0.0;
        for (int jq = 0; jq <= 1; jq++) {
          double facty = c->csu[jrow - 1 + jq - 1] * c->dyu[jrow - 1 + jq - 1];
          for (int kr = 0; kr <= 1; kr++) {
            double syb = fabs(DRODYB(i, k, j, jq, kr) / (DRODZB(i, k, j, kr) + EPSLN));
            double a;
            if (syb > sc) {
              double r = sc / (syb + EPSLN);
              a = Ai0 * TM(i, k + 1, j) * (r * r);
            } else {
              a = Ai0 * TM(i, k + 1, j);
            }
            c->Ai_by[IA(i, k, j, jq, kr)] = a;
            sumy = sumy + facty * a * (syb * syb);
          }
        }
        c->K33[I3(i, k, j)] = c->dxt4r[i - 1] * sumx + c->dyt4r[jrow - 1] * c->cstr[jrow - 1] * sumy;
      }
    }
    for (int b = 0; b < 4; b++) ora_setbcx(&c->Ai_bx[IA(1, 1, j, b & 1, b >> 1)], imt, km);
    for (int b = 0; b < 4; b++) ora_setbcx(&c->Ai_by[IA(1, 1, j, b & 1, b >> 1)], imt, km);
    ora_setbcx(&c->K33[I3(1, 1, j)], imt, km);
  }
  (void)jmt;
}

/* 09/mom/isopyc.F:1140-1576  subroutine isopyc_adv (Gent-McWilliams bolus velocities) */
static void isopyc_adv(ora_ctx *c, int js, int je) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  const int jsmw = 2;
  double top_bc[km + 1], bot_bc[km + 1];
  for (int k = 1; k <= km; k++) { top_bc[k] = 1.0; bot_bc[k] = 1.0; }
  top_bc[1] = 0.0;
  bot_bc[km] = 0.0;

  /* face-averaged density gradients (:1187-1241) */
  for (int j = js; j <= je; j++) {
    int jrow = j;
    for (int i = 1; i <= imt - 1; i++) {
      double at = 0.5 * (ALPHAI(i, 1, j) + ALPHAI(i, 1, j + 1));
      double bt = 0.5 * (BETAI(i, 1, j) + BETAI(i, 1, j + 1));
      c->drodytn[I3(i, 1, jrow)] = at * DDYT(i, 1, j, 1) + bt * DDYT(i, 1, j, 2);
      c->drodztn[I3(i, 1, jrow)] = at * (DDZT(i, 1, j, 1) + DDZT(i, 1, j + 1, 1)) * 0.5 +
                                   bt * (DDZT(i, 1, j, 2) + DDZT(i, 1, j + 1, 2)) * 0.5;
      /* -- zonal -- */
      at = 0.5 * (ALPHAI(i, 1, j) + ALPHAI(i + 1, 1, j));
      bt = 0.5 * (BETAI(i, 1, j) + BETAI(i + 1, 1, j));
      c->drodxte[I3(i, 1, jrow)] = at * DDXT(i, 1, j, 1) + bt * DDXT(i, 1, j, 2);
      c->drodzte[I3(i, 1, jrow)] = at * (DDZT(i, 1, j, 1) + DDZT(i + 1, 1, j, 1)) * 0.5 +
                                   bt * (DDZT(i, 1, j, 2) + DDZT(i + 1, 1, j, 2)) * 0.5;
      for (int k = 1; k <= km; k++) {
        int km1 = imax(k - 1, 1);
        int kp1 = imin(k + 1, km);
        double ab = (ALPHAI(i, k, j) + ALPHAI(i, k, j + 1) + ALPHAI(i, kp1, j) + ALPHAI(i, kp1, j + 1)) * 0.25;
        double bb = (BETAI(i, k, j) + BETAI(i, k, j + 1) + BETAI(i, kp1, j) + BETAI(i, kp1, j + 1)) * 0.25;
        c->drodybn[I3(i, k, jrow)] = ab * 0.5 * (DDYT(i, k, j, 1) + DDYT(i, kp1, j, 1)) +
                                     bb * 0.5 * (DDYT(i, k, j, 2) + DDYT(i, kp1, j, 2));
        c->drodzbn[I3(i, k, jrow)] = ab * 0.5 * (DDZT(i, k, j, 1) + DDZT(i, k, j + 1, 1)) +
                                     bb * 0.5 * (DDZT(i, k, j, 2) + DDZT(i, k, j + 1, 2));
        if (k > 1) {
          c->drodytn[I3(i, k, jrow)] = c->drodybn[I3(i, km1, jrow)];
          c->drodztn[I3(i, k, jrow)] = c->drodzbn[I3(i, km1, jrow)];
        }
        ab = (ALPHAI(i, k, j) + ALPHAI(i + 1, k, j) + ALPHAI(i, kp1, j) + ALPHAI(i + 1, kp1, j)) * 0.25;
        bb = (BETAI(i, k, j) + BETAI(i + 1, k, j) + BETAI(i, kp1, j) + BETAI(i + 1, kp1, j)) * 0.25;
        c->drodxbe[I3(i, k, jrow)] = ab * 0.5 * (DDXT(i, k, j, 1) + DDXT(i, kp1, j, 1)) +
                                     bb * 0.5 * (DDXT(i, k, j, 2) + DDXT(i, kp1, j, 2));
        c->drodzbe[I3(i, k, jrow)] = ab * 0.5 * (DDZT(i, k, j, 1) + DDZT(i + 1, k, j, 1)) +
                                     bb * 0.5 * (DDZT(i, k, j, 2) + DDZT(i + 1, k, j, 2));
        if (k > 1) {
          c->drodxte[I3(i, k, jrow)] = c->drodxbe[I3(i, km1, jrow)];
          c->drodzte[I3(i, k, jrow)] = c->drodzbe[I3(i, km1, jrow)];
        }
      }
    }
  }

  /* meridional component at the northern face (:1404-1440) */
  for (int j = js; j <= je; j++) {
    int jrow = j;
    for (int k = 1; k <= km; k++) {
      double sc = 1.0 / (c->slmxr * c->dtxsqr[k - 1]);
      int kp1 = imin(k + 1, km);
      for (int i = 1; i <= imt; i++) {
        double Ath0 = c->athkdf * 0.5 * (FISOP(i, jrow, k) + FISOP(i, jrow + 1, k));
        double stn = -c->drodytn[I3(i, k, jrow)] / (c->drodztn[I3(i, k, jrow)] + 0.125 * EPSLN);
        double sbn = -c->drodybn[I3(i, k, jrow)] / (c->drodzbn[I3(i, k, jrow)] + 0.125 * EPSLN);
        double absstn = fabs(stn), abssbn = fabs(sbn);
        double ath_t, ath_b;
        if (absstn > sc) {
          double r = sc / (absstn + EPSLN);
          ath_t = Ath0 * TM(i, k, j) * TM(i, k, j + 1) * (r * r);
        } else {
          ath_t = Ath0 * TM(i, k, j) * TM(i, k, j + 1);
        }
        if (abssbn > sc) {
          double r = sc / (abssbn + EPSLN);
          ath_b = Ath0 * TM(i, kp1, j) * TM(i, kp1, j + 1) * (r * r);
        } else {
          ath_b = Ath0 * TM(i, kp1, j) * TM(i, kp1, j + 1);
        }
        c->adv_vntiso[I3(i, k, j)] = -(ath_t * stn * top_bc[k] - ath_b * sbn * bot_bc[k]) * c->dztr[k - 1] * c->csu[jrow - 1];
      }
    }
  }

  /* zonal component at the eastern face (:1446-1484) */
  int jstrt = imax(js, jsmw);
  for (int j = jstrt; j <= je; j++) {
    int jrow = j;
    for (int k = 1; k <= km; k++) {
      double sc = 1.0 / (c->slmxr * c->dtxsqr[k - 1]);
      int kp1 = imin(k + 1, km);
      for (int i = 1; i <= imt - 1; i++) {
        double Ath0 = c->athkdf * 0.5 * (FISOP(i, jrow, k) + FISOP(i + 1, jrow, k));
        double ste = -c->drodxte[I3(i, k, jrow)] / (c->drodzte[I3(i, k, jrow)] + 0.125 * EPSLN);
        double sbe = -c->drodxbe[I3(i, k, jrow)] / (c->drodzbe[I3(i, k, jrow)] + 0.125 * EPSLN);
        double absste = fabs(ste), abssbe = fabs(sbe);
        double ath_t, ath_b;
        if (absste > sc) {
          double r = sc / (absste + EPSLN);
          ath_t = Ath0 * TM(i, k, j) * TM(i + 1, k, j) * (r * r);
        } else {
          ath_t = Ath0 * TM(i, k, j) * TM(i + 1, k, j);
        }
        if (abssbe > sc) {
          double r = sc / (abssbe + EPSLN);
          ath_b = Ath0 * TM(i, kp1, j) * TM(i + 1, kp1, j) * (r * r);
        } else {
          ath_b = Ath0 * TM(i, kp1, j) * TM(i + 1, kp1, j);
        }
        c->adv_vetiso[I3(i, k, j)] = -(ath_t * ste * top_bc[k] - ath_b * sbe * bot_bc[k]) * c->dztr[k - 1];
      }
    }
  }
  for (int j = jstrt; j <= je; j++) ora_setbcx(&c->adv_vetiso[I3(1, 1, j)], imt, km);

  /* vertical component from continuity (:1496-1531) */
  for (int j = jstrt; j <= je; j++)
    for (int i = 1; i <= imt; i++) c->adv_vbtiso[I3Z(i, 0, j)] = 0.0;
  for (int j = jstrt; j <= je; j++) {
    int jrow = j;
    for (int k = 1; k <= km - 1; k++)
      for (int i = 2; i <= imt; i++)
        c->adv_vbtiso[I3Z(i, k, j)] =
            c->dzt[k - 1] * c->cstr[jrow - 1] *
            ((c->adv_vetiso[I3(i, k, j)] - c->adv_vetiso[I3(i - 1, k, j)]) * c->dxtr[i - 1] +
             (c->adv_vntiso[I3(i, k, j)] - c->adv_vntiso[I3(i, k, j - 1)]) * c->dytr[jrow - 1]);
  }
  for (int j = jstrt; j <= je; j++)
    for (int k = 1; k <= km - 1; k++)
      for (int i = 2; i <= imt; i++)
        c->adv_vbtiso[I3Z(i, k, j)] = c->adv_vbtiso[I3Z(i, k, j)] + c->adv_vbtiso[I3Z(i, k - 1, j)];
  for (int j = jstrt; j <= je; j++) {
    int jrow = j;
    for (int i = 2; i <= imt; i++) c->adv_vbtiso[I3Z(i, c->kmt[I2(i, jrow)], j)] = 0.0;
  }
  for (int j = jstrt; j <= je; j++) ora_setbcx(&c->adv_vbtiso[I3Z(1, 0, j)], imt, km + 1);
  (void)jmt;
}

/* 09/mom/isopyc.F:466-557  subroutine isopyc */
void ora_isopyc(ora_ctx *c) {
  const int js = 1, je = c->jmt, is = 2, ie = c->imt - 1;
  elements(c, js, je, is, ie);
  ai_east(c, imax(js - 1, 2), je - 1);
  ai_north(c, imax(js - 1, 1), je - 1);
  ai_bottom(c, imax(js - 1, 2), je - 1);
  isopyc_adv(c, imax(js - 1, 1), je - 1);
}

/* 09/mom/isopyc.F:923-1138  subroutine isoflux (called with js=2, je=jmt-1 from tracer) */
void ora_isoflux(ora_ctx *c, int n) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  const int js = 2, je = jmt - 1;
  /* east face (:950-1002) */
  for (int j = js; j <= je; j++) {
    for (int k = 1; k <= km; k++) {
      double dzt4r = 0.5 * c->dzt2r[k - 1];
      for (int i = 2; i <= imt - 1; i++) {
        double sumz = 0.0;
        for (int kr = 0; kr <= 1; kr++) {
          int km1kr = imax(k - 1 + kr, 1);
          int kpkr = imin(k + kr, km);
          for (int ip = 0; ip <= 1; ip++)
            sumz = sumz - c->Ai_ez[IA(i, k, j, ip, kr)] * (T(i + ip, km1kr, j, n, TAUM1) - T(i + ip, kpkr, j, n, TAUM1)) *
                              DRODXE(i, k, j, ip) / (DRODZE(i, k, j, ip, kr) + EPSLN);
        }
        double flux_x = dzt4r * sumz;
        c->diff_fe[I3(i, k, j)] = c->diff_fe[I3(i, k, j)] +
                                  c->K11[I3(i, k, j)] * c->cstdxur[I2(i, j)] * (T(i + 1, k, j, n, TAUM1) - T(i, k, j, n, TAUM1)) +
                                  flux_x;
      }
    }
    ora_setbcx(&c->diff_fe[I3(1, 1, j)], imt, km);
  }
  /* north face (:1007-1053) */
  for (int j = js - 1; j <= je; j++) {
    int jrow = j;
    for (int k = 1; k <= km; k++) {
      double csu_dzt4r = c->csu[jrow - 1] * 0.5 * c->dzt2r[k - 1];
      for (int i = 2; i <= imt - 1; i++) {
        double sumz = 0.0;
        for (int kr = 0; kr <= 1; kr++) {
          int km1kr = imax(k - 1 + kr, 1);
          int kpkr = imin(k + kr, km);
          for (int jq = 0; jq <= 1; jq++)
            sumz = sumz - c->Ai_nz[IA(i, k, j, jq, kr)] * (T(i, km1kr, j + jq, n, TAUM1) - T(i, kpkr, j + jq, n, TAUM1)) *
                              DRODYN(i, k, j, jq) / (DRODZN(i, k, j, jq, kr) + EPSLN);
        }
        double flux_y = csu_dzt4r * sumz;
        c->diff_fn[I3(i, k, j)] = c->diff_fn[I3(i, k, j)] +
                                  c->K22[I3(i, k, j)] * c->csu_dyur[jrow - 1] * (T(i, k, j + 1, n, TAUM1) - T(i, k, j, n, TAUM1)) +
                                  flux_y;
      }
    }
    ora_setbcx(&c->diff_fn[I3(1, 1, j)], imt, km);
  }
  /* bottom face: K31, K32 explicit part (:1062-1108) */
  for (int j = js; j <= je; j++) {
    int jrow = j;
    for (int k = 1; k <= km - 1; k++)
      for (int i = 2; i <= imt - 1; i++) {
        double sumx = 0.0;
        for (int ip = 0; ip <= 1; ip++)
          for (int kr = 0; kr <= 1; kr++)
            sumx = sumx - c->Ai_bx[IA(i, k, j, ip, kr)] * c->cstr[jrow - 1] *
                              (T(i + ip, k + kr, j, n, TAUM1) - T(i - 1 + ip, k + kr, j, n, TAUM1)) *
                              DRODXB(i, k, j, ip, kr) / (DRODZB(i, k, j, kr) + EPSLN);
        double sumy = 0.0;
        for (int jq = 0; jq <= 1; jq++)
          for (int kr = 0; kr <= 1; kr++)
            sumy = sumy - c->Ai_by[IA(i, k, j, jq, kr)] * c->csu[jrow - 1 + jq - 1] *
                              (T(i, k + kr, j + jq, n, TAUM1) - T(i, k + kr, j - 1 + jq, n, TAUM1)) *
                              DRODYB(i, k, j, jq, kr) / (DRODZB(i, k, j, kr) + EPSLN);
        c->diff_fbiso[I3Z(i, k, j)] = c->dxt4r[i - 1] * sumx + c->dyt4r[jrow - 1] * c->cstr[jrow - 1] * sumy;
      }
    for (int i = 2; i <= imt - 1; i++) {
      c->diff_fbiso[I3Z(i, 0, j)] = 0.0;
      c->diff_fbiso[I3Z(i, km, j)] = 0.0;
    }
    ora_setbcx(&c->diff_fbiso[I3Z(1, 0, j)], imt, km + 1);
  }
  /* GM advective flux through the bottom face (:1110-1134); unused under O_fct because
     totadv already contains the GM velocity (09/mom/tracer.F:1117-1120) */
  for (int j = js; j <= je; j++)
    for (int k = 1; k <= km - 1; k++)
      for (int i = 2; i <= imt - 1; i++)
        c->adv_fbiso[I3Z(i, k, j)] = c->adv_vbtiso[I3Z(i, k, j)] * (T(i, k, j, n, TAUM1) + T(i, k + 1, j, n, TAUM1));
  for (int j = js; j <= je; j++)
    for (int i = 2; i <= imt - 1; i++) {
      c->adv_fbiso[I3Z(i, 0, j)] = 0.0;
      c->adv_fbiso[I3Z(i, km, j)] = 0.0;
    }
}
