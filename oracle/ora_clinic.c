/*
 * ora_clinic.c -- restatement of the baroclinic momentum step with the options of run/mk.in
 * (O_consthmix, O_anisotropic_viscosity, O_constvmix, O_stream_function; no O_biharmonic,
 * O_implicitvmix, O_pressure_gradient_average, O_damp_inertial_oscillation, O_linearized_advection):
 *   ora_adv_vel_u   source/mom/adv_vel.F:160-250   advective velocities on the faces of U cells
 *   ora_setvbc_mom  09/mom/setvbc.F:163-208        surface stress and bottom drag
 *   ora_clinic      09/mom/clinic.F:60-560 with the statement functions of 09/mom/fdifm.h
 * Called as mom does with the memory window fully open: adv_vel(js=1, je=jmt), setvbc(js=1, je=jmt),
 * clinic(js=2, je=jmt-1), istrt=2, iend=imt-1 (source/mom/mom.F:300-390).
 * The velocity filter filuv (ora_filt.c) runs when do_filter is set (O_fourfil).  The diagnostics hooks (diagc1, diagc2)
 * and the ice coupling (isbcu, asbcu) are not part of this restatement.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include "oracle.h"
#include "ora_index.h"

#define U0(i, k, j, n) c->u[I4(i, k, j, n)]     /* tau   */
#define UM(i, k, j, n) c->um1[I4(i, k, j, n)]   /* tau-1 */
#define UP(i, k, j, n) c->up1[I4(i, k, j, n)]   /* tau+1 */

/* source/mom/adv_vel.F:160-250 */
void ora_adv_vel_u(ora_ctx *c) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  const int js = 1, je = jmt, istrt = 2, iend = imt - 1, jsmw = 2;
  /* north face of U cells (:168-186): adv_vnu = LINEAR_INTRP_Y(WT_AVG_X(adv_vnt)) */
  int jsun = imax(js, jsmw) - 1;
  for (int j = jsun; j <= je - 1; j++) {
    int jrow = j;
    double dyr = c->dytr[jrow];
    for (int k = 1; k <= km; k++)
      for (int i = istrt; i <= iend; i++)
        c->adv_vnu[I3(i, k, j)] =
            ((c->adv_vnt[I3(i, k, j)] * c->duw[i - 1] + c->adv_vnt[I3(i + 1, k, j)] * c->due[i - 1]) * c->dus[jrow] +
             (c->adv_vnt[I3(i, k, j + 1)] * c->duw[i - 1] + c->adv_vnt[I3(i + 1, k, j + 1)] * c->due[i - 1]) * c->dun[jrow - 1]) *
            dyr * c->dxur[i - 1];
    ora_setbcx(&c->adv_vnu[I3(1, 1, j)], imt, km);
  }
  /* east face (:195-219): adv_veu = LINEAR_INTRP_X(WT_AVG_Y(adv_vet)), cyclic */
  int jsube = imax(js - 1, jsmw);
  for (int j = jsube; j <= je - 1; j++) {
    int jrow = j;
    double dyr = c->dyur[jrow - 1];
    for (int k = 1; k <= km; k++)
      for (int i = istrt - 1; i <= iend; i++)
        c->adv_veu[I3(i, k, j)] =
            ((c->adv_vet[I3(i, k, j)] * c->dus[jrow - 1] + c->adv_vet[I3(i, k, j + 1)] * c->dun[jrow - 1]) * c->duw[i] +
             (c->adv_vet[I3(i + 1, k, j)] * c->dus[jrow - 1] + c->adv_vet[I3(i + 1, k, j + 1)] * c->dun[jrow - 1]) * c->due[i - 1]) *
            dyr * c->dxtr[i];
    ora_setbcx(&c->adv_veu[I3(1, 1, j)], imt, km);
  }
  /* bottom face (:226-250) */
  for (int j = jsube; j <= je - 1; j++) {
    int jrow = j;
    double dyn = c->dun[jrow - 1] * c->cst[jrow];
    double dys = c->dus[jrow - 1] * c->cst[jrow - 1];
    double dyr = c->dyur[jrow - 1] * c->csur[jrow - 1];
    for (int k = 0; k <= km; k++)
      for (int i = istrt; i <= iend; i++) {
        double asw = c->duw[i - 1] * dys;
        double anw = c->duw[i - 1] * dyn;
        double ase = c->due[i - 1] * dys;
        double ane = c->due[i - 1] * dyn;
        c->adv_vbu[I3Z(i, k, j)] = dyr * c->dxur[i - 1] *
                                   (c->adv_vbt[I3Z(i, k, j)] * asw + c->adv_vbt[I3Z(i + 1, k, j)] * ase +
                                    c->adv_vbt[I3Z(i, k, j + 1)] * anw + c->adv_vbt[I3Z(i + 1, k, j + 1)] * ane);
      }
    ora_setbcx(&c->adv_vbu[I3Z(1, 0, j)], imt, km + 1);
  }
}

/* 09/mom/setvbc.F:163-208: smf from the coupler's wind stress slots, quadratic bottom drag from u(tau-1) */
void ora_setvbc_mom(ora_ctx *c) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  const int js = 1, je = jmt, istrt = 2, iend = imt - 1;
  (void)km;
  for (int j = js; j <= je; j++) {
    int jrow = j;
    for (int i = istrt; i <= iend; i++) {
      c->smf[I2N(i, j, 1)] = c->sbc[I2N(i, jrow, c->itaux)] * c->umask[I3(i, 1, j)];
      c->smf[I2N(i, j, 2)] = c->sbc[I2N(i, jrow, c->itauy)] * c->umask[I3(i, 1, j)];
    }
  }
  for (int n = 1; n <= 2; n++) {
    if (c->cdbot == 0.0) {
      for (int j = js; j <= je; j++)
        for (int i = istrt; i <= iend; i++) c->bmf[I2N(i, j, n)] = 0.0;
    } else {
      for (int j = js; j <= je; j++) {
        int jrow = j;
        for (int i = istrt; i <= iend; i++) {
          int kz = c->kmu[I2(i, jrow)];
          if (kz != 0) {
            double uvmag = sqrt(UM(i, kz, j, 1) * UM(i, kz, j, 1) + UM(i, kz, j, 2) * UM(i, kz, j, 2));
            c->bmf[I2N(i, j, n)] = c->cdbot * UM(i, kz, j, n) * uvmag;
          } else {
            c->bmf[I2N(i, j, n)] = 0.0;
          }
        }
      }
    }
  }
  for (int n = 1; n <= 2; n++) {
    ora_setbcx(&c->smf[I2N(1, js, n)], imt, je - js + 1);
    ora_setbcx(&c->bmf[I2N(1, js, n)], imt, je - js + 1);
  }
}

/* 09/mom/vmixc.F:85 (O_constvmix): visc_cbu = kappa_m on levels 1..kmt-1, zero-initialised COMMON elsewhere */
static void vmixc_mom(ora_ctx *c) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  for (int j = 1; j <= jmt; j++)
    for (int i = 2; i <= imt - 1; i++)
      for (int k = 1; k <= c->kmt[I2(i, j)] - 1; k++) c->visc_cbu[I3(i, k, j)] = c->kappa_m;
}

void ora_clinic(ora_ctx *c) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  const int js = 2, je = jmt - 1, istrt = 2, iend = imt - 1, kmm1 = km - 1;
  const double p5 = 0.5, c0 = 0.0;
  double *adv_fe = c->adv_fe, *adv_fb = c->adv_fb, *diff_fe = c->diff_fe, *diff_fb = c->diff_fb, *grad_p = c->grad_p;
  double *csudxur = c->csudxur, *csudxu2r = c->csudxu2r, *am_csudxtr = c->am_csudxtr, *tempik = c->tempik;
  double *baru = c->baru;

  vmixc_mom(c);

  /* coefficients (:75-82) */
  for (int j = js; j <= je; j++) {
    int jrow = j;
    for (int k = 1; k <= km; k++)
      for (int i = istrt - 1; i <= iend; i++) {
        csudxur[I2(i, j)] = c->csur[jrow - 1] * c->dxur[i - 1];
        csudxu2r[I2(i, j)] = c->csur[jrow - 1] * c->dxur[i - 1] * p5;
        am_csudxtr[I3(i, k, j)] = c->visc_ceu[I3(i, k, j)] * c->csur[jrow - 1] * c->dxtr[i];
      }
  }

  /* hydrostatic pressure gradients (:119-177) */
  double grav_rho0r = c->grav_rho0r;
  for (int j = js; j <= je; j++) {
    int jrow = j;
    double fxa = grav_rho0r * c->dzw[0] * c->csur[jrow - 1];
    double fxb = grav_rho0r * c->dzw[0] * c->dyu2r[jrow - 1];
    for (int i = istrt - 1; i <= iend; i++) {
      double t1 = c->rho[I3(i + 1, 1, j + 1)] - c->rho[I3(i, 1, j)];
      double t2 = c->rho[I3(i, 1, j + 1)] - c->rho[I3(i + 1, 1, j)];
      grad_p[I4(i, 1, j, 1)] = (t1 - t2) * fxa * c->dxu2r[i - 1];
      grad_p[I4(i, 1, j, 2)] = (t1 + t2) * fxb;
    }
  }
  for (int j = js; j <= je + 1; j++)
    for (int k = 2; k <= km; k++)
      for (int i = istrt - 1; i <= iend + 1; i++) tempik[I3(i, k, j)] = c->rho[I3(i, k - 1, j)] + c->rho[I3(i, k, j)];
  for (int j = js; j <= je; j++) {
    int jrow = j;
    double fxa = grav_rho0r * c->csur[jrow - 1] * p5;
    double fxb = grav_rho0r * c->dyu4r[jrow - 1];
    for (int k = 2; k <= km; k++)
      for (int i = istrt - 1; i <= iend; i++) {
        double t1 = tempik[I3(i + 1, k, j + 1)] - tempik[I3(i, k, j)];
        double t2 = tempik[I3(i, k, j + 1)] - tempik[I3(i + 1, k, j)];
        grad_p[I4(i, k, j, 1)] = fxa * (t1 - t2) * c->dzw[k - 1] * c->dxu2r[i - 1];
        grad_p[I4(i, k, j, 2)] = fxb * (t1 + t2) * c->dzw[k - 1];
      }
  }
  for (int j = js; j <= je; j++)
    for (int k = 1; k <= kmm1; k++)
      for (int i = istrt - 1; i <= iend; i++) {
        grad_p[I4(i, k + 1, j, 1)] = grad_p[I4(i, k, j, 1)] + grad_p[I4(i, k + 1, j, 1)];
        grad_p[I4(i, k + 1, j, 2)] = grad_p[I4(i, k, j, 2)] + grad_p[I4(i, k + 1, j, 2)];
      }
  for (int j = js; j <= je; j++) {
    ora_setbcx(&grad_p[I4(1, 1, j, 1)], imt, km);
    ora_setbcx(&grad_p[I4(1, 1, j, 2)], imt, km);
  }

  /* one velocity component at a time (:185-413) */
  for (int n = 1; n <= 2; n++) {
    for (int j = js; j <= je; j++)
      for (int k = 1; k <= km; k++)
        for (int i = istrt - 1; i <= iend; i++) {
          adv_fe[I3(i, k, j)] = c->adv_veu[I3(i, k, j)] * (U0(i, k, j, n) + U0(i + 1, k, j, n));
          diff_fe[I3(i, k, j)] = am_csudxtr[I3(i, k, j)] * (UM(i + 1, k, j, n) - UM(i, k, j, n));
        }
    for (int j = js; j <= je; j++)
      for (int k = 1; k <= kmm1; k++)
        for (int i = istrt; i <= iend; i++) {
          adv_fb[I3Z(i, k, j)] = c->adv_vbu[I3Z(i, k, j)] * (U0(i, k, j, n) + U0(i, k + 1, j, n));
          diff_fb[I3Z(i, k, j)] = c->visc_cbu[I3(i, k, j)] * c->dzwr[k] * (UM(i, k, j, n) - UM(i, k + 1, j, n));
        }
    /* vertical b.c. (:305-315) */
    for (int j = js; j <= je; j++) {
      int jrow = j;
      for (int i = istrt; i <= iend; i++) {
        int kb = c->kmu[I2(i, jrow)];
        diff_fb[I3Z(i, 0, j)] = c->smf[I2N(i, j, n)];
        diff_fb[I3Z(i, kb, j)] = c->bmf[I2N(i, j, n)];
        adv_fb[I3Z(i, 0, j)] = c->adv_vbu[I3Z(i, 0, j)] * (U0(i, 1, j, n) + U0(i, 1, j, n));
        adv_fb[I3Z(i, km, j)] = c->adv_vbu[I3Z(i, km, j)] * U0(i, km, j, n);
      }
    }
    /* O_mobi defines the source term for U cells too; it is zero (:317-330) */
    for (int j = js; j <= je; j++)
      for (int k = 1; k <= km; k++)
        for (int i = istrt; i <= iend; i++) c->source[I3(i, k, j)] = c0;

    /* internal mode part of du/dt (:339-356) with 09/mom/fdifm.h */
    for (int j = js; j <= je; j++) {
      int jrow = j;
      for (int k = 1; k <= km; k++)
        for (int i = istrt; i <= iend; i++) {
          double DIFF_Ux = (diff_fe[I3(i, k, j)] - diff_fe[I3(i - 1, k, j)]) * csudxur[I2(i, j)];
          double DIFF_Uy = c->amc_north[I3(i, k, jrow)] * (UM(i, k, j + 1, n) - UM(i, k, j, n)) -
                           c->amc_south[I3(i, k, jrow)] * (UM(i, k, j, n) - UM(i, k, j - 1, n));
          double DIFF_Uz = (diff_fb[I3Z(i, k - 1, j)] - diff_fb[I3Z(i, k, j)]) * c->dztr[k - 1];
          double DIFF_metric = c->am3[jrow - 1] * UM(i, k, j, n) +
                               c->am4[(jrow - 1) + (size_t)jmt * (n - 1)] * c->dxmetr[i - 1] * (UM(i + 1, k, j, 3 - n) - UM(i - 1, k, j, 3 - n));
          double ADV_Ux = (adv_fe[I3(i, k, j)] - adv_fe[I3(i - 1, k, j)]) * csudxu2r[I2(i, j)];
          double ADV_Uy = (c->adv_vnu[I3(i, k, j)] * (U0(i, k, j, n) + U0(i, k, j + 1, n)) -
                           c->adv_vnu[I3(i, k, j - 1)] * (U0(i, k, j - 1, n) + U0(i, k, j, n))) * c->csudyu2r[jrow - 1];
          double ADV_Uz = (adv_fb[I3Z(i, k - 1, j)] - adv_fb[I3Z(i, k, j)]) * c->dzt2r[k - 1];
          double ADV_metric = c->advmet[(jrow - 1) + (size_t)jmt * (n - 1)] * U0(i, k, j, 1) * U0(i, k, j, 3 - n);
          double CORIOLIS = c->cori[I2N(i, jrow, n)] * U0(i, k, j, 3 - n);
          UP(i, k, j, n) = (DIFF_Ux + DIFF_Uy + DIFF_Uz + DIFF_metric - ADV_Ux - ADV_Uy - ADV_Uz + ADV_metric -
                            grad_p[I4(i, k, j, n)] + CORIOLIS + c->source[I3(i, k, j)]) * c->umask[I3(i, k, j)];
        }
    }
    /* vertical average of du/dt: forcing of the barotropic equation (:378-397) */
    for (int j = js; j <= je; j++)
      for (int i = istrt; i <= iend; i++) c->zu[I2N(i, j, n)] = c0;
    for (int j = js; j <= je; j++)
      for (int k = 1; k <= km; k++) {
        double fx = c->dzt[k - 1];
        for (int i = istrt; i <= iend; i++) c->zu[I2N(i, j, n)] = c->zu[I2N(i, j, n)] + UP(i, k, j, n) * fx;
      }
    for (int j = js; j <= je; j++)
      for (int i = istrt; i <= iend; i++) c->zu[I2N(i, j, n)] = c->zu[I2N(i, j, n)] * c->hr[I2(i, j)];
  }

  /* tau+1 velocities, explicit Coriolis (:440-451) */
  for (int n = 1; n <= 2; n++)
    for (int j = js; j <= je; j++)
      for (int k = 1; k <= km; k++)
        for (int i = istrt; i <= iend; i++) UP(i, k, j, n) = UM(i, k, j, n) + c->c2dtuv * UP(i, k, j, n);

  /* remove the vertical means: pure internal modes (:458-485) */
  for (int n = 1; n <= 2; n++) {
    for (int j = js; j <= je; j++)
      for (int i = istrt; i <= iend; i++) baru[I2N(i, j, n)] = c0;
    for (int j = js; j <= je; j++)
      for (int k = 1; k <= km; k++)
        for (int i = istrt; i <= iend; i++) baru[I2N(i, j, n)] = baru[I2N(i, j, n)] + UP(i, k, j, n) * c->dzt[k - 1];
    for (int j = js; j <= je; j++)
      for (int i = istrt; i <= iend; i++) baru[I2N(i, j, n)] = baru[I2N(i, j, n)] * c->hr[I2(i, j)];
    for (int j = js; j <= je; j++) {
      for (int k = 1; k <= km; k++)
        for (int i = istrt; i <= iend; i++) UP(i, k, j, n) = UP(i, k, j, n) - c->umask[I3(i, k, j)] * baru[I2N(i, j, n)];
      ora_setbcx(&UP(1, 1, j, n), imt, km);
    }
  }
  /* polar filter of the velocities (:494-507) */
  if (c->do_filter) ora_filuv(c);
  /* (:508-511) */
  for (int j = js; j <= je; j++) {
    ora_setbcx(&UP(1, 1, j, 1), imt, km);
    ora_setbcx(&UP(1, 1, j, 2), imt, km);
  }
}
