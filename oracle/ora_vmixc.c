/*
 * ora_vmixc.c -- restatement of 09/mom/vmixc.F:68-188 with O_constvmix O_tidal_kv
 * O_isopycmix (run/mk.in): diff_cbt = background kappa_h + Simmons et al. tidal mixing
 * (four constituents, Schmittner & Egbert 2013) + the K33 isopycnal component.
 * Called as mom does: vmixc(joff=0, js=1, je=jmt, is=2, ie=imt-1) (source/mom/mom.F:347).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include "oracle.h"
#include "ora_index.h"

#define ALPHAI(i, k, j) c->alphai[I3(i, k, j)]
#define BETAI(i, k, j) c->betai[I3(i, k, j)]
#define DDZT(i, k, j, n) c->ddzt[I4Z(i, k, j, n)]
/* 09/common/isopyc.h:135-136 */
#define DRODZB(i, k, j, kr) (ALPHAI(i, (k) + (kr), j) * DDZT(i, k, j, 1) + BETAI(i, (k) + (kr), j) * DDZT(i, k, j, 2))

void ora_vmixc(ora_ctx *c) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  const int js = 1, je = jmt, is = 2, ie = imt - 1;
  int istrt = imax(2, is), iend = imin(imt - 1, ie);
  int jstrt = imax(2, js - 1), jend = je - 1;

  for (int j = jstrt; j <= jend; j++) {
    int jrow = j;
    for (int i = istrt; i <= iend; i++) {
      double qk1, qo1, q2;
      if (c->tidal_kv) {
        /* :85-97 */
        if (fabs(c->tlat[I2(i, jrow)]) < 30.) { qk1 = 0.33; qo1 = 0.33; } else { qk1 = 1.; qo1 = 1.; }
        if (fabs(c->tlat[I2(i, jrow)]) < 70.) q2 = 0.33; else q2 = 1.;
      } else { qk1 = qo1 = q2 = 0.0; }
      for (int k = 1; k <= c->kmt[I2(i, jrow)] - 1; k++) {
        if (c->tidal_kv) {
          /* N^2 on the bottom face of the cell (:106-108) */
          double ZN2 = dmax(-c->gravrho0r * DRODZB(i, k, j, 0), 1e-8);
          /* sum over all levels below k (:110-118) */
          double edr = 0.;
          for (int k1 = k + 1; k1 <= c->kmt[I2(i, jrow)]; k1++) {
            double hab = c->zw[k - 1] - c->zw[k1 - 1];
            edr = edr + (q2 * (c->edrm2[I3(i, k1, jrow)] + c->edrs2[I3(i, k1, jrow)]) + qk1 * c->edrk1[I3(i, k1, jrow)] +
                         qo1 * c->edro1[I3(i, k1, jrow)]) *
                            exp(hab * c->zetar) / (1 - exp(-c->zetar * c->zw[k1 - 1]));
          }
          double zkappa = c->ogamma * edr / ZN2;
          /* :124 */
          c->diff_cbt[I3(i, k, j)] = dmax(c->kappa_h, dmin(100., zkappa + c->kappa_h));
        } else {
          c->diff_cbt[I3(i, k, j)] = c->kappa_h;
        }
      }
    }
  }
  /* add K33 (:182-188) */
  if (c->isopycmix)
    for (int j = jstrt; j <= jend; j++)
      for (int i = istrt; i <= iend; i++)
        for (int k = 1; k <= km; k++) c->diff_cbt[I3(i, k, j)] = c->diff_cbt[I3(i, k, j)] + c->K33[I3(i, k, j)];
}
