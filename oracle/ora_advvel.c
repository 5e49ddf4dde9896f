/*
 * ora_advvel.c -- restatement of the tracer part of adv_vel (source/mom/adv_vel.F:60-131):
 * advective velocities on the east / north / bottom faces of T cells from u(tau) on the
 * B-grid.  Called as mom does: adv_vel(joff=0, js=1, je=jmt, is=2, ie=imt-1).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include "oracle.h"
#include "ora_index.h"

#define U(i, k, j, n) c->u[I4(i, k, j, n)]

void ora_adv_vel(ora_ctx *c) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  const int js = 1, je = jmt, istrt = 2, iend = imt - 1, jsmw = 2;
  /* north face, note the embedded cosine (:66-75) */
  for (int j = js; j <= je; j++) {
    int jrow = j;
    for (int k = 1; k <= km; k++)
      for (int i = istrt; i <= iend; i++)
        c->adv_vnt[I3(i, k, j)] =
            (U(i, k, j, 2) * c->dxu[i - 1] + U(i - 1, k, j, 2) * c->dxu[i - 2]) * c->csu[jrow - 1] * c->dxt2r[i - 1];
    ora_setbcx(&c->adv_vnt[I3(1, 1, j)], imt, km);
  }
  /* east face (:82-91) */
  int jstbe = imax(js, jsmw);
  for (int j = jstbe; j <= je; j++) {
    int jrow = j;
    for (int k = 1; k <= km; k++)
      for (int i = istrt - 1; i <= iend + 1; i++)
        c->adv_vet[I3(i, k, j)] =
            (U(i, k, j, 1) * c->dyu[jrow - 1] + U(i, k, j - 1, 1) * c->dyu[jrow - 2]) * c->dyt2r[jrow - 1];
  }
  /* bottom face from continuity (:97-131) */
  for (int j = jstbe; j <= je; j++) {
    int jrow = j;
    for (int i = istrt; i <= iend; i++) c->adv_vbt[I3Z(i, 0, j)] = 0.0;
    for (int k = 1; k <= km; k++)
      for (int i = istrt; i <= iend; i++)
        c->adv_vbt[I3Z(i, k, j)] = ((c->adv_vet[I3(i, k, j)] - c->adv_vet[I3(i - 1, k, j)]) * c->dxtr[i - 1] +
                                    (c->adv_vnt[I3(i, k, j)] - c->adv_vnt[I3(i, k, j - 1)]) * c->dytr[jrow - 1]) *
                                   c->cstr[jrow - 1] * c->dzt[k - 1];
    for (int k = 1; k <= km; k++)
      for (int i = istrt; i <= iend; i++) c->adv_vbt[I3Z(i, k, j)] = c->adv_vbt[I3Z(i, k, j)] + c->adv_vbt[I3Z(i, k - 1, j)];
    ora_setbcx(&c->adv_vbt[I3Z(1, 0, j)], imt, km + 1);
  }
}

/* source/mom/state.F:1-60 as called from 09/mom/loadmw.F:150-155 with t(tau): the normalised density at T cell centres
   that clinic differentiates, rho(i,k,j) = dens(t-to(k), s-so(k), k), rows 1..jmt, i = istrt-1..iend+1 = 1..imt */
void ora_state(ora_ctx *c) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  for (int j = 1; j <= jmt; j++)
    for (int k = 1; k <= km; k++)
      for (int i = 1; i <= imt; i++)
        c->rho[I3(i, k, j)] = ora_dens(c->eosc, km, c->t[IT(i, k, j, 1, TAU)] - c->to[k - 1], c->t[IT(i, k, j, 2, TAU)] - c->so[k - 1], k);
}
