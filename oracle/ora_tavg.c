/*
 * ora_tavg.c -- time averages of the tracers: the tracer part of
 *   ora_avgvar  09/mom/timeavgs.F:206-375 as called from 09/mom/diag.F:138-146 with vart = t(:,:,:,:,tau)
 *   ora_avgout  09/mom/timeavgs.F:398-420 (the construction of the time means; the netCDF output stays on the host)
 * on the default averaging grid of avgset (the whole model grid: imtav = imt, kmav = km, jmtav = jmt-2, rows 2..jmt-1).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include "oracle.h"
#include "ora_index.h"

void ora_avgvar(ora_ctx *c) {
  const int imt = c->imt, jmt = c->jmt, km = c->km, nt = c->nt;
  /* javgr(jrow) != 0 for jrow = 2..jmt-1 */
  for (int j = 2; j <= jmt - 1; j++) {
    /* three dimensional data (:325-330): spbuf(i,n,jav) = spbuf(i,n,jav) + vart(cvxz(i),j,n), cvxz = identity */
    for (int n = 1; n <= nt; n++)
      for (int k = 1; k <= km; k++)
        for (int i = 1; i <= imt; i++)
          c->spbuf_t[I3(i, k, j) + (size_t)imt * km * jmt * (n - 1)] =
              c->spbuf_t[I3(i, k, j) + (size_t)imt * km * jmt * (n - 1)] + c->t[IT(i, k, j, n, TAU)];
    /* two dimensional fields (:347-362): the surface tracer flux, less the virtual flux for all but T and S */
    for (int n = 1; n <= nt; n++) {
      if (n > 2) {
        for (int i = 1; i <= imt; i++)
          c->spbuf2_stf[I2N(i, j, n)] = c->spbuf2_stf[I2N(i, j, n)] + c->stf[I2N(i, j, n)] - c->vflux[I2(i, j)] * c->gaost[n - 1];
      } else {
        for (int i = 1; i <= imt; i++) c->spbuf2_stf[I2N(i, j, n)] = c->spbuf2_stf[I2N(i, j, n)] + c->stf[I2N(i, j, n)];
      }
    }
  }
  /* integration counter, once per time step on the last row (:371) */
  c->navgts = c->navgts + 1;
}

void ora_avgout(ora_ctx *c) {
  const int imt = c->imt, jmt = c->jmt, km = c->km, nt = c->nt;
  const double rnavgt = 1.0 / c->navgts;   /* :407 */
  for (int j = 2; j <= jmt - 1; j++) {
    for (int n = 1; n <= nt; n++)
      for (int k = 1; k <= km; k++)
        for (int i = 1; i <= imt; i++)
          c->avg_t[I3(i, k, j) + (size_t)imt * km * jmt * (n - 1)] = rnavgt * c->spbuf_t[I3(i, k, j) + (size_t)imt * km * jmt * (n - 1)];
    for (int n = 1; n <= nt; n++)
      for (int i = 1; i <= imt; i++) c->avg_stf[I2N(i, j, n)] = rnavgt * c->spbuf2_stf[I2N(i, j, n)];
  }
}
