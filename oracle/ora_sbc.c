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
