/*
 * ora_core.c -- oracle context: allocation of every COMMON-block array the hot path
 * touches (zero-initialised like static COMMON storage, which the reference relies
 * on, e.g. 09/mom/vmixc.F:84,185), and a by-name registry for the tests.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"
#include "ora_mobi.h"

static void *reg(ora_ctx *c, const char *name, size_t n, int is_int) {
  void *p = calloc(n ? n : 1, is_int ? sizeof(int32_t) : sizeof(double));
  if (!p || c->narr >= ORA_MAXARR) { fprintf(stderr, "oracle: alloc %s failed\n", name); abort(); }
  c->arr[c->narr].name = name;
  c->arr[c->narr].ptr = p;
  c->arr[c->narr].nelem = n;
  c->arr[c->narr].is_int = is_int;
  c->narr++;
  return p;
}
#define RD(field, n) c->field = (double *)reg(c, #field, (size_t)(n), 0)
#define RI(field, n) c->field = (int32_t *)reg(c, #field, (size_t)(n), 1)

ora_ctx *ora_create(int imt, int jmt, int km, int nt, int nsrc) {
  ora_ctx *c = (ora_ctx *)calloc(1, sizeof(ora_ctx));
  c->imt = imt; c->jmt = jmt; c->km = km; c->nt = nt; c->nsrc = nsrc;
  size_t ij = (size_t)imt * jmt, n3 = ij * km, n3z = ij * (km + 1);
  c->fct = 1; c->isopycmix = 1; c->tidal_kv = 1; c->do_convect = 1;

  RI(kmt, ij); RI(itrc, nt); RI(mskhr, ij);
  RD(dxt, imt); RD(dxtr, imt); RD(dxt2r, imt); RD(dxt4r, imt); RD(dxu, imt); RD(dxur, imt);
  RD(dyt, jmt); RD(dytr, jmt); RD(dyt2r, jmt); RD(dyt4r, jmt); RD(dyu, jmt); RD(dyur, jmt);
  RD(cst, jmt); RD(cstr, jmt); RD(csu, jmt); RD(csur, jmt);
  RD(cstdytr, jmt); RD(cstdyt2r, jmt); RD(csu_dyur, jmt);
  RD(dzt, km); RD(dztr, km); RD(dzt2r, km); RD(dztur, km); RD(dztlr, km); RD(zt, km); RD(zw, km);
  RD(dzw, km + 1); RD(dzwr, km + 1);
  RD(dtxcel, km); RD(dtxsqr, km); RD(dztxcl, km); RD(dzwxcl, km);
  RD(tlat, ij);
  RD(duw, imt); RD(due, imt); RD(dus, jmt); RD(dun, jmt);
  RD(eosc, (size_t)km * 9); RD(to, km); RD(so, km);

  RD(t, n3 * nt * 3); RD(u, n3 * 2); RD(rho, n3);
  RD(tmask, n3); RD(umask, n3);
  RD(adv_vet, n3); RD(adv_vnt, n3); RD(adv_vbt, n3z);
  RD(stf, ij * nt); RD(btf, ij * nt);
  RD(src, n3 * (nsrc > 0 ? nsrc : 1));

  RD(alphai, n3); RD(betai, n3);
  RD(ddxt, n3 * 2); RD(ddyt, n3 * 2); RD(ddzt, n3z * 2);
  RD(Ai_ez, n3 * 4); RD(Ai_nz, n3 * 4); RD(Ai_bx, n3 * 4); RD(Ai_by, n3 * 4);
  RD(K11, n3); RD(K22, n3); RD(K33, n3);
  RD(fisop, n3); RD(addisop, n3);
  RD(adv_vetiso, n3); RD(adv_vntiso, n3); RD(adv_vbtiso, n3z); RD(adv_fbiso, n3z);
  RD(drodxte, n3); RD(drodxbe, n3); RD(drodytn, n3); RD(drodybn, n3);
  RD(drodzte, n3); RD(drodzbe, n3); RD(drodztn, n3); RD(drodzbn, n3);

  RD(diff_cbt, n3);
  RD(edrm2, n3); RD(edrs2, n3); RD(edrk1, n3); RD(edro1, n3);

  RD(adv_fe, n3); RD(adv_fn, n3); RD(adv_fb, n3z);
  RD(diff_fe, n3); RD(diff_fn, n3); RD(diff_fb, n3z); RD(diff_fbiso, n3z);
  RD(source, n3);
  RD(anti_fe, n3); RD(anti_fn, n3); RD(anti_fb, n3z);
  RD(R_plusY, n3); RD(R_minusY, n3);
  RD(cstdxtr, ij); RD(cstdxt2r, ij); RD(cstdxur, ij); RD(ah_cstdxur, ij);
  RD(t_lo_dump, n3); RD(texp_dump, n3 * nt);

  RD(tbar, (size_t)km * nt * jmt); RD(travar, (size_t)km * nt * jmt); RD(dtabs, (size_t)km * nt * jmt);
  RD(sumbk, (size_t)3 * km * nt);
  RD(totalk, ij); RD(vdepth, ij); RD(pe, ij);

  RD(dnswr, ij); RD(aice, ij); RD(hice, ij); RD(hsno, ij);
  RD(sg_bathy, n3); RD(fe_hydr, n3); RD(fe_atmdep, ij * 12);
  RD(spbuf_t, n3 * nt); RD(avg_t, n3 * nt); RD(spbuf2_stf, ij * nt); RD(avg_stf, ij * nt);
  RD(vflux, ij); RD(gaost, nt);
  c->numsbc = 2 * nt + 4;
  RD(sbc, ij * c->numsbc); RD(bhf, ij);
  RI(sbc_flx_index, nt); RI(trsbcindex, nt); RI(gas_idx, 15);
  c->ntspos = 1;
  RI(kmu, ij); RD(um1, n3 * 2); RD(up1, n3 * 2);
  RD(adv_veu, n3); RD(adv_vnu, n3); RD(adv_vbu, n3z);
  RD(smf, ij * 2); RD(bmf, ij * 2); RD(hr, ij); RD(cori, ij * 2);
  RD(advmet, jmt * 2); RD(am3, jmt); RD(am4, jmt * 2); RD(dxmetr, imt); RD(dxu2r, imt);
  RD(dyu2r, jmt); RD(dyu4r, jmt); RD(csudyu2r, jmt);
  RD(visc_ceu, n3); RD(amc_north, n3); RD(amc_south, n3); RD(visc_cbu, n3);
  RD(grad_p, n3 * 2); RD(zu, ij * 2); RD(baru, ij * 2);
  RD(csudxur, ij); RD(csudxu2r, ij); RD(am_csudxtr, n3); RD(tempik, n3);
  c->itaux = 1; c->itauy = 2;
  RD(spsin, imt); RD(spcos, imt); RD(phi, jmt);
  RI(mobi_idx, ORA_MOBI_NIDX);
  c->mobi = (ora_mobi_par *)calloc(1, sizeof(ora_mobi_par));
  /* the mobi parameter block is exposed as a flat double array for the tests */
  c->arr[c->narr].name = "mobi_par"; c->arr[c->narr].ptr = c->mobi;
  c->arr[c->narr].nelem = sizeof(ora_mobi_par) / sizeof(double); c->arr[c->narr].is_int = 0;
  c->narr++;
  return c;
}

void ora_destroy(ora_ctx *c) {
  if (!c) return;
  for (int a = 0; a < c->narr; a++) free(c->arr[a].ptr);
  free(c->filt_state);
  free(c->filtu_state);
  free(c);
}

void *ora_array(ora_ctx *c, const char *name, size_t *nelem, int *is_int) {
  for (int a = 0; a < c->narr; a++)
    if (strcmp(c->arr[a].name, name) == 0) {
      if (nelem) *nelem = c->arr[a].nelem;
      if (is_int) *is_int = c->arr[a].is_int;
      return c->arr[a].ptr;
    }
  return NULL;
}
int ora_narrays(const ora_ctx *c) { return c->narr; }
const char *ora_array_name(const ora_ctx *c, int idx) { return c->arr[idx].name; }

#define SCALARS(X) \
  X(dtts) X(c2dtts) X(aidif) X(kappa_h) X(ahisop) X(athkdf) X(slmxr) X(diff_cet) X(diff_cnt) \
  X(zetar) X(ogamma) X(gravrho0r) X(relyr) X(co2ccn) X(dc13ccn) X(dc14ccn) \
  X(c2dtuv) X(kappa_m) X(cdbot) X(grav_rho0r)
#define ISCALARS(X) \
  X(fct) X(isopycmix) X(tidal_kv) X(do_convect) X(do_mobi) X(timavgperts) X(do_filter) \
  X(jfrst) X(jft1) X(jft2) X(jft0) X(eots) X(osegs) X(osege) X(ntspos) X(navgts) X(itaux) X(itauy) X(jfu0) X(jfu1) X(jfu2)

int ora_set_scalar(ora_ctx *c, const char *name, double v) {
#define X(f) if (strcmp(name, #f) == 0) { c->f = v; return 0; }
  SCALARS(X)
#undef X
#define X(f) if (strcmp(name, #f) == 0) { c->f = (int)v; return 0; }
  ISCALARS(X)
#undef X
  return -1;
}
double ora_get_scalar(ora_ctx *c, const char *name) {
#define X(f) if (strcmp(name, #f) == 0) return c->f;
  SCALARS(X)
#undef X
#define X(f) if (strcmp(name, #f) == 0) return (double)c->f;
  ISCALARS(X)
#undef X
  return 0.0 / 0.0;
}
