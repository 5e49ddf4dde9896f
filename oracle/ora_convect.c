/*
 * ora_convect.c -- restatement of convct2 (source/mom/convect.F:99-311), the full
 * convective adjustment (O_fullconvect; Rahmstorf, Ocean Modelling 101), called from
 * tracer as convct2(t(1,1,1,1,taup1), joff=0, js=2, je=jmt-1, is=2, ie=imt-1, kmt)
 * (09/mom/tracer.F:1198).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include "oracle.h"
#include "ora_index.h"

void ora_convct2(ora_ctx *c, double *ts) {
  const int imt = c->imt, km = c->km, jmt = c->jmt, nt = c->nt;
  const int js = 2, je = jmt - 1, is = 2, ie = imt - 1;
  const double *cc = c->eosc, *to = c->to, *so = c->so, *dztxcl = c->dztxcl;
  const double grav = 980.6; /* source/common/pconst.h */
#define TS(i, k, j, n) ts[I3(i, k, j) + (size_t)imt * km * jmt * (size_t)((n)-1)]
#define DENS(tq, sq, k) ora_dens(cc, km, (tq), (sq), (k))
  for (int j = js; j <= je; j++) {
    int jrow = j;
    for (int i = is; i <= ie; i++) {
      int kbo = c->kmt[I2(i, jrow)];
      double ru, rl, zsm, tsm[4], tmx[4];
      int kt, kb, la, lb, chk_la, chk_lb;

      if (c->timavgperts) {
        c->totalk[I2(i, j)] = 0.0;
        c->vdepth[I2(i, j)] = 0.0;
        c->pe[I2(i, j)] = 0.0;
        for (int k = 1; k <= km; k++) {
          ru = DENS(TS(i, k, j, 1) - to[k - 1], TS(i, k, j, 2) - so[k - 1], k);
          c->pe[I2(i, j)] = c->pe[I2(i, j)] + grav * c->zt[k - 1] * ru * dztxcl[k - 1];
        }
      }

      /* search for unstable regions starting from the top (:193-298) */
      kt = 1;
      kb = 2;
      while (kt < kbo) {
        ru = DENS(TS(i, kt, j, 1) - to[kb - 1], TS(i, kt, j, 2) - so[kb - 1], kb);
        rl = DENS(TS(i, kb, j, 1) - to[kb - 1], TS(i, kb, j, 2) - so[kb - 1], kb);
        if (ru > rl) {
          /* sum the first pair found in an unstable region */
          chk_la = 1;
          chk_lb = 1;
          zsm = dztxcl[kt - 1] + dztxcl[kb - 1];
          tsm[1] = TS(i, kt, j, 1) * dztxcl[kt - 1] + TS(i, kb, j, 1) * dztxcl[kb - 1];
          tmx[1] = tsm[1] / zsm;
          tsm[2] = TS(i, kt, j, 2) * dztxcl[kt - 1] + TS(i, kb, j, 2) * dztxcl[kb - 1];
          tmx[2] = tsm[2] / zsm;

          while (chk_lb || chk_la) {
            /* check for an unstable level (lb) below kb */
            if (kb >= kbo) chk_lb = 0;
            while (chk_lb) {
              chk_lb = 0;
              lb = kb + 1;
              ru = DENS(tmx[1] - to[lb - 1], tmx[2] - so[lb - 1], lb);
              rl = DENS(TS(i, lb, j, 1) - to[lb - 1], TS(i, lb, j, 2) - so[lb - 1], lb);
              if (ru > rl) {
                kb = lb;
                zsm = zsm + dztxcl[kb - 1];
                tsm[1] = tsm[1] + TS(i, kb, j, 1) * dztxcl[kb - 1];
                tmx[1] = tsm[1] / zsm;
                tsm[2] = tsm[2] + TS(i, kb, j, 2) * dztxcl[kb - 1];
                tmx[2] = tsm[2] / zsm;
                chk_la = 1;
                if (kb < kbo) chk_lb = 1;
              }
            }
            /* check for an unstable level (la) above kt; the Rahmstorf line is active (:237) */
            chk_la = 1;
            if (kt <= 1) chk_la = 0;
            while (chk_la) {
              chk_la = 0;
              la = kt - 1;
              ru = DENS(TS(i, la, j, 1) - to[kt - 1], TS(i, la, j, 2) - so[kt - 1], kt);
              rl = DENS(tmx[1] - to[kt - 1], tmx[2] - so[kt - 1], kt);
              if (ru > rl) {
                kt = la;
                zsm = zsm + dztxcl[kt - 1];
                tsm[1] = tsm[1] + TS(i, kt, j, 1) * dztxcl[kt - 1];
                tmx[1] = tsm[1] / zsm;
                tsm[2] = tsm[2] + TS(i, kt, j, 2) * dztxcl[kt - 1];
                tmx[2] = tsm[2] / zsm;
                chk_lb = 1;
              }
            }
          }

          /* mix all tracers from kt to kb (:262-277) */
          for (int k = kt; k <= kb; k++) {
            TS(i, k, j, 1) = tmx[1];
            TS(i, k, j, 2) = tmx[2];
          }
          for (int n = 3; n <= nt; n++) {
            tsm[3] = 0.0;
            for (int k = kt; k <= kb; k++) tsm[3] = tsm[3] + TS(i, k, j, n) * dztxcl[k - 1];
            tmx[3] = tsm[3] / zsm;
            for (int k = kt; k <= kb; k++) TS(i, k, j, n) = tmx[3];
          }
          if (c->timavgperts) {
            c->totalk[I2(i, j)] = c->totalk[I2(i, j)] + (double)(kb - kt + 1);
            if (kt == 1) c->vdepth[I2(i, j)] = c->zw[kb - 1];
          }
          kt = kb + 1;
        } else {
          kt = kb;
        }
        /* continue the search for other unstable regions */
        kb = kt + 1;
      }

      if (c->timavgperts) {
        for (int k = 1; k <= km; k++) {
          ru = DENS(TS(i, k, j, 1) - to[k - 1], TS(i, k, j, 2) - so[k - 1], k);
          c->pe[I2(i, j)] = c->pe[I2(i, j)] - grav * c->zt[k - 1] * ru * dztxcl[k - 1];
        }
        c->pe[I2(i, j)] = c->pe[I2(i, j)] / c->c2dtts;
      }
    }
  }
#undef TS
#undef DENS
}
