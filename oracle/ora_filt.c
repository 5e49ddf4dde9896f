/*
 * ora_filt.c -- restatement of the polar Fourier filter of the tracers (O_fourfil):
 *   findex  source/common/findex.F:1-101   ocean strips (is,ie) per filtered row and level
 *   filt    source/common/filt.F:37-115    per row / strip / level / tracer driver
 *   filtr   source/common/filtr.F:1-430    symmetric finite Fourier filter: a dense
 *           im x im array ftarr built from tabulated cosines, s' = fnorm*F*(s-mean),
 *           then the strip sum is restored
 * called from tracer as filt(joff=0, js=2, je=jmt-1) (09/mom/tracer.F:1245-1257).
 * jfrst, jft0, jft1, jft2 come from the host (source/common/setcom.F:37-40,75-85).
 * The fixed table sizes jmtfil=50, lsegf=20 (source/common/index.h:34) are allocated
 * dynamically here.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"
#include "ora_index.h"

typedef struct {
  int imt, km, jjmax, lsegf;
  int *istf, *ietf; /* (jjmax,lsegf,km) */
  /* filtr tables (COMMON /cfilt_i/, /cfilt_d/, /cfilt_r/) */
  int *icbase, *idbase, *ind, *indx;
  double *cossav, *denmsv, *cosnpi, *ftarr, *temp, *cof, *cosine, *denom, *sprime;
  int tables_ready;
} filt_state;

#define ISTF(jj, l, k) fs->istf[((jj)-1) + (size_t)fs->jjmax * (((l)-1) + (size_t)fs->lsegf * ((k)-1))]
#define IETF(jj, l, k) fs->ietf[((jj)-1) + (size_t)fs->jjmax * (((l)-1) + (size_t)fs->lsegf * ((k)-1))]

/* source/common/findex.F:17-93 with O_cyclic */
static void findex(ora_ctx *c, filt_state *fs, const int32_t *kxx, int jf1, int jf2) {
  const int imt = c->imt, jmt = c->jmt, imax = c->imt, kmax = c->km, lsegf = fs->lsegf;
  int *iis = (int *)calloc(lsegf + 3, sizeof(int)), *iie = (int *)calloc(lsegf + 3, sizeof(int));
#define KXX(i, j) kxx[I2(i, j)]
  int jj = 0;
  for (int jrow = c->jfrst; jrow <= jmt - 1; jrow++) {
    if (jrow <= jf1 || jrow >= jf2) {
      jj = jj + 1;
      for (int k = 1; k <= kmax; k++) {
        for (int l = 1; l <= lsegf + 1; l++) { iis[l] = 0; iie[l] = 0; }
        int l = 1;
        if (KXX(2, jrow) >= k) iis[1] = 2;
        for (int i = 2; i <= imax - 1; i++) {
          if (KXX(i - 1, jrow) < k && KXX(i, jrow) >= k) iis[l] = i;
          if (KXX(i, jrow) >= k && KXX(i + 1, jrow) < k) {
            if (i != iis[l] || (i == 2 && KXX(1, jrow) >= k)) {
              iie[l] = i;
              l = l + 1;
            } else {
              iis[l] = 0;
            }
          }
        }
        if (KXX(imax - 1, jrow) >= k && KXX(imax, jrow) >= k) {
          iie[l] = imax - 1;
          l = l + 1;
        }
        int lm = l - 1;
        if (lm > 1) {
          if (iis[1] == 2 && iie[lm] == imax - 1 && KXX(1, jrow) >= k) {
            iis[1] = iis[lm];
            iie[1] = iie[1] + imax - 2;
            iis[lm] = 0;
            iie[lm] = 0;
            lm = lm - 1;
          }
        }
        if (lm > lsegf) { fprintf(stderr, "oracle findex: increase lsegf\n"); abort(); }
        for (int q = 1; q <= lsegf; q++) { ISTF(jj, q, k) = iis[q]; IETF(jj, q, k) = iie[q]; }
      }
    }
  }
#undef KXX
  free(iis); free(iie);
  (void)imt;
}

/* source/common/filtr.F:1-430.  s(1:im) in place; iss > 0 reuses the ftarr of the previous call */
static void filtr(filt_state *fs, double *s /* 1-based: s[1..im] */, int im, int mm, int n, int iss) {
  const int imt = fs->imt, imtp1 = imt + 1;
  const double pi = atan(1.0) * 4.0;
  double *cossav = fs->cossav, *denmsv = fs->denmsv, *cosnpi = fs->cosnpi, *ftarr = fs->ftarr, *temp = fs->temp;
  double *cof = fs->cof, *cosine = fs->cosine, *denom = fs->denom, *sprime = fs->sprime;
  int *icbase = fs->icbase, *idbase = fs->idbase, *ind = fs->ind, *indx = fs->indx;
  const double circle[5] = {0.0, 0.0, -1.0, 0.0, 1.0};
  if (im < 1 || mm < 1 || mm > 3 || n < 0 || iss < 0) { fprintf(stderr, "oracle filtr: bad arguments\n"); abort(); }
  if (!fs->tables_ready) {
    /* :222-259, executed while `first` */
    for (int i = 1; i <= imt * 8; i++) ind[i] = i;
    int ibase = 0, jbase = 0;
    for (int imx = 1; imx <= imtp1; imx++) {
      double fimr = 1.0 / (double)imx;
      int imm1 = imx - 1;
      for (int i = 1; i <= imm1; i++) denmsv[ibase + i] = 1.0 / (1.0 - cos(pi * (double)i * fimr));
      idbase[imx] = ibase;
      ibase = ibase + imm1;
      int imqc = (imx - 1) / 2;
      for (int i = 1; i <= imqc; i++) cossav[jbase + i] = cos(pi * (double)i * fimr);
      icbase[imx] = jbase;
      jbase = jbase + imqc;
    }
    for (int imx = 1; imx <= imt; imx++) cosnpi[imx] = circle[(imx - 1) % 4 + 1];
    fs->tables_ready = 1;
  }
  /* :261-266 */
  if (mm == 2 && n == 0) {
    for (int i = 1; i <= im; i++) s[i] = 0.0;
    return;
  }
  int nmax = (mm == 1) ? n - 1 : n;
  int nmaxp1 = nmax + 1;
  double cc1 = 0.5 * (double)nmax + 0.25;
  double cc2 = (double)nmax + 0.5;
  int lcy;
  double fnorm;
  if (mm == 2) {
    lcy = 2 * (im + 1);
    fnorm = 2.0 / (double)(im + 1);
  } else {
    lcy = 2 * im;
    fnorm = 2.0 / (double)im;
  }
  int lh = lcy / 2, lhm1 = lh - 1, lqm = (lh - 1) / 2, lcyp1 = lcy + 1;
  int imx4 = im * 4, imx8 = im * 8;
  double ssum = 0.0;
  for (int i = 1; i <= im; i++) ssum = ssum + s[i];
  double fim = (double)im;
  double fimr = 1.0 / fim;
  double stemp = ssum * fimr;
  if (!(n > 1 || mm != 1)) {
    for (int i = 1; i <= im; i++) s[i] = stemp;
    return;
  }
  if (mm != 2)
    for (int i = 1; i <= im; i++) s[i] = s[i] - stemp;
  if (iss <= 0) {
    /* build the filter array (:310-377) */
    int jbase = icbase[lh];
    for (int i = 1; i <= lqm; i++) cosine[i] = cossav[jbase + i];
    for (int i = 1; i <= lqm; i++) cosine[lh - i] = -cossav[jbase + i];
    if (2 * (lqm + 1) == lh) cosine[lqm + 1] = 0.0;
    cosine[lh] = -1.0;
    for (int i = 1; i <= lh; i++) cosine[lh + i] = -cosine[i];
    int ibase = idbase[lh];
    for (int i = 1; i <= lhm1; i++) denom[i] = 0.25 * denmsv[ibase + i];
    denom[lh] = 0.125;
    for (int i = 1; i <= lhm1; i++) temp[i] = denom[lh - i];
    for (int i = 1; i <= lhm1; i++) denom[lh + i] = temp[i];
    denom[lcy] = 0.0;
    for (int i = lcyp1; i <= imx4; i++) denom[i] = denom[i - lcy];
    int fact1, fact2;
    if (mm == 3) {
      fact1 = 2 * nmax;
      fact2 = 2 * nmaxp1;
    } else {
      fact1 = nmax;
      fact2 = nmaxp1;
    }
    for (int i = 1; i <= imx4; i++) indx[i] = ind[i] * fact1;
    for (int i = 1; i <= imx4; i++) indx[imx4 + i] = ind[i] * fact2;
    int maxind = imx4 * fact2;
    int ncyc = (maxind - 1) / lcy + 1;
    int maxndx = lcy;
    if (maxndx < maxind) {
      int npwr, found = 0;
      for (npwr = 1; npwr <= ncyc + 2; npwr++) {
        maxndx = 2 * maxndx;
        if (maxndx >= maxind) { found = 1; break; }
      }
      if (!found) { fprintf(stderr, "oracle filtr: cannot reduce indices\n"); abort(); }
      for (int np = 1; np <= npwr; np++) {
        maxndx = maxndx / 2;
        for (int i = 1; i <= imx8; i++)
          if (indx[i] > maxndx) indx[i] = indx[i] - maxndx;
      }
    }
    for (int j = 1; j <= imx8; j++) cof[j] = cosine[indx[j]];
    int ioff1 = lcy, ioff2 = lcy + imx4;
    if (mm == 1) {
      for (int j = 1; j <= im; j++) {
        int joff = (j - 1) * imt;
        for (int i = 1; i <= im; i++)
          ftarr[joff + i] = (cof[i - j + ioff1] - cof[i - j + ioff2]) * denom[i - j + ioff1] +
                            (cof[i + j - 1] - cof[imx4 + i + j - 1]) * denom[i + j - 1] - 0.5;
      }
      for (int j = 1; j <= im; j++) ftarr[j * imtp1 - imt] = ftarr[j * imtp1 - imt] + cc1;
    } else if (mm == 2) {
      for (int j = 1; j <= im; j++) {
        int joff = (j - 1) * imt;
        for (int i = 1; i <= im; i++)
          ftarr[joff + i] = (cof[i - j + ioff1] - cof[i - j + ioff2]) * denom[i - j + ioff1] -
                            (cof[i + j] - cof[imx4 + i + j]) * denom[i + j];
      }
      for (int j = 1; j <= im; j++) ftarr[j * imtp1 - imt] = ftarr[j * imtp1 - imt] + cc1;
    } else {
      double genadj = (2 * n == im) ? 0.5 : 0.0;
      for (int j = 1; j <= im; j++) {
        int joff = (j - 1) * imt;
        for (int i = 1; i <= im; i++)
          ftarr[joff + i] = (2.0 * (cof[i - j + ioff1] - cof[i - j + ioff2])) * denom[2 * i - 2 * j + ioff1] - 0.5 -
                            genadj * cosnpi[i] * cosnpi[j];
      }
      for (int j = 1; j <= im; j++) ftarr[j * imtp1 - imt] = ftarr[j * imtp1 - imt] + cc2;
    }
  }
  /* apply (:379-420) */
  for (int i = 1; i <= im; i++) sprime[i] = 0.0;
  for (int i = 1; i <= im; i++) {
    int ioff = (i - 1) * imt;
    for (int j = 1; j <= im; j++) sprime[j] = sprime[j] + s[i] * ftarr[ioff + j];
  }
  for (int i = 1; i <= im; i++) sprime[i] = fnorm * sprime[i];
  if (mm == 2) {
    for (int i = 1; i <= im; i++) s[i] = sprime[i];
    return;
  }
  double ssm = 0.0;
  for (int i = 1; i <= im; i++) ssm = ssm + sprime[i];
  ssm = (ssum - ssm) * fimr;
  for (int i = 1; i <= im; i++) s[i] = ssm + sprime[i];
}

static filt_state *make_state(ora_ctx *c, const int32_t *kxx, int jf1, int jf2);
static filt_state *get_state(ora_ctx *c) {
  if (c->filt_state) return (filt_state *)c->filt_state;
  c->filt_state = make_state(c, c->kmt, c->jft1, c->jft2);
  return (filt_state *)c->filt_state;
}
static filt_state *make_state(ora_ctx *c, const int32_t *kxx, int jf1, int jf2) {
  filt_state *fs = (filt_state *)calloc(1, sizeof(filt_state));
  const int imt = c->imt, km = c->km;
  fs->imt = imt;
  fs->km = km;
  fs->jjmax = (jf1 - c->jfrst + 1) + (c->jmt - 1 - jf2 + 1);
  if (fs->jjmax < 1) fs->jjmax = 1;
  fs->lsegf = imt / 2 + 2;
  size_t nt = (size_t)fs->jjmax * fs->lsegf * km;
  fs->istf = (int *)calloc(nt, sizeof(int));
  fs->ietf = (int *)calloc(nt, sizeof(int));
  fs->icbase = (int *)calloc(imt + 3, sizeof(int));
  fs->idbase = (int *)calloc(imt + 3, sizeof(int));
  fs->ind = (int *)calloc((size_t)imt * 8 + 2, sizeof(int));
  fs->indx = (int *)calloc((size_t)imt * 8 + 2, sizeof(int));
  int imtd2 = imt / 2;
  fs->cossav = (double *)calloc((size_t)imtd2 * (imt - imtd2) + imt + 2, sizeof(double));
  fs->denmsv = (double *)calloc((size_t)imt * (imt + 1) / 2 + imt + 2, sizeof(double));
  fs->cosnpi = (double *)calloc(imt + 2, sizeof(double));
  fs->ftarr = (double *)calloc((size_t)imt * imt + imt + 2, sizeof(double));
  fs->temp = (double *)calloc((size_t)imt * 4 + 2, sizeof(double));
  fs->cof = (double *)calloc((size_t)imt * 8 + 2, sizeof(double));
  fs->cosine = (double *)calloc((size_t)imt * 8 + 2, sizeof(double));
  fs->denom = (double *)calloc((size_t)imt * 4 + 2, sizeof(double));
  fs->sprime = (double *)calloc(imt + 2, sizeof(double));
  findex(c, fs, kxx, jf1, jf2);
  return fs;
}

/* source/common/filt.F:37-115 (tracer part), filt(joff=0, js=2, je=jmt-1) */
void ora_filt(ora_ctx *c) {
  const int imt = c->imt, km = c->km, jmt = c->jmt, nt = c->nt;
  const int js = 2, je = jmt - 1, imtm1 = imt - 1, imtm2 = imt - 2;
  filt_state *fs = get_state(c);
  const int jskpt = c->jft2 - c->jft1;
  double *tempik = (double *)calloc((size_t)(imt + 2) * (km + 1), sizeof(double));
#define T(i, k, j, n, l) c->t[IT(i, k, j, n, l)]
#define TEMPIK(i, k) tempik[(i) + (size_t)(imt + 2) * (k)]
  for (int n = 1; n <= nt; n++)
    for (int j = js; j <= je; j++) ora_setbcx(&T(1, 1, j, n, TAUP1), imt, km);
  for (int j = js; j <= je; j++) {
    int jrow = j;
    if ((jrow > c->jft1 && jrow < c->jft2) || jrow < c->jfrst) continue;
    int jj = jrow - c->jfrst + 1;
    if (jrow >= c->jft2) jj = jj - jskpt + 1;
    int isave = 0, ieave = 0, im = 0, m = 0, n = 0;
    for (int l = 1; l <= fs->lsegf; l++)
      for (int k = 1; k <= km; k++)
        if (ISTF(jj, l, k) != 0) {
          int is = ISTF(jj, l, k), ie = IETF(jj, l, k);
          int iredo = 0;
          if (is != isave || ie != ieave) {
            iredo = -1;
            isave = is;
            ieave = ie;
            im = ie - is + 1;
            if (im != imtm2 || c->kmt[I2(1, jrow)] < k) {
              m = 1;
              n = (int)round((double)im * c->cst[jrow - 1] * c->cstr[c->jft0 - 1]);
            } else {
              m = 3;
              n = (int)round((double)im * c->cst[jrow - 1] * c->cstr[c->jft0 - 1] * 0.5);
            }
          }
          for (int mm = 1; mm <= nt; mm++) {
            int idx = iredo + mm;
            int ism1 = is - 1;
            int iea = ie;
            if (ie >= imt) iea = imtm1;
            for (int i = is; i <= iea; i++) TEMPIK(i - ism1, k) = T(i, k, j, mm, TAUP1);
            int ieb = 0, ii = 0;
            if (ie >= imt) {
              ieb = ie - imtm2;
              ii = imtm1 - is;
              for (int i = 2; i <= ieb; i++) TEMPIK(i + ii, k) = T(i, k, j, mm, TAUP1);
            }
            filtr(fs, &TEMPIK(0, k), im, m, n, idx);
            for (int i = is; i <= iea; i++) T(i, k, j, mm, TAUP1) = TEMPIK(i - ism1, k);
            if (ie >= imt)
              for (int i = 2; i <= ieb; i++) T(i, k, j, mm, TAUP1) = TEMPIK(i + ii, k);
          }
        }
  }
  free(tempik);
#undef T
#undef TEMPIK
}

/* source/common/filuv.F:1-205 (O_fourfil, O_cyclic): polar filter of the baroclinic velocities u(tau+1), called from
   clinic as filuv(joff=0, js=2, je=jmt-1) (09/mom/clinic.F:494-507).  Strips come from findex on kmu with jfu1, jfu2
   (source/common/setcom.F:158); the components are rotated to polar stereographic ones (spsin, spcos: setcom.F:56-71),
   filtered with m = 2 (land-bounded strip) or m = 3 (full cyclic row), rotated back; on every row that had a strip the
   vertical mean is removed again and the result masked. */
void ora_filuv(ora_ctx *c) {
  const int imt = c->imt, km = c->km, jmt = c->jmt;
  const int js = 2, je = jmt - 1, imtm1 = imt - 1, imtm2 = imt - 2;
  if (!c->filtu_state) c->filtu_state = make_state(c, c->kmu, c->jfu1, c->jfu2);
  filt_state *fs = (filt_state *)c->filtu_state;
  const int jskpu = c->jfu2 - c->jfu1;
  double *tempik = (double *)calloc((size_t)(imt + 2) * (km + 1) * 2, sizeof(double));
#define UP(i, k, j, n) c->up1[I4(i, k, j, n)]
#define TEMPIK(i, k, q) tempik[(i) + (size_t)(imt + 2) * ((k) + (size_t)(km + 1) * ((q)-1))]
  for (int n = 1; n <= 2; n++)
    for (int j = js; j <= je; j++) ora_setbcx(&UP(1, 1, j, n), imt, km);
  for (int j = js; j <= je; j++) {
    int jrow = j;
    if ((jrow > c->jfu1 && jrow < c->jfu2) || jrow < c->jfrst) continue;
    int jj = jrow - c->jfrst + 1;
    if (jrow >= c->jfu2) jj = jj - jskpu + 1;
    double fx = -1.0;
    if (c->phi[jrow - 1] > 0.0) fx = 1.0;
    int isave = 0, ieave = 0, im = 0, m = 0, n = 0;
    for (int l = 1; l <= fs->lsegf; l++)
      for (int k = 1; k <= km; k++)
        if (ISTF(jj, l, k) != 0) {
          int is = ISTF(jj, l, k), ie = IETF(jj, l, k);
          int iredo = 1;
          if (is != isave || ie != ieave) {
            iredo = 0;
            im = ie - is + 1;
            isave = is;
            ieave = ie;
            if (im != imtm2) {
              m = 2;
              n = (int)round((double)im * c->csu[jrow - 1] * c->csur[c->jfu0 - 1]);
            } else {
              m = 3;
              n = (int)round((double)im * c->csu[jrow - 1] * c->csur[c->jfu0 - 1] * 0.5);
            }
          }
          int ism1 = is - 1;
          int iea = ie;
          if (ie >= imt) iea = imtm1;
          for (int i = is; i <= iea; i++) {
            TEMPIK(i - ism1, k, 1) = -fx * UP(i, k, j, 1) * c->spsin[i - 1] - UP(i, k, j, 2) * c->spcos[i - 1];
            TEMPIK(i - ism1, k, 2) = fx * UP(i, k, j, 1) * c->spcos[i - 1] - UP(i, k, j, 2) * c->spsin[i - 1];
          }
          int ieb = 0, ii = 0;
          if (ie >= imt) {
            ieb = ie - imtm2;
            ii = imtm1 - is;
            for (int i = 2; i <= ieb; i++) {
              TEMPIK(i + ii, k, 1) = -fx * UP(i, k, j, 1) * c->spsin[i - 1] - UP(i, k, j, 2) * c->spcos[i - 1];
              TEMPIK(i + ii, k, 2) = fx * UP(i, k, j, 1) * c->spcos[i - 1] - UP(i, k, j, 2) * c->spsin[i - 1];
            }
          }
          filtr(fs, &TEMPIK(0, k, 1), im, m, n, iredo);
          filtr(fs, &TEMPIK(0, k, 2), im, m, n, 1);
          for (int i = is; i <= iea; i++) {
            UP(i, k, j, 1) = fx * (-TEMPIK(i - ism1, k, 1) * c->spsin[i - 1] + TEMPIK(i - ism1, k, 2) * c->spcos[i - 1]);
            UP(i, k, j, 2) = -TEMPIK(i - ism1, k, 1) * c->spcos[i - 1] - TEMPIK(i - ism1, k, 2) * c->spsin[i - 1];
          }
          if (ie >= imt)
            for (int i = 2; i <= ieb; i++) {
              UP(i, k, j, 1) = fx * (-TEMPIK(i + ii, k, 1) * c->spsin[i - 1] + TEMPIK(i + ii, k, 2) * c->spcos[i - 1]);
              UP(i, k, j, 2) = -TEMPIK(i + ii, k, 1) * c->spcos[i - 1] - TEMPIK(i + ii, k, 2) * c->spsin[i - 1];
            }
        }
    if (isave != 0 && ieave != 0) {
      /* :140-166: remove the vertical mean of the filtered row, then mask */
      for (int i = 1; i <= imt; i++) { TEMPIK(i, 1, 1) = 0.0; TEMPIK(i, 1, 2) = 0.0; }
      for (int k = 1; k <= km; k++)
        for (int i = 1; i <= imt; i++) {
          TEMPIK(i, 1, 1) = TEMPIK(i, 1, 1) + UP(i, k, j, 1) * c->dzt[k - 1];
          TEMPIK(i, 1, 2) = TEMPIK(i, 1, 2) + UP(i, k, j, 2) * c->dzt[k - 1];
        }
      for (int i = 1; i <= imt; i++) {
        TEMPIK(i, 1, 1) = TEMPIK(i, 1, 1) * c->hr[I2(i, jrow)];
        TEMPIK(i, 1, 2) = TEMPIK(i, 1, 2) * c->hr[I2(i, jrow)];
      }
      for (int k = 1; k <= km; k++)
        for (int i = 1; i <= imt; i++) {
          UP(i, k, j, 1) = UP(i, k, j, 1) - TEMPIK(i, 1, 1);
          UP(i, k, j, 2) = UP(i, k, j, 2) - TEMPIK(i, 1, 2);
        }
      for (int k = 1; k <= km; k++)
        for (int i = 1; i <= imt; i++) {
          UP(i, k, j, 1) = UP(i, k, j, 1) * c->umask[I3(i, k, j)];
          UP(i, k, j, 2) = UP(i, k, j, 2) * c->umask[I3(i, k, j)];
        }
    }
  }
  free(tempik);
#undef UP
#undef TEMPIK
}
