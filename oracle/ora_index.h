/*
 * ora_index.h -- Fortran-style (1-based, column-major) index macros for the oracle.
 * Each routine declares `const int imt=c->imt, km=c->km, jmt=c->jmt;` and then reads
 * exactly like the reference.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#ifndef UVIC_ORA_INDEX_H
#define UVIC_ORA_INDEX_H
#include <math.h>

/* (imt,km,jmt) */
#define I3(i, k, j) ((size_t)((i)-1) + (size_t)imt * ((size_t)((k)-1) + (size_t)km * (size_t)((j)-1)))
/* (imt,0:km,jmt) */
#define I3Z(i, k, j) ((size_t)((i)-1) + (size_t)imt * ((size_t)(k) + (size_t)(km + 1) * (size_t)((j)-1)))
/* (imt,jmt) */
#define I2(i, j) ((size_t)((i)-1) + (size_t)imt * (size_t)((j)-1))
/* (imt,jmt,km)  -- fisop, sg_bathy, fe_hydr */
#define IJK(i, j, k) ((size_t)((i)-1) + (size_t)imt * ((size_t)((j)-1) + (size_t)jmt * (size_t)((k)-1)))
/* (imt,km,jmt,2) */
#define I4(i, k, j, n) (I3(i, k, j) + (size_t)imt * km * jmt * (size_t)((n)-1))
#define I4Z(i, k, j, n) (I3Z(i, k, j) + (size_t)imt * (km + 1) * jmt * (size_t)((n)-1))
/* (imt,km,jmt,0:1,0:1) */
#define IA(i, k, j, a, b) (I3(i, k, j) + (size_t)imt * km * jmt * (size_t)((a) + 2 * (b)))
/* t(imt,km,jmt,nt,-1:1) */
#define IT(i, k, j, n, l) (I3(i, k, j) + (size_t)imt * km * jmt * ((size_t)((n)-1) + (size_t)c->nt * (size_t)((l) + 1)))
/* (imt,jmt,nt) */
#define I2N(i, j, n) (I2(i, j) + (size_t)imt * jmt * (size_t)((n)-1))
/* src(imt,km,jmt,nsrc) */
#define IS(i, k, j, m) (I3(i, k, j) + (size_t)imt * km * jmt * (size_t)((m)-1))

#define TAUM1 (-1)
#define TAU 0
#define TAUP1 1

/* source/common/pconst.h:20 */
#define EPSLN 1.0e-20

static inline double dmax(double a, double b) { return a > b ? a : b; }
static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* source/common/util.F:789-812  setbcx: cyclic boundary on the first index */
static inline void ora_setbcx(double *a, int imt_, int n) {
  for (int k = 0; k < n; k++) {
    a[(size_t)k * imt_] = a[(size_t)k * imt_ + imt_ - 2];
    a[(size_t)k * imt_ + imt_ - 1] = a[(size_t)k * imt_ + 1];
  }
}

/* source/mom/dens.h:18-22 statement functions; cc = c(km,9) column-major, k 1-based */
#define EC(k, m) cc[((k)-1) + (size_t)km * ((m)-1)]
static inline double ora_dens(const double *cc, int km, double tq, double sq, int k) {
  return (EC(k, 1) + (EC(k, 4) + EC(k, 7) * sq) * sq + (EC(k, 3) + EC(k, 8) * sq + EC(k, 6) * tq) * tq) * tq +
         (EC(k, 2) + (EC(k, 5) + EC(k, 9) * sq) * sq) * sq;
}
static inline double ora_drodt(const double *cc, int km, double tq, double sq, int k) {
  return EC(k, 1) + (EC(k, 4) + EC(k, 7) * sq) * sq + (2.0 * EC(k, 3) + 2.0 * EC(k, 8) * sq + 3.0 * EC(k, 6) * tq) * tq;
}
static inline double ora_drods(const double *cc, int km, double tq, double sq, int k) {
  return (EC(k, 4) + 2.0 * EC(k, 7) * sq + EC(k, 8) * tq) * tq + EC(k, 2) + (2.0 * EC(k, 5) + 3.0 * EC(k, 9) * sq) * sq;
}
#undef EC
#endif
