// k_isopyc.cu -- Redi / Gent-McWilliams isopycnal mixing coefficients on the device.
//
// Replaces `call isopyc` (source/mom/mom.F:340 -> 09/mom/isopyc.F:466-557):
//   k_elements   elements      09/mom/isopyc.F:363-464   alpha, beta, masked T/S gradients
//   k_isocoef    ai_east/ai_north/ai_bottom  :559-921    tapered Ai, K11, K22, K33
//   k_gm_faces   isopyc_adv    :1187-1484                GM velocities on east / north faces
//   k_gm_column  isopyc_adv    :1496-1531                vertical GM velocity (column scan)
//
// Everything here is tracer independent and runs once per step.  Unlike the reference,
// which stores the 16 Ai_* arrays and re-derives drodx/(drodz+eps) for every tracer
// inside isoflux (:963-971), the coefficient kernel stores the slope-weighted products
//   ce = Ai_ez*drodxe/(drodze+eps),  cn = Ai_nz*drodyn/(drodzn+eps),
//   cbx = Ai_bx*cstr*drodxb/(drodzb+eps),  cby = Ai_by*csu*drodyb/(drodzb+eps)
// so the per-tracer flux kernel needs no divides.
#include "ctx.h"

__device__ __forceinline__ double tmask_of(const DevView &v, int i, int k, int j) {
  // 09/mom/loadmw.F:60-77: tmask(i,k,j) = 1 if kmt(i,jrow) >= k
  return (v.kmt[X2(i, j)] >= k) ? 1.0 : 0.0;
}

// source/mom/dens.h:20-22
#define EC(k, m) v.eosc[((k)-1) + v.km * ((m)-1)]
__device__ __forceinline__ double drodt_f(const DevView &v, double tq, double sq, int k) {
  return EC(k, 1) + (EC(k, 4) + EC(k, 7) * sq) * sq + (2.0 * EC(k, 3) + 2.0 * EC(k, 8) * sq + 3.0 * EC(k, 6) * tq) * tq;
}
__device__ __forceinline__ double drods_f(const DevView &v, double tq, double sq, int k) {
  return (EC(k, 4) + 2.0 * EC(k, 7) * sq + EC(k, 8) * tq) * tq + EC(k, 2) + (2.0 * EC(k, 5) + 3.0 * EC(k, 9) * sq) * sq;
}

// decode a flat thread index into (i in 2..imt-1, k in 1..km, j in jbase..jtop)
__device__ __forceinline__ bool decode_ikj(const DevView &v, long long idx, int &i, int &k, int &j) {
  int ni = v.imt - 2;
  long long tot = (long long)ni * v.km * v.jl;
  if (idx >= tot) return false;
  i = (int)(idx % ni) + 2;
  long long r = idx / ni;
  k = (int)(r % v.km) + 1;
  j = (int)(r / v.km) + v.jbase;
  return true;
}

// store with the cyclic boundary of setbcx (source/common/util.F:789-812)
__device__ __forceinline__ void store_cyc(const DevView &v, double *a, long long base_i1, int i, double val) {
  // base_i1 = index of element i=1 in this (k,j) line
  a[base_i1 + (i - 1)] = val;
  if (i == 2) a[base_i1 + (v.imt - 1)] = val;
  if (i == v.imt - 1) a[base_i1] = val;
}

__global__ void __launch_bounds__(256) k_elements(const DevView v) {
  int i, k, j;
  if (!decode_ikj(v, (long long)blockIdx.x * blockDim.x + threadIdx.x, i, k, j)) return;
  const int jtop = v.jbase + v.jl - 1;
  const double *T = v.t_m1;            // temperature is tracer 1, salinity tracer 2
  const double *S = v.t_m1 + v.n3;
  long long c = X3(i, k, j);
  long long line = X3(1, k, j);
  // alpha, beta at T-cell centres (09/mom/isopyc.F:391-402)
  double tprime = T[c] - v.to[k - 1];
  double sprime = S[c] - v.so[k - 1];
  store_cyc(v, v.alphai, line, i, drodt_f(v, tprime, sprime, k));
  store_cyc(v, v.betai, line, i, drods_f(v, tprime, sprime, k));
  // gradients at the bottom face (:408-422)
  int kp1 = min(k + 1, v.km);
  double mkp1 = tmask_of(v, i, kp1, j);
  long long linez = X3Z(1, k, j);
  long long ckp1 = X3(i, kp1, j);
  store_cyc(v, v.ddzt, linez, i, mkp1 * v.dzwr[k] * (T[c] - T[ckp1]));
  store_cyc(v, v.ddzt + v.n3z, linez, i, mkp1 * v.dzwr[k] * (S[c] - S[ckp1]));
  if (k == 1) {
    long long l0 = X3Z(1, 0, j);
    store_cyc(v, v.ddzt, l0, i, 0.0);
    store_cyc(v, v.ddzt + v.n3z, l0, i, 0.0);
  }
  double m = tmask_of(v, i, k, j);
  // gradients at the eastern face (:428-440), rows max(js-1,2)..je-1
  if (j >= 2 && j <= v.jmt - 1) {
    double me = tmask_of(v, i + 1, k, j);
    long long ce = X3(i + 1, k, j);
    store_cyc(v, v.ddxt, line, i, m * me * v.cstr[j - 1] * v.dxur[i - 1] * (T[ce] - T[c]));
    store_cyc(v, v.ddxt + v.n3, line, i, m * me * v.cstr[j - 1] * v.dxur[i - 1] * (S[ce] - S[c]));
  }
  // gradients at the northern face (:446-460), rows max(js-1,1)..je-1
  if (j <= v.jmt - 1 && j + 1 <= jtop) {
    double mn = tmask_of(v, i, k, j + 1);
    long long cn = X3(i, k, j + 1);
    store_cyc(v, v.ddyt, line, i, m * mn * v.dyur[j - 1] * (T[cn] - T[c]));
    store_cyc(v, v.ddyt + v.n3, line, i, m * mn * v.dyur[j - 1] * (S[cn] - S[c]));
  }
}

// statement functions of 09/common/isopyc.h:121-136
#define AL(i, k, j) v.alphai[X3(i, k, j)]
#define BE(i, k, j) v.betai[X3(i, k, j)]
#define DX(i, k, j, n) v.ddxt[X3(i, k, j) + ((n)-1) * v.n3]
#define DY(i, k, j, n) v.ddyt[X3(i, k, j) + ((n)-1) * v.n3]
#define DZ(i, k, j, n) v.ddzt[X3Z(i, k, j) + ((n)-1) * v.n3z]

__global__ void __launch_bounds__(256) k_isocoef(const DevView v) {
  int i, k, j;
  if (!decode_ikj(v, (long long)blockIdx.x * blockDim.x + threadIdx.x, i, k, j)) return;
  const int jtop = v.jbase + v.jl - 1;
  const double sc = 1.0 / (v.slmxr * v.dtxsqr[k - 1]);
  const double dzt4r = 0.5 * v.dzt2r[k - 1];
  const long long c = X3(i, k, j);
  const double m = tmask_of(v, i, k, j);

  // ---- east face: ai_east (09/mom/isopyc.F:559-665), rows 2..jmt-1 ----
  if (j >= 2 && j <= v.jmt - 1) {
    double Ai0 = .5 * (v.fisop[XIJK(i, j, k)] + v.fisop[XIJK(i + 1, j, k)]) * v.ahisop + v.addisop[c];
    double me = tmask_of(v, i + 1, k, j);
    double sumz = 0.0;
#pragma unroll
    for (int kr = 0; kr <= 1; kr++)
#pragma unroll
      for (int ip = 0; ip <= 1; ip++) {
        double al = AL(i + ip, k, j), be = BE(i + ip, k, j);
        double drodxe = al * DX(i, k, j, 1) + be * DX(i, k, j, 2);
        double drodze = al * DZ(i + ip, k - 1 + kr, j, 1) + be * DZ(i + ip, k - 1 + kr, j, 2);
        double den = drodze + UVIC_EPSLN;
        double sxe = fabs(qdiv(drodxe, den));
        double a;
        if (sxe > sc) {
          double r = qdiv(sc, sxe + UVIC_EPSLN);
          a = Ai0 * m * me * (r * r);
        } else {
          a = Ai0 * m * me;
        }
        sumz = sumz + v.dzw[k - 1 + kr] * a;
        v.ce[c + (ip + 2 * kr) * v.n3] = qdiv(a * drodxe, den);
      }
    v.K11[c] = dzt4r * sumz;
  }

  // ---- north face: ai_north (:667-771), rows 1..jmt-1 ----
  if (j <= v.jmt - 1 && j + 1 <= jtop) {
    double Ai0 = 0.5 * (v.fisop[XIJK(i, j, k)] + v.fisop[XIJK(i, j + 1, k)]) * v.ahisop;
    double mn = tmask_of(v, i, k, j + 1);
    double sumz = 0.0;
#pragma unroll
    for (int kr = 0; kr <= 1; kr++)
#pragma unroll
      for (int jq = 0; jq <= 1; jq++) {
        double al = AL(i, k, j + jq), be = BE(i, k, j + jq);
        double drodyn = al * DY(i, k, j, 1) + be * DY(i, k, j, 2);
        double drodzn = al * DZ(i, k - 1 + kr, j + jq, 1) + be * DZ(i, k - 1 + kr, j + jq, 2);
        double den = drodzn + UVIC_EPSLN;
        double syn = fabs(qdiv(drodyn, den));
        double a;
        if (syn > sc) {
          double r = qdiv(sc, syn + UVIC_EPSLN);
          a = Ai0 * m * mn * (r * r);
        } else {
          a = Ai0 * m * mn;
        }
        sumz = sumz + v.dzw[k - 1 + kr] * a;
        v.cn[c + (jq + 2 * kr) * v.n3] = qdiv(a * drodyn, den);
      }
    v.K22[c] = dzt4r * sumz;
  }

  // ---- bottom face: ai_bottom (:773-921), rows 2..jmt-1, k = 1..km-1 ----
  if (j >= 2 && j <= v.jmt - 1 && j - 1 >= v.jbase && j + 1 <= jtop) {
    if (k <= v.km - 1) {
      double Ai0 = 0.5 * (v.fisop[XIJK(i, j, k + 1)] + v.fisop[XIJK(i, j, k)]) * v.ahisop;
      double mb = tmask_of(v, i, k + 1, j);
      double sumx = 0.0, sumy = 0.0;
#pragma unroll
      for (int ip = 0; ip <= 1; ip++)
#pragma unroll
        for (int kr = 0; kr <= 1; kr++) {
          double al = AL(i, k + kr, j), be = BE(i, k + kr, j);
          double drodxb = al * DX(i - 1 + ip, k + kr, j, 1) + be * DX(i - 1 + ip, k + kr, j, 2);
          double drodzb = al * DZ(i, k, j, 1) + be * DZ(i, k, j, 2);
          double den = drodzb + UVIC_EPSLN;
          double sxb = fabs(qdiv(drodxb, den));
          double a;
          if (sxb > sc) {
            double r = qdiv(sc, sxb + UVIC_EPSLN);
            a = Ai0 * mb * (r * r);
          } else {
            a = Ai0 * mb;
          }
          sumx = sumx + v.dxu[i - 1 + ip - 1] * a * (sxb * sxb);
          v.cbx[c + (ip + 2 * kr) * v.n3] = qdiv(a * v.cstr[j - 1] * drodxb, den);
        }
#pragma unroll
      for (int jq = 0; jq <= 1; jq++) {
        double facty = v.csu[j - 1 + jq - 1] * v.dyu[j - 1 + jq - 1];
#pragma unroll
        for (int kr = 0; kr <= 1; kr++) {
          double al = AL(i, k + kr, j), be = BE(i, k + kr, j);
          double drodyb = al * DY(i, k + kr, j - 1 + jq, 1) + be * DY(i, k + kr, j - 1 + jq, 2);
          double drodzb = al * DZ(i, k, j, 1) + be * DZ(i, k, j, 2);
          double den = drodzb + UVIC_EPSLN;
          double syb = fabs(qdiv(drodyb, den));
          double a;
          if (syb > sc) {
            double r = qdiv(sc, syb + UVIC_EPSLN);
            a = Ai0 * mb * (r * r);
          } else {
            a = Ai0 * mb;
          }
          sumy = sumy + facty * a * (syb * syb);
          v.cby[c + (jq + 2 * kr) * v.n3] = qdiv(a * v.csu[j - 1 + jq - 1] * drodyb, den);
        }
      }
      v.K33[c] = v.dxt4r[i - 1] * sumx + v.dyt4r[j - 1] * v.cstr[j - 1] * sumy;
    } else {
      // K33(i,km,j) is never assigned in the reference (zero-initialised COMMON)
      v.K33[c] = 0.0;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        v.cbx[c + q * v.n3] = 0.0;
        v.cby[c + q * v.n3] = 0.0;
      }
    }
  }
}

// face-averaged density gradients on the north face, bottom edge (09/mom/isopyc.F:1204-1213)
__device__ __forceinline__ void gm_bn(const DevView &v, int i, int k, int j, double &dy, double &dz) {
  int kp1 = min(k + 1, v.km);
  double ab = (AL(i, k, j) + AL(i, k, j + 1) + AL(i, kp1, j) + AL(i, kp1, j + 1)) * 0.25;
  double bb = (BE(i, k, j) + BE(i, k, j + 1) + BE(i, kp1, j) + BE(i, kp1, j + 1)) * 0.25;
  dy = ab * 0.5 * (DY(i, k, j, 1) + DY(i, kp1, j, 1)) + bb * 0.5 * (DY(i, k, j, 2) + DY(i, kp1, j, 2));
  dz = ab * 0.5 * (DZ(i, k, j, 1) + DZ(i, k, j + 1, 1)) + bb * 0.5 * (DZ(i, k, j, 2) + DZ(i, k, j + 1, 2));
}
// east face, bottom edge (:1224-1233)
__device__ __forceinline__ void gm_be(const DevView &v, int i, int k, int j, double &dx, double &dz) {
  int kp1 = min(k + 1, v.km);
  double ab = (AL(i, k, j) + AL(i + 1, k, j) + AL(i, kp1, j) + AL(i + 1, kp1, j)) * 0.25;
  double bb = (BE(i, k, j) + BE(i + 1, k, j) + BE(i, kp1, j) + BE(i + 1, kp1, j)) * 0.25;
  dx = ab * 0.5 * (DX(i, k, j, 1) + DX(i, kp1, j, 1)) + bb * 0.5 * (DX(i, k, j, 2) + DX(i, kp1, j, 2));
  dz = ab * 0.5 * (DZ(i, k, j, 1) + DZ(i + 1, k, j, 1)) + bb * 0.5 * (DZ(i, k, j, 2) + DZ(i + 1, k, j, 2));
}

__global__ void __launch_bounds__(256) k_gm_faces(const DevView v) {
  int i, k, j;
  if (!decode_ikj(v, (long long)blockIdx.x * blockDim.x + threadIdx.x, i, k, j)) return;
  const int jtop = v.jbase + v.jl - 1;
  const double sc = 1.0 / (v.slmxr * v.dtxsqr[k - 1]);
  const double top_bc = (k == 1) ? 0.0 : 1.0, bot_bc = (k == v.km) ? 0.0 : 1.0;
  const int kp1 = min(k + 1, v.km);
  const long long line = X3(1, k, j);
  const double m = tmask_of(v, i, k, j), mk1 = tmask_of(v, i, kp1, j);

  // ---- meridional GM velocity on the north face (:1404-1440), rows 1..jmt-1 ----
  if (j <= v.jmt - 1 && j + 1 <= jtop) {
    double tn_y, tn_z, bn_y, bn_z;
    gm_bn(v, i, k, j, bn_y, bn_z);
    if (k > 1) {
      gm_bn(v, i, k - 1, j, tn_y, tn_z);  // top face of this cell = bottom face of the cell above (:1215-1218)
    } else {
      double at = 0.5 * (AL(i, 1, j) + AL(i, 1, j + 1));
      double bt = 0.5 * (BE(i, 1, j) + BE(i, 1, j + 1));
      tn_y = at * DY(i, 1, j, 1) + bt * DY(i, 1, j, 2);
      tn_z = at * (DZ(i, 1, j, 1) + DZ(i, 1, j + 1, 1)) * 0.5 + bt * (DZ(i, 1, j, 2) + DZ(i, 1, j + 1, 2)) * 0.5;
    }
    double Ath0 = v.athkdf * 0.5 * (v.fisop[XIJK(i, j, k)] + v.fisop[XIJK(i, j + 1, k)]);
    double stn = qdiv(-tn_y, tn_z + 0.125 * UVIC_EPSLN);
    double sbn = qdiv(-bn_y, bn_z + 0.125 * UVIC_EPSLN);
    double absstn = fabs(stn), abssbn = fabs(sbn);
    double mn = tmask_of(v, i, k, j + 1), mnk1 = tmask_of(v, i, kp1, j + 1);
    double ath_t, ath_b;
    if (absstn > sc) {
      double r = qdiv(sc, absstn + UVIC_EPSLN);
      ath_t = Ath0 * m * mn * (r * r);
    } else {
      ath_t = Ath0 * m * mn;
    }
    if (abssbn > sc) {
      double r = qdiv(sc, abssbn + UVIC_EPSLN);
      ath_b = Ath0 * mk1 * mnk1 * (r * r);
    } else {
      ath_b = Ath0 * mk1 * mnk1;
    }
    store_cyc(v, v.adv_vntiso, line, i, -(ath_t * stn * top_bc - ath_b * sbn * bot_bc) * v.dztr[k - 1] * v.csu[j - 1]);
  }

  // ---- zonal GM velocity on the east face (:1446-1484), rows 2..jmt-1 ----
  if (j >= 2 && j <= v.jmt - 1) {
    double te_x, te_z, be_x, be_z;
    gm_be(v, i, k, j, be_x, be_z);
    if (k > 1) {
      gm_be(v, i, k - 1, j, te_x, te_z);
    } else {
      double at = 0.5 * (AL(i, 1, j) + AL(i + 1, 1, j));
      double bt = 0.5 * (BE(i, 1, j) + BE(i + 1, 1, j));
      te_x = at * DX(i, 1, j, 1) + bt * DX(i, 1, j, 2);
      te_z = at * (DZ(i, 1, j, 1) + DZ(i + 1, 1, j, 1)) * 0.5 + bt * (DZ(i, 1, j, 2) + DZ(i + 1, 1, j, 2)) * 0.5;
    }
    double Ath0 = v.athkdf * 0.5 * (v.fisop[XIJK(i, j, k)] + v.fisop[XIJK(i + 1, j, k)]);
    double ste = qdiv(-te_x, te_z + 0.125 * UVIC_EPSLN);
    double sbe = qdiv(-be_x, be_z + 0.125 * UVIC_EPSLN);
    double absste = fabs(ste), abssbe = fabs(sbe);
    double me = tmask_of(v, i + 1, k, j), mek1 = tmask_of(v, i + 1, kp1, j);
    double ath_t, ath_b;
    if (absste > sc) {
      double r = qdiv(sc, absste + UVIC_EPSLN);
      ath_t = Ath0 * m * me * (r * r);
    } else {
      ath_t = Ath0 * m * me;
    }
    if (abssbe > sc) {
      double r = qdiv(sc, abssbe + UVIC_EPSLN);
      ath_b = Ath0 * mk1 * mek1 * (r * r);
    } else {
      ath_b = Ath0 * mk1 * mek1;
    }
    store_cyc(v, v.adv_vetiso, line, i, -(ath_t * ste * top_bc - ath_b * sbe * bot_bc) * v.dztr[k - 1]);
  }
}

// total (resolved + GM) velocities on the east and north faces used by the FCT kernels
// (totadv = adv_v*t + adv_v*tiso, 09/mom/tracer_adv_flx.F:498-499,512-513), one thread per cell
__global__ void __launch_bounds__(256) k_gm_total(const DevView v) {
  int i, k, j;
  if (!decode_ikj(v, (long long)blockIdx.x * blockDim.x + threadIdx.x, i, k, j)) return;
  const int jtop = v.jbase + v.jl - 1;
  const bool gm = v.isopycmix != 0;
  const long long line = X3(1, k, j);
  if (j <= v.jmt - 1 && j + 1 <= jtop)   // north face, rows 1..jmt-1
    store_cyc(v, v.vn, line, i, v.adv_vnt[line + i - 1] + (gm ? v.adv_vntiso[line + i - 1] : 0.0));
  if (j >= 2 && j <= v.jmt - 1)
    store_cyc(v, v.ue, line, i, v.adv_vet[line + i - 1] + (gm ? v.adv_vetiso[line + i - 1] : 0.0));
}

// one thread per column: vertical GM velocity by continuity + downward cumulative sum
// (09/mom/isopyc.F:1496-1531) and the total vertical velocity (09/mom/tracer_adv_flx.F:524-525)
__global__ void __launch_bounds__(128) k_gm_column(const DevView v) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int ni = v.imt - 2;
  if (idx >= (long long)ni * v.jl) return;
  int i = (int)(idx % ni) + 2;
  int j = (int)(idx / ni) + v.jbase;
  const bool gm = v.isopycmix != 0;
  if (j >= 2 && j <= v.jmt - 1 && j - 1 >= v.jbase) {
    int kb = v.kmt[X2(i, j)];
    double acc = 0.0;  // adv_vbtiso(i,0,j) = 0
    {
      long long l0 = X3Z(1, 0, j);
      if (gm) store_cyc(v, v.adv_vbtiso, l0, i, 0.0);
      store_cyc(v, v.wb, l0, i, v.adv_vbt[l0 + i - 1] + 0.0);
    }
    // loads of 8 levels at a time, ahead of the running sum (few columns per SM: latency, not bandwidth)
    const double *__restrict__ vet = v.adv_vetiso, *__restrict__ vnt = v.adv_vntiso, *__restrict__ vbt = v.adv_vbt;
    const int c1 = (int)X3(i, 1, j), sk3 = v.imt, sj3 = v.imt * v.km;
    const double cstr_j = v.cstr[j - 1], dxtr_i = v.dxtr[i - 1], dytr_j = v.dytr[j - 1];
    for (int k0 = 1; k0 <= v.km; k0 += 8) {
      double ec[8], ew[8], nc[8], ns[8], vb[8], dz[8];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int k = min(k0 + q, v.km);
        const int c = c1 + (k - 1) * sk3;
        vb[q] = vbt[X3Z(1, k, j) + i - 1];
        dz[q] = v.dzt[k - 1];
        if (gm) { ec[q] = vet[c]; ew[q] = vet[c - 1]; nc[q] = vnt[c]; ns[q] = vnt[c - sj3]; }
        else { ec[q] = ew[q] = nc[q] = ns[q] = 0.0; }
      }
    asm volatile("" ::: "memory");   // keep the chunk's loads together
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int k = k0 + q;
        if (k <= v.km) {
          long long lz = X3Z(1, k, j);
          double w = 0.0;
          if (gm && k <= v.km - 1) {
            double d = dz[q] * cstr_j * ((ec[q] - ew[q]) * dxtr_i + (nc[q] - ns[q]) * dytr_j);
            acc = d + acc;
            w = (k == kb) ? 0.0 : acc;   // adv_vbtiso(i,kmt,j) = 0 (:1521-1525); the running sum keeps going
          }
          if (gm) store_cyc(v, v.adv_vbtiso, lz, i, w);
          store_cyc(v, v.wb, lz, i, vb[q] + w);
        }
      }
    }
  }
}

// The GM velocity chain only shares k_elements' output with the Redi coefficients: it runs on its own stream, beside
// k_isocoef / vmixc / the diffusion pass (none of which fills the GPU on small grids), and the main stream joins it
// right before the first kernel that reads the total velocities.  In the per-kernel profiling pass everything stays on
// the launch stream so that every CUDA-event interval is one kernel.
static bool gm_side(uvic_b200_ctx *c) { return c->stream3 != nullptr && !c->prof_on; }

void gm_join(uvic_b200_ctx *c) {
  if (!c->gm_inflight) return;
  cudaStreamWaitEvent(c->stream, c->ev_gm, 0);
  c->gm_inflight = false;
}

// the part of isopyc that depends on t(tau-1) only
void launch_isopyc_coef(uvic_b200_ctx *c) {
  DevView &v = c->v;
  long long ncell = (long long)(v.imt - 2) * v.km * v.jl;
  if (v.isopycmix) {
    gm_join(c);   // an unread chain of an earlier call still writes the arrays k_gm_faces is about to write
    KLAUNCH("k_elements", k_elements, cdiv(ncell, 256), 256, v);
    if (gm_side(c)) {
      cudaEventRecord(c->ev_elem, c->stream);
      cudaStreamWaitEvent(c->stream3, c->ev_elem, 0);
      cudaStream_t main_stream = c->stream;
      c->stream = c->stream3;
      KLAUNCH("k_gm_faces", k_gm_faces, cdiv(ncell, 256), 256, v);
      c->stream = main_stream;
      cudaEventRecord(c->ev_gm, c->stream3);
      c->gm_inflight = true;
      KLAUNCH("k_isocoef", k_isocoef, cdiv(ncell, 256), 256, v);
    } else {
      KLAUNCH("k_isocoef", k_isocoef, cdiv(ncell, 256), 256, v);
      KLAUNCH("k_gm_faces", k_gm_faces, cdiv(ncell, 256), 256, v);
    }
  }
}
// the part that also needs the resolved advective velocities of this step; `after` (may be null) is an event the
// velocities' upload signals
void launch_isopyc_vel_after(uvic_b200_ctx *c, cudaEvent_t after) {
  DevView &v = c->v;
  long long ncol = (long long)(v.imt - 2) * v.jl;
  if (gm_side(c)) {
    // everything the main stream holds so far (an upload of the velocities on it, the previous step's readers of ue / vn /
    // wb) comes first
    cudaEventRecord(c->ev_elem, c->stream);
    cudaStreamWaitEvent(c->stream3, c->ev_elem, 0);
    if (after) cudaStreamWaitEvent(c->stream3, after, 0);
    cudaStream_t main_stream = c->stream;
    c->stream = c->stream3;
    KLAUNCH("k_gm_total", k_gm_total, cdiv(ncol * v.km, 256), 256, v);
    KLAUNCH("k_gm_column", k_gm_column, cdiv(ncol, 128), 128, v);
    c->stream = main_stream;
    cudaEventRecord(c->ev_gm, c->stream3);
    c->gm_inflight = true;
    // with the FCT the advection kernels read the TOTAL velocities ue / vn / wb only: from here on the device copies of
    // adv_vet / adv_vnt / adv_vbt may be overwritten by the next step's upload (uvic_b200_tracer_step_coupled)
    if (c->vel_free && v.fct) { cudaEventRecord(c->vel_free, c->stream3); c->vel_free_valid = true; }
  } else {
    if (after) cudaStreamWaitEvent(c->stream, after, 0);
    KLAUNCH("k_gm_total", k_gm_total, cdiv(ncol * v.km, 256), 256, v);
    KLAUNCH("k_gm_column", k_gm_column, cdiv(ncol, 128), 128, v);
    if (c->vel_free && v.fct) { cudaEventRecord(c->vel_free, c->stream); c->vel_free_valid = true; }
  }
}
void launch_isopyc_vel(uvic_b200_ctx *c) { launch_isopyc_vel_after(c, nullptr); }
void launch_isopyc(uvic_b200_ctx *c) {
  launch_isopyc_coef(c);
  launch_isopyc_vel(c);
}
