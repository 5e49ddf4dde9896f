// k_vmix.cu -- vertical diffusivity and the tracer-independent part of the implicit solve.
//
// Replaces `call vmixc` (source/mom/mom.F:347 -> 09/mom/vmixc.F:68-188):
//   diff_cbt = max(kappa_h, min(100, kappa_h + ogamma*edr/N2)) + K33
// with edr the Simmons et al. tidal sum over all levels below k.  The geometry factors
// exp(hab*zetar) and 1-exp(-zetar*zw(k1)) of 09/mom/vmixc.F:103 are time invariant and
// come from host tables (edr_e1, edr_den), as does the latitude-weighted constituent sum
// q2*(edrm2+edrs2)+qk1*edrk1+qo1*edro1 (edrsum), so the step itself evaluates no exp.
//
// A column pass then factorises the tridiagonal matrix of invtri
// (source/mom/invtri.F:55-100): a, c, b, e and bet depend only on diff_cbt, the masks,
// tdt and aidif, not on the tracer, so they are built once per step (the reference
// rebuilds them for each of the nt tracers) and the per-tracer solve is two sweeps.
#include "ctx.h"

// diff_cbt, one thread per cell: the tidal sum of a cell runs over the levels below it in the reference's order
__global__ void __launch_bounds__(256) k_vmix_cbt(const DevView v) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ni = v.imt - 2, km = v.km;
  const int nrow = v.jhi - v.jlo + 1;
  if (idx >= (long long)ni * km * nrow) return;
  const int i = (int)(idx % ni) + 2;
  const long long r = idx / ni;
  const int k = (int)(r % km) + 1;
  const int j = (int)(r / km) + v.jlo;
  const int kb = v.kmt[X2(i, j)];
  const long long c = X3(i, k, j);
  // ---- diff_cbt (09/mom/vmixc.F:68-124, 182-188) ----
  double d = 0.0;  // diff_cbt(i,k>=kmt,j) is never assigned in the reference: zero COMMON
  if (k <= kb - 1) {
    if (v.tidal_kv) {
      double drodzb = v.alphai[c] * v.ddzt[X3Z(i, k, j)] + v.betai[c] * v.ddzt[X3Z(i, k, j) + v.n3z];
      double zn2 = fmax(-v.gravrho0r * drodzb, 1e-8);
      double edr = 0.;
      for (int k1 = k + 1; k1 <= kb; k1++)
        edr = edr + qdiv(v.edrsum[c + (long long)(k1 - k) * v.imt] * v.edr_e1[(k - 1) + km * (k1 - 1)], v.edr_den[k1 - 1]);
      double zkappa = qdiv(v.ogamma * edr, zn2);
      d = fmax(v.kappa_h, fmin(100., zkappa + v.kappa_h));
    } else {
      d = v.kappa_h;
    }
  }
  if (v.isopycmix) d = d + v.K33[c];
  v.diff_cbt[c] = d;
}

// the tracer-independent Thomas factors, one thread per column
__global__ void __launch_bounds__(128) k_vmix_factor(const DevView v) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int ni = v.imt - 2;
  int nrow = v.jhi - v.jlo + 1;
  if (idx >= (long long)ni * nrow) return;
  int i = (int)(idx % ni) + 2;
  int j = (int)(idx / ni) + v.jlo;
  const int km = v.km;
  const int kb = v.kmt[X2(i, j)];

  // ---- invtri factorisation (source/mom/invtri.F:55-100) ----
  // The loads of a chunk of levels are issued together, ahead of the (serial, divide-bound) recurrence: with ~2 warps
  // per SM on the 100x100 grid the kernel is otherwise one memory round trip per level.
  const double eps = 1.e-30;
  const double *__restrict__ dcb = v.diff_cbt;
  double *__restrict__ ta = v.tri_a, *__restrict__ te = v.tri_e, *__restrict__ tb = v.tri_bet;
  const int c1 = (int)X3(i, 1, j), sk = v.imt;
  double bet = 0.0, cprev = 0.0;
  double dprev = dcb[c1];   // diff_cbt(i,max(1,k-1),j) of level 1
  for (int k0 = 1; k0 <= km; k0 += 8) {
    double dq[8], tq[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int k = min(k0 + q, km);
      dq[q] = dcb[c1 + (k - 1) * sk];
      tq[q] = v.dtxcel[k - 1];
    }
    asm volatile("" ::: "memory");   // keep the chunk's loads together
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int k = k0 + q;
      if (k <= km) {
        const int kp1 = min(k + 1, km);
        double tdt = v.c2dtts * tq[q];
        double factu = v.dztur[k - 1] * tdt * v.aidif;
        double factl = v.dztlr[k - 1] * tdt * v.aidif;
        double mk = (kb >= k) ? 1.0 : 0.0, mkp1 = (kb >= kp1) ? 1.0 : 0.0;
        double a = -dprev * factu * mk;
        double cc = -dq[q] * factl * mkp1;
        if (k == 1) a = 0.0;
        if (k == km) cc = 0.0;
        double b = 1.0 - a - cc;
        double e = 0.0;
        if (k == 1) {
          bet = qdiv(mk, b + eps);
        } else {
          e = cprev * bet;
          bet = qdiv(mk, b - a * e + eps);
        }
        const int c = c1 + (k - 1) * sk;
        ta[c] = a;
        te[c] = e;
        tb[c] = bet;
        cprev = cc;
        dprev = dq[q];
      }
    }
  }
}

void launch_vmixc(uvic_b200_ctx *c) {
  DevView &v = c->v;
  long long ncol = (long long)(v.imt - 2) * (v.jhi - v.jlo + 1);
  KLAUNCH("k_vmix_cbt", k_vmix_cbt, cdiv(ncol * v.km, 256), 256, v);
  KLAUNCH("k_vmix_factor", k_vmix_factor, cdiv(ncol, 128), 128, v);
}
