// k_reduce.cu -- conservation / basin-mean reductions of diagt1
// (09/mom/tracer.F:1516-1539 tbar, :1548-1565 sumbk).
//
// Deterministic by construction: one warp owns one (k, tracer, row) line, lanes stride
// over i in a fixed pattern and combine with a fixed shuffle tree; the inventory of a
// tracer is the sum of its tbar entries taken in a fixed (row, level) order by a single
// block.  No atomics, so repeated runs (and slab partitions combined in rank order)
// reproduce bit for bit.
#include "ctx.h"

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
  return x;
}

// out(k,n,jloc) over the owned rows jlo..jhi: sum_i t*dzt(k)*dxt(i)*cst*dyt*tmask
__global__ void __launch_bounds__(256) k_tbar(const DevView v, const double *t, double *out) {
  int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  int nrow = v.jhi - v.jlo + 1;
  long long nline = (long long)v.km * v.nt * nrow;
  if (warp >= nline) return;
  int k = warp % v.km + 1;
  int n0 = (warp / v.km) % v.nt;
  int j = warp / (v.km * v.nt) + v.jlo;
  double cosdyt = v.cst[j - 1] * v.dyt[j - 1];
  const double *tn = t + (long long)n0 * v.n3;
  double s = 0.0;
  for (int i = 2 + lane; i <= v.imt - 1; i += 32) {
    double m = (v.kmt[X2(i, j)] >= k) ? 1.0 : 0.0;
    double darea = v.dzt[k - 1] * v.dxt[i - 1] * cosdyt * m;
    s += tn[X3(i, k, j)] * darea;
  }
  s = warp_sum(s);
  if (lane == 0) out[warp] = s;
}

// travar(k,n,jloc) = sum_i t(tau)^2 darea and dtabs(k,n,jloc) = sum_i |t(tau+1) - t(tau-1)| darea fx, fx = 1/(c2dtts dtxcel(k))
// (09/mom/tracer.F:1521-1536); same line ownership and summation tree as k_tbar
__global__ void __launch_bounds__(256) k_travar_dtabs(const DevView v, double *travar, double *dtabs) {
  int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  int nrow = v.jhi - v.jlo + 1;
  long long nline = (long long)v.km * v.nt * nrow;
  if (warp >= nline) return;
  int k = warp % v.km + 1;
  int n0 = (warp / v.km) % v.nt;
  int j = warp / (v.km * v.nt) + v.jlo;
  const double cosdyt = v.cst[j - 1] * v.dyt[j - 1];
  const double r2dt = 1.0 / v.c2dtts;
  const double fx = r2dt / v.dtxcel[k - 1];
  const double *t0 = v.t_0 + (long long)n0 * v.n3, *tp = v.t_p1 + (long long)n0 * v.n3, *tm = v.t_m1 + (long long)n0 * v.n3;
  double s1 = 0.0, s2 = 0.0;
  for (int i = 2 + lane; i <= v.imt - 1; i += 32) {
    const double m = (v.kmt[X2(i, j)] >= k) ? 1.0 : 0.0;
    const double darea = v.dzt[k - 1] * v.dxt[i - 1] * cosdyt * m;
    const double tc = t0[X3(i, k, j)];
    s1 += tc * tc * darea;
    s2 += fabs(tp[X3(i, k, j)] - tm[X3(i, k, j)]) * darea * fx;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    travar[warp] = s1;
    dtabs[warp] = s2;
  }
}

// inv(n) = sum over (row, level) of tbar(k,n,row), fixed order, one block per tracer
__global__ void __launch_bounds__(256) k_inventory(const DevView v, const double *tbar, double *inv) {
  __shared__ double sh[256];
  int n0 = blockIdx.x;
  int nrow = v.jhi - v.jlo + 1;
  int tot = v.km * nrow;
  double s = 0.0;
  for (int e = threadIdx.x; e < tot; e += 256) {
    int k0 = e % v.km, r = e / v.km;
    s += tbar[k0 + (long long)v.km * (n0 + (long long)v.nt * r)];
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) inv[n0] = sh[0];
}

// sumbk(mask,k,n), nhreg = 3 (09/common/param.h:28): one block per (k,n)
__global__ void __launch_bounds__(256) k_sumbk(const DevView v, const double *t, double *out) {
  __shared__ double sh[3][256];
  int k = blockIdx.x % v.km + 1;
  int n0 = blockIdx.x / v.km;
  const double *tn = t + (long long)n0 * v.n3;
  int ni = v.imt - 2, nrow = v.jhi - v.jlo + 1;
  double s[3] = {0.0, 0.0, 0.0};
  for (int e = threadIdx.x; e < ni * nrow; e += 256) {
    int i = e % ni + 2, j = e / ni + v.jlo;
    int mask = v.mskhr ? v.mskhr[X2(i, j)] : 0;
    if (mask >= 1 && mask <= 3) {
      double m1 = (v.kmt[X2(i, j)] >= 1) ? 1.0 : 0.0, mk = (v.kmt[X2(i, j)] >= k) ? 1.0 : 0.0;
      double boxar = v.cst[j - 1] * v.dxt[i - 1] * v.dyt[j - 1] * m1 * 0.0001;
      s[mask - 1] += tn[X3(i, k, j)] * boxar * v.dzt[k - 1] * mk * 0.01;
    }
  }
  for (int q = 0; q < 3; q++) sh[q][threadIdx.x] = s[q];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int q = 0; q < 3; q++) sh[q][threadIdx.x] += sh[q][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    for (int q = 0; q < 3; q++) out[q + 3 * ((k - 1) + (long long)v.km * n0)] = sh[q][0];
}

void launch_inventory(uvic_b200_ctx *c, const double *t, double *out_dev) {
  DevView &v = c->v;
  long long nline = (long long)v.km * v.nt * (v.jhi - v.jlo + 1);
  KLAUNCH("k_tbar", k_tbar, cdiv(nline * 32, 256), 256, v, t, c->tbar);
  KLAUNCH("k_inventory", k_inventory, v.nt, 256, v, c->tbar, out_dev);
}
void launch_tbar(uvic_b200_ctx *c) {
  DevView &v = c->v;
  long long nline = (long long)v.km * v.nt * (v.jhi - v.jlo + 1);
  KLAUNCH("k_tbar", k_tbar, cdiv(nline * 32, 256), 256, v, v.t_0, c->tbar);
}
void launch_travar_dtabs(uvic_b200_ctx *c, double *travar, double *dtabs) {
  DevView &v = c->v;
  long long nline = (long long)v.km * v.nt * (v.jhi - v.jlo + 1);
  KLAUNCH("k_travar_dtabs", k_travar_dtabs, cdiv(nline * 32, 256), 256, v, travar, dtabs);
}
void launch_sumbk(uvic_b200_ctx *c) {
  DevView &v = c->v;
  KLAUNCH("k_sumbk", k_sumbk, v.km * v.nt, 256, v, v.t_0, c->sumbk);
}
