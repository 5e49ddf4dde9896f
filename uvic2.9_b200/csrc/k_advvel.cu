// k_advvel.cu -- tracer part of adv_vel (source/mom/adv_vel.F:60-131) on the device:
// advective velocities on the east / north / bottom faces of T cells from the B-grid
// velocity u(tau).  SURVEY.md section 8(f) rank 1: keeps the three velocity fields off
// the per-step host-to-device path.
#include "ctx.h"

#define UU(i, k, j, n) v.u[X3(i, k, j) + ((n)-1) * v.n3]

__global__ void __launch_bounds__(256) k_advvel_faces(const DevView v) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long tot = (long long)v.imt * v.km * v.jl;
  if (idx >= tot) return;
  int i = (int)(idx % v.imt) + 1;
  long long r = idx / v.imt;
  int k = (int)(r % v.km) + 1;
  int j = (int)(r / v.km) + v.jbase;
  // north face, note the embedded cosine (:66-75); i = 2..imt-1 then setbcx
  if (i >= 2 && i <= v.imt - 1) {
    double val = (UU(i, k, j, 2) * v.dxu[i - 1] + UU(i - 1, k, j, 2) * v.dxu[i - 2]) * v.csu[j - 1] * v.dxt2r[i - 1];
    long long line = X3(1, k, j);
    v.adv_vnt[line + i - 1] = val;
    if (i == 2) v.adv_vnt[line + v.imt - 1] = val;
    if (i == v.imt - 1) v.adv_vnt[line] = val;
  }
  // east face (:82-91); i = 1..imt, rows >= 2
  if (j >= 2 && j - 1 >= v.jbase)
    v.adv_vet[X3(i, k, j)] = (UU(i, k, j, 1) * v.dyu[j - 1] + UU(i, k, j - 1, 1) * v.dyu[j - 2]) * v.dyt2r[j - 1];
}

// bottom face from continuity, one thread per column (:97-131)
__global__ void __launch_bounds__(128) k_advvel_column(const DevView v) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int ni = v.imt - 2;
  if (idx >= (long long)ni * v.jl) return;
  int i = (int)(idx % ni) + 2;
  int j = (int)(idx / ni) + v.jbase;
  if (j < 2 || j - 1 < v.jbase) return;
  double acc = 0.0;
  for (int k = 0; k <= v.km; k++) {
    if (k >= 1) {
      double d = ((v.adv_vet[X3(i, k, j)] - v.adv_vet[X3(i - 1, k, j)]) * v.dxtr[i - 1] +
                  (v.adv_vnt[X3(i, k, j)] - v.adv_vnt[X3(i, k, j - 1)]) * v.dytr[j - 1]) *
                 v.cstr[j - 1] * v.dzt[k - 1];
      acc = d + acc;
    }
    long long line = X3Z(1, k, j);
    v.adv_vbt[line + i - 1] = acc;
    if (i == 2) v.adv_vbt[line + v.imt - 1] = acc;
    if (i == v.imt - 1) v.adv_vbt[line] = acc;
  }
}

void launch_adv_vel(uvic_b200_ctx *c) {
  DevView &v = c->v;
  long long tot = (long long)v.imt * v.km * v.jl;
  KLAUNCH("k_advvel_faces", k_advvel_faces, cdiv(tot, 256), 256, v);
  long long ncol = (long long)(v.imt - 2) * v.jl;
  KLAUNCH("k_advvel_column", k_advvel_column, cdiv(ncol, 128), 128, v);
}

// state (source/mom/state.F:1-60, called from 09/mom/loadmw.F:150-155 with t(tau)): the normalised density clinic
// differentiates, rho(i,k,j) = dens(t - to(k), s - so(k), k) (source/mom/dens.h:18-19), all local rows, i = 1..imt.
// With it on the device the host reads one 3-D field per step instead of T and S.
__global__ void __launch_bounds__(256) k_state(const DevView v, const double *__restrict__ t, double *__restrict__ rho) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= v.n3) return;
  const int k = (int)((idx / v.imt) % v.km) + 1;
  const double tq = t[idx] - v.to[k - 1], sq = t[idx + v.n3] - v.so[k - 1];
#define ECS(m) v.eosc[(k - 1) + v.km * ((m)-1)]
  rho[idx] = (ECS(1) + (ECS(4) + ECS(7) * sq) * sq + (ECS(3) + ECS(8) * sq + ECS(6) * tq) * tq) * tq + (ECS(2) + (ECS(5) + ECS(9) * sq) * sq) * sq;
#undef ECS
}

void launch_state(uvic_b200_ctx *c, const double *t, double *rho) {
  DevView &v = c->v;
  KLAUNCH("k_state", k_state, cdiv(v.n3, 256), 256, v, t, rho);
}
