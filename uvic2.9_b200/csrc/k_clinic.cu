// k_clinic.cu -- the baroclinic momentum step on the device, with the options of
// run/mk.in: O_consthmix + O_anisotropic_viscosity, O_constvmix (explicit vertical viscosity), O_stream_function
// (explicit Coriolis), no O_biharmonic / O_implicitvmix / O_pressure_gradient_average.
//
//   k_clinic_advvel   source/mom/adv_vel.F:160-250   advective velocities on the east / north / bottom faces of U cells
//   k_setvbc_mom      09/mom/setvbc.F:163-208        surface stress (coupler slots) and quadratic bottom drag
//   k_clinic_column   09/mom/clinic.F:60-511 + 09/mom/fdifm.h
//
// Two variants of the step itself (launch_clinic): three kernels with a cell-parallel tendency (default, see below), or
// k_clinic_column: one thread per U column (i fastest, so every level is one coalesced row segment), both velocity
// components at once (one thread per column AND component was measured slower, 947 against 592 us on 0.5 degree: the
// kernel is bound by L1/L2 load traffic, and the split repeats the rho / velocity / viscosity loads).  The hydrostatic pressure gradient is the running sum the reference builds in grad_p (:150-177) and
// stays in two registers; the vertical fluxes through the bottom face of level k are the top-face fluxes of level k+1 and
// are carried, not recomputed.  Pass 1 writes u(tau-1) + c2dtuv * du/dt and accumulates the two depth sums (zu: forcing
// of the barotropic equation, :378-397; baru: the vertical mean that is removed, :458-485) in the reference's k order;
// pass 2 subtracts the mean from the wet levels (the column is L1/L2 resident).  Operation order inside every expression
// is the reference's (the library is compiled without FMA contraction), so the result is bit-identical to the oracle.
// The polar filter of the velocities (filuv, k_filter.cu) follows when the context was set up with fourfil.  The
// diagnostics hooks (diagc1, diagc2) and the ice coupling (isbcu, asbcu) stay on the host.
#include <stdlib.h>
#include "ctx.h"

#define U0(i, k, j, n) v.u[X3(i, k, j) + (long long)((n)-1) * v.n3]
#define UM(i, k, j, n) cv.u_m1[X3(i, k, j) + (long long)((n)-1) * v.n3]
#define UP(i, k, j, n) cv.u_p1[X3(i, k, j) + (long long)((n)-1) * v.n3]

// one thread per (i, k = 0..km, j); rows jc0-1..jc1 (adv_vnu) and jc0..jc1 (adv_veu, adv_vbu)
__global__ void __launch_bounds__(128) k_clinic_advvel(const DevView v, const ClinicView cv) {
  // grid: x over i, y = level 0..km, z = row jc0-1..jc1 (no index division)
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 2;
  if (i > v.imt - 1) return;
  const int k = blockIdx.y;
  const int j = blockIdx.z + cv.jc0 - 1;
  const bool west = (i == 2), east = (i == v.imt - 1);
  if (k >= 1) {
    // adv_vnu = LINEAR_INTRP_Y(WT_AVG_X(adv_vnt)) (:173-186)
    const double dyr = v.dytr[j];
    const double val = ((v.adv_vnt[X3(i, k, j)] * v.duw[i - 1] + v.adv_vnt[X3(i + 1, k, j)] * v.due[i - 1]) * v.dus[j] +
                        (v.adv_vnt[X3(i, k, j + 1)] * v.duw[i - 1] + v.adv_vnt[X3(i + 1, k, j + 1)] * v.due[i - 1]) * v.dun[j - 1]) *
                       dyr * v.dxur[i - 1];
    const long long line = X3(1, k, j);
    cv.adv_vnu[line + i - 1] = val;
    if (west) cv.adv_vnu[line + v.imt - 1] = val;
    if (east) cv.adv_vnu[line] = val;
  }
  if (j < cv.jc0) return;
  if (k >= 1) {
    // adv_veu = LINEAR_INTRP_X(WT_AVG_Y(adv_vet)), cyclic (:200-219)
    const double dyr = v.dyur[j - 1];
    const double val = ((v.adv_vet[X3(i, k, j)] * v.dus[j - 1] + v.adv_vet[X3(i, k, j + 1)] * v.dun[j - 1]) * v.duw[i] +
                        (v.adv_vet[X3(i + 1, k, j)] * v.dus[j - 1] + v.adv_vet[X3(i + 1, k, j + 1)] * v.dun[j - 1]) * v.due[i - 1]) *
                       dyr * v.dxtr[i];
    const long long line = X3(1, k, j);
    cv.adv_veu[line + i - 1] = val;
    if (west) cv.adv_veu[line + v.imt - 1] = val;
    if (east) cv.adv_veu[line] = val;
  }
  {
    // bottom face (:226-250)
    const double dyn = v.dun[j - 1] * v.cst[j];
    const double dys = v.dus[j - 1] * v.cst[j - 1];
    const double dyr = v.dyur[j - 1] * v.csur[j - 1];
    const double asw = v.duw[i - 1] * dys, anw = v.duw[i - 1] * dyn, ase = v.due[i - 1] * dys, ane = v.due[i - 1] * dyn;
    const double val = dyr * v.dxur[i - 1] *
                       (v.adv_vbt[X3Z(i, k, j)] * asw + v.adv_vbt[X3Z(i + 1, k, j)] * ase + v.adv_vbt[X3Z(i, k, j + 1)] * anw +
                        v.adv_vbt[X3Z(i + 1, k, j + 1)] * ane);
    const long long line = X3Z(1, k, j);
    cv.adv_vbu[line + i - 1] = val;
    if (west) cv.adv_vbu[line + v.imt - 1] = val;
    if (east) cv.adv_vbu[line] = val;
  }
}

// 09/mom/setvbc.F:163-208: every local row, i = 2..imt-1 then the cyclic copies.  itaux = 0 keeps an uploaded smf.
__global__ void __launch_bounds__(128) k_setvbc_mom(const DevView v, const ClinicView cv, int itaux, int itauy) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ni = v.imt - 2;
  if (idx >= (long long)ni * v.jl) return;
  const int i = (int)(idx % ni) + 2;
  const int j = (int)(idx / ni) + v.jbase;
  const int kz = cv.kmu[X2(i, j)];
  double s[2], b[2];
  if (itaux > 0) {
    const double um = (kz >= 1) ? 1.0 : 0.0;   // umask(i,1,j)
    s[0] = v.sbc[X2(i, j) + (long long)(itaux - 1) * v.n2] * um;
    s[1] = v.sbc[X2(i, j) + (long long)(itauy - 1) * v.n2] * um;
  }
  if (cv.cdbot == 0.0 || kz == 0) {
    b[0] = b[1] = 0.0;
  } else {
    const double u1 = UM(i, kz, j, 1), u2 = UM(i, kz, j, 2);
    const double uvmag = sqrt(u1 * u1 + u2 * u2);
    b[0] = cv.cdbot * u1 * uvmag;
    b[1] = cv.cdbot * u2 * uvmag;
  }
  for (int n = 0; n < 2; n++) {
    const long long line = X2(1, j) + (long long)n * v.n2;
    if (itaux > 0) {
      cv.smf[line + i - 1] = s[n];
      if (i == 2) cv.smf[line + v.imt - 1] = s[n];
      if (i == v.imt - 1) cv.smf[line] = s[n];
    }
    cv.bmf[line + i - 1] = b[n];
    if (i == 2) cv.bmf[line + v.imt - 1] = b[n];
    if (i == v.imt - 1) cv.bmf[line] = b[n];
  }
}

// CTA = 32 columns in i x CL_TJ rows: the north / south neighbours of the inner rows are lines the CTA's other warps load
// at the same level, so they come from L1 instead of L2
#define CL_TJ 4
template <int MINB>
__global__ void __launch_bounds__(32 * CL_TJ, MINB) k_clinic_column(const DevView v, const ClinicView cv) {
  const int i = blockIdx.x * 32 + (threadIdx.x & 31) + 2;
  const int j = blockIdx.y * CL_TJ + (threadIdx.x >> 5) + cv.jc0;
  if (i > v.imt - 1 || j > cv.jc1) return;
  const int km = v.km;
  const bool west = (i == 2), east = (i == v.imt - 1);
  const int kb = cv.kmu[X2(i, j)];
  const int kmt_ij = v.kmt[X2(i, j)];          // visc_cbu lives on T-cell levels (09/mom/vmixc.F:84-85)
  const double hr = cv.hr[X2(i, j)];
  const double c2dtuv = cv.c2dtuv;

  // row / column constants (09/mom/clinic.F:75-82, 119-147)
  const double csur = v.csur[j - 1];
  const double csudxur = csur * v.dxur[i - 1];
  const double csudxu2r = csur * v.dxur[i - 1] * 0.5;
  const double csudyu2r = cv.csudyu2r[j - 1];
  const double dxtr_e = v.dxtr[i], dxtr_w = v.dxtr[i - 1];
  const double dxu2r = cv.dxu2r[i - 1];
  const double g = cv.grav_rho0r;
  const double am3 = cv.am3[j - 1];
  const double am4d[2] = {cv.am4[j - 1] * cv.dxmetr[i - 1], cv.am4[(j - 1) + v.jmt] * cv.dxmetr[i - 1]};
  const double advmet[2] = {cv.advmet[j - 1], cv.advmet[(j - 1) + v.jmt]};
  const double cori[2] = {cv.cori[X2(i, j)], cv.cori[X2(i, j) + v.n2]};
  const double smf[2] = {cv.smf[X2(i, j)], cv.smf[X2(i, j) + v.n2]};
  const double bmf[2] = {cv.bmf[X2(i, j)], cv.bmf[X2(i, j) + v.n2]};

  double gp[2] = {0.0, 0.0};          // grad_p(i,k,j,:) integrated downward
  double rp[4] = {0.0, 0.0, 0.0, 0.0};   // rho of the level above at (i,j) (i+1,j) (i,j+1) (i+1,j+1)
  double afb_up[2], dfb_up[2];        // fluxes through the top face of the current level
  double zu[2] = {0.0, 0.0}, bar[2] = {0.0, 0.0};
  double u0d_prev[2] = {0.0, 0.0}, umd_prev[2] = {0.0, 0.0};
  double um_k[2];

  for (int k = 1; k <= km; k++) {
    double tend[2] = {0.0, 0.0};
    if (k <= kb) {
      // ---- pressure gradient at this level ----
      const double r00 = cv.rho[X3(i, k, j)], r10 = cv.rho[X3(i + 1, k, j)], r01 = cv.rho[X3(i, k, j + 1)], r11 = cv.rho[X3(i + 1, k, j + 1)];
      if (k == 1) {
        const double fxa = g * v.dzw[0] * csur, fxb = g * v.dzw[0] * cv.dyu2r[j - 1];
        const double t1 = r11 - r00, t2 = r01 - r10;
        gp[0] = (t1 - t2) * fxa * dxu2r;
        gp[1] = (t1 + t2) * fxb;
      } else {
        const double fxa = g * csur * 0.5, fxb = g * cv.dyu4r[j - 1];
        const double e00 = rp[0] + r00, e10 = rp[1] + r10, e01 = rp[2] + r01, e11 = rp[3] + r11;   // tempik (:150-157)
        const double t1 = e11 - e00, t2 = e01 - e10;
        gp[0] = gp[0] + fxa * (t1 - t2) * v.dzw[k - 1] * dxu2r;
        gp[1] = gp[1] + fxb * (t1 + t2) * v.dzw[k - 1];
      }
      rp[0] = r00; rp[1] = r10; rp[2] = r01; rp[3] = r11;
      if (cv.grad_p) {
        cv.grad_p[X3(i, k, j)] = gp[0];
        cv.grad_p[X3(i, k, j) + v.n3] = gp[1];
      }
      // ---- operands ----
      double u0c[2], u0e[2], u0w[2], u0n[2], u0s[2], u0d[2], umc[2], ume[2], umw[2], umn[2], ums[2], umd[2];
      const int kd = (k < km) ? k + 1 : km;
#pragma unroll
      for (int n = 0; n < 2; n++) {
        // the centre values of level k are the "level below" values of level k-1 (both wet): carried, not reloaded
        u0c[n] = (k == 1) ? U0(i, k, j, n + 1) : u0d_prev[n];
        umc[n] = (k == 1) ? UM(i, k, j, n + 1) : umd_prev[n];
        u0e[n] = U0(i + 1, k, j, n + 1); u0w[n] = U0(i - 1, k, j, n + 1);
        u0n[n] = U0(i, k, j + 1, n + 1); u0s[n] = U0(i, k, j - 1, n + 1); u0d[n] = U0(i, kd, j, n + 1);
        ume[n] = UM(i + 1, k, j, n + 1); umw[n] = UM(i - 1, k, j, n + 1);
        umn[n] = UM(i, k, j + 1, n + 1); ums[n] = UM(i, k, j - 1, n + 1); umd[n] = UM(i, kd, j, n + 1);
        u0d_prev[n] = u0d[n]; umd_prev[n] = umd[n];
      }
      const double veu_e = cv.adv_veu[X3(i, k, j)], veu_w = cv.adv_veu[X3(i - 1, k, j)];
      const double vnu_n = cv.adv_vnu[X3(i, k, j)], vnu_s = cv.adv_vnu[X3(i, k, j - 1)];
      const double vbu_lo = cv.adv_vbu[X3Z(i, k, j)];
      const double amx_e = cv.visc_ceu[X3(i, k, j)] * csur * dxtr_e;      // am_csudxtr(i,k,j)   (:80)
      const double amx_w = cv.visc_ceu[X3(i - 1, k, j)] * csur * dxtr_w;  // am_csudxtr(i-1,k,j)
      const double amcn = cv.amc_north[X3(i, k, j)], amcs = cv.amc_south[X3(i, k, j)];
      const double visc_lo = (k <= kmt_ij - 1) ? cv.kappa_m : 0.0;
#pragma unroll
      for (int n = 0; n < 2; n++) {
        const int o = 1 - n;
        if (k == 1) {
          // surface b.c. (:309-313)
          dfb_up[n] = smf[n];
          afb_up[n] = cv.adv_vbu[X3Z(i, 0, j)] * (u0c[n] + u0c[n]);
        }
        // bottom face of level k (:283-293, 310-314)
        double afb_lo, dfb_lo;
        if (k < km) {
          afb_lo = vbu_lo * (u0c[n] + u0d[n]);
          dfb_lo = visc_lo * v.dzwr[k] * (umc[n] - umd[n]);
        } else {
          afb_lo = vbu_lo * u0c[n];
          dfb_lo = 0.0;                 // diff_fb(i,km,j) is only ever set through kb = km
        }
        const double dfb_reg = dfb_lo;
        if (k == kb) dfb_lo = bmf[n];
        const double afe_e = veu_e * (u0c[n] + u0e[n]), afe_w = veu_w * (u0w[n] + u0c[n]);
        const double dfe_e = amx_e * (ume[n] - umc[n]), dfe_w = amx_w * (umc[n] - umw[n]);
        // 09/mom/fdifm.h
        const double DIFF_Ux = (dfe_e - dfe_w) * csudxur;
        const double DIFF_Uy = amcn * (umn[n] - umc[n]) - amcs * (umc[n] - ums[n]);
        const double DIFF_Uz = (dfb_up[n] - dfb_lo) * v.dztr[k - 1];
        const double DIFF_metric = am3 * umc[n] + am4d[n] * (ume[o] - umw[o]);
        const double ADV_Ux = (afe_e - afe_w) * csudxu2r;
        const double ADV_Uy = (vnu_n * (u0c[n] + u0n[n]) - vnu_s * (u0s[n] + u0c[n])) * csudyu2r;
        const double ADV_Uz = (afb_up[n] - afb_lo) * v.dzt2r[k - 1];
        const double ADV_metric = advmet[n] * u0c[0] * u0c[o];
        const double CORIOLIS = cori[n] * u0c[o];
        tend[n] = DIFF_Ux + DIFF_Uy + DIFF_Uz + DIFF_metric - ADV_Ux - ADV_Uy - ADV_Uz + ADV_metric - gp[n] + CORIOLIS;
        um_k[n] = umc[n];
        afb_up[n] = afb_lo;
        dfb_up[n] = dfb_reg;
      }
    }
    if (k > kb) {
      um_k[0] = UM(i, k, j, 1);
      um_k[1] = UM(i, k, j, 2);
    }
    // ---- tau+1 before the mean is removed (:444-451), depth sums in the reference's order ----
#pragma unroll
    for (int n = 0; n < 2; n++) {
      zu[n] = zu[n] + tend[n] * v.dzt[k - 1];
      const double up = um_k[n] + c2dtuv * tend[n];
      bar[n] = bar[n] + up * v.dzt[k - 1];
      UP(i, k, j, n + 1) = up;
    }
  }
#pragma unroll
  for (int n = 0; n < 2; n++) {
    cv.zu[X2(i, j) + (long long)n * v.n2] = zu[n] * hr;
    bar[n] = bar[n] * hr;
  }
  // ---- pure internal modes (:476-485) and the cyclic boundary ----
  for (int k = 1; k <= km; k++) {
#pragma unroll
    for (int n = 0; n < 2; n++) {
      double up = UP(i, k, j, n + 1);
      if (k <= kb) {
        up = up - bar[n];
        UP(i, k, j, n + 1) = up;
      }
      if (west) UP(v.imt, k, j, n + 1) = up;
      if (east) UP(1, k, j, n + 1) = up;
    }
  }
}


// ---- the cell-parallel variant (UVIC_B200_CLINIC=cell): the same arithmetic in three kernels -----------------------
// k_clinic_gradp  column threads, 4 loads per level: the downward integral of the pressure gradient into grad_p
// k_clinic_tend   one thread per wet U cell (no serial chain): du/dt of both components into u_p1; the vertical face
//                 fluxes are evaluated from their definitions (the same expressions the column kernel carries)
// k_clinic_finish column threads: zu, u(tau-1) + c2dtuv du/dt, the vertical mean and its removal, cyclic columns
__global__ void __launch_bounds__(128) k_clinic_gradp(const DevView v, const ClinicView cv) {
  const int i = blockIdx.x * 32 + (threadIdx.x & 31) + 2;
  const int j = blockIdx.y * 4 + (threadIdx.x >> 5) + cv.jc0;
  if (i > v.imt - 1 || j > cv.jc1) return;
  const int kb = cv.kmu[X2(i, j)];
  const double *__restrict__ rho = cv.rho;
  double *__restrict__ gpo = cv.grad_p;
  const double csur = v.csur[j - 1], dxu2r = cv.dxu2r[i - 1], g = cv.grav_rho0r;
  const long long jst = (long long)v.imt * v.km;
  double gp0 = 0.0, gp1 = 0.0, rp0 = 0.0, rp1 = 0.0, rp2 = 0.0, rp3 = 0.0;
  for (int k = 1; k <= kb; k++) {
    const long long x = X3(i, k, j);
    const double r00 = rho[x], r10 = rho[x + 1], r01 = rho[x + jst], r11 = rho[x + jst + 1];
    if (k == 1) {
      const double fxa = g * v.dzw[0] * csur, fxb = g * v.dzw[0] * cv.dyu2r[j - 1];
      const double t1 = r11 - r00, t2 = r01 - r10;
      gp0 = (t1 - t2) * fxa * dxu2r;
      gp1 = (t1 + t2) * fxb;
    } else {
      const double fxa = g * csur * 0.5, fxb = g * cv.dyu4r[j - 1];
      const double e00 = rp0 + r00, e10 = rp1 + r10, e01 = rp2 + r01, e11 = rp3 + r11;
      const double t1 = e11 - e00, t2 = e01 - e10;
      gp0 = gp0 + fxa * (t1 - t2) * v.dzw[k - 1] * dxu2r;
      gp1 = gp1 + fxb * (t1 + t2) * v.dzw[k - 1];
    }
    rp0 = r00; rp1 = r10; rp2 = r01; rp3 = r11;
    gpo[x] = gp0;
    gpo[x + v.n3] = gp1;
  }
}

__global__ void __launch_bounds__(128) k_clinic_tend(const DevView v, const ClinicView cv) {
  // grid: x over i, y = level 1..km, z = row jc0..jc1
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 2;
  if (i > v.imt - 1) return;
  const int k = blockIdx.y + 1;
  const int j = blockIdx.z + cv.jc0;
  const int kb = cv.kmu[X2(i, j)];
  if (k > kb) return;
  const int km = v.km;
  const int kmt_ij = v.kmt[X2(i, j)];
  const long long x = X3(i, k, j), lev = (long long)v.imt, jst = (long long)v.imt * v.km;
  const double csur = v.csur[j - 1];
  const double csudxur = csur * v.dxur[i - 1];
  const double csudxu2r = csur * v.dxur[i - 1] * 0.5;
  const double csudyu2r = cv.csudyu2r[j - 1];
  const double am3 = cv.am3[j - 1];
  const double veu_e = cv.adv_veu[x], veu_w = cv.adv_veu[x - 1];
  const double vnu_n = cv.adv_vnu[x], vnu_s = cv.adv_vnu[x - jst];
  const double vbu_lo = cv.adv_vbu[X3Z(i, k, j)], vbu_up = cv.adv_vbu[X3Z(i, k - 1, j)];
  const double amx_e = cv.visc_ceu[x] * csur * v.dxtr[i];
  const double amx_w = cv.visc_ceu[x - 1] * csur * v.dxtr[i - 1];
  const double amcn = cv.amc_north[x], amcs = cv.amc_south[x];
  const double visc_lo = (k <= kmt_ij - 1) ? cv.kappa_m : 0.0;
  const double visc_up = (k - 1 <= kmt_ij - 1) ? cv.kappa_m : 0.0;
  const long long xd = (k < km) ? x + lev : x, xu = (k > 1) ? x - lev : x;
  double u0c[2], u0e[2], u0w[2], u0n[2], u0s[2], u0d[2], u0u[2], umc[2], ume[2], umw[2], umn[2], ums[2], umd[2], umu[2];
#pragma unroll
  for (int n = 0; n < 2; n++) {
    const double *__restrict__ u0 = v.u + (long long)n * v.n3;
    const double *__restrict__ um = cv.u_m1 + (long long)n * v.n3;
    u0c[n] = u0[x]; u0e[n] = u0[x + 1]; u0w[n] = u0[x - 1]; u0n[n] = u0[x + jst]; u0s[n] = u0[x - jst]; u0d[n] = u0[xd]; u0u[n] = u0[xu];
    umc[n] = um[x]; ume[n] = um[x + 1]; umw[n] = um[x - 1]; umn[n] = um[x + jst]; ums[n] = um[x - jst]; umd[n] = um[xd]; umu[n] = um[xu];
  }
#pragma unroll
  for (int n = 0; n < 2; n++) {
    const int o = 1 - n;
    // top face (:283-293 for the face k-1; :309-313 at the surface)
    double afb_up, dfb_up;
    if (k == 1) {
      dfb_up = cv.smf[X2(i, j) + (long long)n * v.n2];
      afb_up = vbu_up * (u0c[n] + u0c[n]);
    } else {
      afb_up = vbu_up * (u0u[n] + u0c[n]);
      dfb_up = visc_up * v.dzwr[k - 1] * (umu[n] - umc[n]);
    }
    // bottom face
    double afb_lo, dfb_lo;
    if (k < km) {
      afb_lo = vbu_lo * (u0c[n] + u0d[n]);
      dfb_lo = visc_lo * v.dzwr[k] * (umc[n] - umd[n]);
    } else {
      afb_lo = vbu_lo * u0c[n];
      dfb_lo = 0.0;
    }
    if (k == kb) dfb_lo = cv.bmf[X2(i, j) + (long long)n * v.n2];
    const double afe_e = veu_e * (u0c[n] + u0e[n]), afe_w = veu_w * (u0w[n] + u0c[n]);
    const double dfe_e = amx_e * (ume[n] - umc[n]), dfe_w = amx_w * (umc[n] - umw[n]);
    const double DIFF_Ux = (dfe_e - dfe_w) * csudxur;
    const double DIFF_Uy = amcn * (umn[n] - umc[n]) - amcs * (umc[n] - ums[n]);
    const double DIFF_Uz = (dfb_up - dfb_lo) * v.dztr[k - 1];
    const double DIFF_metric = am3 * umc[n] + cv.am4[(j - 1) + n * v.jmt] * cv.dxmetr[i - 1] * (ume[o] - umw[o]);
    const double ADV_Ux = (afe_e - afe_w) * csudxu2r;
    const double ADV_Uy = (vnu_n * (u0c[n] + u0n[n]) - vnu_s * (u0s[n] + u0c[n])) * csudyu2r;
    const double ADV_Uz = (afb_up - afb_lo) * v.dzt2r[k - 1];
    const double ADV_metric = cv.advmet[(j - 1) + n * v.jmt] * u0c[0] * u0c[o];
    const double CORIOLIS = cv.cori[X2(i, j) + (long long)n * v.n2] * u0c[o];
    cv.u_p1[x + (long long)n * v.n3] = DIFF_Ux + DIFF_Uy + DIFF_Uz + DIFF_metric - ADV_Ux - ADV_Uy - ADV_Uz + ADV_metric -
                                       cv.grad_p[x + (long long)n * v.n3] + CORIOLIS;
  }
}

__global__ void __launch_bounds__(128) k_clinic_finish(const DevView v, const ClinicView cv) {
  const int i = blockIdx.x * 32 + (threadIdx.x & 31) + 2;
  const int j = blockIdx.y * 4 + (threadIdx.x >> 5) + cv.jc0;
  if (i > v.imt - 1 || j > cv.jc1) return;
  const int n = blockIdx.z;
  const int kb = cv.kmu[X2(i, j)];
  const double hr = cv.hr[X2(i, j)], c2dtuv = cv.c2dtuv;
  const double *__restrict__ um = cv.u_m1 + (long long)n * v.n3;
  double *up = cv.u_p1 + (long long)n * v.n3;
  const bool west = (i == 2), east = (i == v.imt - 1);
  double zu = 0.0, bar = 0.0;
  for (int k = 1; k <= v.km; k++) {
    const long long x = X3(i, k, j);
    const double tend = (k <= kb) ? up[x] : 0.0;
    zu = zu + tend * v.dzt[k - 1];
    const double upk = um[x] + c2dtuv * tend;
    bar = bar + upk * v.dzt[k - 1];
  }
  cv.zu[X2(i, j) + (long long)n * v.n2] = zu * hr;
  bar = bar * hr;
  for (int k = 1; k <= v.km; k++) {
    const long long line = X3(1, k, j);
    const double tend = (k <= kb) ? up[line + i - 1] : 0.0;
    double upk = um[line + i - 1] + c2dtuv * tend;
    if (k <= kb) upk = upk - bar;
    up[line + i - 1] = upk;
    if (west) up[line + v.imt - 1] = upk;
    if (east) up[line] = upk;
  }
}

void launch_setvbc_mom(uvic_b200_ctx *c, int itaux, int itauy) {
  DevView &v = c->v;
  const ClinicView &cv = *c->clinic;
  const long long ncol = (long long)(v.imt - 2) * v.jl;
  KLAUNCH("k_setvbc_mom", k_setvbc_mom, cdiv(ncol, 128), 128, v, cv, itaux, itauy);
}

void launch_clinic(uvic_b200_ctx *c) {
  DevView &v = c->v;
  const ClinicView &cv = *c->clinic;
  {
    ProfScope ps_(c, "k_clinic_advvel");
    k_clinic_advvel<<<dim3(cdiv(v.imt - 2, 128), v.km + 1, cv.jc1 - cv.jc0 + 2), 128, 0, c->stream>>>(v, cv);
  }
  // default: the cell-parallel variant (whole call 0.623 against 0.674 ms on 0.5 degree x 40 levels, 0.044 against 0.048 ms
  // on 100x100x19); UVIC_B200_CLINIC=column selects the single marching kernel.  Both are bit-identical to the oracle.
  static const bool cell_variant = !(getenv("UVIC_B200_CLINIC") && std::string(getenv("UVIC_B200_CLINIC")) == "column");
  if (cell_variant) {
    const dim3 cols(cdiv(v.imt - 2, 32), cdiv(cv.jc1 - cv.jc0 + 1, 4));
    { ProfScope ps_(c, "k_clinic_gradp"); k_clinic_gradp<<<cols, 128, 0, c->stream>>>(v, cv); }
    {
      ProfScope ps_(c, "k_clinic_tend");
      k_clinic_tend<<<dim3(cdiv(v.imt - 2, 128), v.km, cv.jc1 - cv.jc0 + 1), 128, 0, c->stream>>>(v, cv);
    }
    { ProfScope ps_(c, "k_clinic_finish"); k_clinic_finish<<<dim3(cols.x, cols.y, 2), 128, 0, c->stream>>>(v, cv); }
  } else {
    ProfScope ps_(c, "k_clinic_column");
    // resident CTAs per SM the register allocation is capped for: 2 = 240 registers, 3 = 166, 4 = 128 (with 144 B of spills)
    // Measured (whole clinic call): 0.5 degree x 40 levels 0.713 / 0.699 / 0.673 ms, 100x100x19 (80 CTAs, less than one
    // per SM) 0.049 / 0.057 / 0.062 ms -- so the cap follows the grid size.
    const dim3 grid(cdiv(v.imt - 2, 32), cdiv(cv.jc1 - cv.jc0 + 1, CL_TJ));
    static const int occ_env = getenv("UVIC_B200_CLINIC_OCC") ? atoi(getenv("UVIC_B200_CLINIC_OCC")) : 0;
    const int occ = occ_env ? occ_env : ((long long)grid.x * grid.y >= 4 * 148 ? 4 : 2);
    if (occ <= 2) k_clinic_column<2><<<grid, 32 * CL_TJ, 0, c->stream>>>(v, cv);
    else if (occ == 3) k_clinic_column<3><<<grid, 32 * CL_TJ, 0, c->stream>>>(v, cv);
    else k_clinic_column<4><<<grid, 32 * CL_TJ, 0, c->stream>>>(v, cv);
  }
  // O_fourfil: filuv on the polar rows (09/mom/clinic.F:494-507), including the final setbcx of those rows
  launch_filuv(c, cv.u_p1, cv.spsin, cv.spcos, cv.kmu, cv.hr);
}
