/*
 * uvic_b200.h -- C ABI of the B200-native ocean tracer step for the UVic ESCM 2.9.
 *
 * This is the drop-in boundary.  The reference has no FFI: its
 * boundary is a set of Fortran subroutine call sites inside `mom` plus the COMMON blocks
 * they share.  Each entry point below replaces one of those call sites; the Fortran side
 * binds them with ISO_C_BINDING from an overlay tracer.F / isopyc.F / vmixc.F
 * (INTEGRATION.md shows the shim).  Plain pointers and sizes only; no exceptions cross
 * the ABI; every function returns 0 on success and non-zero on error, with a message
 * retrievable through uvic_b200_last_error (the reference convention is a message on
 * stdout followed by `stop '=>routine'`, e.g. 09/mom/tracer.F:1248-1250; the Fortran
 * wrapper prints the message and stops).
 *
 * Layout: every field is the reference's Fortran column-major array, unchanged:
 * (i,k,j[,n]) with i fastest (09/mom/mw.h:76-77).  Arrays documented "0:km" carry km+1
 * levels (09/mom/mw.h:249,302,309).  All reals are FP64 (the reference builds with -r8,
 * run/mk.ver:39-43).
 *
 * Latitude slabs: a context owns the rows jrow_lo..jrow_hi of the global grid (2..jmt-1
 * for a single GPU) and keeps a 2-row halo on each side resident, so every array argument
 * with a j extent covers the `jl` rows jbase..jbase+jl-1, jbase = max(1, jrow_lo-2),
 * jbase+jl-1 = min(jmt, jrow_hi+2).  With one GPU that is simply rows 1..jmt.
 * Paths are relative to /root/reference; "09/" = updates/09/source/.
 */
#ifndef UVIC_B200_H
#define UVIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct uvic_b200_ctx uvic_b200_ctx;

/* sizes: 09/common/size.h:27-144 */
typedef struct uvic_b200_dims {
  int32_t imt, jmt, km;     /* global grid incl. the boundary columns/rows */
  int32_t nt, nsrc;         /* tracers, tracers with sources */
  int32_t jrow_lo, jrow_hi; /* first / last global row this context computes */
} uvic_b200_dims;

/* grid metrics: source/common/grdvar.h:64-86, coord.h, accel.h (host pointers, copied) */
typedef struct uvic_b200_grid {
  const double *dxt, *dxtr, *dxt2r, *dxt4r, *dxu, *dxur;                 /* (imt) */
  const double *dyt, *dytr, *dyt2r, *dyt4r, *dyu, *dyur;                 /* (jmt) global */
  const double *cst, *cstr, *csu, *csur, *cstdytr, *cstdyt2r, *csu_dyur; /* (jmt) global */
  const double *dzt, *dztr, *dzt2r, *dztur, *dztlr, *zt, *zw;            /* (km) */
  const double *dzw, *dzwr;                                              /* (0:km) */
  const double *dtxcel, *dtxsqr, *dztxcl, *dzwxcl;                       /* (km) */
  const double *tlat;                                                    /* (imt,jl) */
  const double *duw, *due;                                               /* (imt) adv_vel */
  const double *dus, *dun;                                               /* (jmt) adv_vel */
  const double *eosc;                                                    /* c(km,9) source/mom/state.h:15-16 */
  const double *to, *so;                                                 /* (km) */
} uvic_b200_grid;

/* run-time parameters: &mixing / &isopyc of run/control.in, 09/mom/setmom.F:80-82,
 * and the cpp options of run/mk.in that select code paths */
typedef struct uvic_b200_params {
  double aidif, kappa_h;        /* run/control.in:8 */
  double ahisop, athkdf, slmxr; /* 09/mom/isopyc.F:70-110 */
  double diff_cet, diff_cnt;    /* 09/mom/hmixc.F:192-200 */
  double zetar, ogamma, gravrho0r;
  int32_t fct;                  /* O_fct */
  int32_t isopycmix;            /* O_isopycmix + O_gent_mcwilliams */
  int32_t tidal_kv;             /* O_tidal_kv */
  int32_t fullconvect;          /* O_fullconvect */
  int32_t mobi;                 /* O_mobi (+ the O_mobi_* / O_carbon* set of run/mk.in) */
  int32_t fourfil;              /* O_fourfil */
  int32_t jfrst, jft0, jft1, jft2; /* Fourier-filter rows (global), source/common/setcom.F:37-40,75-85 */
  const int32_t *itrc;          /* (nt) source slot per tracer, 0 = none (09/mom/mw.h:125-221) */
  const int32_t *mobi_index;    /* tracer / source index maps for MOBI (layout: uvic2.9_b200/csrc/mobi_par.h); may be NULL.  The
                                 * isotope options are selected independently, as in run/mk.in: the tracers of O_carbon_13,
                                 * O_carbon_14, O_mobi_nitrogen_15 that a build does not carry (09/common/size.h:31-144) have
                                 * index 0 here -- e.g. nt = 21, nsrc = 19 for "full MOBI, no isotopes" */
  const double *mobi_par;       /* MOBI parameter block after mobi_init's unit conversion; may be NULL */
  int32_t n_mobi_index, n_mobi_par;
} uvic_b200_params;

/* time-invariant 2-D / 3-D inputs (host pointers, copied at create) */
typedef struct uvic_b200_static {
  const int32_t *kmt;     /* (imt,jl)  09/common/levind.h */
  const int32_t *mskhr;   /* (imt,jl)  horizontal regions for basin means, may be NULL */
  const double *fisop;    /* (imt,jl,km)  note (i,j,k)  09/common/isopyc.h:42 */
  const double *addisop;  /* (imt,km,jl) */
  const double *edrm2, *edrs2, *edrk1, *edro1; /* (imt,km,jl) 09/mom/tidal_kv.h */
  const double *sg_bathy; /* (imt,jl,km) may be NULL without MOBI */
  const double *fe_hydr;  /* (imt,jl,km) */
  const double *fe_atmdep;/* (imt,jl,12) */
} uvic_b200_static;

/* per-step switches and scalars (source/mom/mom.F:111-146, 09/common/switch.F:217-224) */
typedef struct uvic_b200_stepinfo {
  double dtts;       /* tracer time step (s) */
  int32_t leapfrog;  /* 1: leapfrog (c2dtts = 2 dtts); 0: forward mixing step (tau-1 := tau) */
  int32_t diag;      /* 1: also form tbar / sumbk inventories (tsiperts .and. eots) */
  double relyr;      /* model time in years (MOBI light / month index) */
  double co2ccn;     /* atmospheric CO2 (ppmv) for co2calc */
} uvic_b200_stepinfo;

/* ---- life cycle ---------------------------------------------------------------- */
int uvic_b200_create(const uvic_b200_dims *dims, const uvic_b200_grid *grid, const uvic_b200_params *par,
                     const uvic_b200_static *st, int device, uvic_b200_ctx **out);
int uvic_b200_destroy(uvic_b200_ctx *ctx);
const char *uvic_b200_last_error(const uvic_b200_ctx *ctx); /* ctx may be NULL: last create error */
int uvic_b200_set_stream(uvic_b200_ctx *ctx, void *cuda_stream);
int uvic_b200_synchronize(uvic_b200_ctx *ctx);

/* ---- state movement (replaces loadmw / putmw, 09/mom/loadmw.F:102-117,717-744) ---- */
/* level: -1 = tau-1, 0 = tau, +1 = tau+1.  host arrays are t(imt,km,jl,nt). */
int uvic_b200_upload_t(uvic_b200_ctx *ctx, int level, const double *t_host);
int uvic_b200_download_t(uvic_b200_ctx *ctx, int level, double *t_host);
int uvic_b200_download_tracer(uvic_b200_ctx *ctx, int level, int n, double *field_host); /* lazy D2H of one tracer (1-based n) */
/* advective velocities from adv_vel (source/mom/adv_vel.F:60-131): (imt,km,jl),(imt,km,jl),(imt,0:km,jl) */
/* The reference dimensions the advective velocities (imt,km,jsmw:jmw) and (imt,0:km,jsmw:jmw) (09/mom/mw.h:246-263): with
 * the fully opened memory window jsmw = 2, so the COMMON arrays have NO row 1.  A host that passes c_loc of those arrays
 * says so once: every host velocity pointer (uvic_b200_upload_adv_vel, _tracer_step, _tracer_step_coupled) is then read as
 * starting at global row jrow_first (the device row before it keeps its zeros: row 1 is a closed wall).  Default 1. */
int uvic_b200_set_host_window(uvic_b200_ctx *ctx, int jrow_first);
int uvic_b200_upload_adv_vel(uvic_b200_ctx *ctx, const double *adv_vet, const double *adv_vnt, const double *adv_vbt);
/* or: B-grid velocity u(imt,km,jl,2) at tau, and adv_vel on the device */
int uvic_b200_upload_u(uvic_b200_ctx *ctx, const double *u_host);
int uvic_b200_adv_vel(uvic_b200_ctx *ctx);
/* call state (source/mom/state.F:1-60, from 09/mom/loadmw.F:150-155): rho(imt,km,jl) = dens(T - to(k), S - so(k), k) of the
 * time level (-1, 0, +1) to the host -- the one 3-D field clinic needs from the tracers every step */
int uvic_b200_state(uvic_b200_ctx *ctx, int level, double *rho_host);
/* surface / bottom tracer fluxes from setvbc (09/mom/setvbc.F): stf, btf (imt,jl,nt) */
int uvic_b200_upload_vbc(uvic_b200_ctx *ctx, const double *stf, const double *btf);
/* MOBI 2-D forcing: dnswr, aice, hice, hsno (imt,jl) (09/mom/tracer.F:370-390) */
int uvic_b200_upload_forcing(uvic_b200_ctx *ctx, const double *dnswr, const double *aice, const double *hice, const double *hsno);
/* rotate time levels after a step: tau-1 <- tau <- tau+1 (source/mom/mom.F:210-212) */
/* Surface boundary conditions on the device.  sbc(imt,jl,numsbc) is this slab of the coupler's
 * array (09/common/csbc.h); flx_index[n-1] / acc_index[n-1] are the 1-based sbc slots that hold tracer n's surface flux
 * (the right-hand sides of the assignment list in 09/mom/setvbc.F:83-126) and its surface accumulator (trsbcindex(n),
 * 09/mom/tracer.F:1270-1288); 0 = none.  With these the host moves 2-D fields per step: the flux slots in, and the
 * accumulator slots out at the end of an ocean segment. */
int uvic_b200_sbc_setup(uvic_b200_ctx *ctx, int numsbc, const int32_t *flx_index, const int32_t *acc_index);
int uvic_b200_upload_sbc(uvic_b200_ctx *ctx, const double *sbc, const double *bhf);            /* whole array / (imt,jl); NULL = keep */
int uvic_b200_upload_sbc_slot(uvic_b200_ctx *ctx, int slot, const double *field);              /* one (imt,jl) slot, async */
int uvic_b200_download_sbc(uvic_b200_ctx *ctx, double *sbc);
int uvic_b200_download_sbc_slot(uvic_b200_ctx *ctx, int slot, double *field);
/* Air-sea gas exchange: the flux loop of gasbc (09/common/gasbc.F:148-266) on the sbc array the device holds: DIC,
 * DI13C, 14C and O2 fluxes from the segment-mean surface state (the accumulators of uvic_b200_set_sbc), co2calc_SWS at
 * the surface; land points take the land carbon fluxes when inpp > 0.  Slots are 1-based indices into sbc. */
typedef struct uvic_b200_gasbc_par {
  int32_t isst, isss, issdic, issalk, issdic13, issc14, isso2, iws; /* read: SST, SSS, surface DIC, ALK, DI13C, 14C, O2, wind speed */
  int32_t inpp, isr, iburn;                                         /* read on land: NPP, soil respiration, burning (0 = none) */
  int32_t idicflx, idic13flx, ic14flx, io2flx;                      /* written */
  double co2ccn, dc13ccn, dc14ccn;                                  /* atmospheric CO2 (ppmv), delta13C, delta14C (permil) */
} uvic_b200_gasbc_par;
int uvic_b200_gasbc(uvic_b200_ctx *ctx, const uvic_b200_gasbc_par *par);
int uvic_b200_setvbc(uvic_b200_ctx *ctx);   /* call setvbc (source/mom/mom.F:360, 09/mom/setvbc.F:60-140): fills stf, btf */
/* call set_sbc for every tracer with an accumulator slot (09/mom/tracer.F:1270-1288, 09/mom/set_sbc.F:36-83) on the
 * t(tau+1) the last uvic_b200_tracer produced; eots/osegs/osege/ntspos are the switches of source/common/switch.h */
int uvic_b200_set_sbc(uvic_b200_ctx *ctx, int eots, int osegs, int osege, int ntspos);

/* Time averages of the tracers on the device: the tracer part of avgvar / avgout
 * (09/mom/timeavgs.F:206-375, 398-420; avgvar is called from 09/mom/diag.F:138-146 with t(tau) on timavgperts steps) on
 * the default averaging grid of avgset (the whole model grid).  accumulate adds t(tau) and stf (less vflux*gaost(n) for
 * all tracers but T and S; NULL = no virtual flux) of the owned rows to running sums; fetch returns rnavgt*sum
 * (avg_t(imt,km,jl,nt), avg_stf(imt,jl,nt); NULL = skip), the step count, and optionally starts a new period. */
int uvic_b200_tavg_accumulate(uvic_b200_ctx *ctx, const double *vflux, const double *gaost);
int uvic_b200_tavg_fetch(uvic_b200_ctx *ctx, double *avg_t, double *avg_stf, int32_t *navgts, int reset);

/* ---- baroclinic momentum step -----------------------------------------------------------
 * Replaces `call clinic (joff, jscalc, jecalc, is, ie)` (source/mom/mom.F:390; 09/mom/clinic.F:1-511) together with the
 * U-cell half of adv_vel (source/mom/adv_vel.F:160-250) and the momentum half of setvbc (09/mom/setvbc.F:163-208), for
 * the options of run/mk.in: O_consthmix + O_anisotropic_viscosity, O_constvmix, O_stream_function.  The diagnostics
 * hooks (diagc1 / diagc2) and the ice coupling (isbcu / asbcu) stay with the caller; the polar filter of the velocities
 * (filuv, source/common/filuv.F) runs on the device when `fourfil` is set.  Time-invariant inputs, host pointers copied at setup; 2-D / 3-D arrays cover the context's local rows. */
typedef struct uvic_b200_clinic_static {
  const int32_t *kmu;                              /* (imt,jl)    09/common/levind.h */
  const double *hr;                                /* (imt,jl)    1/depth at U points, 09/mom/setmom.F:1108-1111 */
  const double *cori;                              /* (imt,jl,2)  09/mom/setmom.F:777-778 */
  const double *advmet, *am3, *am4;                /* (jmt,2), (jmt), (jmt,2) global, 09/mom/setmom.F:791-802 */
  const double *dxmetr, *dxu2r;                    /* (imt)       source/common/grids.F:470-530 */
  const double *dyu2r, *dyu4r, *csudyu2r;          /* (jmt) global */
  const double *visc_ceu, *amc_north, *amc_south;  /* (imt,km,jl) 09/mom/hmixc.F:62-150 */
  double kappa_m, cdbot, grav_rho0r;               /* 09/common/UVic_ESCM.F:1682-1684; grav*rho0r */
  int32_t fourfil;                                 /* O_fourfil: run filuv (source/common/filuv.F) on u(tau+1) */
  int32_t jfrst, jfu0, jfu1, jfu2;                 /* filter rows (global), source/common/setcom.F:75-83 */
  const double *spsin, *spcos;                     /* (imt) source/common/setcom.F:56-71; may be NULL without fourfil */
  const double *phi;                               /* (jmt) latitude of the U rows in radians (coord.h); may be NULL without fourfil */
} uvic_b200_clinic_static;
int uvic_b200_clinic_setup(uvic_b200_ctx *ctx, const uvic_b200_clinic_static *cs);
/* u(imt,km,jl,2) of a time level: upload level -1 or 0 (0 is the field uvic_b200_upload_u sets), download -1, 0 or +1 */
int uvic_b200_upload_u_level(uvic_b200_ctx *ctx, int level, const double *u_host);
int uvic_b200_download_u(uvic_b200_ctx *ctx, int level, double *u_host);
/* surface momentum flux smf(imt,jl,2) when the wind stress does not come from the coupler array on the device */
int uvic_b200_upload_smf(uvic_b200_ctx *ctx, const double *smf);
/* One momentum step from the resident u(tau-1), u(tau), t(tau) and T-cell advective velocities (uvic_b200_adv_vel or
 * uvic_b200_upload_adv_vel): rho = state(t(tau)), smf / bmf, U-cell advective velocities, u(tau+1) with its vertical
 * mean removed and the cyclic columns set, and zu (the vertical mean of du/dt that forces the barotropic equation).
 * itaux / itauy: 1-based slots of the wind stress in the coupler array of uvic_b200_sbc_setup, or 0, 0 to keep the smf
 * of uvic_b200_upload_smf.  Asynchronous on the context's stream. */
int uvic_b200_clinic(uvic_b200_ctx *ctx, double c2dtuv, int itaux, int itauy);
int uvic_b200_download_zu(uvic_b200_ctx *ctx, double *zu_host);   /* (imt,jl,2) */
int uvic_b200_rotate_u(uvic_b200_ctx *ctx);                        /* tau-1 <- tau <- tau+1 */

int uvic_b200_rotate(uvic_b200_ctx *ctx);

/* ---- the hot path, one entry per reference call site (device resident) ------------ */
int uvic_b200_isopyc(uvic_b200_ctx *ctx);                              /* call isopyc  source/mom/mom.F:340 */
int uvic_b200_vmixc(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *si); /* call vmixc   source/mom/mom.F:347 */
int uvic_b200_tracer(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *si);/* call tracer  source/mom/mom.F:389 */
/* isopyc + vmixc + tracer in one call */
int uvic_b200_step(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *si);

/* Optional look-ahead: the stepinfo of the step AFTER the next call of uvic_b200_tracer / _step / _tracer_step.  The
 * MOBI source terms of a step depend only on t(tau-1) of that step and the 2-D forcing, all of which are final once
 * the current step's kernels have run (source/mom/mom.F:111-146 fixes the schedule: mixing steps every nmix-th itt),
 * so with a hint the library computes them on a side stream while the current t(tau+1) travels to the host.  The
 * result is used only if the next call's stepinfo equals the hint bit for bit and no upload_t / upload_forcing came
 * in between; otherwise it is discarded and recomputed.  Results are identical with and without hints. */
int uvic_b200_hint_next_step(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *next);
/* How often the sources computed ahead were adopted (hit) or had to be recomputed because the step that came was not the
 * step that was hinted (miss: relyr by more than 1e-9 years, co2ccn, dtts, leapfrog or the t(tau-1) slot differ).  `next` of the hint must carry
 * the NEXT step's relyr / co2ccn -- a driver advances relyr every step (source/mom/mom.F). */
int uvic_b200_lookahead_stats(uvic_b200_ctx *ctx, int64_t *hits, int64_t *misses);
/* Drop sources computed ahead: required after writing t, the forcing or the vertical b.c. through raw device pointers
 * (uvic_b200_t_ptr / uvic_b200_device_ptr); the upload_* entry points do it themselves. */
int uvic_b200_invalidate_lookahead(uvic_b200_ctx *ctx);
/* Orders the context's launch stream behind its side streams (look-ahead MOBI, GM velocities), so that an event the caller
 * records next covers all the work issued so far. */
int uvic_b200_join_streams(uvic_b200_ctx *ctx);
/* Latitude slabs: `cuda_event` (a cudaEvent_t) marks the end of the halo exchange of the newest time level.  The next
 * uvic_b200_step lets its coefficient / diffusion kernels (which only read t(tau-1)) run beside the exchange and waits
 * for the event right before its first advection kernel; a mixing step, or a step driven through the separate
 * isopyc / vmixc / tracer entry points, waits at once.  The event is consumed by the wait. */
int uvic_b200_wait_before_advection(uvic_b200_ctx *ctx, void *cuda_event);

/* page-lock / release a host array passed every step (the Fortran COMMON storage), so H2D / D2H copies overlap kernels */
int uvic_b200_pin_host(void *host, size_t bytes);
int uvic_b200_unpin_host(void *host);

/* one synchronous step with HOST buffers (what the Fortran shim calls): uploads
 * t(tau-1), t(tau), the advective velocities and the vertical b.c., runs the step and
 * returns t(tau+1).  NULL inputs keep the resident copy.  Velocities and b.c. are copied on a separate stream under
 * the kernels that do not need them, and t(tau+1) leaves in tracer batches (T,S first) as soon as each is final. */
int uvic_b200_tracer_step(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *si, const double *t_taum1, const double *t_tau,
                          const double *adv_vet, const double *adv_vnt, const double *adv_vbt, const double *stf,
                          const double *btf, double *t_taup1);

/* ---- diagnostics (09/mom/tracer.F:1516-1565) ---------------------------------------- */
/* volume-weighted inventories sum(t*dV) of every tracer at one time level over the rows
 * this context owns, deterministic fixed-order reduction; out(nt) */
/* One ocean step with setvbc / set_sbc on the device (needs uvic_b200_sbc_setup).  Host -> device: the advective
 * velocities (every step) and, when they changed, the coupler's sbc array and the bottom heat flux (NULL = keep).
 * Device -> host: T and S of t(tau+1), ts_taup1(imt,km,jl,2), for the density in clinic / loadmw (NULL = skip), and the
 * whole sbc array with the averaged surface accumulators when this step ends an ocean segment (eots && osege).  All
 * other tracers stay resident; output steps fetch them with uvic_b200_download_tracer.  On return the host velocity
 * buffers have been read and every requested output is on the host ("complete for the outputs requested");
 * the resident tracers of t(tau+1) may still be in flight -- every later call on the context is ordered behind them --
 * so the host's work between two tracer steps overlaps the rest of this one, and the next step's velocity upload starts
 * as soon as this step has formed its total velocities.  Replaces setvbc source/mom/mom.F:360, isopyc :340, vmixc :347,
 * tracer :389. */
int uvic_b200_tracer_step_coupled(uvic_b200_ctx *ctx, const uvic_b200_stepinfo *si, const double *adv_vet,
                                  const double *adv_vnt, const double *adv_vbt, const double *sbc_in, const double *bhf,
                                  int eots, int osegs, int osege, int ntspos, double *ts_taup1, double *sbc_out);

int uvic_b200_inventory(uvic_b200_ctx *ctx, int level, double *out_nt);
/* tbar(km,nt,jl): sum_i t(tau)*dzt*dxt*cst*dyt*tmask per level, tracer and row */
int uvic_b200_tbar(uvic_b200_ctx *ctx, double *tbar_host);
/* basin sums sumbk(3,km,nt) by mskhr */
/* travar(k,n,jrow) = sum_i t(tau)^2 dV and dtabs(k,n,jrow) = sum_i |t(tau+1)-t(tau-1)| dV / (c2dtts dtxcel(k)) of the last step made
 * with diag = 1 (09/mom/tracer.F:1521-1536), shaped like uvic_b200_tbar's output; either pointer may be NULL */
int uvic_b200_travar_dtabs(uvic_b200_ctx *ctx, double *travar_host, double *dtabs_host);
int uvic_b200_sumbk(uvic_b200_ctx *ctx, double *sumbk_host);

/* ---- introspection (tests, halo exchange, profiling) ---------------------------------- */
/* copy a named device array to the host; returns the element count through *nelem when
 * host == NULL.  Names follow the reference's COMMON variables (alphai, K11, diff_cbt ...) */
int uvic_b200_fetch(uvic_b200_ctx *ctx, const char *name, double *host, size_t *nelem);
/* raw device pointer of a named array (for NCCL halo exchange driven by the host) */
void *uvic_b200_device_ptr(uvic_b200_ctx *ctx, const char *name, size_t *nelem);
/* device pointer of t at a time level: t(imt,km,jl,nt) */
void *uvic_b200_t_ptr(uvic_b200_ctx *ctx, int level);
/* number of kernels this context has launched so far */
/* ---- several devices, ONE host thread ------------------------------------------------------------------------------
 * The reference's host is a single serial process (source/mom/mom.F:289-407; SURVEY.md 8b "one process, 8 devices").  A
 * group holds one context per device, each owning a latitude slab of equal estimated work (rows 2..jmt-1 cut by wet-cell
 * count).  Every array argument is the GLOBAL array exactly as the COMMON blocks hold it (j extent jmt; `tlat`, `kmt`, ...
 * of uvic_b200_grid / uvic_b200_static likewise): the library slices it.  group_step queues one ocean step on every device and
 * then lets every device pull the 2-row halos of t(tau+1) from its neighbours' memory (peer copies over NVLink, ordered
 * by events; the consumer waits right before its first advection kernel).  Inventories are per-device partial sums added
 * on the host in device order (fixed order).  Calls are asynchronous unless they return data. */
typedef struct uvic_b200_group uvic_b200_group;
int uvic_b200_group_create(const uvic_b200_dims *dims, const uvic_b200_grid *grid, const uvic_b200_params *par,
                           const uvic_b200_static *st, int ndev, const int32_t *devices /* NULL: 0..ndev-1 */, uvic_b200_group **out);
int uvic_b200_group_destroy(uvic_b200_group *g);
const char *uvic_b200_group_last_error(const uvic_b200_group *g);   /* g may be NULL: last create error */
int uvic_b200_group_size(const uvic_b200_group *g);
uvic_b200_ctx *uvic_b200_group_ctx(uvic_b200_group *g, int r);      /* the r-th context, for the per-context entry points */
int uvic_b200_group_rows(const uvic_b200_group *g, int r, int32_t *jrow_lo, int32_t *jrow_hi);
int uvic_b200_group_upload_t(uvic_b200_group *g, int level, const double *t_global);       /* t(imt,km,jmt,nt) */
int uvic_b200_group_download_t(uvic_b200_group *g, int level, double *t_global);           /* synchronous */
int uvic_b200_group_upload_adv_vel(uvic_b200_group *g, const double *adv_vet, const double *adv_vnt, const double *adv_vbt);
int uvic_b200_group_upload_vbc(uvic_b200_group *g, const double *stf, const double *btf);  /* (imt,jmt,nt) */
int uvic_b200_group_upload_forcing(uvic_b200_group *g, const double *dnswr, const double *aice, const double *hice, const double *hsno);
int uvic_b200_group_step(uvic_b200_group *g, const uvic_b200_stepinfo *si, const uvic_b200_stepinfo *next /* may be NULL */);
int uvic_b200_group_rotate(uvic_b200_group *g);
int uvic_b200_group_inventory(uvic_b200_group *g, int level, double *out_nt);              /* synchronous */
int uvic_b200_group_synchronize(uvic_b200_group *g);

/* Measurement helper (bench.py): the FP64 instruction rates of this device -- thread-level DFMA, DADD, DMUL per second from
 * dependent-chain kernels with no memory traffic, best of three launches -- the roof the MOBI and flux kernels are measured
 * against ("reported against both roofs").  Not part of the reference interface. */
int uvic_b200_measure_fp64_peak(int device, double *dfma_per_s, double *dadd_per_s, double *dmul_per_s, double *sm_clock_mhz);

int64_t uvic_b200_kernel_launches(const uvic_b200_ctx *ctx);
/* per-kernel device time, measured with CUDA events on the launch stream while enabled */
int uvic_b200_profile_enable(uvic_b200_ctx *ctx, int on);
int uvic_b200_profile_count(uvic_b200_ctx *ctx); /* synchronises; number of distinct kernels seen */
int uvic_b200_profile_get(uvic_b200_ctx *ctx, int idx, char *name, int name_len, double *total_ms, int64_t *count);
int uvic_b200_profile_reset(uvic_b200_ctx *ctx);
int uvic_b200_local_rows(const uvic_b200_ctx *ctx, int32_t *jbase, int32_t *jl);
const char *uvic_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
