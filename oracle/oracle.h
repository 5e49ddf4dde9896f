/*
 * oracle.h -- CPU restatement of the UVic ESCM 2.9 ocean tracer step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (uvic2.9_b200/, the C-ABI
 * library, bench.py's GPU arm) may include, link or call this code; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY PINNED against the reference's own code: the reference ships no tests, golden vectors or
 * fixtures and no Fortran compiler exists here, so oracle/refgen cpp-expands the reference sources
 * exactly as `mk` does and translates them to C MECHANICALLY (f2c.py, no hand editing) into
 * oracle/_ref/libref_<tag>.so; tests/test_cpu_refpin.py drives that library and this restatement
 * with the same inputs and requires BITWISE equality (isopyc, vmixc, adv_flux, isoflux, mobi_init,
 * mobi_src on 10^4 random cells, co2calc_SWS, state, adv_vel, diagt1, filt/filtr/findex, whole
 * `tracer` steps with MOBI over leapfrog and mixing steps, set_sbc, setvbc, gasbc's flux loop, clinic + filuv).  Every routine follows the reference's
 * array shapes, index ranges, loop order and operation order, and cites the
 * reference file:line it restates (paths relative to /root/reference; "09/" means
 * updates/09/source/, the update level run/mk.in:204 selects).
 *
 * Arrays are held exactly in the Fortran column-major layouts of 09/mom/mw.h and
 * 09/common/isopyc.h, with one simplification the reference itself allows: the memory
 * window is fully open (jmw=jmt, 09/common/size.h:154), so arrays dimensioned
 * jsmw:jemw or 1:jemw are simply allocated 1:jmt, and the (...,nt) extent on the FCT
 * scratch arrays (09/mom/mw.h:389-393) is dropped because with one window it is pure
 * scratch (nothing persists across the n loop).
 */
#ifndef UVIC_ORACLE_H
#define UVIC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_MAXARR 320

typedef struct ora_arr {
  const char *name;
  void *ptr;
  size_t nelem;   /* number of elements */
  int is_int;     /* 1: int32, 0: double */
} ora_arr;

/* MOBI parameters after the unit conversion done in mobi_init (09/mom/mobi.F:209-294). */
typedef struct ora_mobi_par ora_mobi_par;

typedef struct ora_ctx {
  int imt, jmt, km, nt, nsrc;

  /* ---- scalars (09/common/scalar.h, vmixc.h, isopyc.h, tidal_kv.h, hmixc.h) ---- */
  double dtts, c2dtts;          /* source/mom/mom.F:111-146 */
  double aidif, kappa_h;        /* source/common/vmixc.h */
  double ahisop, athkdf, slmxr; /* 09/mom/isopyc.F:70-110 */
  double diff_cet, diff_cnt;    /* 09/mom/hmixc.F:192-200 */
  double zetar, ogamma, gravrho0r; /* 09/mom/setmom.F:80-82 */
  int fct;                      /* 1: O_fct branch, 0: 2nd-order centred */
  int isopycmix;                /* 1: O_isopycmix + O_gent_mcwilliams on */
  int tidal_kv;                 /* 1: O_tidal_kv on (else diff_cbt = kappa_h) */
  int do_convect;               /* O_fullconvect */
  int do_mobi;                  /* O_mobi */
  int timavgperts;              /* diagnostics switch used by convct2 */

  /* ---- registry of every array, so tests can fill / read by name ---- */
  ora_arr arr[ORA_MAXARR];
  int narr;

  /* ---- integer maps ---- */
  int32_t *kmt;     /* (imt,jmt)            09/common/levind.h */
  int32_t *itrc;    /* (nt) source slot of tracer n, 0 = none   09/mom/mw.h:125-221 */
  int32_t *mskhr;   /* (imt,jmt) horizontal region mask  source/common/cregin.h */

  /* ---- grid (source/common/grdvar.h, coord.h, accel.h) ---- */
  double *dxt, *dxtr, *dxt2r, *dxt4r, *dxu, *dxur;                 /* (imt) */
  double *dyt, *dytr, *dyt2r, *dyt4r, *dyu, *dyur;                 /* (jmt) */
  double *cst, *cstr, *csu, *csur, *cstdytr, *cstdyt2r, *csu_dyur; /* (jmt) */
  double *dzt, *dztr, *dzt2r, *dztur, *dztlr, *zt, *zw;            /* (km)  */
  double *dzw, *dzwr;                                              /* (0:km) */
  double *dtxcel, *dtxsqr, *dztxcl, *dzwxcl;                       /* (km)  */
  double *tlat;                                                    /* (imt,jmt) */
  double *duw, *due, *dus, *dun;                                   /* adv_vel only */

  /* ---- equation of state (source/mom/state.h) ---- */
  double *eosc;  /* c(km,9) */
  double *to, *so;

  /* ---- prognostic + forcing (09/mom/mw.h) ---- */
  double *t;        /* t(imt,km,jmt,nt,-1:1) */
  double *u;        /* u(imt,km,jmt,2) at tau (only for adv_vel) */
  double *rho;      /* (imt,km,jmt) normalised density of t(tau), source/mom/state.F */
  double *tmask, *umask;    /* (imt,km,jmt) */
  double *adv_vet, *adv_vnt;/* (imt,km,jmt) */
  double *adv_vbt;          /* (imt,0:km,jmt) */
  double *stf, *btf;        /* (imt,jmt,nt) */
  double *src;              /* src(imt,km,jmt,nsrc) */

  /* ---- isopycnal mixing (09/common/isopyc.h) ---- */
  double *alphai, *betai;            /* (imt,km,jmt) */
  double *ddxt, *ddyt;               /* (imt,km,jmt,2) */
  double *ddzt;                      /* (imt,0:km,jmt,2) */
  double *Ai_ez, *Ai_nz, *Ai_bx, *Ai_by; /* (imt,km,jmt,0:1,0:1) */
  double *K11, *K22, *K33;           /* (imt,km,jmt) */
  double *fisop;                     /* (imt,jmt,km)  note (i,j,k) order */
  double *addisop;                   /* (imt,km,jmt) */
  double *adv_vetiso, *adv_vntiso;   /* (imt,km,jmt) */
  double *adv_vbtiso, *adv_fbiso;    /* (imt,0:km,jmt) */
  double *drodxte, *drodxbe, *drodytn, *drodybn;
  double *drodzte, *drodzbe, *drodztn, *drodzbn; /* (imt,km,jmt) */

  /* ---- vertical mixing (source/common/vmixc.h, 09/mom/tidal_kv.h) ---- */
  double *diff_cbt;                         /* (imt,km,jmt) */
  double *edrm2, *edrs2, *edrk1, *edro1;    /* (imt,km,jmt) */

  /* ---- work arrays of tracer / adv_flux (09/mom/mw.h:246-393) ---- */
  double *adv_fe, *adv_fn;      /* (imt,km,jmt) */
  double *adv_fb;               /* (imt,0:km,jmt) */
  double *diff_fe, *diff_fn;    /* (imt,km,jmt) */
  double *diff_fb, *diff_fbiso; /* (imt,0:km,jmt) */
  double *source;               /* (imt,km,jmt) */
  double *anti_fe, *anti_fn;    /* (imt,km,jmt) */
  double *anti_fb;              /* (imt,0:km,jmt) */
  double *R_plusY, *R_minusY;   /* (imt,km,jmt) */
  double *cstdxtr, *cstdxt2r, *cstdxur, *ah_cstdxur; /* (imt,jmt) */
  double *t_lo_dump;            /* (imt,km,jmt) copy of t_lo per row, diagnostics for tests */
  double *texp_dump;            /* (imt,km,jmt,nt) explicit t(tau+1) before invtri, for tests */

  /* ---- diagnostics ---- */
  double *tbar, *travar, *dtabs; /* (km,nt,jmt)  09/mom/tracer.F:1516-1539 */
  double *sumbk;                 /* (nhreg=3,km,nt) */
  double *totalk, *vdepth, *pe;  /* (imt,jmt)  convct2 diagnostics */

  /* ---- MOBI inputs (09/mom/mobi.h, 09/mom/tracer.F:310-545) ---- */
  double *dnswr, *aice, *hice, *hsno;  /* (imt,jmt) */
  double *sg_bathy;                    /* (imt,jmt,km) */
  double *fe_hydr;                     /* (imt,jmt,km) */
  double *fe_atmdep;                   /* (imt,jmt,12) */
  double relyr, co2ccn;
  ora_mobi_par *mobi;
  int32_t *mobi_idx;   /* tracer index maps, see ora_mobi.h */

  /* ---- surface boundary conditions (09/common/csbc.h, 09/mom/setvbc.F, 09/mom/set_sbc.F) ---- */
  int numsbc;
  double *sbc;             /* (imt,jmt,numsbc) */
  double *bhf;             /* (imt,jmt) bottom heat flux */
  int32_t *sbc_flx_index;  /* (nt) 1-based sbc slot of tracer n's surface flux, 0 = none */
  int32_t *trsbcindex;     /* (nt) 1-based sbc slot of tracer n's surface accumulator, 0 = none */
  int eots, osegs, osege, ntspos;   /* source/common/switch.h */
  int32_t *gas_idx;        /* (15) sbc slots of the gas exchange, see ora_gasbc */
  double dc13ccn, dc14ccn; /* atmospheric delta 13C, delta 14C (09/common/cembm.h) */

  /* ---- time averages of the tracers (09/mom/timeavgs.F) ---- */
  double *spbuf_t, *avg_t;       /* (imt,km,jmt,nt) running sums / means of t(tau) */
  double *spbuf2_stf, *avg_stf;  /* (imt,jmt,nt) running sums / means of the surface tracer flux */
  double *vflux, *gaost;         /* (imt,jmt) virtual flux, (nt) its tracer factors */
  int navgts;

  /* ---- baroclinic momentum (09/mom/clinic.F, source/mom/adv_vel.F:160-250, 09/mom/setvbc.F:163-208) ---- */
  int32_t *kmu;                         /* (imt,jmt) levels at U points, 09/common/levind.h */
  double *um1, *up1;                    /* u(imt,km,jmt,2) at tau-1 and tau+1 (c->u is tau) */
  double *adv_veu, *adv_vnu;            /* (imt,km,jmt) */
  double *adv_vbu;                      /* (imt,0:km,jmt) */
  double *smf, *bmf;                    /* (imt,jmt,2) surface / bottom momentum flux */
  double *hr;                           /* (imt,jmt) 1/depth at U points, 09/mom/setmom.F:1108-1111 */
  double *cori;                         /* (imt,jmt,2) 09/mom/setmom.F:777-778 */
  double *advmet, *am3, *am4;           /* (jmt,2), (jmt), (jmt,2) 09/mom/setmom.F:791-802 */
  double *dxmetr, *dxu2r;               /* (imt) source/common/grids.F */
  double *dyu2r, *dyu4r, *csudyu2r;     /* (jmt) */
  double *visc_ceu, *amc_north, *amc_south, *visc_cbu; /* (imt,km,jmt) 09/mom/hmixc.F:62-150, 09/mom/vmixc.F:85 */
  double *grad_p;                       /* (imt,km,jmt,2) */
  double *zu;                           /* (imt,jmt,2) vertical mean of du/dt, forcing of tropic */
  double *baru;                         /* (imt,jmt,2) */
  double *csudxur, *csudxu2r;           /* (imt,jmt) 09/mom/mw.h */
  double *am_csudxtr, *tempik;          /* (imt,km,jmt) */
  double c2dtuv, kappa_m, cdbot, grav_rho0r;
  int itaux, itauy;                     /* 1-based sbc slots of the wind stress */

  /* ---- Fourier filter (source/common/index.h) ---- */
  int do_filter;
  int jfrst, jft1, jft2, jft0;  /* source/common/setcom.F */
  void *filt_state;
  int jfu0, jfu1, jfu2;         /* source/common/setcom.F:79-81 */
  double *spsin, *spcos;        /* (imt) source/common/setcom.F:56-71 */
  double *phi;                  /* (jmt) latitude of U rows in radians, source/common/coord.h */
  void *filtu_state;
} ora_ctx;

/* context management (ora_core.c) */
ora_ctx *ora_create(int imt, int jmt, int km, int nt, int nsrc);
void ora_destroy(ora_ctx *c);
void *ora_array(ora_ctx *c, const char *name, size_t *nelem, int *is_int);
int ora_narrays(const ora_ctx *c);
const char *ora_array_name(const ora_ctx *c, int idx);
int ora_set_scalar(ora_ctx *c, const char *name, double v);
double ora_get_scalar(ora_ctx *c, const char *name);

/* hot-path routines (one translation unit per reference file) */
void ora_make_masks(ora_ctx *c);                       /* 09/mom/loadmw.F:60-77 */
void ora_adv_vel(ora_ctx *c);                          /* source/mom/adv_vel.F:60-131 */
void ora_state(ora_ctx *c);                            /* source/mom/state.F:1-60 via 09/mom/loadmw.F:150-155 */
void ora_isopyc(ora_ctx *c);                           /* 09/mom/isopyc.F:466-557 */
void ora_vmixc(ora_ctx *c);                            /* 09/mom/vmixc.F:68-188 */
void ora_adv_flux(ora_ctx *c, int n);                  /* 09/mom/tracer_adv_flx.F */
void ora_isoflux(ora_ctx *c, int n);                   /* 09/mom/isopyc.F:923-1138 */
void ora_invtri(ora_ctx *c, double *z, const double *topbc, const double *botbc,
                const double *dcb, const double *tdt); /* source/mom/invtri.F */
void ora_convct2(ora_ctx *c, double *ts);              /* source/mom/convect.F:99-311 */
void ora_tracer(ora_ctx *c);                           /* 09/mom/tracer.F:214-1364 */
void ora_diag_tbar(ora_ctx *c, int n);                 /* 09/mom/tracer.F:1516-1565 */
void ora_mobi_columns(ora_ctx *c);                     /* 09/mom/tracer.F:310-545,848-867 */
void ora_filt(ora_ctx *c);                             /* source/common/filt.F */
void ora_filuv(ora_ctx *c);                            /* source/common/filuv.F */
void ora_avgvar(ora_ctx *c);                           /* 09/mom/timeavgs.F:206-375 (tracer part) */
void ora_avgout(ora_ctx *c);                           /* 09/mom/timeavgs.F:398-420 (time means) */
void ora_setvbc(ora_ctx *c);                           /* 09/mom/setvbc.F:60-140 */
void ora_gasbc(ora_ctx *c);                            /* 09/common/gasbc.F:148-266 (gas exchange loop) */
void ora_set_sbc(ora_ctx *c);                          /* 09/mom/set_sbc.F:36-83 via 09/mom/tracer.F:1270-1288 */
void ora_adv_vel_u(ora_ctx *c);                        /* source/mom/adv_vel.F:160-250 */
void ora_setvbc_mom(ora_ctx *c);                       /* 09/mom/setvbc.F:163-208 */
void ora_clinic(ora_ctx *c);                           /* 09/mom/clinic.F:60-560 (run/mk.in options; filuv when do_filter) */

/* one full step as mom.F sequences it: isopyc -> vmixc -> tracer (source/mom/mom.F:340-389) */
void ora_step(ora_ctx *c);

#ifdef __cplusplus
}
#endif
#endif
