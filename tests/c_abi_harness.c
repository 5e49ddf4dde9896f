/*
 * c_abi_harness.c -- a host program written against include/uvic_b200.h alone (no Python, no ctypes prototypes): the
 * arrays live in STATIC storage laid out the way the reference's COMMON blocks are (Fortran column-major, sizes fixed at
 * compile time like size.h's parameters: -DIMT -DJMT -DKM -DNT -DNSRC), filled from a flat input file, and the program
 * drives the library as the Fortran shim does:
 *     uvic_b200_create -> uvic_b200_upload_t x2 -> uvic_b200_upload_forcing -> uvic_b200_sbc_setup
 *     -> { uvic_b200_hint_next_step, uvic_b200_tracer_step_coupled, uvic_b200_rotate } x NSTEP -> uvic_b200_download_t
 * or, with `group N`, through the one-host-thread multi-device interface (uvic_b200_group_*; N slabs).
 * tests/test_gpu_cabi.py compares its output with the Python path bit for bit.  TEST INFRASTRUCTURE.
 *
 * Input file (doubles and int32 in this order; written by tests/test_gpu_cabi.py): every array of uvic_b200_grid, the
 * scalars of uvic_b200_params, itrc, mobi_index[128], mobi_par[n], kmt, mskhr, the static 3-D fields, t(tau-1), t(tau),
 * adv_vet/vnt/vbt, dnswr, aice, hice, hsno, dtts, relyr, co2ccn.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "uvic_b200.h"

#ifndef IMT
#error "compile with -DIMT= -DJMT= -DKM= -DNT= -DNSRC= -DNMOBIPAR="
#endif
#define N2 (IMT * JMT)
#define N3 (IMT * KM * JMT)
#define N3Z (IMT * (KM + 1) * JMT)
#define NSBC (2 * NT + 4)
#define NSTEP 4
#define NTSPOS 4

/* grdvar.h / coord.h / accel.h / state.h */
static double dxt[IMT], dxtr[IMT], dxt2r[IMT], dxt4r[IMT], dxu[IMT], dxur[IMT];
static double dyt[JMT], dytr[JMT], dyt2r[JMT], dyt4r[JMT], dyu[JMT], dyur[JMT];
static double cst[JMT], cstr[JMT], csu[JMT], csur[JMT], cstdytr[JMT], cstdyt2r[JMT], csu_dyur[JMT];
static double dzt[KM], dztr[KM], dzt2r[KM], dztur[KM], dztlr[KM], zt[KM], zw[KM], dzw[KM + 1], dzwr[KM + 1];
static double dtxcel[KM], dtxsqr[KM], dztxcl[KM], dzwxcl[KM], tlat[N2], duw[IMT], due[IMT], dus[JMT], dun[JMT];
static double eosc[KM * 9], to[KM], so[KM];
static double scal[10];           /* aidif kappa_h ahisop athkdf slmxr diff_cet diff_cnt zetar ogamma gravrho0r */
static int32_t itrc[NT], mobi_index[128], kmt[N2], mskhr[N2];
static double mobi_par[NMOBIPAR];
/* isopyc.h, tidal_kv.h, levind.h, mobi.h */
static double fisop[N3], addisop[N3], edrm2[N3], edrs2[N3], edrk1[N3], edro1[N3], sg_bathy[N3], fe_hydr[N3], fe_atmdep[N2 * 12];
/* mw.h */
static double t_m1[(size_t)N3 * NT], t_0[(size_t)N3 * NT], t_out[(size_t)N3 * NT];
static double adv_vet[N3], adv_vnt[N3], adv_vbt[N3Z], stf[(size_t)N2 * NT], btf[(size_t)N2 * NT];
static double dnswr[N2], aice[N2], hice[N2], hsno[N2], step_scal[3];
/* csbc.h */
static double sbc[(size_t)N2 * NSBC], sbc_out[(size_t)N2 * NSBC], bhf[N2], ts_p1[(size_t)N3 * 2];
static int32_t flx_index[NT], acc_index[NT];

static FILE *fin;
#define RD(a) do { if (fread((a), sizeof((a)[0]), sizeof(a) / sizeof((a)[0]), fin) != sizeof(a) / sizeof((a)[0])) { fprintf(stderr, "short read: %s\n", #a); return 2; } } while (0)
#define CALL(x) do { if ((x) != 0) { fprintf(stderr, "%s failed: %s\n", #x, uvic_b200_last_error(ctx)); return 1; } } while (0)
#define GCALL(x) do { if ((x) != 0) { fprintf(stderr, "%s failed: %s\n", #x, uvic_b200_group_last_error(grp)); return 1; } } while (0)

static int leapfrog_at(int itt) { return (itt % 3) != 0; }   /* a mixing step every third step (nmix = 3) */

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s input.bin output.bin [group N]\n", argv[0]); return 2; }
  const int ngroup = (argc >= 5 && strcmp(argv[3], "group") == 0) ? atoi(argv[4]) : 0;
  fin = fopen(argv[1], "rb");
  if (!fin) { perror(argv[1]); return 2; }
  RD(dxt); RD(dxtr); RD(dxt2r); RD(dxt4r); RD(dxu); RD(dxur);
  RD(dyt); RD(dytr); RD(dyt2r); RD(dyt4r); RD(dyu); RD(dyur);
  RD(cst); RD(cstr); RD(csu); RD(csur); RD(cstdytr); RD(cstdyt2r); RD(csu_dyur);
  RD(dzt); RD(dztr); RD(dzt2r); RD(dztur); RD(dztlr); RD(zt); RD(zw); RD(dzw); RD(dzwr);
  RD(dtxcel); RD(dtxsqr); RD(dztxcl); RD(dzwxcl); RD(tlat); RD(duw); RD(due); RD(dus); RD(dun); RD(eosc); RD(to); RD(so);
  RD(scal); RD(itrc); RD(mobi_index); RD(mobi_par); RD(kmt); RD(mskhr);
  RD(fisop); RD(addisop); RD(edrm2); RD(edrs2); RD(edrk1); RD(edro1); RD(sg_bathy); RD(fe_hydr); RD(fe_atmdep);
  RD(t_m1); RD(t_0); RD(adv_vet); RD(adv_vnt); RD(adv_vbt); RD(dnswr); RD(aice); RD(hice); RD(hsno); RD(step_scal);
  fclose(fin);

  uvic_b200_dims d = {IMT, JMT, KM, NT, NSRC, 2, JMT - 1};
  uvic_b200_grid g;
  memset(&g, 0, sizeof g);
  g.dxt = dxt; g.dxtr = dxtr; g.dxt2r = dxt2r; g.dxt4r = dxt4r; g.dxu = dxu; g.dxur = dxur;
  g.dyt = dyt; g.dytr = dytr; g.dyt2r = dyt2r; g.dyt4r = dyt4r; g.dyu = dyu; g.dyur = dyur;
  g.cst = cst; g.cstr = cstr; g.csu = csu; g.csur = csur; g.cstdytr = cstdytr; g.cstdyt2r = cstdyt2r; g.csu_dyur = csu_dyur;
  g.dzt = dzt; g.dztr = dztr; g.dzt2r = dzt2r; g.dztur = dztur; g.dztlr = dztlr; g.zt = zt; g.zw = zw; g.dzw = dzw; g.dzwr = dzwr;
  g.dtxcel = dtxcel; g.dtxsqr = dtxsqr; g.dztxcl = dztxcl; g.dzwxcl = dzwxcl; g.tlat = tlat;
  g.duw = duw; g.due = due; g.dus = dus; g.dun = dun; g.eosc = eosc; g.to = to; g.so = so;
  uvic_b200_params p;
  memset(&p, 0, sizeof p);
  p.aidif = scal[0]; p.kappa_h = scal[1]; p.ahisop = scal[2]; p.athkdf = scal[3]; p.slmxr = scal[4];
  p.diff_cet = scal[5]; p.diff_cnt = scal[6]; p.zetar = scal[7]; p.ogamma = scal[8]; p.gravrho0r = scal[9];
  p.fct = 1; p.isopycmix = 1; p.tidal_kv = 1; p.fullconvect = 1; p.mobi = 1; p.fourfil = 0;
  p.itrc = itrc; p.mobi_index = mobi_index; p.mobi_par = mobi_par; p.n_mobi_index = 128; p.n_mobi_par = NMOBIPAR;
  uvic_b200_static s;
  memset(&s, 0, sizeof s);
  s.kmt = kmt; s.mskhr = mskhr; s.fisop = fisop; s.addisop = addisop; s.edrm2 = edrm2; s.edrs2 = edrs2; s.edrk1 = edrk1;
  s.edro1 = edro1; s.sg_bathy = sg_bathy; s.fe_hydr = fe_hydr; s.fe_atmdep = fe_atmdep;
  uvic_b200_stepinfo si, nx;
  memset(&si, 0, sizeof si);
  si.dtts = step_scal[0]; si.relyr = step_scal[1]; si.co2ccn = step_scal[2];
  nx = si;

  if (ngroup > 0) {
    /* one host thread, `ngroup` slabs (all on device 0 when the box has one GPU) */
    uvic_b200_group *grp = NULL;
    int32_t devs[16];
    int ndev_box = 1;
    if (getenv("HARNESS_NDEV")) ndev_box = atoi(getenv("HARNESS_NDEV"));
    for (int r = 0; r < ngroup && r < 16; r++) devs[r] = r % ndev_box;
    if (uvic_b200_group_create(&d, &g, &p, &s, ngroup, devs, &grp) != 0) { fprintf(stderr, "group_create: %s\n", uvic_b200_group_last_error(NULL)); return 1; }
    GCALL(uvic_b200_group_upload_t(grp, -1, t_m1));
    GCALL(uvic_b200_group_upload_t(grp, 0, t_0));
    GCALL(uvic_b200_group_upload_adv_vel(grp, adv_vet, adv_vnt, adv_vbt));
    GCALL(uvic_b200_group_upload_vbc(grp, stf, btf));
    GCALL(uvic_b200_group_upload_forcing(grp, dnswr, aice, hice, hsno));
    for (int itt = 1; itt <= NSTEP; itt++) {
      si.leapfrog = leapfrog_at(itt);
      nx.leapfrog = leapfrog_at(itt + 1);
      GCALL(uvic_b200_group_step(grp, &si, &nx));
      GCALL(uvic_b200_group_rotate(grp));
    }
    GCALL(uvic_b200_group_download_t(grp, 0, t_out));
    uvic_b200_group_destroy(grp);
  } else {
    uvic_b200_ctx *ctx = NULL;
    if (uvic_b200_create(&d, &g, &p, &s, 0, &ctx) != 0) { fprintf(stderr, "create: %s\n", uvic_b200_last_error(NULL)); return 1; }
    CALL(uvic_b200_upload_t(ctx, -1, t_m1));
    CALL(uvic_b200_upload_t(ctx, 0, t_0));
    CALL(uvic_b200_upload_forcing(ctx, dnswr, aice, hice, hsno));
    for (int n = 0; n < NT; n++) { flx_index[n] = n + 1; acc_index[n] = NT + n + 1; }
    CALL(uvic_b200_sbc_setup(ctx, NSBC, flx_index, acc_index));
    for (int itt = 1; itt <= NSTEP; itt++) {
      const int pseg = (itt - 1) % NTSPOS;
      si.leapfrog = leapfrog_at(itt);
      nx.leapfrog = leapfrog_at(itt + 1);
      CALL(uvic_b200_hint_next_step(ctx, &nx));
      CALL(uvic_b200_tracer_step_coupled(ctx, &si, adv_vet, adv_vnt, adv_vbt, pseg == 0 ? sbc : NULL, pseg == 0 ? bhf : NULL, 1, pseg == 0,
                                         pseg == NTSPOS - 1, NTSPOS, ts_p1, sbc_out));
      CALL(uvic_b200_rotate(ctx));
    }
    CALL(uvic_b200_download_t(ctx, 0, t_out));
    uvic_b200_destroy(ctx);
  }
  FILE *fo = fopen(argv[2], "wb");
  if (!fo) { perror(argv[2]); return 2; }
  fwrite(t_out, sizeof(double), (size_t)N3 * NT, fo);
  if (ngroup == 0) {
    fwrite(ts_p1, sizeof(double), (size_t)N3 * 2, fo);
    fwrite(sbc_out, sizeof(double), (size_t)N2 * NSBC, fo);
  }
  fclose(fo);
  printf("c_abi_harness: %d steps, %s, ok\n", NSTEP, ngroup ? "group" : "single context");
  return 0;
}
