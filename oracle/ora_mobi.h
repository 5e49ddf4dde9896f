/*
 * ora_mobi.h -- parameter block and index maps of the MOBI restatement (ora_mobi.c).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * The parameter block holds the &mobi namelist values after mobi_init's unit conversion
 * (09/mom/mobi.F:209-294), as plain doubles in the order below; the product library
 * receives the same block through uvic_b200_params.mobi_par (include/uvic_b200_mobi.h
 * lists the same order), so the two sides cannot disagree on a parameter value.
 */
#ifndef UVIC_ORA_MOBI_H
#define UVIC_ORA_MOBI_H

#define ORA_MOBI_NVAR 32   /* ntnpzd with the shipped options, 09/mom/mobi.h:104-142 */
#define ORA_MOBI_NIDX 128
#define ORA_MOBI_KMAX 128

/* MOBI-internal state order fixed by mobi_init's setimobi sequence (09/mom/mobi.F:440-497), 0-based here */
enum {
  M_PO4 = 0, M_PHYT, M_PHYT_PHOS, M_ZOOP, M_DETR, M_DETR_PHOS, M_DIC, M_DIC13, M_PHYTC13, M_ZOOPC13,
  M_DETRC13, M_DOC13, M_DIAZC13, M_DIATC13, M_CACO3C13, M_DOP, M_NO3, M_DON, M_DIAZ, M_DIN15, M_DON15,
  M_PHYTN15, M_ZOOPN15, M_DETRN15, M_DIAZN15, M_DIATN15, M_CACO3, M_DIAT, M_SIL, M_OPL, M_DFE, M_DETRFE
};
/* mobi_idx layout: [0..31] tracer index (1-based) of state m; [32..63] source slot (1-based) of state m;
 * then the extra tracers / sources */
enum { MI_TR = 0, MI_SRC = 32, MI_ITEMP = 64, MI_ISALT, MI_IALK, MI_IO2, MI_IC14, MI_ISALK, MI_ISO2, MI_ISC14, MI_N };

typedef struct ora_mobi_par {
  double kw, kc, ki, tap, abio_P, bbio, cbio, nup, nup_D, nupt0, nupt0_D, gamma1, gbio, nuz, nud0, nudon0, nudop0;
  double dtnpzd, redctn, redptn, redotn, redotc, redntp, redntc, diazptn, diazntp, caprmax, kcapr, dissk0, kc_c;
  double jdiar, dbct_D, kzoo, geZ, dfr, pfr, dfrt, hdop, abiodiat, nu_diat, nudt0, opl_disk0;
  double zprefP, zprefDiat, zprefDiaz, zprefZ, zprefDet;
  double eps_assim, eps_excr, eps_nfix, eps_wcdeni, eps_bdeni0, eps_recy;
  double kfemin, kfemax, knmin, knmax, pmax, kfe_D, kfemin_Diat, kfemax_Diat, knmin_Diat, knmax_Diat, pmax_Diat;
  double kfeleq, thetamaxhi, thetamaxlo, alphamax, alphamin, mc, kfeorg, rfeton, iscr, kfecol;
  double reserved[6];
  double wd[ORA_MOBI_KMAX], wc[ORA_MOBI_KMAX], wo[ORA_MOBI_KMAX], ztt[ORA_MOBI_KMAX];
} ora_mobi_par;

#endif
