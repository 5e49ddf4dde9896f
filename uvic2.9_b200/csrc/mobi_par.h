// mobi_par.h -- layout of the MOBI parameter block and index maps passed through
// uvic_b200_params.mobi_par / mobi_index (values after mobi_init's unit conversion,
// 09/mom/mobi.F:191-262; built on the host by uvic2.9_b200/mobi_params.py).
#pragma once

#define MOBI_NVAR 32   // ntnpzd with the options of run/mk.in (09/mom/mobi.h:104-142)
#define MOBI_KMAX 128
#define MOBI_NIDX 128
#define MOBI_NPRE 16   // per-cell fields of the MOBI pre-pass (k_mobi.cu, enum MobiPre)

// MOBI-internal state order = mobi_init's setimobi sequence (09/mom/mobi.F:440-497)
enum MobiVar {
  V_PO4 = 0, V_PHYT, V_PHYT_PHOS, V_ZOOP, V_DETR, V_DETR_PHOS, V_DIC, V_DIC13, V_PHYTC13, V_ZOOPC13, V_DETRC13,
  V_DOC13, V_DIAZC13, V_DIATC13, V_CACO3C13, V_DOP, V_NO3, V_DON, V_DIAZ, V_DIN15, V_DON15, V_PHYTN15, V_ZOOPN15,
  V_DETRN15, V_DIAZN15, V_DIATN15, V_CACO3, V_DIAT, V_SIL, V_OPL, V_DFE, V_DETRFE
};
// index map: [0..31] tracer index (1-based) of state m, [32..63] source slot (1-based) of state m, then extras
enum MobiIdx { IX_TR = 0, IX_SRC = 32, IX_ITEMP = 64, IX_ISALT, IX_IALK, IX_IO2, IX_IC14, IX_ISALK, IX_ISO2, IX_ISC14, IX_N };

struct MobiPar {
  double kw, kc, ki, tap, abio_P, bbio, cbio, nup, nup_D, nupt0, nupt0_D, gamma1, gbio, nuz, nud0, nudon0, nudop0;
  double dtnpzd, redctn, redptn, redotn, redotc, redntp, redntc, diazptn, diazntp, caprmax, kcapr, dissk0, kc_c;
  double jdiar, dbct_D, kzoo, geZ, dfr, pfr, dfrt, hdop, abiodiat, nu_diat, nudt0, opl_disk0;
  double zprefP, zprefDiat, zprefDiaz, zprefZ, zprefDet;
  double eps_assim, eps_excr, eps_nfix, eps_wcdeni, eps_bdeni0, eps_recy;
  double kfemin, kfemax, knmin, knmax, pmax, kfe_D, kfemin_Diat, kfemax_Diat, knmin_Diat, knmax_Diat, pmax_Diat;
  double kfeleq, thetamaxhi, thetamaxlo, alphamax, alphamin, mc, kfeorg, rfeton, iscr, kfecol;
  double reserved[6];
  double wd[MOBI_KMAX], wc[MOBI_KMAX], wo[MOBI_KMAX], ztt[MOBI_KMAX];
};
