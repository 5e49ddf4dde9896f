/*
 * ora_mobi.c -- restatement of the MOBI biogeochemistry path with the options of
 * run/mk.in (O_mobi, O_mobi_alk, _caco3, _o2, _nitrogen, _nitrogen_15, _silicon, _iron,
 * O_carbon, O_carbon_13, O_carbon_14):
 *   ora_mobi_columns  the column prologue in tracer   09/mom/tracer.F:310-545, 848-867
 *   mobi_driver       one water column                09/mom/mobi.F:519-1483
 *   mobi_src          the ecosystem ODE per cell      09/mom/mobi.F:1485-3313
 * COMMON-block quantities the reference rewrites per cell (ptn_P, k1n, k1p_P,
 * alpha_Diat, sipr0, capr) are locals here.  A real exponent with
 * the literal value 2. is evaluated as x*x (what gcc / gfortran make of pow(x, 2.0) at
 * every optimisation level); x**(0.5) stays pow(x, 0.5), as in oracle/_ref (the pin
 * decided: glibc's pow(x, 0.5) differs from sqrt(x) by one ulp for 0.08 % of arguments).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include <math.h>
#include <string.h>
#include "oracle.h"
#include "ora_index.h"
#include "ora_mobi.h"

void ora_co2calc_SWS(double t, double s, double dic_in, double ta_in, double co2_in, double atmpres, double depth, double *ph,
                     double *co2star, double *dco2star, double *pCO2, double *dpco2, double *CO3, double *Omega_c, double *Omega_a);

#define TRCMIN 5e-12          /* 09/mom/mobi.h:199 */
#define RN15STD 0.0036765
#define RC13STD 0.0112372
#define RC14STD 1.176e-12
#define DAYLEN 86400.0

static inline double sgn(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); } /* Fortran sign(a,b) */
static inline double flag_of(double x) { return 0.5 + sgn(0.5, x - TRCMIN); }

typedef struct {
  /* outputs of mobi_src besides bioout */
  double nfix, expo, expo_phos, calpro, dissl, expocaco3, expoopl, rn15expo, rc13expo, rcaco3c13expo, expofe, remife;
} src_out;

/* 09/mom/mobi.F:1485-3313.  bioin is clipped in place (:1894), which the caller sees. */
static void mobi_src(const ora_mobi_par *P, int nbio, double dtbio, double capr, double *bioin, double gl, double bct, double impo,
                     double dzt, double impo_phos, double dayfrac, double wwd, double nud, double impocaco3, double wwc, double dissk1,
                     double impoopl, double wwo, double opl_disk1, double nudop, double nudon, double bctz, double rn15impo,
                     double rc13impo, double ac13b, double rcaco3c13impo, double impofe, double o2, double aou, double *bioout,
                     src_out *O) {
  const double kw = P->kw, kc = P->kc, kc_c = P->kc_c, gamma1 = P->gamma1, redptn = P->redptn, redctn = P->redctn;
  const double redntp = P->redntp, diazntp = P->diazntp, diazptn = P->diazptn, dfr = P->dfr, pfr = P->pfr, dfrt = P->dfrt;
  const double geZ = P->geZ, rfeton = P->rfeton;
  double biopo4 = bioin[M_PO4], biophyt = bioin[M_PHYT], biophyt_phos = bioin[M_PHYT_PHOS], biozoop = bioin[M_ZOOP];
  double biodetr = bioin[M_DETR], biodetr_phos = bioin[M_DETR_PHOS];
  /* :1781-1784 ratios from the raw (unclipped) inputs */
  double ptn_P = biophyt_phos / biophyt;
  double ptn_detr = biodetr_phos / biodetr;
  double biodic = bioin[M_DIC], biodop = bioin[M_DOP], biono3 = bioin[M_NO3], biodon = bioin[M_DON], biodiaz = bioin[M_DIAZ];
  double biodin15 = bioin[M_DIN15], biodon15 = bioin[M_DON15], biophytn15 = bioin[M_PHYTN15], biozoopn15 = bioin[M_ZOOPN15];
  double biodetrn15 = bioin[M_DETRN15], biodiazn15 = bioin[M_DIAZN15], biodiatn15 = bioin[M_DIATN15];
  double biodic13 = bioin[M_DIC13], biophytc13 = bioin[M_PHYTC13], biozoopc13 = bioin[M_ZOOPC13], biodetrc13 = bioin[M_DETRC13];
  double biodoc13 = bioin[M_DOC13], biodiazc13 = bioin[M_DIAZC13], biodiatc13 = bioin[M_DIATC13], biocaco3c13 = bioin[M_CACO3C13];
  double biocaco3 = bioin[M_CACO3], biodiat = bioin[M_DIAT], biosil = bioin[M_SIL], bioopl = bioin[M_OPL];
  double biodfe = bioin[M_DFE], biodetrfe = bioin[M_DETRFE];

  /* flags (:1814-1890) */
  double po4flag = flag_of(biopo4), phytflag = flag_of(biophyt), zoopflag = flag_of(biozoop), detrflag = flag_of(biodetr);
  double phyt_phosflag = flag_of(biophyt_phos), detr_phosflag = flag_of(biodetr_phos);
  double sf_P_phosflag = 0.5 + sgn(0.5, ptn_P - gamma1 * redptn);
  double sf_detr_phosflag = 0.5 + sgn(0.5, ptn_detr - gamma1 * redptn);
  double dopflag = flag_of(biodop), no3flag = flag_of(biono3), donflag = flag_of(biodon), diazflag = flag_of(biodiaz);
  double din15flag = flag_of(biodin15), don15flag = flag_of(biodon15), phytn15flag = flag_of(biophytn15);
  double diatn15flag = flag_of(biodiatn15), zoopn15flag = flag_of(biozoopn15), detrn15flag = flag_of(biodetrn15);
  double diazn15flag = flag_of(biodiazn15), dic13flag = flag_of(biodic13), phytc13flag = flag_of(biophytc13);
  double diatc13flag = flag_of(biodiatc13), caco3c13flag = flag_of(biocaco3c13), zoopc13flag = flag_of(biozoopc13);
  double detrc13flag = flag_of(biodetrc13), doc13flag = flag_of(biodoc13), diazc13flag = flag_of(biodiazc13);
  double dfeflag = flag_of(biodfe), detrfeflag = flag_of(biodetrfe), caco3flag = flag_of(biocaco3);
  double diatflag = flag_of(biodiat), silflag = flag_of(biosil), oplflag = flag_of(bioopl);

  /* limit tracers to positive values (:1893-1926); the clip of bioin is visible to the caller */
  for (int m = 0; m < ORA_MOBI_NVAR; m++) bioin[m] = dmax(bioin[m], TRCMIN);
  biopo4 = dmax(biopo4, TRCMIN); biophyt = dmax(biophyt, TRCMIN); biozoop = dmax(biozoop, TRCMIN); biodetr = dmax(biodetr, TRCMIN);
  biophyt_phos = dmax(biophyt_phos, TRCMIN); biodetr_phos = dmax(biodetr_phos, TRCMIN); biodic = dmax(biodic, TRCMIN);
  biono3 = dmax(biono3, TRCMIN); biodop = dmax(biodop, TRCMIN); biodon = dmax(biodon, TRCMIN); biodiaz = dmax(biodiaz, TRCMIN);
  biodin15 = dmax(biodin15, TRCMIN); biodon15 = dmax(biodon15, TRCMIN); biophytn15 = dmax(biophytn15, TRCMIN);
  biodiatn15 = dmax(biodiatn15, TRCMIN); biozoopn15 = dmax(biozoopn15, TRCMIN); biodetrn15 = dmax(biodetrn15, TRCMIN);
  biodiazn15 = dmax(biodiazn15, TRCMIN); biodic13 = dmax(biodic13, TRCMIN); biophytc13 = dmax(biophytc13, TRCMIN);
  biodiatc13 = dmax(biodiatc13, TRCMIN); biocaco3c13 = dmax(biocaco3c13, TRCMIN); biozoopc13 = dmax(biozoopc13, TRCMIN);
  biodetrc13 = dmax(biodetrc13, TRCMIN); biodoc13 = dmax(biodoc13, TRCMIN); biodiazc13 = dmax(biodiazc13, TRCMIN);
  biocaco3 = dmax(biocaco3, TRCMIN); biodiat = dmax(biodiat, TRCMIN); biosil = dmax(biosil, TRCMIN); bioopl = dmax(bioopl, TRCMIN);
  biodfe = dmax(biodfe, TRCMIN); biodetrfe = dmax(biodetrfe, TRCMIN);

  /* iron-dependent light harvesting (:1928-1952) */
  double p1 = dmin(biophyt, P->pmax);
  double p2 = dmax(0.0, biophyt - P->pmax);
  double kfevar = (P->kfemin * p1 + P->kfemax * p2) / (p1 + p2);
  double deffe = biodfe / (kfevar + biodfe);
  double thetamax = P->thetamaxlo + (P->thetamaxhi - P->thetamaxlo) * deffe;
  double alpha_O = P->alphamin + (P->alphamax - P->alphamin) * deffe;
  double gl_O = gl * thetamax * alpha_O;
  p1 = dmin(biodiat, P->pmax_Diat);
  p2 = dmax(0.0, biodiat - P->pmax_Diat);
  double kfevar_Diat = (P->kfemin_Diat * p1 + P->kfemax_Diat * p2) / (p1 + p2);
  double deffe_Diat = biodfe / (kfevar_Diat + biodfe);
  double thetamax_Diat = P->thetamaxlo + (P->thetamaxhi - P->thetamaxlo) * deffe_Diat;
  double alpha_Diat = P->alphamin + (P->alphamax - P->alphamin) * deffe_Diat;
  double gl_Diat = gl * thetamax_Diat * alpha_Diat;
  double deffe_D = biodfe / (P->kfe_D + biodfe);
  double thetamax_D = P->thetamaxlo + (P->thetamaxhi - P->thetamaxlo) * deffe_D;
  double alpha_D = P->alphamin + (P->alphamax - P->alphamin) * deffe_D;
  double gl_D = gl * thetamax_D * alpha_D;

  /* photosynthesis after Evans & Parslow (:1954-2003) */
  double kirr = -kw - kc * (biophyt + biodiaz + biodiat) - kc_c * biocaco3;
  double f1 = exp(kirr * dzt);
  double jmax = P->abio_P * bct * deffe;
  double gd = jmax * dayfrac;
  double u1 = dmax(gl_O / gd, 1.e-6);
  double u2 = u1 * f1;
  double phi1 = log(u1 + sqrt(1. + u1 * u1)) - (sqrt(1. + u1 * u1) - 1.) / u1;
  double phi2 = log(u2 + sqrt(1. + u2 * u2)) - (sqrt(1. + u2 * u2) - 1.) / u2;
  double avej = gd * (phi1 - phi2) / (-kirr * dzt);
  double gmax = P->gbio * bctz;
  double jmax_D = dmax(0., P->abio_P * (bct - P->dbct_D) * deffe_D) * P->jdiar;
  double gd_D = dmax(1.e-14, jmax_D * dayfrac);
  u1 = dmax(gl_D / gd_D, 1.e-6);
  u2 = u1 * f1;
  phi1 = log(u1 + sqrt(1. + u1 * u1)) - (sqrt(1. + u1 * u1) - 1.) / u1;
  phi2 = log(u2 + sqrt(1. + u2 * u2)) - (sqrt(1. + u2 * u2) - 1.) / u2;
  double avej_D = gd_D * (phi1 - phi2) / (-kirr * dzt);
  double jmax_Diat = P->abiodiat * bct * deffe_Diat;
  double gd_Diat = jmax_Diat * dayfrac;
  u1 = dmax(gl_Diat / gd_Diat, 1.e-6);
  u2 = u1 * f1;
  phi1 = log(u1 + sqrt(1. + u1 * u1)) - (sqrt(1. + u1 * u1) - 1.) / u1;
  phi2 = log(u2 + sqrt(1. + u2 * u2)) - (sqrt(1. + u2 * u2) - 1.) / u2;
  double avej_Diat = gd_Diat * (phi1 - phi2) / (-kirr * dzt);

  double nupt = P->nupt0 * bct;
  double nupt_D = P->nupt0_D * bct;
  double nudt = P->nudt0 * bct;
  double nfixout = 0.0, expoout = 0.0, expo_phosout = 0.0, rn15expoout = 0.0, rc13expoout = 0.0, rcaco3c13expoout = 0.0;
  double calproout = 0.0, disslout = 0.0, expocaco3out = 0.0, expooplout = 0.0, expofeout = 0.0, remifeout = 0.0;

  for (int n = 1; n <= nbio; n++) {
    /* :2150-2166 */
    p1 = dmin(biophyt, P->pmax);
    p2 = dmax(0.0, biophyt - P->pmax);
    double k1n = (P->knmin * p1 + P->knmax * p2) / (p1 + p2);
    double k1p_P = k1n * ptn_P;
    kfevar = (P->kfemin * p1 + P->kfemax * p2) / (p1 + p2);
    deffe = biodfe / (kfevar + biodfe);
    jmax = P->abio_P * bct * deffe;
    p1 = dmin(biodiat, P->pmax_Diat);
    p2 = dmax(0.0, biodiat - P->pmax_Diat);
    kfevar_Diat = (P->kfemin_Diat * p1 + P->kfemax_Diat * p2) / (p1 + p2);
    double k1n_Diat = (P->knmin_Diat * p1 + P->knmax_Diat * p2) / (p1 + p2);
    double k1p_Diat = k1n_Diat * redptn;
    deffe_Diat = biodfe / (kfevar_Diat + biodfe);
    jmax_Diat = P->abiodiat * bct * deffe_Diat;
    deffe_D = biodfe / (P->kfe_D + biodfe);
    jmax_D = dmax(0., P->abio_P * (bct - P->dbct_D) * deffe_D) * P->jdiar;

    /* growth rates (:2168-2206) */
    double limP_dop = P->hdop * biodop / (k1p_P + biodop);
    double limP_po4 = biopo4 / (k1p_P + biopo4);
    double dopupt_flag = 0.5 + sgn(0.5, limP_dop - limP_po4);
    double limP = limP_dop * dopupt_flag + limP_po4 * (1. - dopupt_flag);
    double u_P = dmin(avej, jmax * limP);
    double k1si = 5.e-3;
    double limSi = biosil / (k1si + biosil);
    limP_dop = P->hdop * biodop / (k1p_Diat + biodop);
    limP_po4 = biopo4 / (k1p_Diat + biopo4);
    double dopupt_Diat_flag = 0.5 + sgn(0.5, limP_dop - limP_po4);
    double limP_Diat = limP_dop * dopupt_Diat_flag + limP_po4 * (1. - dopupt_Diat_flag);
    double u_Diat = dmin(avej_Diat, jmax_Diat * limSi);
    u_Diat = dmin(u_Diat, jmax_Diat * limP_Diat);
    u_P = dmin(u_P, jmax * biono3 / (k1n + biono3));
    u_Diat = dmin(u_Diat, jmax_Diat * biono3 / (k1n_Diat + biono3));
    double u_D = dmin(avej_D, jmax_D * limP);
    double dopupt_D_flag = dopupt_flag;
    /* grazing coefficients (:2208-2216) */
    double thetaZ = P->zprefP * biophyt + P->zprefDet * biodetr + P->zprefZ * biozoop + P->zprefDiaz * biodiaz + P->kzoo +
                    P->zprefDiat * biodiat;
    double ing_P = P->zprefP / thetaZ, ing_Det = P->zprefDet / thetaZ, ing_Z = P->zprefZ / thetaZ;
    double ing_D = P->zprefDiaz / thetaZ, ing_Diat = P->zprefDiat / thetaZ;
    double npp = u_P * biophyt;
    double npp_Diat = u_Diat * biodiat;
    double dopupt = npp * dopupt_flag;
    double dopupt_Diat = npp_Diat * dopupt_Diat_flag;
    double npp_D = dmax(0., u_D * biodiaz);
    double g_D = gmax * ing_D * biodiaz;
    double graz_D = g_D * biozoop;
    double morpt_D = nupt_D * biodiaz;
    double morp_D = P->nup_D * biodiaz * biodiaz;
    double no3upt_D = (0.5 + 0.5 * tanh(biono3 - 5.)) * npp_D;
    double dopupt_D = npp_D * dopupt_D_flag;
    double g_P = gmax * ing_P * biophyt;
    double graz = g_P * biozoop;
    double g_Z = gmax * ing_Z * biozoop;
    double graz_Z = g_Z * biozoop;
    double g_Det = gmax * ing_Det * biodetr;
    double graz_Det = g_Det * biozoop;
    double morp = P->nup * biophyt;
    double morpt = nupt * biophyt;
    double recy_don = nudon * bct * biodon;
    double recy_dop = nudop * bct * biodop;
    double morz = P->nuz * biozoop * biozoop;
    double remi = nud * bct * biodetr;
    double expo = wwd * biodetr;
    double expo_phos = wwd * biodetr_phos;
    double dissl = biocaco3 * dissk1;
    double expocaco3 = wwc * biocaco3;
    double g_Diat = gmax * ing_Diat * biodiat;
    double graz_Diat = g_Diat * biozoop;
    double morp_Diat = P->nu_diat * biodiat;
    double morpt_Diat = nudt * biodiat;
    double opldis = bioopl * opl_disk1;
    double expoopl = wwo * bioopl;
    double remife = nud * bct * biodetrfe;
    /* iron scavenging (:2262-2283) */
    double o2flag = tanh(dmax(o2, 0.));
    double ligand = dmax(pow(dmax(aou, 40.), 0.8) / 66. + pow(biodon, 0.8) / 4.8, 0.5) / 1000.;
    double fepa = (1.0 + P->kfeleq * (ligand - biodfe)) * o2flag;
    double feprime = ((-fepa + pow(fepa * fepa + 4.0 * P->kfeleq * biodfe, 0.5)) / (2.0 * P->kfeleq)) * o2flag;   /* :2217-2219 */
    double feorgads = (P->kfeorg * (pow(((biodetr * detrflag) * P->mc * redctn), 0.58)) * feprime) * o2flag;
    double fecol = P->kfecol * (feprime * feprime) * o2flag;
    double expofe = wwd * biodetrfe;
    /* apply the flags (:2284-2334) */
    graz = graz * phytflag * phyt_phosflag * sf_P_phosflag * phytn15flag;
    graz_Z = graz_Z * zoopflag * zoopn15flag;
    graz_Det = graz_Det * detrflag * detr_phosflag * sf_detr_phosflag * detrn15flag;
    morp = morp * phytflag * phyt_phosflag * phytn15flag;
    morpt = morpt * phytflag * phyt_phosflag * phytn15flag;
    morz = morz * zoopflag * zoopn15flag;
    remi = remi * detrflag * detr_phosflag * detrn15flag;
    expo = expo * detrflag * detrn15flag;
    expo_phos = expo_phos * detr_phosflag;
    recy_dop = recy_dop * dopflag;
    npp = npp * no3flag * (dopupt_flag * dopflag + (1. - dopupt_flag) * po4flag) * din15flag;
    npp_Diat = npp_Diat * no3flag * (dopupt_Diat_flag * dopflag + (1. - dopupt_Diat_flag) * po4flag) * din15flag;
    npp_D = npp_D * (dopupt_D_flag * dopflag + (1. - dopupt_D_flag) * po4flag) * din15flag;
    graz_D = graz_D * diazflag * diazn15flag;
    morpt_D = morpt_D * diazflag * diazn15flag;
    morp_D = morp_D * diazflag * diazn15flag;
    no3upt_D = no3upt_D * no3flag * din15flag;
    recy_don = recy_don * donflag * don15flag;
    dissl = dissl * caco3flag;
    expocaco3 = expocaco3 * caco3flag;
    graz_Diat = graz_Diat * diatflag;
    morp_Diat = morp_Diat * diatflag;
    morpt_Diat = morpt_Diat * diatflag;
    remife = remife * detrfeflag;
    feorgads = feorgads * dfeflag;
    expofe = expofe * detrfeflag;
    fecol = fecol * dfeflag;
    /* digestion, excretion, sloppy feeding (:2335-2440) */
    double dig_P = gamma1 * graz, dig_Z = gamma1 * graz_Z, dig_Det = gamma1 * graz_Det, dig_Diat = gamma1 * graz_Diat;
    double dig = dig_Z + dig_P + dig_Det + dig_Diat;
    double excr_P = gamma1 * (1 - geZ) * graz, excr_Z = gamma1 * (1 - geZ) * graz_Z, excr_Det = gamma1 * (1 - geZ) * graz_Det;
    double excr_Diat = gamma1 * (1 - geZ) * graz_Diat;
    double excr = excr_Z + excr_P + excr_Det + excr_Diat;
    double sf_P = (1. - gamma1) * graz, sf_Z = (1. - gamma1) * graz_Z, sf_Det = (1. - gamma1) * graz_Det;
    double sf_Diat = (1. - gamma1) * graz_Diat;
    double sf = sf_P + sf_Z + sf_Det + sf_Diat;
    double sf_P_phos = (graz * ptn_P - dig_P * redptn);
    double sf_Det_phos = (graz_Det * ptn_detr - dig_Det * redptn);
    double nr_excr_P = 0.0, nr_excr_detr = 0.0;
    double sf_phos = sf_P_phos + sf_Z * redptn + sf_Det_phos + sf_Diat * redptn;
    double dig_D = gamma1 * graz_D * (redntp / diazntp);
    dig = dig + dig_D;
    double excr_D = gamma1 * (1 - geZ) * graz_D * (redntp / diazntp);
    excr = excr + excr_D;
    double nr_excr_D = gamma1 * graz_D * (1 - (redntp / diazntp)) + (1 - gamma1) * graz_D * (1 - (redntp / diazntp));
    double sf_D = (1 - gamma1) * graz_D * (redntp / diazntp);
    sf = sf + sf_D;
    sf_phos = sf_phos + sf_D * redptn;
    /* isotope parameters (:2441-2530) */
    double uno3 = npp * dtbio / biono3;
    uno3 = dmin(uno3, 0.999);
    uno3 = dmax(uno3, TRCMIN);
    double rno3 = biodin15 / (biono3 - biodin15);
    rno3 = dmin(rno3, 2 * RN15STD);
    rno3 = dmax(rno3, RN15STD / 2.);
    double bassim = rno3 + P->eps_assim * (1 - uno3) / uno3 * log(1 - uno3) * rno3 / 1000.;
    double fcassim = bassim / (1 + bassim);
    double udon = recy_don * dtbio / biodon;
    udon = dmin(udon, 0.999);
    udon = dmax(udon, TRCMIN);
    double rdon = biodon15 / (biodon - biodon15);
    rdon = dmin(rdon, 2 * RN15STD);
    rdon = dmax(rdon, RN15STD / 2.);
    double brecy = rdon + P->eps_recy * (1 - udon) / udon * log(1 - udon) * rdon / 1000.;
    double fcrecy = brecy / (1 + brecy);
    double rzoop = biozoopn15 / (biozoop - biozoopn15);
    rzoop = dmin(rzoop, 2. * RN15STD);
    rzoop = dmax(rzoop, RN15STD / 2.);
    double bexcr = rzoop - P->eps_excr * rzoop / 1000.;
    double fcexcr = bexcr / (1 + bexcr);
    double bnfix = RN15STD - P->eps_nfix * RN15STD / 1000.;
    double fcnfix = bnfix / (1 + bnfix);
#define CLAMP15(x) (dmax(dmin((x), 2. * RN15STD / (1 + RN15STD)), RN15STD / (1 + RN15STD) / 2.))
#define CLAMP13(x) (dmax(dmin((x), 2. * RC13STD / (1 + RC13STD)), 0.5 * RC13STD / (1 + RC13STD)))
    double rtdin15 = CLAMP15(biodin15 / biono3);
    double rtdon15 = CLAMP15(biodon15 / biodon);
    double rtphytn15 = CLAMP15(biophytn15 / biophyt);
    double rtdiatn15 = CLAMP15(biodiatn15 / biodiat);
    double rtzoopn15 = CLAMP15(biozoopn15 / biozoop);
    double rtdetrn15 = CLAMP15(biodetrn15 / biodetr);
    double rtdiazn15 = CLAMP15(biodiazn15 / biodiaz);
    double rdic13 = biodic13 / (biodic - biodic13);
    rdic13 = dmin(rdic13, 2. * RC13STD);
    rdic13 = dmax(rdic13, 0.5 * RC13STD);
    double bc13npp = ac13b * rdic13;
    double fcnpp = bc13npp / (1 + bc13npp);
    double rtdic13 = CLAMP13(biodic13 / biodic);
    double rtphytc13 = CLAMP13(biophytc13 / (biophyt * redctn));
    double rtdiatc13 = CLAMP13(biodiatc13 / (biodiat * redctn));
    double rtcaco3c13 = CLAMP13(biocaco3c13 / biocaco3);
    double rtzoopc13 = CLAMP13(biozoopc13 / (biozoop * redctn));
    double rtdetrc13 = CLAMP13(biodetrc13 / (biodetr * redctn));
    double rtdoc13 = CLAMP13(biodoc13 / (biodon * redctn));
    double rtdiazc13 = CLAMP13(biodiazc13 / (biodiaz * redctn));
    (void)rtdin15; (void)rtdon15;

    /* CaCO3 and opal production (:2532-2548) */
    double calpro = ((sf_Z + morz) * capr + (sf_P + morp) * capr) * redctn * 1.e3;
    double negcoeff = -0.46204044117647, VTP = 1.60266544117647, tanh_m = 6.9, tanh_b = -3.673092;
    double sipr0 = (negcoeff * tanh(tanh_m * biodfe * 1.e3 + tanh_b) + VTP);
    double oplpro = (morp_Diat + sf_Diat) * sipr0 * silflag * (1.e-3);
    opldis = opldis * oplflag;
    expoopl = expoopl * oplflag;
    double GM15ptc = 0.0060 + 0.0069 * biopo4;
    double GM15ptn = GM15ptc * redctn * 1.e3;

    /* prognostic equations (:2552-2680) */
    biopo4 = biopo4 + dtbio * (dopupt * ptn_P - GM15ptn * npp + (1. - dfrt) * morpt * ptn_P + (1. - pfr) * remi * ptn_detr +
                               diazptn * (morpt_D - (npp_D - dopupt_D)) + recy_dop +
                               redptn * (excr + (1. - dfrt) * morpt_Diat - (npp_Diat - dopupt_Diat)));
    biodop = biodop + dtbio * (dfr * morp * ptn_P + redptn * (dfr * morp_Diat + dfrt * morpt_Diat - dopupt_Diat) +
                               dfrt * morpt * ptn_P + pfr * remi * ptn_detr - ptn_P * dopupt - diazptn * dopupt_D - recy_dop);
    biophyt = biophyt + dtbio * (npp - morp - graz - morpt);
    biophyt_phos = biophyt_phos + dtbio * (npp * GM15ptn - morp * ptn_P - graz * ptn_P - morpt * ptn_P);
    biozoop = biozoop + dtbio * (dig - morz - graz_Z - excr);
    biodetr = biodetr + dtbio * ((1. - dfr) * morp + sf + morz - remi - graz_Det - expo + impo + morp_D * (redntp / diazntp) +
                                 (1. - dfr) * morp_Diat);
    biodetr_phos = biodetr_phos + dtbio * ((1. - dfr) * morp * ptn_P + sf_phos + morz * redptn - remi * ptn_detr - graz_Det * ptn_detr -
                                           expo_phos + impo_phos + morp_D * (redntp / diazntp) * redptn + (1. - dfr) * morp_Diat * redptn);
    biodic = biodic + dtbio * redctn *
                          (excr + (1. - pfr) * remi + (1. - dfrt) * morpt - npp + (1. - dfrt) * morpt_Diat - npp_Diat + morpt_D - npp_D +
                           recy_don + nr_excr_D + nr_excr_P + nr_excr_detr + morp_D * (1. - (redntp / diazntp)));
    biono3 = biono3 + dtbio * (excr + (1. - pfr) * remi + (1. - dfrt) * morpt - npp + (1. - dfrt) * morpt_Diat - npp_Diat + morpt_D -
                               no3upt_D + recy_don + nr_excr_D + nr_excr_P + nr_excr_detr + morp_D * (1. - (redntp / diazntp)));
    biodon = biodon + dtbio * (dfr * morp + dfrt * morpt + pfr * remi - recy_don + dfr * morp_Diat + dfrt * morpt_Diat);
    biodiaz = biodiaz + dtbio * (npp_D - morp_D - morpt_D - graz_D);
    ptn_P = biophyt_phos / biophyt;
    ptn_detr = biodetr_phos / biodetr;
    biocaco3 = biocaco3 + dtbio * (calpro - dissl - expocaco3 + impocaco3);
    biodiat = biodiat + dtbio * (npp_Diat - morp_Diat - graz_Diat - morpt_Diat);
    biosil = biosil + dtbio * (opldis - oplpro);
    bioopl = bioopl + dtbio * (oplpro - opldis - expoopl + impoopl);
    biodfe = biodfe + dtbio * (rfeton * (excr + (1. - dfrt) * morpt - npp + morpt_D - npp_D + recy_don + nr_excr_D + nr_excr_P +
                                         nr_excr_detr + morp_D * (1. - (redntp / diazntp))) -
                               feorgads + remife - fecol + rfeton * ((1. - dfrt) * morpt_Diat - npp_Diat));
    biodetrfe = biodetrfe + dtbio * (rfeton * (sf + (1. - dfr) * morp + morp_D * (redntp / diazntp) + morz - graz_Det) + feorgads +
                                     P->iscr * fecol - remife - expofe + impofe + rfeton * (1. - dfr) * morp_Diat);
    /* isotope equations (:2682-2760) */
    biodin15 = biodin15 + dtbio * (rtphytn15 * (1. - dfrt) * morpt + rtphytn15 * nr_excr_P + rtdiatn15 * (1. - dfrt) * morpt_Diat -
                                   fcassim * npp_Diat + fcexcr * excr + rtdiazn15 * morpt_D + rtdiazn15 * nr_excr_D +
                                   rtdiazn15 * morp_D * (1. - (redntp / diazntp)) + rtdetrn15 * (1. - pfr) * remi +
                                   rtdetrn15 * nr_excr_detr + fcrecy * recy_don - fcassim * npp - fcassim * no3upt_D);
    biodon15 = biodon15 + dtbio * (dfr * rtphytn15 * morp + dfr * rtdiatn15 * morp_Diat + dfrt * rtdiatn15 * morpt_Diat +
                                   dfrt * rtphytn15 * morpt + rtdetrn15 * pfr * remi - fcrecy * recy_don);
    biophytn15 = biophytn15 + dtbio * (fcassim * npp - rtphytn15 * morp - rtphytn15 * graz - rtphytn15 * morpt);
    biodiatn15 = biodiatn15 + dtbio * (fcassim * npp_Diat - rtdiatn15 * morp_Diat - rtdiatn15 * graz_Diat - rtdiatn15 * morpt_Diat);
    biozoopn15 = biozoopn15 + dtbio * (rtphytn15 * dig_P + rtdiatn15 * dig_Diat + rtzoopn15 * dig_Z + rtdetrn15 * dig_Det +
                                       rtdiazn15 * dig_D - rtzoopn15 * morz - rtzoopn15 * graz_Z - fcexcr * excr);
    biodetrn15 = biodetrn15 + dtbio * (rtphytn15 * (1. - dfr) * morp + rtdiatn15 * (1. - dfr) * morp_Diat + rtdiatn15 * sf_Diat +
                                       rtphytn15 * sf_P + rtzoopn15 * sf_Z + rtdetrn15 * sf_Det + rtdiazn15 * sf_D + rtzoopn15 * morz -
                                       rtdetrn15 * remi - rtdetrn15 * graz_Det - rtdetrn15 * expo + rn15impo * impo +
                                       rtdiazn15 * morp_D * (redntp / diazntp));
    biodiazn15 = biodiazn15 + dtbio * (fcnfix * (npp_D - no3upt_D) + fcassim * no3upt_D - rtdiazn15 * morp_D - rtdiazn15 * graz_D -
                                       rtdiazn15 * morpt_D);
    biodic13 = biodic13 + dtbio * redctn *
                              (rtphytc13 * (1. - dfrt) * morpt + rtphytc13 * nr_excr_P + rtzoopc13 * excr + rtdiazc13 * morpt_D +
                               rtdiazc13 * nr_excr_D + rtdiazc13 * morp_D * (1 - (redntp / diazntp)) + rtdetrc13 * (1. - pfr) * remi +
                               rtdetrc13 * nr_excr_detr + rtdiatc13 * (1. - dfrt) * morpt_Diat - fcnpp * npp_Diat + rtdoc13 * recy_don -
                               fcnpp * npp - fcnpp * npp_D);
    biodoc13 = biodoc13 + dtbio * redctn *
                              (dfr * rtphytc13 * morp + rtdiatc13 * (dfr * morp_Diat + dfrt * morpt_Diat) + rtphytc13 * dfrt * morpt +
                               rtdetrc13 * pfr * remi - rtdoc13 * recy_don);
    biophytc13 = biophytc13 + dtbio * redctn * (fcnpp * npp - rtphytc13 * morp - rtphytc13 * graz - rtphytc13 * morpt);
    biozoopc13 = biozoopc13 + dtbio * redctn *
                                  (rtphytc13 * dig_P + rtdiatc13 * dig_Diat + rtzoopc13 * dig_Z + rtdetrc13 * dig_Det + rtdiazc13 * dig_D -
                                   rtzoopc13 * morz - rtzoopc13 * graz_Z - rtzoopc13 * excr);
    biodetrc13 = biodetrc13 + dtbio * redctn *
                                  (rtphytc13 * (1. - dfr) * morp + rtdiatc13 * (1. - dfr) * morp_Diat + rtdiatc13 * sf_Diat +
                                   rtphytc13 * sf_P + rtzoopc13 * sf_Z + rtdetrc13 * sf_Det + rtdiazc13 * sf_D + rtzoopc13 * morz -
                                   rtdetrc13 * remi - rtdetrc13 * graz_Det - rtdetrc13 * expo + rc13impo +
                                   rtdiazc13 * morp_D * (redntp / diazntp));
    biodiazc13 = biodiazc13 + dtbio * redctn * (fcnpp * npp_D - rtdiazc13 * (morp_D + graz_D + morpt_D));
    biocaco3c13 = biocaco3c13 + dtbio * (rtdic13 * calpro - rtcaco3c13 * dissl - rtcaco3c13 * expocaco3 + rcaco3c13impo);
    biodiatc13 = biodiatc13 + dtbio * redctn * (fcnpp * npp_Diat - rtdiatc13 * (morp_Diat + graz_Diat + morpt_Diat));
    /* accumulate outputs (:2762-2777) */
    expoout = expoout + expo;
    expo_phosout = expo_phosout + expo_phos;
    rn15expoout = rn15expoout + rtdetrn15;
    rc13expoout = rc13expoout + rtdetrc13 * expo;
    rcaco3c13expoout = rcaco3c13expoout + rtcaco3c13 * expocaco3;
    calproout = calproout + calpro;
    disslout = disslout + dissl;
    expocaco3out = expocaco3out + expocaco3;
    expooplout = expooplout + expoopl;
    nfixout = nfixout + npp_D - no3upt_D;
    expofeout = expofeout + expofe;
    remifeout = remifeout + remife;
    /* re-evaluate the flags that are still 1 (:3175-3251) */
#define REFLAG(f, x) if ((f) == 1) (f) = flag_of(x)
    REFLAG(po4flag, biopo4); REFLAG(phytflag, biophyt); REFLAG(zoopflag, biozoop); REFLAG(detrflag, biodetr);
    REFLAG(phyt_phosflag, biophyt_phos); REFLAG(detr_phosflag, biodetr_phos); REFLAG(no3flag, biono3); REFLAG(dopflag, biodop);
    REFLAG(donflag, biodon); REFLAG(diazflag, biodiaz); REFLAG(din15flag, biodin15); REFLAG(don15flag, biodon15);
    REFLAG(phytn15flag, biophytn15); REFLAG(diatn15flag, biodiatn15); REFLAG(zoopn15flag, biozoopn15);
    REFLAG(detrn15flag, biodetrn15); REFLAG(diazn15flag, biodiazn15); REFLAG(caco3flag, biocaco3); REFLAG(diatflag, biodiat);
    REFLAG(silflag, biosil); REFLAG(oplflag, bioopl); REFLAG(dfeflag, biodfe); REFLAG(detrfeflag, biodetrfe);
    REFLAG(dic13flag, biodic13); REFLAG(phytc13flag, biophytc13); REFLAG(diatc13flag, biodiatc13);
    REFLAG(caco3c13flag, biocaco3c13); REFLAG(zoopc13flag, biozoopc13); REFLAG(detrc13flag, biodetrc13);
    REFLAG(doc13flag, biodoc13); REFLAG(diazc13flag, biodiazc13);
  }
  (void)dic13flag; (void)phytc13flag; (void)diatc13flag; (void)caco3c13flag; (void)zoopc13flag; (void)detrc13flag;
  (void)doc13flag; (void)diazc13flag; (void)don15flag; (void)alpha_Diat;

  /* increments (:3254-3311) */
  bioout[M_PO4] = biopo4 - bioin[M_PO4]; bioout[M_PHYT] = biophyt - bioin[M_PHYT];
  bioout[M_PHYT_PHOS] = biophyt_phos - bioin[M_PHYT_PHOS]; bioout[M_ZOOP] = biozoop - bioin[M_ZOOP];
  bioout[M_DETR] = biodetr - bioin[M_DETR]; bioout[M_DETR_PHOS] = biodetr_phos - bioin[M_DETR_PHOS];
  bioout[M_DIC] = biodic - bioin[M_DIC]; bioout[M_DOP] = biodop - bioin[M_DOP]; bioout[M_NO3] = biono3 - bioin[M_NO3];
  bioout[M_DON] = biodon - bioin[M_DON]; bioout[M_DIAZ] = biodiaz - bioin[M_DIAZ]; bioout[M_DIN15] = biodin15 - bioin[M_DIN15];
  bioout[M_DON15] = biodon15 - bioin[M_DON15]; bioout[M_PHYTN15] = biophytn15 - bioin[M_PHYTN15];
  bioout[M_ZOOPN15] = biozoopn15 - bioin[M_ZOOPN15]; bioout[M_DETRN15] = biodetrn15 - bioin[M_DETRN15];
  bioout[M_DIAZN15] = biodiazn15 - bioin[M_DIAZN15]; bioout[M_DIATN15] = biodiatn15 - bioin[M_DIATN15];
  bioout[M_CACO3] = biocaco3 - bioin[M_CACO3]; bioout[M_DIAT] = biodiat - bioin[M_DIAT]; bioout[M_SIL] = biosil - bioin[M_SIL];
  bioout[M_OPL] = bioopl - bioin[M_OPL]; bioout[M_DFE] = biodfe - bioin[M_DFE]; bioout[M_DETRFE] = biodetrfe - bioin[M_DETRFE];
  bioout[M_DIC13] = biodic13 - bioin[M_DIC13]; bioout[M_PHYTC13] = biophytc13 - bioin[M_PHYTC13];
  bioout[M_ZOOPC13] = biozoopc13 - bioin[M_ZOOPC13]; bioout[M_DETRC13] = biodetrc13 - bioin[M_DETRC13];
  bioout[M_DOC13] = biodoc13 - bioin[M_DOC13]; bioout[M_DIAZC13] = biodiazc13 - bioin[M_DIAZC13];
  bioout[M_DIATC13] = biodiatc13 - bioin[M_DIATC13]; bioout[M_CACO3C13] = biocaco3c13 - bioin[M_CACO3C13];

  O->nfix = nfixout; O->expo = expoout; O->expo_phos = expo_phosout; O->calpro = calproout; O->dissl = disslout;
  O->expocaco3 = expocaco3out; O->expoopl = expooplout; O->rn15expo = rn15expoout; O->rc13expo = rc13expoout;
  O->rcaco3c13expo = rcaco3c13expoout; O->expofe = expofeout; O->remife = remifeout;
}

/* 09/mom/mobi.F:519-1483.  tnpzd(km,ntnpzd) column-major [m*km + (k-1)]; src(km,nsrc) [(s-1)*km + (k-1)] */
static void mobi_driver(const ora_ctx *c, int kmx, double twodt, double rctheta, double dayfrac, double swr, double *tnpzd,
                        const double *t_in, const double *o2_in, const double *aou_in, const double *s_in, const double *dic_in,
                        const double *alk_in, double co2_in, const double *sgb_in, double *src, int nbio, double dtbio, double rdtts,
                        double rnbio) {
  const ora_mobi_par *P = c->mobi;
  const int km = c->km, nsrc = c->nsrc;
  /* option subsets (O_carbon_13 / O_carbon_14 / O_mobi_nitrogen_15 off): a source slot of 0 means "absent"; its source
   * goes to the scratch slot nsrc + 1 of the column buffer (the caller sizes it), which nothing reads */
  int32_t ixl[ORA_MOBI_NIDX];
  for (int q = 0; q < ORA_MOBI_NIDX; q++) ixl[q] = c->mobi_idx[q];
  for (int q = MI_SRC; q < MI_SRC + ORA_MOBI_NVAR; q++) if (ixl[q] == 0) ixl[q] = nsrc + 1;
  for (int q = MI_ISALK; q < MI_N; q++) if (ixl[q] == 0) ixl[q] = nsrc + 1;
  const int32_t *ix = ixl;
  const double redctn = P->redctn;
#define TN(k, m) tnpzd[(size_t)(m)*km + ((k)-1)]
#define SRC(k, s) src[(size_t)((s)-1) * km + ((k)-1)]
#define ISM(m) ix[MI_SRC + (m)]
  double snpzd[ORA_MOBI_NVAR];
  double dic_npzd_sms[km + 1], nfix[km + 1], bdeni[km + 1], rtdic13[km + 1], rtcaco3c13[km + 1];
  double rcalpro[km + 1], rdissl[km + 1], rexpocaco3[km + 1], rexpoopl[km + 1];
  double expo = 0.0, impo = 0.0, expo_phos = 0.0, impo_phos = 0.0, phin = 0.0, prca = 0.0, sedrr = 0.0;
  double rn15impo = 0.0, rn15expo = 0.0, rc13impo = 0.0, rc13expo = 0.0, prca13 = 0.0, rcaco3c13impo = 0.0, rcaco3c13expo = 0.0;
  double expofe = 0.0, impofe = 0.0, calpro = 0.0, caco3in = 0.0, dissl = 0.0, impocaco3 = 0.0, expocaco3 = 0.0, dissk1 = 0.0;
  double expoopl = 0.0, impoopl = 0.0;
  memset(src, 0, sizeof(double) * (size_t)km * (nsrc + 1));
  memset(snpzd, 0, sizeof snpzd);
  for (int k = 0; k <= km; k++) { rcalpro[k] = rdissl[k] = rexpocaco3[k] = rexpoopl[k] = 0.0; nfix[k] = bdeni[k] = 0.0; }

  /* 1111 main k-loop (:763-1289) */
  for (int k = 1; k <= kmx; k++) {
    double pH, co2star, dco2star, pCO2, dpco2, CO3, Omega_c, Omega_a;
    rn15impo = rn15expo;
    double atmpres = 1.0;
    double depth = c->zt[k - 1] / 100.;
    ora_co2calc_SWS(t_in[k - 1], s_in[k - 1], dic_in[k - 1], alk_in[k - 1], co2_in, atmpres, depth, &pH, &co2star, &dco2star, &pCO2,
                    &dpco2, &CO3, &Omega_c, &Omega_a);
    double ac13_DIC_aq = -1.0512994e-4 * t_in[k - 1] + 1.011765;
    double ac13_aq_POC = -0.017 * log10(dmin(dmax(co2star * 1000., 2.), 74.)) + 1.0034;
    double ac13b = ac13_aq_POC / ac13_DIC_aq;
    rc13impo = rc13expo * c->dztr[k - 1];
    rcaco3c13impo = rcaco3c13expo * c->dztr[k - 1];
    dissk1 = P->dissk0 * dmax(0., (1. - Omega_c));
    double capr = P->caprmax * dmax(0., (Omega_c - 1.) / (P->kcapr + Omega_c - 1.));
    double opl_disk1 = P->opl_disk0;
    swr = swr * exp(-P->kc * phin - P->kc_c * caco3in);
    phin = dmax(TN(k, M_PHYT), TRCMIN) * c->dzt[k - 1] + dmax(TN(k, M_DIAZ), TRCMIN) * c->dzt[k - 1] +
           dmax(TN(k, M_DIAT), TRCMIN) * c->dzt[k - 1];
    caco3in = caco3in + TN(k, M_CACO3) * c->dzt[k - 1];
    impocaco3 = expocaco3 * c->dztr[k - 1];
    double gl = swr * exp(P->ztt[k - 1] * rctheta);
    impo = expo * c->dztr[k - 1];
    impo_phos = expo_phos * c->dztr[k - 1];
    impofe = expofe * c->dztr[k - 1];
    double bct = pow(P->bbio, (P->cbio * t_in[k - 1]));
    impoopl = expoopl * c->dztr[k - 1];
    double bctz = (0.5 * (tanh(o2_in[k - 1] - 8.) + 1)) * pow(P->bbio, (P->cbio * t_in[k - 1]));
    double nud = P->nud0 * (0.6 + 0.4 * tanh(0.22 * dmax(o2_in[k - 1], 0.)));
    double nudon = P->nudon0, nudop = P->nudop0;

    /* the section actual argument tnpzd(k,:) is copied in and back (:853) */
    double bioin[ORA_MOBI_NVAR];
    for (int m = 0; m < ORA_MOBI_NVAR; m++) bioin[m] = TN(k, m);
    src_out so;
    mobi_src(P, nbio, dtbio, capr, bioin, gl, bct, impo, c->dzt[k - 1], impo_phos, dayfrac, P->wd[k - 1], nud, impocaco3, P->wc[k - 1],
             dissk1, impoopl, P->wo[k - 1], opl_disk1, nudop, nudon, bctz, rn15impo, rc13impo, ac13b, rcaco3c13impo, impofe,
             o2_in[k - 1], aou_in[k - 1], snpzd, &so);
    for (int m = 0; m < ORA_MOBI_NVAR; m++) TN(k, m) = bioin[m];
    nfix[k] = so.nfix; expo = so.expo; expo_phos = so.expo_phos; calpro = so.calpro; dissl = so.dissl; expocaco3 = so.expocaco3;
    expoopl = so.expoopl; rn15expo = so.rn15expo; rc13expo = so.rc13expo; rcaco3c13expo = so.rcaco3c13expo; expofe = so.expofe;

    /* source/sink terms (:880-895) */
    for (int m = 0; m < ORA_MOBI_NVAR; m++) snpzd[m] = snpzd[m] * rdtts;
    expofe = expofe * rnbio;
    expocaco3 = expocaco3 * rnbio;
    expoopl = expoopl * rnbio;
    rexpoopl[k] = expoopl;
    expo = expo * rnbio;
    expo_phos = expo_phos * rnbio;
    rn15expo = rn15expo * rnbio;
    rc13expo = rc13expo * rnbio;
    rcaco3c13expo = rcaco3c13expo * rnbio;
    rcalpro[k] = calpro * rnbio;
    rdissl[k] = dissl * rnbio;
    rexpocaco3[k] = expocaco3;

    sedrr = sgb_in[k - 1] * expo * c->dzt[k - 1];
    /* benthic denitrification, Bohlen et al. 2012 (:1035-1075) */
    double no3flag = 0.5 + sgn(0.5, TN(k, M_NO3) - TRCMIN);
    double din15flag = 0.5 + sgn(0.5, TN(k, M_DIN15) - TRCMIN);
    double lno3 = 0.5 * tanh(TN(k, M_NO3) * 10 - 5.0);
    double sg_bdeni = (0.06 + 0.19 * pow(0.99, (dmax(o2_in[k - 1], TRCMIN) - dmax(TN(k, M_NO3), TRCMIN)))) *
                      dmax(expo * sgb_in[k - 1], TRCMIN) * redctn * 1.e3;
    sg_bdeni = dmin(sg_bdeni, sgb_in[k - 1] * expo);
    sg_bdeni = dmax(sg_bdeni, 0.);
    sg_bdeni = sg_bdeni * (0.5 + lno3) * no3flag * din15flag;
    bdeni[k] = sg_bdeni;
    snpzd[M_NO3] = snpzd[M_NO3] + sgb_in[k - 1] * expo - sg_bdeni;
    double rno3 = dmax(TN(k, M_DIN15), TRCMIN * RN15STD / (1 + RN15STD)) /
                  dmax(TN(k, M_NO3) - TN(k, M_DIN15), TRCMIN * RN15STD / (1 + RN15STD));
    rno3 = dmin(rno3, 2. * RN15STD);
    rno3 = dmax(rno3, RN15STD / 2.);
    double eps_bdeni = P->eps_bdeni0 * exp(-2.5e-6 * (c->zt[k - 1]));
    double bbdeni = rno3 - eps_bdeni * rno3 / 1000.;
    snpzd[M_DIN15] = snpzd[M_DIN15] + rn15expo * sgb_in[k - 1] * expo - bbdeni / (1 + bbdeni) * sg_bdeni;
    /* sediment carbon oxidation and iron release (:1076-1110) */
    double coxdepth = dmin(dmax(c->zt[k - 1], 50000.), 150000.);
    double oblinc = -1.26e-6 * coxdepth + 0.203;
    double obexpc = -6.e-7 * coxdepth + 1.14;
    double nburial = (oblinc * pow((expo * sgb_in[k - 1] * c->dzt[k - 1] / 100 * 86400. * 365. * redctn * 1000.), obexpc)) /
                     (86400. * 365. * c->dzt[k - 1] / 100 * redctn * 1000.);
    double coxsed = expo * sgb_in[k - 1] - nburial;
    double fesedmax = 85.;
    double fesed = fesedmax * tanh(coxsed * redctn * 1000 * c->dzt[k - 1] / 100 * 86400. / o2_in[k - 1]) /
                   (c->dzt[k - 1] / 100 * 86400 * 1000);
    snpzd[M_DFE] = snpzd[M_DFE] + fesed;
    snpzd[M_PO4] = snpzd[M_PO4] + sgb_in[k - 1] * expo_phos;
    snpzd[M_DIC] = snpzd[M_DIC] + sgb_in[k - 1] * expo * redctn;
    snpzd[M_DIC13] = snpzd[M_DIC13] + rc13expo * sgb_in[k - 1] * redctn;
    rc13expo = rc13expo - sgb_in[k - 1] * rc13expo;
    expo = expo - sgb_in[k - 1] * expo;
    expo_phos = expo_phos - sgb_in[k - 1] * expo_phos;

    /* set source/sink terms (:1149-1204) */
    for (int m = 0; m < ORA_MOBI_NVAR; m++) SRC(k, ISM(m)) = snpzd[m];

    dic_npzd_sms[k] = snpzd[M_DIC];
    double dprca = rcalpro[k] * 1e-3;
    prca = prca + dprca * c->dzt[k - 1];
    rtdic13[k] = dmax(TN(k, M_DIC13), TRCMIN * RC13STD / (1 + RC13STD)) / dmax(dic_in[k - 1], TRCMIN);
    rtdic13[k] = dmin(rtdic13[k], 2. * RC13STD / (1 + RC13STD));
    rtdic13[k] = dmax(rtdic13[k], 0.5 * RC13STD / (1 + RC13STD));
    prca13 = prca13 + dprca * c->dzt[k - 1] * rtdic13[k];
    rtcaco3c13[k] = dmax(TN(k, M_CACO3C13), TRCMIN * RC13STD / (1 + RC13STD)) / dmax(TN(k, M_CACO3), TRCMIN);
    rtcaco3c13[k] = dmin(rtcaco3c13[k], 2. * RC13STD / (1 + RC13STD));
    rtcaco3c13[k] = dmax(rtcaco3c13[k], 0.5 * RC13STD / (1 + RC13STD));
    SRC(k, ix[MI_ISALK]) = -snpzd[M_DIC] * P->redntc * 1.e-3;
    /* total export -> import for the next layer (:1280-1288) */
    expo = expo * c->dzt[k - 1];
    expo_phos = expo_phos * c->dzt[k - 1];
    rc13expo = rc13expo * c->dzt[k - 1];
    rcaco3c13expo = rcaco3c13expo * c->dzt[k - 1];
    expofe = expofe * c->dzt[k - 1];
    expocaco3 = expocaco3 * c->dzt[k - 1];
    expoopl = expoopl * c->dzt[k - 1];
  }
  (void)sedrr; (void)prca; (void)prca13;

  /* 2222 second k-loop: O2, water-column denitrification, ALK (:1301-1366) */
  for (int k = 1; k <= kmx; k++) {
    double fo2 = tanh(0.22 * dmax(o2_in[k - 1], 0.));
    double so2 = dic_npzd_sms[k] * P->redotc + nfix[k] * rnbio * 1.25e-3;
    double no3flag = 0.5 + sgn(0.5, TN(k, M_NO3) - TRCMIN);
    double din15flag = 0.5 + sgn(0.5, TN(k, M_DIN15) - TRCMIN);
    double lno3 = 0.5 * tanh(TN(k, M_NO3) - 2.5);
    double lntp = 0.5 * tanh(TN(k, M_NO3) / (P->redntp * TN(k, M_PO4)) * 100. - 60.);
    (void)lntp;
    double wcdeni = 800. * no3flag * so2 * (1.0 - fo2) * (0.5 + lno3) * din15flag;
    wcdeni = dmax(wcdeni, 0.);
    SRC(k, ISM(M_NO3)) = SRC(k, ISM(M_NO3)) - wcdeni;
    double uno3 = wcdeni * twodt / TN(k, M_NO3);
    uno3 = dmin(uno3, 0.999);
    uno3 = dmax(uno3, TRCMIN);
    double rno3 = dmax(TN(k, M_DIN15), TRCMIN * RN15STD / (1 + RN15STD)) /
                  dmax(TN(k, M_NO3) - TN(k, M_DIN15), TRCMIN * RN15STD / (1 + RN15STD));
    rno3 = dmin(rno3, 2. * RN15STD);
    rno3 = dmax(rno3, RN15STD / 2.);
    double bwcdeni = rno3 + P->eps_wcdeni * (1 - uno3) / uno3 * log(1 - uno3) * rno3 / 1000.;
    SRC(k, ISM(M_DIN15)) = SRC(k, ISM(M_DIN15)) - (bwcdeni / (1 + bwcdeni)) * wcdeni;
    SRC(k, ix[MI_ISALK]) = SRC(k, ix[MI_ISALK]) + wcdeni * 1.e-3;
    SRC(k, ix[MI_ISALK]) = SRC(k, ix[MI_ISALK]) + bdeni[k] * 1.e-3;
    SRC(k, ix[MI_ISALK]) = SRC(k, ix[MI_ISALK]) - nfix[k] * rnbio * 1.e-3;
    SRC(k, ix[MI_ISO2]) = -so2 * fo2;
  }

  /* 3333 third k-loop: calcite dissolution / production (:1372-1400) */
  for (int k = 1; k <= kmx - 1; k++) {
    SRC(k, ISM(M_DIC)) = SRC(k, ISM(M_DIC)) + rdissl[k] * 1.e-3 - rcalpro[k] * 1.e-3;
    SRC(k, ISM(M_DIC13)) = SRC(k, ISM(M_DIC13)) + rdissl[k] * 1.e-3 * rtcaco3c13[k] - rcalpro[k] * 1.e-3 * rtdic13[k];
    SRC(k, ix[MI_ISALK]) = SRC(k, ix[MI_ISALK]) + 2. * rdissl[k] * 1.e-3 - 2. * rcalpro[k] * 1.e-3;
  }
  SRC(kmx, ISM(M_DIC)) = SRC(kmx, ISM(M_DIC)) + rdissl[kmx] * 1.e-3 - rcalpro[kmx] * 1.e-3 + rexpocaco3[kmx] * 1.e-3;
  SRC(kmx, ISM(M_DIC13)) = SRC(kmx, ISM(M_DIC13)) + rdissl[kmx] * 1.e-3 * rtcaco3c13[kmx] - rcalpro[kmx] * 1.e-3 * rtdic13[kmx] +
                           rexpocaco3[kmx] * 1.e-3 * rtcaco3c13[kmx];
  SRC(kmx, ix[MI_ISALK]) = SRC(kmx, ix[MI_ISALK]) + 2. * rdissl[kmx] * 1.e-3 - 2. * rcalpro[kmx] * 1.e-3 + 2. * rexpocaco3[kmx] * 1.e-3;
  /* put the opal leftovers back into the ocean (:1478) */
  SRC(kmx, ISM(M_SIL)) = SRC(kmx, ISM(M_SIL)) + rexpoopl[kmx];
#undef TN
#undef SRC
#undef ISM
}

/* Test entry point (tests/test_cpu_refpin.py): one call of mobi_src with explicit arguments, so that the restatement can
 * be driven with the same random cells as the translated reference routine.  in[24] = gl, bct, impo, dzt, impo_phos,
 * dayfrac, wwd, nud, impocaco3, wwc, dissk1, impoopl, wwo, opl_disk1, nudop, nudon, bctz, rn15impo, rc13impo, ac13b,
 * rcaco3c13impo, impofe, o2, aou; out[12] = nfix, expo, expo_phos, calpro, dissl, expocaco3, expoopl, rn15expo, rc13expo,
 * rcaco3c13expo, expofe, remife. */
void ora_test_mobi_src(ora_ctx *c, int nbio, double dtbio, double capr, double *bioin, const double *in, double *bioout, double *out) {
  src_out so;
  mobi_src(c->mobi, nbio, dtbio, capr, bioin, in[0], in[1], in[2], in[3], in[4], in[5], in[6], in[7], in[8], in[9], in[10], in[11],
           in[12], in[13], in[14], in[15], in[16], in[17], in[18], in[19], in[20], in[21], in[22], in[23], bioout, &so);
  out[0] = so.nfix; out[1] = so.expo; out[2] = so.expo_phos; out[3] = so.calpro; out[4] = so.dissl; out[5] = so.expocaco3;
  out[6] = so.expoopl; out[7] = so.rn15expo; out[8] = so.rc13expo; out[9] = so.rcaco3c13expo; out[10] = so.expofe;
  out[11] = so.remife;
}

/* 09/mom/tracer.F:310-545 (column prologue + mobi_driver) and :848-867 (c14 source) */
void ora_mobi_columns(ora_ctx *c) {
  const int imt = c->imt, km = c->km, jmt = c->jmt, nsrc = c->nsrc;
  const int js = 2, je = jmt - 1, is = 2, ie = imt - 1;
  const ora_mobi_par *P = c->mobi;
  const int32_t *ix = c->mobi_idx;
  const double pi = atan(1.0) * 4.0;
  const double radian = 360. / (2. * pi);
#define T(i, k, j, n, l) c->t[IT(i, k, j, n, l)]
  /* month index for the monthly deposition input (:310-336) */
  double yrtime = fmod(c->relyr, 1.);
  int mi = 12;
  for (int m = 1; m <= 12; m++)
    if (yrtime <= m / 12.) { mi = m; break; }
  double declin = sin((fmod(c->relyr, 1.) - 0.22) * 2. * pi) * 0.4;
  int nbio = (int)(c->c2dtts / P->dtnpzd);
  double dtbio = c->c2dtts / nbio;
  double rdtts = 1. / c->c2dtts;
  double rnbio = 1. / nbio;

  double tnpzd[ORA_MOBI_NVAR * km], t_in[km], o2_in[km], aou_in[km], s_in[km], dic_in[km], alk_in[km], sgb_in[km];
  double srccol[(size_t)km * (nsrc + 1)];   /* + the scratch slot of an option subset */
  for (int j = js; j <= je; j++) {
    int jrow = j;
    for (int i = is; i <= ie; i++) {
      if (c->kmt[I2(i, jrow)] > 0) {
        double ai = c->aice[I2(i, jrow)], hi = c->hice[I2(i, jrow)], hs = c->hsno[I2(i, jrow)];
        /* day fraction and incoming solar (:370-390) */
        double rctheta = dmax(-1.5, dmin(1.5, c->tlat[I2(i, jrow)] / radian - declin));
        double cr = cos(rctheta);
        rctheta = P->kw / sqrt(1. - (1. - cr * cr) / (1.33 * 1.33));
        double dayfrac = dmin(1., -tan(c->tlat[I2(i, jrow)] / radian) * tan(declin));
        dayfrac = dmax(1e-12, acos(dmax(-1., dayfrac)) / pi);
        double swr = P->tap * c->dnswr[I2(i, jrow)] * 1e-3 * (1. + ai * (exp(-P->ki * (hi + hs)) - 1.));
        /* gather (:393-503) */
        /* an absent isotope tracer (index 0) reads as 1: its equations are computed and discarded, and no equation of a
         * present variable reads an isotope variable */
        for (int m = 0; m < ORA_MOBI_NVAR; m++)
          for (int k = 1; k <= km; k++) tnpzd[(size_t)m * km + (k - 1)] = ix[MI_TR + m] ? T(i, k, j, ix[MI_TR + m], TAUM1) : 1.0;
        for (int k = 1; k <= km; k++) {
          t_in[k - 1] = T(i, k, j, ix[MI_ITEMP], TAUM1);
          o2_in[k - 1] = T(i, k, j, ix[MI_IO2], TAUM1) * 1000.;
          s_in[k - 1] = 1.e3 * T(i, k, j, ix[MI_ISALT], TAUM1) + 35.0;
        }
        for (int k = 1; k <= c->kmt[I2(i, jrow)]; k++) {
          double f1 = log((298.15 - t_in[k - 1]) / (273.15 + t_in[k - 1]));
          double f2 = f1 * f1, f3 = f2 * f1, f4 = f3 * f1, f5 = f4 * f1;
          double o2sat = exp(2.00907 + 3.22014 * f1 + 4.05010 * f2 + 4.94457 * f3 - 2.56847E-1 * f4 + 3.88767 * f5 +
                             s_in[k - 1] * (-6.24523e-3 - 7.37614e-3 * f1 - 1.03410e-2 * f2 - 8.17083E-3 * f3) -
                             4.88682E-7 * s_in[k - 1] * s_in[k - 1]);
          o2sat = o2sat / 22391.6 * 1000.0 * 1000.;
          aou_in[k - 1] = o2sat - o2_in[k - 1];
        }
        for (int k = 1; k <= km; k++) {
          dic_in[k - 1] = T(i, k, j, ix[MI_TR + M_DIC], TAUM1);
          alk_in[k - 1] = T(i, k, j, ix[MI_IALK], TAUM1);
          sgb_in[k - 1] = c->sg_bathy[IJK(i, j, k)];
        }
        mobi_driver(c, c->kmt[I2(i, jrow)], c->c2dtts, rctheta, dayfrac, swr, tnpzd, t_in, o2_in, aou_in, s_in, dic_in, alk_in,
                    c->co2ccn, sgb_in, srccol, nbio, dtbio, rdtts, rnbio);
        /* dust and hydrothermal iron (:536-545) */
        int isdfe = ix[MI_SRC + M_DFE];
        srccol[(size_t)(isdfe - 1) * km + 0] =
            srccol[(size_t)(isdfe - 1) * km + 0] + c->fe_atmdep[I2(i, j) + (size_t)imt * jmt * (mi - 1)] * 1000 / (c->dzt[0] / 100.);
        for (int k = 1; k <= c->kmt[I2(i, j)]; k++)
          srccol[(size_t)(isdfe - 1) * km + (k - 1)] = srccol[(size_t)(isdfe - 1) * km + (k - 1)] + c->fe_hydr[IJK(i, j, k)];
        for (int s = 1; s <= nsrc; s++)
          for (int k = 1; k <= km; k++) c->src[IS(i, k, j, s)] = srccol[(size_t)(s - 1) * km + (k - 1)];
      }
    }
  }
  /* source for c14 (:848-867) */
  if (ix[MI_IC14] == 0) return;   /* O_carbon_14 off */
  for (int j = js; j <= je; j++) {
    int jrow = j;
    for (int i = is; i <= ie; i++)
      if (c->kmt[I2(i, jrow)] > 0)
        for (int k = 1; k <= c->kmt[I2(i, jrow)]; k++)
          c->src[IS(i, k, j, ix[MI_ISC14])] =
              c->src[IS(i, k, j, ix[MI_SRC + M_DIC])] * RC14STD - 3.836e-12 * T(i, k, j, ix[MI_IC14], TAUM1);
  }
#undef T
}
