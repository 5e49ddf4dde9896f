"""MOBI parameters and index maps handed to the device library.

Restates the host-side set-up of 09/mom/mobi.F `mobi_init` (:40-438): namelist defaults
(:58-183), the `&mobi` overrides of run/control.in (:40-52), the conversion to model units
(:191-245), the sinking speeds per level (:247-262), and the grazing-preference
renormalisation (:264-281, including its quirk: the sum counts zprefDiaz twice, adds the
never-initialised zprefC (= 0, zero COMMON) and omits zprefDiat).  The result is a flat
float64 block whose order is shared with oracle/ora_mobi.h and csrc/mobi_par.h.
"""
from __future__ import annotations

import numpy as np

DAYLEN = 86400.0
KMAX = 128

# MOBI-internal state order, fixed by the setimobi sequence (09/mom/mobi.F:440-497)
MOBI_STATE = [
    "po4", "phyt", "phyt_phos", "zoop", "detr", "detr_phos", "dic", "dic13", "phytc13", "zoopc13", "detrc13", "doc13",
    "diazc13", "diatc13", "caco3c13", "dop", "no3", "don", "diaz", "din15", "don15", "phytn15", "zoopn15", "detrn15",
    "diazn15", "diatn15", "caco3", "diat", "sil", "opl", "dfe", "detrfe",
]
# source slots in the order tracer_init assigns them (09/common/UVic_ESCM.F:1375-1483)
SOURCE_ORDER = [
    "c14", "dic", "dic13", "alk", "o2", "po4", "phyt", "phyt_phos", "zoop", "detr", "detr_phos", "dfe", "detrfe", "caco3",
    "diat", "sil", "opl", "dop", "no3", "don", "diaz", "din15", "don15", "phytn15", "diatn15", "zoopn15", "detrn15",
    "diazn15", "phytc13", "diatc13", "caco3c13", "zoopc13", "detrc13", "doc13", "diazc13",
]

PAR_ORDER = (
    "kw kc ki tap abio_P bbio cbio nup nup_D nupt0 nupt0_D gamma1 gbio nuz nud0 nudon0 nudop0 "
    "dtnpzd redctn redptn redotn redotc redntp redntc diazptn diazntp caprmax kcapr dissk0 kc_c "
    "jdiar dbct_D kzoo geZ dfr pfr dfrt hdop abiodiat nu_diat nudt0 opl_disk0 "
    "zprefP zprefDiat zprefDiaz zprefZ zprefDet "
    "eps_assim eps_excr eps_nfix eps_wcdeni eps_bdeni0 eps_recy "
    "kfemin kfemax knmin knmax pmax kfe_D kfemin_Diat kfemax_Diat knmin_Diat knmax_Diat pmax_Diat "
    "kfeleq thetamaxhi thetamaxlo alphamax alphamin mc kfeorg rfeton iscr kfecol"
).split()
N_SCALAR = len(PAR_ORDER) + 6       # + reserved[6]
N_PAR = N_SCALAR + 4 * KMAX

# &mobi of run/control.in:40-52
CONTROL_IN = dict(
    abio_P=0.4, kfemin=0.04e-3, kfemax=0.4e-3, abiodiat=0.7, kfemin_Diat=0.08e-3, kfemax_Diat=0.8e-3, nup=0.03,
    nu_diat=0.03, nuz=0.06, nupt0=0.01, nudt0=0.01, kfe_D=0.08e-3, wo0=50., sipr0=0.15, sildustfluxfac=0., zprefP=0.29,
    zprefDiat=0.24, zprefDet=0.14, zprefZ=0.29, zprefDiaz=0.04, nupt0_D=0.0032, nup_D=0.36, gbio=0.4, geZ=0.56, jdiar=0.08,
    dbct_D=0., sgbdfac=2., nud0=0.07, wd0=18, mwz=100000., mw=0.05, dfr=0.055, dfrt=0., hdop=0.5, nudop0=2.e-5,
    nudon0=1.e-5, eps_assim=6., pfr=0.02, eps_excr=4., eps_recy=1.3, eps_wcdeni=25., eps_bdeni0=4., eps_nfix=1.,
    caprmax=0.035, kcapr=0.4, dissk0=0.013, redctn=7, diazntp=32., redotn=11.0, dtnpzd=27000.,
)


def mobi_defaults(dtts):
    """namelist defaults, 09/mom/mobi.F:58-183"""
    return dict(
        alpha=0.16, kw=0.04, kc=0.047, ki=5.0, abio_P=0.6, bbio=1.066, cbio=1.0, nup=0.03, nup_D=0.0001, nupt0=0.015,
        nupt0_D=0.001, gamma1=0.70, gbio=0.38, epsbio=1.6, nuz=0.06, nud0=0.07, nudon0=2.33e-5, nudop0=7.e-5, wd0=16.0,
        mwz=100000., mw=0.02, mw_c=0.06, par=0.43, dtnpzd=dtts / 4., redctn=7.1, redptn=1. / 16., caprmax=0.022, kcapr=0.4,
        dcaco3=650000.0, redotn=10.6, jdiar=0.08, dbct_D=2.6, kzoo=0.15, geZ=0.6, sgbdfac=1.0, diazntp=28., dfr=0.08,
        pfr=0.03, dfrt=0.01, hdop=0.4, abiodiat=3.45, nu_diat=0.03, nudt0=0.015, wo0=50., opl_disk0=8.3e-3, kc_c=0.047,
        wc0=35., kcal=100., dissk0=0.013, zprefP=0.18, zprefDiat=0.18, zprefDiaz=0.1, zprefZ=0.18, zprefDet=0.18,
        eps_assim=6., eps_excr=4., eps_nfix=1., eps_wcdeni=25., eps_bdeni0=6., eps_recy=1., kfemin=0.04e-3, kfemax=0.2e-3,
        knmin=0.15, knmax=1.5, pmax=0.15, kfe_D=0.1e-3, kfemin_Diat=0.04e-3, kfemax_Diat=0.8e-3, knmin_Diat=0.3,
        knmax_Diat=3.0, pmax_Diat=0.15, kfeleq=10. ** 5.5, lig=1.0e-3, thetamaxhi=0.04, thetamaxlo=0.01,
        alphamax=73.6e-6 * 86400, alphamin=18.4e-6 * 86400, mc=12.011, fetopsed=0.076, o2min=5.,
        kfeorg=2.8 * (1. / 86400.), rfeton=5.5e-6 * 7, iscr=0.6, kfecol=900 / 86400.,
    )


def mobi_par_block(case, overrides=None):
    """Flat parameter block after mobi_init's unit conversion (09/mom/mobi.F:191-262)."""
    a = case.arrays
    dtts = case.scalars["dtts"]
    p = mobi_defaults(dtts)
    p.update(CONTROL_IN)
    if dtts != 108000.0:
        # synthetic fine grids run a shorter tracer step; keep nbio = c2dtts/dtnpzd = 8
        p["dtnpzd"] = dtts / 4.0
    if overrides:
        p.update(overrides)
    # unit conversions (:191-245)
    p["redctn"] = p["redctn"] * 1.e-3
    p["redotn"] = p["redotn"] * 1.e-3
    p["redotp"] = p["redotn"] / p["redptn"]
    p["redctp"] = p["redctn"] / p["redptn"]
    p["redotc"] = p["redotn"] / p["redctn"]
    p["redntp"] = 1. / p["redptn"]
    p["redntc"] = 1. / p["redctn"]
    p["diazptn"] = 1. / p["diazntp"]
    p["wc0"] = p["wc0"] * 1.e2
    p["kc_c"] = p["kc_c"] * 1.e-2
    p["dissk0"] = p["dissk0"] / DAYLEN
    for k in ("abiodiat", "nu_diat", "nudt0", "opl_disk0", "abio_P", "nup", "nup_D", "nupt0", "nupt0_D", "gbio", "epsbio",
              "nuz", "nud0", "nudop0", "nudon0", "alphamax", "alphamin"):
        p[k] = p[k] / DAYLEN
    p["wo0"] = p["wo0"] * 1.e2
    p["kw"] = p["kw"] * 1.e-2
    p["kc"] = p["kc"] * 1.e-2
    p["ki"] = p["ki"] * 1.e-2
    p["wd0"] = p["wd0"] * 1.e2
    p["tap"] = 2. * p["par"]
    # grazing preferences (:264-281) -- zprefC is never set (0), zprefDiaz counted twice
    zprefC = 0.0
    sumz = p["zprefP"] + p["zprefDet"] + p["zprefZ"] + p["zprefDiaz"] + zprefC + p["zprefDiaz"]
    if sumz != 1.:
        for k in ("zprefP", "zprefZ", "zprefDet", "zprefDiaz", "zprefDiat"):
            p[k] = p[k] / sumz
    km = case.km
    zt, zw, dzt = a["zt"], a["zw"], a["dzt"]
    assert km <= KMAX
    blk = np.zeros(N_PAR)
    for n, name in enumerate(PAR_ORDER):
        blk[n] = p[name]
    wd, wc, wo, ztt = (np.zeros(KMAX) for _ in range(4))
    for k in range(km):
        if zt[k] < p["mwz"]:
            wd[k] = (p["wd0"] + p["mw"] * zt[k]) / DAYLEN / dzt[k]
            wc[k] = (p["wc0"] + p["mw_c"] * zt[k]) / DAYLEN / dzt[k]
        else:
            wd[k] = (p["wd0"] + p["mw"] * p["mwz"]) / DAYLEN / dzt[k]
            wc[k] = (p["wc0"] + p["mw_c"] * p["mwz"]) / DAYLEN / dzt[k]
        wo[k] = p["wo0"] / DAYLEN / dzt[k]
    ztt[0] = 0.0
    ztt[1:km] = (-1) * zw[0:km - 1]
    o = N_SCALAR
    blk[o:o + KMAX] = wd
    blk[o + KMAX:o + 2 * KMAX] = wc
    blk[o + 2 * KMAX:o + 3 * KMAX] = wo
    blk[o + 3 * KMAX:o + 4 * KMAX] = ztt
    return blk


MI_TR, MI_SRC, MI_ITEMP, MI_ISALT, MI_IALK, MI_IO2, MI_IC14, MI_ISALK, MI_ISO2, MI_ISC14, MI_N = 0, 32, 64, 65, 66, 67, 68, 69, 70, 71, 72
N_IDX = 128


# the option groups of run/mk.in that add MOBI state variables (09/mom/mobi.h:104-142, 09/common/size.h:31-144)
ISOTOPE_GROUPS = {
    "O_carbon_13": ["dic13", "phytc13", "zoopc13", "detrc13", "doc13", "diazc13", "diatc13", "caco3c13"],
    "O_carbon_14": ["c14"],
    "O_mobi_nitrogen_15": ["din15", "don15", "phytn15", "zoopn15", "detrn15", "diazn15", "diatn15"],
}
OPTIONAL = {nm for g in ISOTOPE_GROUPS.values() for nm in g}


def tracer_names_for(options=("O_carbon_13", "O_carbon_14", "O_mobi_nitrogen_15")):
    """tracer order of tracer_init with the given isotope options on: all three -> the 37 tracers of the
    shipped run/mk.in (BASELINE config 3); none -> the 21 tracers of "full MOBI, no isotopes" (config 2)"""
    from .synthetic import MOBI_TRACERS_37
    off = {nm for opt, g in ISOTOPE_GROUPS.items() if opt not in options for nm in g}
    return [nm for nm in MOBI_TRACERS_37 if nm not in off]


def mobi_index_maps(tracer_names):
    """(itrc(nt), mobi_idx(128), nsrc): tracer -> source slot, and the MOBI gather/scatter maps
    (09/mom/tracer.F:393-447, 09/mom/mobi.F:1149-1204).  Isotope tracers that are not in `tracer_names` (options
    O_carbon_13 / O_carbon_14 / O_mobi_nitrogen_15 off) get index 0 = absent; an option group is all or nothing."""
    pos = {nm: n + 1 for n, nm in enumerate(tracer_names)}
    missing = [s for s in MOBI_STATE + ["temp", "salt", "alk", "o2", "c14"] if s not in pos and s not in OPTIONAL]
    if missing:
        raise ValueError(f"MOBI needs tracers {missing}")
    for opt, g in ISOTOPE_GROUPS.items():
        have = [nm in pos for nm in g]
        if any(have) and not all(have):
            raise ValueError(f"{opt}: its tracers {g} must be all present or all absent")
    order = [nm for nm in SOURCE_ORDER if nm in pos]
    slot = {nm: s + 1 for s, nm in enumerate(order)}
    itrc = np.zeros(len(tracer_names), dtype=np.int32)
    for nm, s in slot.items():
        itrc[pos[nm] - 1] = s
    idx = np.zeros(N_IDX, dtype=np.int32)
    for m, nm in enumerate(MOBI_STATE):
        idx[MI_TR + m] = pos.get(nm, 0)
        idx[MI_SRC + m] = slot.get(nm, 0)
    idx[MI_ITEMP], idx[MI_ISALT], idx[MI_IALK], idx[MI_IO2], idx[MI_IC14] = pos["temp"], pos["salt"], pos["alk"], pos["o2"], pos.get("c14", 0)
    idx[MI_ISALK], idx[MI_ISO2], idx[MI_ISC14] = slot["alk"], slot["o2"], slot.get("c14", 0)
    return itrc, idx, len(order)
