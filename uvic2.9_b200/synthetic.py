"""Synthetic grids, bathymetry, forcing and initial tracer fields for the tracer step.

The reference's input data set (data.100.100.19/*.nc, run/mk.in:196) is not in the
repository, so every test and benchmark runs on seeded synthetic inputs of the named grid
shapes.  The derived metric arrays follow the formulas of
source/common/grids.F:470-577 exactly (same operations, numpy float64), and are handed
unchanged to both the CUDA library and the CPU oracle, so neither recomputes them.

All arrays use the reference's Fortran layouts, expressed as C-ordered numpy arrays with
the dimensions reversed: a Fortran ``t(imt,km,jmt,nt,-1:1)`` is ``t[3, nt, jmt, km, imt]``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SEED = 2901

# tracer names in the order tracer_init assigns them with the shipped run/mk.in
# (09/common/UVic_ESCM.F:1282-1362)
MOBI_TRACERS_37 = [
    "temp", "salt", "dic", "dic13", "c14", "alk", "o2", "po4", "phyt", "phyt_phos", "zoop", "detr",
    "detr_phos", "caco3", "diat", "sil", "opl", "dop", "no3", "don", "diaz", "din15", "don15",
    "phytn15", "diatn15", "zoopn15", "detrn15", "diazn15", "dfe", "detrfe", "phytc13", "diatc13",
    "caco3c13", "zoopc13", "detrc13", "doc13", "diazc13",
]

# typical magnitudes (model units: C, alk, O2, Si in mol m-3; N, P, Fe, plankton in mmol m-3;
# 09/mom/mobi.F:207-208)
_TYPICAL = {
    "dic": 2.2, "alk": 2.4, "o2": 0.2, "po4": 1.5, "phyt": 0.2, "phyt_phos": 0.0125, "zoop": 0.1,
    "detr": 0.05, "detr_phos": 0.003, "caco3": 0.02, "diat": 0.1, "sil": 0.05, "opl": 0.005,
    "dop": 0.1, "no3": 15.0, "don": 3.0, "diaz": 0.01, "dfe": 5e-4, "detrfe": 1e-5,
}
_RN15STD = 0.0036765   # 09/mom/mobi.h
_RC13STD = 0.0112372
_RC14STD = 1.176e-12


@dataclass
class Case:
    """One synthetic configuration: dims, scalars and every array the step needs."""

    imt: int
    jmt: int
    km: int
    nt: int
    nsrc: int
    scalars: dict = field(default_factory=dict)
    arrays: dict = field(default_factory=dict)
    tracer_names: list = field(default_factory=list)
    has_mobi: bool = False

    def __getitem__(self, k):
        return self.arrays[k]


def _rng(name: str, seed: int) -> np.random.Generator:
    # one independent substream per field name
    h = np.frombuffer(name.encode(), dtype=np.uint8).astype(np.uint64)
    key = int((h * np.arange(1, len(h) + 1, dtype=np.uint64)).sum() % (2**31))
    return np.random.default_rng(np.random.SeedSequence([seed, key]))


def _smooth2d(rng, imt, jmt, lam, phi, nmodes=6, kmax=4):
    """Smooth field in [-1,1], exactly cyclic in i (columns 1 == imt-1, imt == 2)."""
    f = np.zeros((jmt, imt))
    for _ in range(nmodes):
        kx = int(rng.integers(0, kmax + 1))
        ky = float(rng.uniform(0.5, kmax))
        a = float(rng.uniform(0.3, 1.0))
        p1, p2 = rng.uniform(0, 2 * math.pi, 2)
        f += a * np.cos(kx * np.deg2rad(lam)[None, :] + p1) * np.cos(ky * np.deg2rad(phi)[:, None] * 2 + p2)
    m = np.abs(f).max()
    return f / (m if m > 0 else 1.0)


def make_grid(imt, jmt, km, arrays, scalars):
    """Uniform lat-lon grid, stretched levels; derived metrics as in grids.F:470-577."""
    c0, c1, c2, p5, p25 = 0.0, 1.0, 2.0, 0.5, 0.25
    radius = 6370.0e5                      # 09/common/UVic_ESCM.F:1647
    pi = math.atan(1.0) * 4.0
    radian = 360.0 / (c2 * pi)             # grids.F:415
    degtcm = radius / radian

    dlam = 360.0 / (imt - 2)
    dphi = 178.0 / jmt                     # rows span (-89, 89): cos(phi) stays positive
    xt = (np.arange(1, imt + 1) - 1.5) * dlam
    xu = xt + 0.5 * dlam
    yt = -89.0 + (np.arange(1, jmt + 1) - 0.5) * dphi
    yu = yt + 0.5 * dphi
    dxtdeg = np.full(imt, dlam)
    dxudeg = np.full(imt, dlam)
    dytdeg = np.full(jmt, dphi)
    dyudeg = np.full(jmt, dphi)

    # levels: 50 m at the surface stretched to 500 m at depth (cm units)
    dzt = np.array([50.0e2 + (500.0e2 - 50.0e2) * (k / max(km - 1, 1)) ** 1.5 for k in range(km)])
    zw = np.cumsum(dzt)
    zt = zw - p5 * dzt
    dzw = np.zeros(km + 1)
    dzw[1:km] = zt[1:] - zt[:-1]
    dzw[0] = zt[0]                         # grids.F:147-148
    dzw[km] = zw[km - 1] - zt[km - 1]

    dxt = dxtdeg * degtcm
    dxu = dxudeg * degtcm
    dyt = dytdeg * degtcm
    dyu = dyudeg * degtcm
    dxt[0], dxt[-1] = dxt[-2], dxt[1]
    dxu[0], dxu[-1] = dxu[-2], dxu[1]

    a = arrays
    a["dxt"], a["dxu"], a["dyt"], a["dyu"] = dxt, dxu, dyt, dyu
    a["dzt"], a["dzw"], a["zt"], a["zw"] = dzt, dzw, zt, zw
    c2dzt = c2 * dzt
    a["dzt2r"] = c1 / c2dzt
    a["dzwr"] = c1 / dzw
    a["dztur"] = c1 / (dzw[0:km] * dzt)
    a["dztlr"] = c1 / (dzw[1:km + 1] * dzt)
    a["dztr"] = c1 / dzt
    a["dytr"] = c1 / dyt
    a["dyt2r"] = p5 / dyt
    a["dyt4r"] = p25 / dyt
    a["dyur"] = c1 / dyu
    phi = yu / radian
    phit = yt / radian
    cst = np.cos(phit)
    csu = np.cos(phi)
    a["cst"], a["csu"] = cst, csu
    a["cstr"] = c1 / cst
    a["csur"] = c1 / csu
    a["cstdytr"] = c1 / (cst * dyt)
    a["cstdyt2r"] = a["cstdytr"] * p5
    a["csu_dyur"] = csu / dyu
    a["dxtr"] = c1 / dxt
    a["dxt2r"] = p5 / dxt
    a["dxt4r"] = p25 / dxt
    a["dxur"] = c1 / dxu
    duw = (xu - xt) * degtcm
    due = np.empty(imt)
    due[:-1] = (xt[1:] - xu[:-1]) * degtcm
    due[-1] = due[1]
    dus = (yu - yt) * degtcm
    dun = np.empty(jmt)
    dun[:-1] = (yt[1:] - yu[:-1]) * degtcm
    dun[-1] = dun[-2]
    a["duw"], a["due"], a["dus"], a["dun"] = duw, due, dus, dun
    # tracer time step acceleration off (MOBI forces dtxcel == 1, 09/mom/setmom.F:969-974)
    dtxcel = np.ones(km)
    a["dtxcel"] = dtxcel
    a["dtxsqr"] = np.sqrt(dtxcel)
    dztxcl = dzt / dtxcel
    dzwxcl = np.zeros(km)
    dzwxcl[:-1] = c1 / (dztxcl[:-1] + dztxcl[1:])
    a["dztxcl"], a["dzwxcl"] = dztxcl, dzwxcl
    a["tlat"] = np.broadcast_to(yt[:, None], (jmt, imt)).copy()
    a["_xt"], a["_yt"], a["_yu"], a["_xu"] = xt, yt, yu, xu
    scalars["radian"] = radian
    scalars["pi"] = pi
    return a


def make_bathymetry(imt, jmt, km, arrays, seed, land_lat=72.0, land_frac=0.30):
    """kmt in {0} u [2,km], closed walls at rows 1 and jmt, cyclic in i; kmu per topog.F:145."""
    rng = _rng("kmt", seed)
    xt, yt = arrays["_xt"], arrays["_yt"]
    d = _smooth2d(rng, imt, jmt, xt, yt, nmodes=8, kmax=5)
    thr = np.quantile(d, land_frac)
    depthfrac = np.clip((d - thr) / (d.max() - thr + 1e-30), 0.0, 1.0)
    kmt = np.where(d > thr, np.clip(np.rint(2 + (km - 2) * np.sqrt(depthfrac)), 2, km), 0).astype(np.int32)
    kmt[np.abs(yt) > land_lat, :] = 0
    kmt[0, :] = 0
    kmt[-1, :] = 0
    kmt[:, 0] = kmt[:, -2]
    kmt[:, -1] = kmt[:, 1]
    kmu = np.zeros_like(kmt)
    kmu[:-1, :-1] = np.minimum(np.minimum(kmt[:-1, :-1], kmt[:-1, 1:]), np.minimum(kmt[1:, :-1], kmt[1:, 1:]))
    kmu[:, -1] = kmu[:, 1]
    kmu[:, 0] = kmu[:, -2]
    arrays["kmt"], arrays["kmu"] = kmt, kmu
    k = np.arange(1, km + 1)[None, :, None]
    arrays["tmask"] = (kmt[:, None, :] >= k).astype(np.float64)   # 09/mom/loadmw.F:60-77
    arrays["umask"] = (kmu[:, None, :] >= k).astype(np.float64)
    return arrays


def make_eos(km, arrays):
    """Stand-in for the eqstate polynomial fit (source/mom/denscoef.F): c(km,9), to, so."""
    zt_m = arrays["zt"] / 100.0
    to = 2.0 + 12.0 * np.exp(-zt_m / 900.0)
    so = (34.6 + 0.3 * (1 - np.exp(-zt_m / 1500.0)) - 35.0) / 1000.0
    c = np.zeros((9, km))          # Fortran c(km,9) -> C order [9][km]
    c[0] = -(0.7e-4 + 1.1e-5 * to + 2.5e-8 * zt_m)
    c[1] = 0.78 - 1.0e-3 * to
    c[2] = -(5.5e-6 - 4.0e-8 * to)
    c[3] = -2.5e-3
    c[4] = 0.1
    c[5] = 4.0e-8
    c[6] = 1.0e-2
    c[7] = 2.0e-5
    c[8] = 0.5
    arrays["eosc"], arrays["to"], arrays["so"] = c, to, so
    return arrays


def make_velocity(imt, jmt, km, arrays, seed, u0=8.0):
    """B-grid u,v (cm/s) with zero depth integral at every U point, so that adv_vbt from
    continuity (source/mom/adv_vel.F:97-131) closes at kmt to round-off."""
    rng = _rng("u", seed)
    xu, yu = arrays["_xu"], arrays["_yu"]
    kmu = arrays["kmu"]
    dzt, zt = arrays["dzt"], arrays["zt"]
    u = np.zeros((2, jmt, km, imt))
    for n in range(2):
        w = _smooth2d(rng, imt, jmt, xu, yu, nmodes=6, kmax=4) * np.cos(np.deg2rad(yu))[:, None]
        for kk in np.unique(kmu):
            if kk < 2:
                continue
            sel = kmu == kk
            h = np.cos(math.pi * zt[:kk] / zt[kk - 1] * (1.0 + 0.5 * n))
            g = h - (h * dzt[:kk]).sum() / dzt[:kk].sum()
            prof = np.zeros(km)
            prof[:kk] = g
            u[n] += (w * sel)[:, None, :] * prof[None, :, None]
        u[n] *= u0
    u *= arrays["umask"][None]
    u[..., 0] = u[..., -2]
    u[..., -1] = u[..., 1]
    arrays["u"] = u
    return arrays


def adv_vel_numpy(imt, jmt, km, a):
    """Tracer part of adv_vel (source/mom/adv_vel.F:60-131), used only to build synthetic
    inputs in numpy; the device kernel and the oracle each have their own restatement."""
    u = a["u"]
    dxu, dyu, csu = a["dxu"], a["dyu"], a["csu"]
    vnt = np.zeros((jmt, km, imt))
    vet = np.zeros((jmt, km, imt))
    vbt = np.zeros((jmt, km + 1, imt))
    vnt[:, :, 1:-1] = (u[1][:, :, 1:-1] * dxu[1:-1] + u[1][:, :, 0:-2] * dxu[0:-2]) * csu[:, None, None] * a["dxt2r"][1:-1]
    vnt[..., 0] = vnt[..., -2]
    vnt[..., -1] = vnt[..., 1]
    vet[1:, :, :] = (u[0][1:] * dyu[1:, None, None] + u[0][:-1] * dyu[:-1, None, None]) * a["dyt2r"][1:, None, None]
    div = np.zeros((jmt, km, imt))
    div[1:, :, 1:-1] = ((vet[1:, :, 1:-1] - vet[1:, :, 0:-2]) * a["dxtr"][1:-1]
                        + (vnt[1:, :, 1:-1] - vnt[:-1, :, 1:-1]) * a["dytr"][1:, None, None]) \
        * a["cstr"][1:, None, None] * a["dzt"][None, :, None]
    for k in range(1, km + 1):
        vbt[:, k, :] = div[:, k - 1, :] + vbt[:, k - 1, :]
    vbt[..., 0] = vbt[..., -2]
    vbt[..., -1] = vbt[..., 1]
    return vet, vnt, vbt


def make_tracers(imt, jmt, km, nt, names, arrays, seed, noise=1.0):
    rng = _rng("t", seed)
    xt, yt = arrays["_xt"], arrays["_yt"]
    zt_m = arrays["zt"] / 100.0
    tmask = arrays["tmask"]
    t = np.zeros((3, nt, jmt, km, imt))
    cosphi = np.cos(np.deg2rad(yt))[:, None, None]
    prof = np.exp(-zt_m / 800.0)[None, :, None]
    for n, name in enumerate(names):
        sm = _smooth2d(rng, imt, jmt, xt, yt)[:, None, :]
        sm2 = _smooth2d(rng, imt, jmt, xt, yt, kmax=8)[:, None, :]
        wn = rng.standard_normal((jmt, km, imt))
        if name == "temp":
            f = 25.0 * prof * (0.4 + 0.6 * cosphi) + 2.0 * sm * np.exp(-zt_m / 1000.0)[None, :, None] + 1.0 * cosphi \
                + noise * 0.15 * wn * np.exp(-zt_m / 1500.0)[None, :, None]
        elif name == "salt":
            f = (34.7 - 35.0) / 1000.0 + 5.0e-4 * sm * prof + 2.0e-4 * sm2 + noise * 2.0e-5 * wn
        else:
            f = None
        if f is None:
            base = name
            ratio = 1.0
            for suf, r in (("n15", _RN15STD), ("c13", _RC13STD)):
                if name.endswith(suf):
                    base, ratio = name[: -len(suf)], r
            if name == "din15":
                base, ratio = "no3", _RN15STD
            elif name == "dic13":
                base, ratio = "dic", _RC13STD
            elif name == "doc13":
                base, ratio = "don", _RC13STD * 7.0
            elif name == "c14":
                base, ratio = "dic", _RC14STD * 0.9
            mag = _TYPICAL.get(base, 1.0)
            # positive log-normal about the typical magnitude, smooth + small-scale part
            vert = 0.6 + 0.4 * (prof if base in ("phyt", "zoop", "diat", "diaz", "detr", "o2") else (1.0 - 0.7 * prof))
            f = mag * ratio * vert * np.exp(0.35 * sm + 0.15 * sm2 + noise * 0.05 * wn)
        if name.startswith("passive"):
            f = 1.0 + 0.5 * sm + 0.2 * sm2 + noise * 0.05 * wn
        if name == "alk" and "dic" in names:
            # keep the carbonate system realistic: alkalinity tracks DIC with a positive excess,
            # so calcite stays near saturation (Omega_c ~ 1-6) instead of wandering through the
            # singularity of the PIC:POC formula at Omega_c = 1 - kcapr (09/mom/mobi.F:786)
            f = t[0, names.index("dic")] + 0.22 + 0.04 * sm - 0.08 * (1.0 - prof) + noise * 0.003 * wn
        f = f * tmask
        f[..., 0] = f[..., -2]
        f[..., -1] = f[..., 1]
        t[0, n] = f                          # tau-1
        dn = 1.0 + 1.0e-3 * sm2
        g = f * dn
        g[..., 0] = g[..., -2]
        g[..., -1] = g[..., 1]
        t[1, n] = g                          # tau
    arrays["t"] = t
    return arrays


def default_tracer_names(nt):
    if nt == 2:
        return ["temp", "salt"]
    names = list(MOBI_TRACERS_37[: min(nt, 37)])
    names += [f"passive{m}" for m in range(nt - len(names))]
    return names


def indp(value, array):
    """index (1-based) of the array element nearest to value (source/common/util.F `indp`)."""
    return int(np.argmin(np.abs(np.asarray(array) - value))) + 1


def filter_rows(yt):
    """Fourier-filter rows as setcom derives them (source/common/setcom.F:37-40,75-85)."""
    return dict(jfrst=indp(-87.3, yt), jft0=indp(-67.5, yt), jft1=indp(-69.3, yt), jft2=indp(69.3, yt))


def make_case(imt=102, jmt=102, km=19, nt=2, seed=SEED, names=None, dtts=None, noise=1.0,
              with_static_mobi=True, land_lat=72.0, land_frac=0.30) -> Case:
    """Build one complete synthetic configuration."""
    names = list(names) if names is not None else default_tracer_names(nt)
    assert len(names) == nt
    arrays: dict = {}
    scalars: dict = {}
    make_grid(imt, jmt, km, arrays, scalars)
    make_bathymetry(imt, jmt, km, arrays, seed, land_lat=land_lat, land_frac=land_frac)
    make_eos(km, arrays)
    make_velocity(imt, jmt, km, arrays, seed)
    vet, vnt, vbt = adv_vel_numpy(imt, jmt, km, arrays)
    arrays["adv_vet"], arrays["adv_vnt"], arrays["adv_vbt"] = vet, vnt, vbt
    make_tracers(imt, jmt, km, nt, names, arrays, seed, noise=noise)

    xt, yt = arrays["_xt"], arrays["_yt"]
    rng = _rng("static", seed)
    # isopycnal structure function and anisotropic equatorial addition (09/mom/isopyc.F:228-262)
    arrays["fisop"] = 1.0 + 0.3 * np.broadcast_to(_smooth2d(rng, imt, jmt, xt, yt)[None], (km, jmt, imt)).copy()
    arrays["addisop"] = np.broadcast_to((5.0e6 * np.exp(-(yt / 7.0) ** 2))[:, None, None], (jmt, km, imt)).copy()
    # tidal energy dissipation (09/mom/tidal_kv.h): decays with height above the bottom
    kmt = arrays["kmt"]
    zw = arrays["zw"]
    depth = np.where(kmt > 0, zw[np.maximum(kmt, 1) - 1], 0.0)
    hab = np.maximum(depth[:, None, :] - arrays["zt"][None, :, None], 0.0)
    base = np.exp(-hab / 500.0e2) * arrays["tmask"]
    for nm, amp in (("edrm2", 1.0e-3), ("edrs2", 4.0e-4), ("edrk1", 3.0e-4), ("edro1", 2.0e-4)):
        arrays[nm] = amp * base * (1.0 + 0.5 * _smooth2d(rng, imt, jmt, xt, yt)[:, None, :])
    # surface / bottom fluxes zero by default (conservation runs)
    arrays["stf"] = np.zeros((nt, jmt, imt))
    arrays["btf"] = np.zeros((nt, jmt, imt))
    # horizontal regions for basin means (source/common/cregin.h; nhreg=3)
    msk = np.zeros((jmt, imt), dtype=np.int32)
    third = (imt - 2) / 3.0
    ii = np.arange(imt)
    reg = np.clip(((ii - 1) // third).astype(np.int32), 0, 2) + 1
    msk[:, :] = reg[None, :]
    msk[kmt == 0] = 0
    arrays["mskhr"] = msk
    if with_static_mobi:
        # MOBI static inputs (09/mom/mobi.F:296-419, 09/common/topog.F:255-289)
        sgb = np.zeros((km, jmt, imt))
        jj, iii = np.nonzero(kmt > 0)
        sgb[kmt[jj, iii] - 1, jj, iii] = 1.0        # sg_bathy(kmt)=1 closes the sinking fluxes
        arrays["sg_bathy"] = sgb
        arrays["fe_hydr"] = 1.0e-9 * sgb * (1.0 + _smooth2d(rng, imt, jmt, xt, yt)[None])
        arrays["fe_atmdep"] = 2.0e-12 * (1.0 + 0.8 * np.stack([_smooth2d(rng, imt, jmt, xt, yt) for _ in range(12)]))
        arrays["dnswr"] = np.broadcast_to((200.0e3 * np.cos(np.deg2rad(yt)))[:, None], (jmt, imt)).copy()
        arrays["aice"] = np.zeros((jmt, imt))
        arrays["hice"] = np.zeros((jmt, imt))
        arrays["hsno"] = np.zeros((jmt, imt))

    # time stepping: run/control.in (dtts=108000 s at 3.6 deg); finer synthetic grids scale dt
    # with the zonal spacing so the explicit lateral terms stay stable
    if dtts is None:
        dtts = 108000.0 * min(1.0, (360.0 / (imt - 2)) / 3.6)
    rho0 = 1.035
    grav = 980.6
    zetar = 1.0 / 500.0e2                       # 09/mom/setmom.F:80-82
    scalars.update(
        dtts=dtts, c2dtts=2.0 * dtts, aidif=0.5, kappa_h=0.35,     # run/control.in:3,8
        ahisop=1.2e7, athkdf=8.0e6, slmxr=1.0 / 0.01,                # 09/mom/isopyc.F:70-110
        diff_cet=0.0, diff_cnt=0.0,                                  # ah=ahbkg=0 with O_isopycmix
        zetar=zetar, ogamma=0.2 * (1.0 / rho0) * zetar, gravrho0r=grav * (1.0 / rho0),
        relyr=0.37, co2ccn=280.0,
        **filter_rows(arrays["_yt"]),
    )
    itrc = np.zeros(nt, dtype=np.int32)
    nsrc = 0
    for n, nm in enumerate(names):
        if nm not in ("temp", "salt") and not nm.startswith("passive"):
            nsrc += 1
            itrc[n] = nsrc
    arrays["itrc"] = itrc
    case = Case(imt=imt, jmt=jmt, km=km, nt=nt, nsrc=nsrc, scalars=scalars, arrays=arrays, tracer_names=names)
    from . import mobi_params as mp
    if all(s in names for s in mp.MOBI_STATE + ["alk", "o2", "c14"] if s not in mp.OPTIONAL):
        # full MOBI tracer set: source slots in tracer_init's order, gather/scatter maps, parameters
        itrc, idx, nsrc = mp.mobi_index_maps(names)
        arrays["itrc"], arrays["mobi_idx"] = itrc, idx
        case.nsrc = nsrc
        arrays["mobi_par"] = mp.mobi_par_block(case)
        case.has_mobi = True
    return case


def add_momentum(case: Case, seed=SEED, am=1.5e9, kappa_m=10.0, cdbot=1.3e-3, dtuv=None) -> Case:
    """Inputs of the baroclinic momentum step (09/mom/clinic.F) for a case made by make_case: u(tau-1), wind stress,
    the Coriolis / metric factors of setmom (09/mom/setmom.F:777-802), the grid reciprocals of grids.F:470-530, 1/depth at
    U points (09/mom/setmom.F:1108-1111) and the anisotropic viscosity coefficients in the shape hmixc leaves them
    (09/mom/hmixc.F:62-150: Large et al. 2001 in the tropics above 550 m, `am` elsewhere).  Synthetic inputs only: the
    oracle and the device read these arrays, neither derives them."""
    a = case.arrays
    imt, jmt, km = case.imt, case.jmt, case.km
    rng = _rng("momentum", seed)
    radius = 6370.0e5
    pi = case.scalars["pi"]
    radian = case.scalars["radian"]
    omega = pi / 43082.0                                        # 09/common/UVic_ESCM.F: earth's rotation rate
    xu, yu, xt, yt = a["_xu"], a["_yu"], a["_xt"], a["_yt"]
    phi = yu / radian
    sine, csu = np.sin(phi), a["csu"]
    tng = sine / csu
    cor1 = np.broadcast_to((2.0 * omega * sine)[:, None], (jmt, imt))
    a["cori"] = np.stack([cor1, -cor1]).copy()
    a["advmet"] = np.stack([tng / radius, -(tng / radius)])
    a["am3"] = am * (1.0 - tng * tng) / (radius ** 2)
    am41 = -am * 2.0 * sine / (radius * csu * csu)
    a["am4"] = np.stack([am41, -am41])
    dxt, dxu, dyu = a["dxt"], a["dxu"], a["dyu"]
    dxmetr = np.empty(imt)
    dxmetr[:-1] = 1.0 / (dxt[:-1] + dxt[1:])
    dxmetr[-1] = dxmetr[1]
    a["dxmetr"] = dxmetr
    a["dxu2r"] = 0.5 / dxu
    a["dyu2r"] = 0.5 / dyu
    a["dyu4r"] = 0.25 / dyu
    a["csudyu2r"] = 0.5 / (csu * dyu)
    kmu, zw = a["kmu"], a["zw"]
    a["hr"] = np.where(kmu > 0, 1.0 / zw[np.maximum(kmu, 1) - 1], 0.0)
    # anisotropic viscosity
    tropics = (np.abs(yu) <= 20.0)[:, None, None] & (zw <= 55000.0)[None, :, None]
    dlam = 360.0 / (imt - 2)
    dphi = 178.0 / jmt
    coslat = np.abs(np.cos(pi / 180.0 * yu))
    beddy = 1.0e7 * (1.0 + 24.5 * (1.0 - np.abs(np.cos(2.0 * pi / 180.0 * yu))))
    delx = dlam * 1.11e7 * coslat
    bmunk = 0.2 * (0.0228e-11 * coslat) * delx ** 3
    cnu = np.where(tropics, np.maximum(bmunk, beddy)[:, None, None], am) * np.ones((jmt, km, imt))
    gridlen = np.maximum(delx, dphi * 1.1e7)
    ceu = np.where(tropics, (0.5 * 100.0 * gridlen)[:, None, None], am) * np.ones((jmt, km, imt))
    cst, dytr, csur, dyur = a["cst"], a["dytr"], a["csur"], a["dyur"]
    jp1 = np.minimum(np.arange(jmt) + 1, jmt - 1)
    a["visc_ceu"] = ceu
    a["amc_north"] = cnu * (cst[jp1] * dytr[jp1] * csur * dyur)[:, None, None]
    a["amc_south"] = cnu * (cst * dytr * csur * dyur)[:, None, None]
    # u(tau-1): u(tau) plus a masked perturbation (the vertical mean is not removed: clinic itself does that for tau+1)
    u = a["u"]
    pert = np.stack([_smooth2d(rng, imt, jmt, xu, yu)[:, None, :] * np.cos(pi * a["zt"] / a["zt"][-1])[None, :, None] for _ in range(2)])
    um1 = (0.97 * u + 0.4 * pert) * a["umask"][None]
    um1[..., 0] = um1[..., -2]
    um1[..., -1] = um1[..., 1]
    a["um1"] = um1
    # wind stress (dyn cm-2): trades / westerlies plus noise
    taux = -0.8 * np.cos(3.0 * phi)[:, None] * np.ones((jmt, imt)) + 0.2 * _smooth2d(rng, imt, jmt, xu, yu)
    tauy = 0.2 * _smooth2d(rng, imt, jmt, xu, yu)
    a["taux"], a["tauy"] = taux, tauy
    # polar filter of the velocities (source/common/setcom.F:41-43,56-71,79-81)
    fxa = dxt[0] / radius
    ang = fxa * (np.arange(1, imt + 1) - 2.0)
    spsin, spcos = np.sin(ang), np.cos(ang)
    spsin[np.abs(spsin) < 1.0e-10] = 0.0
    spcos[np.abs(spcos) < 1.0e-10] = 0.0
    spsin[0] = spcos[0] = spsin[-1] = spcos[-1] = 0.0
    a["spsin"], a["spcos"], a["phi"] = spsin, spcos, phi
    case.scalars.update(jfu0=indp(-68.4, yu), jfu1=indp(-70.2, yu), jfu2=indp(70.2, yu))
    if dtuv is None:
        dtuv = case.scalars["dtts"] / 96.0                       # run/control.in:3: dtts=108000, dtuv=1125
    case.scalars.update(c2dtuv=2.0 * dtuv, kappa_m=kappa_m, cdbot=cdbot, grav_rho0r=980.6 * (1.0 / 1.035))
    return case


# arrays with a j extent: axis of j in the C-ordered numpy array (1-D metric arrays of length jmt: axis 0)
_J_1D = ("dyt", "dyu", "dytr", "dyt2r", "dyt4r", "dyur", "cst", "csu", "cstr", "csur", "cstdytr", "cstdyt2r", "csu_dyur", "dus",
         "dun", "_yt", "_yu", "am3", "dyu2r", "dyu4r", "csudyu2r", "phi")
_J_ND = {"kmt": 0, "kmu": 0, "mskhr": 0, "tlat": 0, "tmask": 0, "umask": 0, "fisop": 1, "sg_bathy": 1, "fe_hydr": 1, "fe_atmdep": 1,
         "addisop": 0, "edrm2": 0, "edrs2": 0, "edrk1": 0, "edro1": 0, "adv_vet": 0, "adv_vnt": 0, "adv_vbt": 0, "stf": 1,
         "btf": 1, "u": 1, "t": 2, "dnswr": 0, "aice": 0, "hice": 0, "hsno": 0,
         # inputs of the momentum step (add_momentum)
         "um1": 1, "hr": 0, "cori": 1, "visc_ceu": 0, "amc_north": 0, "amc_south": 0, "taux": 0, "tauy": 0, "advmet": 1, "am4": 1}


def stack_bands(case: Case, nbands: int, lazy: bool = False) -> Case:
    """Weak-scaling grid: `nbands` copies of the interior rows 2..jmt-1 of `case` stacked in latitude between one pair of
    closed walls (global rows 1 and jmt_new), every copy with the geometry, bathymetry, tracers and velocities of the
    original.  The polar rows of the synthetic bathymetry are land, so the copies are separate oceans that only exchange
    land rows: one copy per GPU does exactly the work of the one-GPU case, with the real 2-row halo exchange between
    neighbours.  (Stretching -89..89 degrees over nbands times as many rows instead would shrink dy by nbands and push the
    synthetic meridional velocities past the CFL limit of the unchanged time step.)

    lazy=True keeps every array with three or more dimensions at the size of ONE band and records the row map instead
    (out.lazy, out.row_map; api.slab_slice gathers a slab's rows through it): a rank then holds one band on the host,
    not the whole stack -- what the 0.1 degree grid needs (47 GB per band, 8 bands)."""
    if nbands == 1:
        return case
    jmt = case.jmt
    new_jmt = 2 + (jmt - 2) * nbands

    def tile(a, ax):
        a = np.asarray(a)
        idx = [slice(None)] * a.ndim
        parts = []
        idx[ax] = slice(0, 1)
        parts.append(a[tuple(idx)])
        idx[ax] = slice(1, jmt - 1)
        parts.extend([a[tuple(idx)]] * nbands)
        idx[ax] = slice(jmt - 1, jmt)
        parts.append(a[tuple(idx)])
        return np.ascontiguousarray(np.concatenate(parts, axis=ax))

    arrays = {}
    lazy_names = set()
    for k, v in case.arrays.items():
        if k in _J_1D:
            arrays[k] = tile(v, 0)
        elif k in _J_ND and np.asarray(v).ndim > _J_ND[k] and np.asarray(v).shape[_J_ND[k]] == jmt:
            if lazy and np.asarray(v).ndim >= 3:
                arrays[k] = v
                lazy_names.add(k)
            else:
                arrays[k] = tile(v, _J_ND[k])
        else:
            arrays[k] = v
    out = Case(imt=case.imt, jmt=new_jmt, km=case.km, nt=case.nt, nsrc=case.nsrc, scalars=dict(case.scalars), arrays=arrays,
               tracer_names=list(case.tracer_names))
    out.has_mobi = case.has_mobi
    if lazy:
        out.lazy = lazy_names
        # global row g (0-based) -> row of the band: the walls map to the band's walls, the interior rows repeat
        g = np.arange(new_jmt)
        out.row_map = np.where(g == 0, 0, np.where(g == new_jmt - 1, jmt - 1, 1 + (g - 1) % (jmt - 2)))
    return out
