"""TEST INFRASTRUCTURE — CPU restatement ("port") of SimplyP v0-2A's daily mass-balance path.

This is the parity oracle for the CUDA path.  It restates, with plain Python
floats and numpy (no pandas inside the loops), exactly the algorithm of the
reference's ``Current_Release/v0-2A/simplyP/model.py`` and solves each
(sub-catchment, day) with the same third-party solver the reference calls:
``scipy.integrate.odeint`` (ODEPACK LSODA; the reference pins ``scipy=1.2.0`` in
its README, this image has scipy 1.18.1), cold-started every day over
``t = [0, step_len]`` (``model.py:640``).  ``rtol``/``atol`` are arguments here
because parity at 1e-5 is only meaningful at matched *tight* tolerances
(SURVEY.md §7 "Hard parts"); ``rtol=0.01, atol=None`` reproduces the reference's
own setting.

Pinned (tests/test_oracle_*.py) against
* the reference's shipped golden outputs (``Example_Data/Example_Output/*.csv``,
  committed as ``tests/golden/shipped_*.npz``) — at the reference tolerance,
* known-answer vectors of ``f_x``/``ode_f``/``discretized_soilP`` and whole runs of
  the UNMODIFIED reference executed in the build container
  (``oracle/reference_live.py`` -> ``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.  It is never on the product path.
"""
from __future__ import annotations

import math

import numpy as np

try:  # scipy is only needed by the solver call, keep import errors readable
    from scipy.integrate import odeint as _odeint
except Exception as _e:  # pragma: no cover
    _odeint = None
    _odeint_err = _e

# Column orders of the reference's raw outputs (model.py:737-739 and :743-745).
ODE_COLS = ["VsA", "VsS", "Vg", "Vr", "Qr_EndOfDay", "Qr", "Msus_EndOfDay", "Msus_kg/day",
            "TDPr_EndOfDay", "TDP_kg/day", "PPr_EndOfDay", "PP_kg/day"]
NONODE_COLS = ["Qq", "QsA", "QsS", "Qg", "C_cover_A", "EPC0_A_kgmm", "EPC0_NC_kgmm",
               "TDPs_A_kg", "P_labile_A_kg", "conc_TDPs_A_kgmm",
               "TDPs_NC_kgmm", "P_labile_NC_kg", "conc_TDPs_NC_kgmm"]


# --------------------------------------------------------------------------- A1  model.py:23-37
def f_x(x, threshold, reld):
    """Smoothstep gate: 0 below ``threshold``, 1 above ``threshold*(1+reld)``, -2u^3+3u^2 between."""
    d = threshold * reld
    if x < threshold:
        return 0
    if x > threshold + d:
        return 1
    u = (x - threshold) / d
    return -2 * u ** 3 + 3 * u ** 2


# --------------------------------------------------------------------------- A4  model.py:39-56
def discretized_soilP(P_netInput, catchment_area, SC, Kf, Msoil, EPC0, Qs, Qq, Vs, TDPs, Plab):
    """One-day closed-form update of soil-water TDP mass and labile soil P."""
    a = P_netInput * catchment_area * 100 / 365. + Kf * Msoil * EPC0
    b = (Kf * Msoil + Qs + Qq) / Vs
    TDPs = a / b + (TDPs - a / b) * np.exp(-b)          # :44
    b0 = b * Vs
    if Vs > 0:                                             # :50
        # NB uses the already-updated TDPs (:44 then :51)
        sorp = Kf * Msoil * (a / b0 - EPC0 + (1 / b) * (TDPs / Vs - a / b0) * (1 - np.exp(-b)))
    else:
        sorp = 0.
    Plab = Plab + sorp
    return (TDPs, Plab)


# --------------------------------------------------------------------------- A2  model.py:58-187
def ode_f(y, t, q):
    """Right-hand side of the 12-state ODE.  ``q`` is a plain tuple (see ``_pack_rhs_params``)."""
    (P, E, mu, Qq_i, Qr_US_i, EsA, EsS, EsIG, Msus_US_i, TDPr_US_i, PPr_US_i,
     f_A, f_Ar, f_IG, f_S, f_NC_A, f_NC_Ar, f_NC_IG, f_NC_S, nc_is_A,
     f_quick, alpha, beta, T_sA, T_sS, T_g, fc, L_reach, A_catch,
     a_Q, b_Q, k_M, cA, cNC, PlabA, PlabNC, Msoil, TDPeff, TDPg, E_PP, P_inactive, Qg_min) = q

    VsA, VsS, Vg, Vr, Qr = y[0], y[1], y[2], y[3], y[4]
    Msus, TDPr, PPr = y[6], y[8], y[10]

    # soil boxes (:105-110)
    QsA = (VsA - fc) * f_x(VsA, fc, 0.01) / T_sA
    dVsA = P * (1 - f_quick) - alpha * E * (1 - np.exp(-mu * VsA)) - QsA
    QsS = (VsS - fc) * f_x(VsS, fc, 0.01) / T_sS
    dVsS = P * (1 - f_quick) - alpha * E * (1 - np.exp(-mu * VsS)) - QsS
    # newly-converted land shares the hydrology of its new class (:113-118)
    QsNC = QsA if nc_is_A else QsS
    # groundwater (:121-124)
    f_Qg = f_x(Vg / T_g, Qg_min, 0.01)
    Qg = (1 - f_Qg) * Qg_min + f_Qg * (Vg / T_g)
    dVg = beta * (f_A * QsA + f_S * QsS) - Qg
    # reach (:127-132)
    net = Qq_i + (1 - beta) * (f_A * QsA + f_S * QsS) + Qg + Qr_US_i - Qr
    dQr = net * a_Q * (Qr ** b_Q) * (8.64 * 10 ** 4) / ((1 - b_Q) * L_reach)
    dVr = net
    dQr_av = Qr
    # sediment (:138-147)
    qk = Qr ** k_M
    MinA, MinS, MinIG = EsA * qk, EsS * qk, EsIG * qk
    dMsus = f_Ar * MinA + f_IG * MinIG + f_S * MinS + Msus_US_i - (Msus / Vr) * Qr
    dMsus_out = (Msus / Vr) * Qr
    # TDP (:154-168)
    dTDPr = ((1 - beta) * (f_A * (1 - f_NC_A) * QsA * cA
                           + f_A * f_NC_A * QsNC * cNC
                           + f_S * f_NC_S * QsNC * cNC)
             + f_A * (1 - f_NC_A) * Qq_i * cA
             + f_A * f_NC_A * Qq_i * cNC
             + f_S * f_NC_S * Qq_i * cNC
             + Qg * (TDPg * A_catch)
             + TDPeff
             + TDPr_US_i
             - Qr * (TDPr / Vr))
    dTDPr_out = Qr * (TDPr / Vr)
    # PP (:171-180)
    dPPr = (E_PP * (f_Ar * (1 - f_NC_Ar) * MinA * (PlabA + P_inactive) / Msoil
                    + f_IG * (1 - f_NC_IG) * MinIG * (PlabA + P_inactive) / Msoil
                    + f_S * (1 - f_NC_S) * MinS * P_inactive / Msoil
                    + f_Ar * f_NC_Ar * MinA * (PlabNC + P_inactive) / Msoil
                    + f_IG * f_NC_IG * MinIG * (PlabNC + P_inactive) / Msoil
                    + f_S * f_NC_S * MinS * (PlabNC + P_inactive) / Msoil)
            + PPr_US_i
            - Qr * (PPr / Vr))
    dPPr_out = Qr * PPr / Vr
    return [dVsA, dVsS, dVg, dVr, dQr, dQr_av, dMsus, dMsus_out, dTDPr, dTDPr_out, dPPr, dPPr_out]


# --------------------------------------------------------------------------- helpers
def parse_upstream(cell):
    """Reach-structure cell -> list of upstream SC ids (model.py:480-487)."""
    if isinstance(cell, str):
        return [int(x.strip()) for x in cell.split(",")]
    if isinstance(cell, (int, np.integer)):
        return [int(cell)]
    try:
        if not np.isnan(cell):
            return [int(cell)]
    except TypeError:
        pass
    return []


def crop_cover_factor(dayNo, C_cover_A, f_spr, d_maxE_spr, d_maxE_aut):
    """Dynamic arable cover factor, triangular wave (model.py:354-361, :563-580)."""
    E_risk_period = 60.0
    out = {}
    for s, mid in (("spr", d_maxE_spr), ("aut", d_maxE_aut)):
        d_start, d_end = mid - E_risk_period / 2., mid + E_risk_period / 2.
        # `dayNo in np.arange(d_start, d_end)`: membership in {d_start, d_start+1, ...} below d_end (:567)
        k = dayNo - d_start
        inside = (k >= 0) and (dayNo < d_end) and (k == math.floor(k))
        if inside:
            if dayNo < mid:
                C = C_cover_A + (1.0 - C_cover_A) * (dayNo - d_start) / (mid - d_start)
            else:
                C = 1.0 + (C_cover_A - 1.0) * (dayNo - mid) / (d_end - mid)
        else:
            C = C_cover_A - (E_risk_period * (1 - C_cover_A) / (2 * (365 - E_risk_period)))
        out[s] = C
    return f_spr * out["spr"] + (1 - f_spr) * out["aut"]


def snow_hydrol_inputs(D_snow_0, f_DDSM, precip, t_air):
    """Degree-day snow module on plain arrays (inputs.py:159-210).  Returns (P, D_snow_end, P_melt)."""
    n = len(precip)
    P = np.empty(n)
    D_end = np.empty(n)
    melt_out = np.empty(n)
    depth = float(D_snow_0)
    for i in range(n):
        p_snow = precip[i] if t_air[i] < 0 else 0.0        # :183-184
        p_rain = precip[i] - p_snow                        # :187
        melt = f_DDSM * (t_air[i] - 0)                     # :190
        if melt < 0:
            melt = 0.0                                     # :191
        melt = min(melt, depth)                            # :198,204
        depth = depth + p_snow - melt                      # :199-200,205
        P[i] = p_rain + melt                               # :208
        D_end[i] = depth
        melt_out[i] = melt
    return P, D_end, melt_out


# --------------------------------------------------------------------------- driver  model.py:193-827
def run_network(forcing_P, forcing_PET, doy, p, p_LU, p_SC, upstream, run_mode="cal",
                dynamic_EPC0=False, dynamic_erodibility=False, step_len=1.0,
                rtol=0.01, atol=None, mxstep=5000, strict_quirks=True, n_days=None, only=None, preset=None):
    """Integrate a whole reach network, SC-major then day-minor like the reference.

    Plain-python inputs:
      forcing_P, forcing_PET, doy : sequences of length D
      p      : dict of scalars (f_quick, alpha, fc, beta, T_g, Qg_min, a_Q, b_Q, Qr0_init, SC_Qr0,
               Msoil_m2, Kf, TDPg, E_PP, E_M, k_M, d_maxE_spr, d_maxE_aut)
      p_LU   : dict name -> dict class -> value (classes 'A','S','IG','NC')
      p_SC   : dict sc_id -> dict name -> value
      upstream : dict sc_id -> list of directly-upstream sc ids; iteration order of ``p_SC`` is the
                 run order and must be upstream-first (model.py:524 raises KeyError otherwise)

    ``only`` / ``preset`` serve ``oracle/parallel.py`` (the same computation spread over processes): ``preset`` maps
    sub-catchment ids to ``(ode [D][12], nonode [D][13])`` computed earlier, which are taken as they are, and only the
    ids in ``only`` are integrated (their direct parents must be preset or in ``only``).

    Returns dict with ``ode`` [S][D][12], ``nonode`` [S][D][13], ``Kf`` {sc: Kf}, ``nfe`` total RHS calls,
    ``sc_ids``.
    """
    if _odeint is None:  # pragma: no cover
        raise RuntimeError("scipy.integrate.odeint unavailable: %r" % (_odeint_err,))
    D = len(forcing_P) if n_days is None else int(n_days)
    sc_ids = list(p_SC.keys())

    # derived SC params + validation (:318-335)
    der = {}
    last_nc_type = "None"
    for SC in sc_ids:
        q = p_SC[SC]
        f_A = q["f_IG"] + q["f_Ar"]
        f_NC_A = (q["f_Ar"] * q["f_NC_Ar"]) + (q["f_NC_IG"] * q["f_IG"])
        if (f_A + q["f_S"]) != 1:
            raise ValueError("Land use proportions do not add to 1 in SC %s" % SC)
        if f_NC_A > 0:
            if q["f_NC_S"] > 0:
                raise ValueError("Sub-catchment %s has 2 kinds of newly-converted land" % SC)
            nc = "A"
        elif q["f_NC_S"] > 0:
            nc = "S"
        else:
            nc = "None"
        der[SC] = (f_A, f_NC_A, nc)
        last_nc_type = nc  # the python variable that leaks out of the loop (Appendix D.1)

    mu = -math.log(0.01) / p["fc"]                                        # :349
    for season in ("spr", "aut"):
        assert 30 < p["d_maxE_%s" % season] < 335, "'d_maxE_%s' must be between 30 and 335" % season

    ode_out = np.zeros((len(sc_ids), D, 12))
    non_out = np.zeros((len(sc_ids), D, 13))
    Kf_out = {}
    pos = {SC: i for i, SC in enumerate(sc_ids)}
    nfe = 0
    nst = 0

    for SC in sc_ids:
        if preset is not None and SC in preset:
            ode_out[pos[SC]], non_out[pos[SC]] = preset[SC]
            continue
        if only is not None and SC not in only:
            continue
        q = p_SC[SC]
        f_A, f_NC_A, nc_type = der[SC]
        post_nc_type = last_nc_type if strict_quirks else nc_type         # :442,676 use the leaked name
        A_catch = q["A_catch"]

        # initial conditions (:377-396)
        VsA0 = p["fc"]
        VsS0 = VsA0
        Qr0 = p["Qr0_init"] * 86400 / (1000 * p_SC[p["SC_Qr0"]]["A_catch"])   # UC_Qinv, :386
        Vg0 = (p["beta"] * Qr0) * p["T_g"]
        TDPr0, PPr0, Msus0 = 0.0, 0.0, 0.0
        # soil P (:404-446)
        Msoil = p["Msoil_m2"] * 10 ** 6 * A_catch
        P_inactive = 10 ** -6 * p_LU["SoilPconc"]["S"] * Msoil
        EPC0_0 = {LU: p_LU["EPC0_init_mgl"][LU] * A_catch for LU in ("A", "S")}
        Plab0 = {LU: 10 ** -6 * (p_LU["SoilPconc"][LU] - p_LU["SoilPconc"]["S"]) * Msoil for LU in ("A", "S")}
        TDPs0 = {"A": EPC0_0["A"] * VsA0, "S": 0}
        Plab0_A, TDPs0_A = Plab0["A"], TDPs0["A"]
        if nc_type == "S":
            Plab0_NC, TDPs0_NC = Plab0_A, TDPs0_A
        else:
            Plab0_NC, TDPs0_NC = 0.0, TDPs0["S"]
        conc_TDPs_A = TDPs0_A / VsA0
        VsNC0 = VsA0 if post_nc_type == "A" else VsS0
        conc_TDPs_NC = TDPs0_NC / VsNC0
        if run_mode == "cal":                                              # :449-453
            Kf = 10 ** -6 * (p_LU["SoilPconc"]["A"] - p_LU["SoilPconc"]["S"]) / EPC0_0["A"]
        else:
            Kf = p["Kf"]
        Kf_out[SC] = Kf
        Tr0 = q["L_reach"] / (p["a_Q"] * Qr0 ** p["b_Q"] * 8.64 * 10 ** 4)    # :457-459
        Vr0 = Tr0 * Qr0
        TDPeff = q["TDPeff"]
        if TDPeff is None or (isinstance(TDPeff, float) and math.isnan(TDPeff)):
            TDPeff = 0.                                                    # :462-463
        slope = {"A": q["S_Ar"], "IG": q["S_IG"], "S": q["S_SN"]}          # :469
        ups = upstream.get(SC, [])
        isc = pos[SC]

        for idx in range(D):
            P = float(forcing_P[idx])
            E = float(forcing_PET[idx])
            Qq_i = p["f_quick"] * P                                        # :501
            # upstream inputs, same day (:508-544)
            Qr_US_i = Msus_US_i = TDPr_US_i = PPr_US_i = 0.0
            if len(ups) > 0:
                Qr_li, Ms_li, TD_li, PP_li = [], [], [], []
                for up in ups:
                    row = ode_out[pos[up], idx]
                    Qr_li.append(row[5] * (p_SC[up]["A_catch"] / A_catch))
                    Ms_li.append(row[7])
                    TD_li.append(row[9])
                    PP_li.append(row[11])
                Qr_US_i, Msus_US_i, TDPr_US_i, PPr_US_i = sum(Qr_li), sum(Ms_li), sum(TD_li), sum(PP_li)
            # erodibility (:549-594)
            dayNo = doy[idx]
            if dynamic_erodibility:
                C_cover_A = crop_cover_factor(dayNo, p_LU["C_cover"]["A"], q["f_spr"],
                                              p["d_maxE_spr"], p["d_maxE_aut"])
            else:
                C_cover_A = p_LU["C_cover"]["A"]
            Es = {}
            for LU in ("A", "S", "IG"):
                C = C_cover_A if LU == "A" else p_LU["C_cover"][LU]
                Es[LU] = p["E_M"] * q["S_reach"] * slope[LU] * C * (1 - p_LU["C_measures"][LU])
            # EPC0 (:600-611)
            if dynamic_EPC0:
                EPC0_A_i = max(Plab0_A / (Kf * Msoil), 0)
                EPC0_NC_i = max(Plab0_NC / (Kf * Msoil), 0)
            else:
                EPC0_A_i = EPC0_0["A"]
                EPC0_NC_i = EPC0_0["A"] if nc_type == "S" else EPC0_0["S"]

            y0 = [VsA0, VsS0, Vg0, Vr0, Qr0, 0.0, Msus0, 0.0, TDPr0, 0.0, PPr0, 0.0]   # :618
            prm = (P, E, mu, Qq_i, Qr_US_i, Es["A"], Es["S"], Es["IG"], Msus_US_i, TDPr_US_i, PPr_US_i,
                   f_A, q["f_Ar"], q["f_IG"], q["f_S"], f_NC_A, q["f_NC_Ar"], q["f_NC_IG"], q["f_NC_S"],
                   nc_type == "A",
                   p["f_quick"], p["alpha"], p["beta"], p_LU["T_s"]["A"], p_LU["T_s"]["S"], p["T_g"], p["fc"],
                   q["L_reach"], A_catch, p["a_Q"], p["b_Q"], p["k_M"], conc_TDPs_A, conc_TDPs_NC,
                   Plab0_A, Plab0_NC, Msoil, TDPeff, p["TDPg"], p["E_PP"], P_inactive, p["Qg_min"])
            kw = {"rtol": rtol, "mxstep": mxstep}
            if atol is not None:
                kw["atol"] = atol
            y, info = _odeint(ode_f, y0, [0, step_len], args=(prm,), full_output=1, **kw)   # :640
            nfe += int(info["nfe"][-1])
            nst += int(info["nst"][-1])
            res = y[1]
            ode_out[isc, idx] = res

            VsA0, VsS0, Vg0, Vr0, Qr0 = res[0], res[1], res[2], res[3], res[4]   # :648-658
            Msus0, TDPr0, PPr0 = res[6], res[8], res[10]
            QsA0 = (VsA0 - p["fc"]) * f_x(VsA0, p["fc"], 0.01) / p_LU["T_s"]["A"]   # :663-664
            QsS0 = (VsS0 - p["fc"]) * f_x(VsS0, p["fc"], 0.01) / p_LU["T_s"]["S"]
            f_Qg = f_x(Vg0 / p["T_g"], p["Qg_min"], 0.01)                   # :668-670
            Qg0 = (1 - f_Qg) * p["Qg_min"] + f_Qg * (Vg0 / p["T_g"])
            Vg0 = Qg0 * p["T_g"]
            if post_nc_type == "A":                                         # :676-681
                VsNC0, QsNC0 = VsA0, QsA0
            else:
                VsNC0, QsNC0 = VsS0, QsS0
            if dynamic_EPC0:                                                # :684-703
                TDPs0_A, Plab0_A = discretized_soilP(p_LU["P_netInput"]["A"], A_catch, SC, Kf, Msoil,
                                                     EPC0_A_i, QsA0, Qq_i, VsA0, TDPs0_A, Plab0_A)
                TDPs0_NC, Plab0_NC = discretized_soilP(p_LU["P_netInput"]["NC"], A_catch, SC, Kf, Msoil,
                                                       EPC0_NC_i, QsNC0, Qq_i, VsNC0, TDPs0_NC, Plab0_NC)
                TDPs0_A = max(TDPs0_A, 0.)
                Plab0_A = max(Plab0_A, 0.)
                TDPs0_NC = max(TDPs0_NC, 0.)
                Plab0_NC = max(Plab0_NC, 0.)
                conc_TDPs_A = TDPs0_A / VsA0
                conc_TDPs_NC = TDPs0_NC / VsNC0
            else:                                                           # :707-715
                conc_TDPs_A = EPC0_A_i
                conc_TDPs_NC = EPC0_NC_i
            non_out[isc, idx] = [Qq_i, QsA0, QsS0, Qg0, C_cover_A, EPC0_A_i, EPC0_NC_i,
                                 TDPs0_A, Plab0_A, conc_TDPs_A, TDPs0_NC, Plab0_NC, conc_TDPs_NC]   # :721-723

    return {"ode": ode_out, "nonode": non_out, "Kf": Kf_out, "nfe": nfe, "nst": nst, "sc_ids": sc_ids,
            "nc_type": {SC: der[SC][2] for SC in sc_ids}}


# --------------------------------------------------------------------------- pandas adaptor
def _to_float(x):
    try:
        return float(x)
    except (TypeError, ValueError):
        return float("nan")


def unpack_pandas(p_struc, p_LU, p_SC, p):
    """pandas parameter objects (reference layout) -> the plain dicts ``run_network`` takes."""
    sc_ids = [int(s) for s in p["SC_list"]]
    pd_ = {k: _to_float(p[k]) for k in p.index if k != "SC_list"}
    pd_["SC_Qr0"] = int(p["SC_Qr0"])
    lu = {}
    for name in p_LU.index:
        lu[name] = {c: _to_float(p_LU.loc[name, c]) for c in p_LU.columns}
    if "P_netInput" in lu:
        for c in ("A", "NC"):
            if math.isnan(lu["P_netInput"].get(c, float("nan"))):
                pass  # a blank stays NaN, exactly like the reference
    sc = {}
    for s in sc_ids:
        sc[s] = {name: _to_float(p_SC.loc[name, s]) for name in p_SC.index if name != "NC_type"}
    ups = {s: parse_upstream(p_struc.loc[s, "Upstream_SCs"]) for s in sc_ids}
    return pd_, lu, sc, ups


def run_simply_p(met_df, p_struc, p_SU, p_LU, p_SC, p, dynamic_options, step_len=1.,
                 rtol=0.01, atol=None, mxstep=5000):
    """Same signature/return as the reference's ``run_simply_p`` (model.py:193, :827), built on
    ``run_network``; does not mutate its arguments.  Returns (df_TC_dict, df_R_dict, Kf, info)."""
    import pandas as pd

    pd_, lu, sc, ups = unpack_pandas(p_struc, p_LU, p_SC, p)
    raw = run_network(met_df["P"].to_numpy(), met_df["PET"].to_numpy(), met_df.index.dayofyear.to_numpy(),
                      pd_, lu, sc, ups, run_mode=p_SU["run_mode"],
                      dynamic_EPC0=(dynamic_options["Dynamic_EPC0"] == "y"),
                      dynamic_erodibility=(dynamic_options["Dynamic_erodibility"] == "y"),
                      step_len=step_len, rtol=rtol, atol=atol, mxstep=mxstep)
    df_TC_dict, df_R_dict = {}, {}
    for i, SC in enumerate(raw["sc_ids"]):
        A = sc[SC]["A_catch"]
        df_ODE = pd.DataFrame(raw["ode"][i], columns=ODE_COLS, index=met_df.index)
        df_non = pd.DataFrame(raw["nonode"][i], columns=NONODE_COLS, index=met_df.index)
        df_TC = pd.concat([df_ODE[["VsA", "VsS", "Vg"]], df_non], axis=1)              # :755
        df_TC["TDPs_A_mgl"] = df_TC["conc_TDPs_A_kgmm"] / A                             # :758-761
        df_TC["EPC0_A_mgl"] = df_TC["EPC0_A_kgmm"] / A
        df_TC["Plabile_A_mgkg"] = 10 ** 6 * df_TC["P_labile_A_kg"] / (pd_["Msoil_m2"] * 10 ** 6 * A)
        nc = raw["nc_type"][SC]
        if nc != "None":                                                                # :764-773
            src = "A" if nc == "A" else "S"
            df_TC["VsNC"] = df_TC["Vs" + src]
            df_TC["QsNC"] = df_TC["Qs" + src]
            df_TC["TDPs_NC_mgl"] = df_TC["conc_TDPs_NC_kgmm"] / A
            df_TC["Plabile_NC_mgkg"] = 10 ** 6 * df_TC["P_labile_NC_kg"] / (pd_["Msoil_m2"] * 10 ** 6 * A)
        if p_SU["inc_snowmelt"] == "y":
            df_TC["D_snow"] = met_df["D_snow_end"]                                      # :775-776
        df_R = df_ODE.drop(["VsA", "VsS", "Vg"], axis=1)                                # :779-793
        df_R["Q_cumecs"] = df_R["Qr"] * A * 1000 / 86400
        df_R["SS_mgl"] = (df_R["Msus_kg/day"] / df_R["Qr"]) / A
        df_R["TDP_mgl"] = (df_R["TDP_kg/day"] / df_R["Qr"]) / A
        df_R["PP_mgl"] = (df_R["PP_kg/day"] / df_R["Qr"]) / A
        df_R["TP_mgl"] = df_R["TDP_mgl"] + df_R["PP_mgl"]                               # :840-845
        df_R["TP_kg/day"] = df_R["TDP_kg/day"] + df_R["PP_kg/day"]
        df_R["SRP_mgl"] = df_R["TDP_mgl"] * pd_["f_TDP"]
        df_R["SRP_kg/day"] = df_R["TDP_kg/day"] * pd_["f_TDP"]
        df_TC_dict[SC] = df_TC.sort_index(axis=1)
        df_R_dict[SC] = df_R.sort_index(axis=1)
    Kf = raw["Kf"][raw["sc_ids"][-1]]
    return df_TC_dict, df_R_dict, Kf, {"nfe": raw["nfe"], "nst": raw["nst"], "message": "oracle (LSODA)"}


# --------------------------------------------------------------------------- A10  visualise_results.py:387-474
def gof_stats(obs, sim):
    """Goodness-of-fit of one variable: arrays aligned on date; NaN pairs dropped.

    Returns dict(n, NSE, log_NSE, spearman_r, r2, pbias, nRMSD) following
    ``visualise_results.py:441-449`` (np.std is the population std, ddof=0).  ``n`` is the number of
    non-null observations *before* alignment (:429); statistics are None if ``n <= 10`` (:431).
    """
    obs = np.asarray(obs, dtype=float)
    sim = np.asarray(sim, dtype=float)
    n_obs = int(np.sum(~np.isnan(obs)))
    if n_obs <= 10:
        return {"n": n_obs}
    ok = ~np.isnan(obs) & ~np.isnan(sim)
    o, s = obs[ok], sim[ok]
    lo, ls = np.log(o), np.log(s)
    NSE = 1 - np.sum((o - s) ** 2) / np.sum((o - np.mean(o)) ** 2)
    log_NSE = 1 - np.sum((lo - ls) ** 2) / np.sum((lo - np.mean(lo)) ** 2)
    from scipy.stats import rankdata
    ro, rs = rankdata(o), rankdata(s)
    spearman = float(np.corrcoef(ro, rs)[0, 1])
    r2 = float(np.corrcoef(o, s)[0, 1] ** 2)
    pbias = 100 * np.sum(s - o) / np.sum(o)
    nrmsd = 100 * np.mean(np.abs(s - o)) / np.std(o)
    return {"n": n_obs, "NSE": float(NSE), "log_NSE": float(log_NSE), "spearman_r": spearman,
            "r2": r2, "pbias": float(pbias), "nRMSD": float(nrmsd)}


# --------------------------------------------------------------------------- A11  Development/2016/MCMC.ipynb:213-242
def gaussian_log_likelihood(obs, sim, m):
    """Heteroscedastic Gaussian log-likelihood, sigma = m*sim:
    sum(-0.5 ln(2 pi) - ln(m sim) - (obs-sim)^2 / (2 m^2 sim^2)) over pairs with an observation."""
    obs = np.asarray(obs, dtype=float)
    sim = np.asarray(sim, dtype=float)
    ok = ~np.isnan(obs)
    o, s = obs[ok], sim[ok]
    sigma = m * s
    ll = -0.5 * math.log(2 * math.pi) - np.log(sigma) - (o - s) ** 2 / (2 * sigma ** 2)
    tot = float(np.sum(ll))
    return -math.inf if math.isnan(tot) else tot
