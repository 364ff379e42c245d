"""Synthetic reach networks and forcing for the scale configurations.

The reference ships no generator; these build parameter objects in the reference's own pandas layout
(``p_SC`` columns = sub-catchment ids in upstream-first order, ``p_struc['Upstream_SCs']`` as blank /
number / "a, b" strings, ``model.py:480-487``) so that both the CUDA path and the oracle consume them.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def random_network(p, base_sc, n_sc=256, seed=3, all_land_uses=False, nc_fraction=0.25, max_jump=16):
    """Random branching tree: sub-catchment i drains into one j drawn from (i, min(i+max_jump, n)], the last
    one is the outlet, so 1..n is already upstream-first.

    ``A_catch ~ U[5, 60] km2``; ``L_reach ~ U[2, 15] km`` (bounded below: the reach rate constant scales
    with 1/L and an explicit integrator becomes stability-limited for very short reaches — SURVEY.md §7);
    land-use fractions on a 0.05 grid so that ``f_A + f_S == 1`` exactly; about ``nc_fraction`` of the
    sub-catchments have one kind of newly-converted land.
    """
    rng = np.random.default_rng(seed)
    ids = list(range(1, n_sc + 1))
    down = {}
    for i in ids[:-1]:
        down[i] = int(rng.integers(i + 1, min(i + max_jump, n_sc) + 1))
    ups = {i: [] for i in ids}
    for i, j in down.items():
        ups[j].append(i)
    cols = {}
    for i in ids:
        col = base_sc.copy()
        col["A_catch"] = float(np.round(rng.uniform(5, 60), 2))
        col["L_reach"] = float(np.round(rng.uniform(2000, 15000), 0))
        g = int(rng.integers(2, 15)) if all_land_uses else int(rng.integers(0, 17))     # f_S on the 0.05 grid
        f_S = g * 0.05 if not all_land_uses else g * 0.05
        n_agri = 20 - int(round(f_S / 0.05))
        k = int(rng.integers(1, n_agri)) if (all_land_uses and n_agri > 1) else int(rng.integers(0, n_agri + 1))
        f_Ar, f_IG = k * 0.05, (n_agri - k) * 0.05
        f_S = 1.0 - (f_Ar + f_IG)
        if (f_Ar + f_IG) + f_S != 1.0:      # keep the reference's exact check satisfiable
            f_Ar, f_IG, f_S = 0.25, 0.25, 0.5
        col["f_Ar"], col["f_IG"], col["f_S"] = f_Ar, f_IG, f_S
        col["f_NC_Ar"] = col["f_NC_IG"] = col["f_NC_S"] = 0.0
        if rng.random() < nc_fraction:
            if rng.random() < 0.5 and f_Ar > 0:
                col["f_NC_Ar"] = float(np.round(rng.uniform(0.05, 0.5), 2))
            elif f_S > 0:
                col["f_NC_S"] = float(np.round(rng.uniform(0.05, 0.5), 2))
        col["S_Ar"] = float(np.round(rng.uniform(1, 12), 1))
        col["S_IG"] = float(np.round(rng.uniform(1, 12), 1))
        col["S_SN"] = float(np.round(rng.uniform(1, 12), 1))
        col["f_spr"] = float(np.round(rng.uniform(0.2, 0.8), 2))
        col["TDPeff"] = float(np.round(rng.uniform(0, 0.3), 3))
        cols[i] = col
    p_SC = pd.DataFrame(cols)
    cells = []
    for i in ids:
        u = sorted(ups[i])
        cells.append(np.nan if not u else (float(u[0]) if len(u) == 1 else ", ".join(str(x) for x in u)))
    p_struc = pd.DataFrame({"Upstream_SCs": pd.Series(cells, index=ids, dtype=object),
                            "In_final_flux?": [np.nan] * (n_sc - 1) + [1.0]}, index=pd.Index(ids, name="Reach"))
    p = p.copy(deep=True)
    p["SC_list"] = np.arange(1, n_sc + 1)
    p["SC_Qr0"] = 1
    return p, p_SC, p_struc


def synthetic_met(n_days, start="1981-01-01", seed=11, p_wet=0.55):
    """Seeded daily forcing: gamma wet-day rain, sinusoidal air temperature 3 +- 9 C + N(0, 3), and a smooth
    seasonal PET.  Returns a DataFrame with T_air, PET, Precipitation (feed it to snow_hydrol_inputs)."""
    rng = np.random.default_rng(seed)
    idx = pd.date_range(start, periods=n_days, freq="D", name="Date")
    doy = idx.dayofyear.to_numpy()
    t_air = 3.0 + 9.0 * np.sin(2 * np.pi * (doy - 110) / 365.25) + rng.normal(0, 3, n_days)
    wet = rng.random(n_days) < p_wet
    rain = np.where(wet, rng.gamma(0.8, 6.0, n_days), 0.0)
    pet = np.clip(1.5 + 1.4 * np.sin(2 * np.pi * (doy - 105) / 365.25), 0.1, None)
    return pd.DataFrame({"T_air": np.round(t_air, 2), "PET": np.round(pet, 2), "Precipitation": np.round(rain, 2)},
                        index=idx)


def scale_config(cfg, members=1, n_days=None):
    """BASELINE.json configs 3 and 5 as the reference's pandas objects plus the packed arrays.

    config 3: 256-sub-catchment branching network (seed 3), 30 years (10,958 days) of seeded synthetic forcing, both
    dynamic options on; config 5: 4096 sub-catchments with all three land-use classes present, 50 years (18,262
    days).  Parameters are Tarland's; with ``members`` > 1 the members differ in ``a_Q`` (0.8x .. 1.2x).  ``n_days``
    truncates the record (parity windows).  Returns a dict: p_SU, dyn, p, p_LU, p_SC, p_struc, met (pandas), topo,
    opt, member [M][40], sc [1][S][16], forcing [D][4].
    """
    from . import inputs as spi, model as spm, packing as pk, tarland
    p_SU, dyn, p, p_LU, p_SC0, _p_struc0, _met, _obs = tarland.load(dynamic="y")
    if cfg == 3:
        n_sc, days, all_lu = 256, 10958, False
    elif cfg == 5:
        n_sc, days, all_lu = 4096, 18262, True
    else:
        raise ValueError("scale_config: cfg must be 3 or 5")
    p, p_SC, p_struc = random_network(p, p_SC0[1], n_sc=n_sc, seed=3, all_land_uses=all_lu)
    met = synthetic_met(days, seed=11)
    met = spi.snow_hydrol_inputs(p["D_snow_0"], p["f_DDSM"], met)
    if n_days is not None:
        met = met.iloc[:int(n_days)]
    pk.validate_land_use(p_SC, p["SC_list"])
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    member = np.repeat(pk.member_vector(p, p_LU)[None], members, axis=0)
    if members > 1:
        member[:, pk.MEMBER_INDEX["a_Q"]] *= np.linspace(0.8, 1.2, members)
    sc = pk.sc_matrix(p_SC, topo.sc_ids)[None]
    return dict(p_SU=p_SU, dyn=dyn, p=p, p_LU=p_LU, p_SC=p_SC, p_struc=p_struc, met=met, topo=topo, opt=opt,
                member=member, sc=sc, forcing=pk.forcing_matrix(met))
