"""The Tarland (Scotland) example set-up, rebuilt from the package data in ``simplyp_b200/data/tarland/``.

The reference ships this set-up as an Excel workbook, a met CSV and two observation workbooks
(``Example_Data/Tarland_Scotland``).  Those files are not available where the GPU tests and the
benchmark run, so ``tests/golden/make_golden.py`` stores their contents as small JSON/npz files (the example
data of this package) and this module turns them back into exactly the pandas objects ``read_input_data``
returns.
"""
from __future__ import annotations

import json
import os

import numpy as np
import pandas as pd

from .inputs import snow_hydrol_inputs

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "tarland")


def _nan(x):
    return np.nan if x is None else x


def load_parameters(data_dir=DATA_DIR):
    with open(os.path.join(data_dir, "tarland_inputs.json")) as f:
        d = json.load(f)
    p_SU = pd.Series({k: _nan(v) for k, v in d["p_SU"].items()}, name="Value", dtype=object)
    p_SU["n_SC"] = int(p_SU["n_SC"])
    p = pd.Series({k: v for k, v in d["p"].items()}, name="Value", dtype=object)
    p["SC_list"] = np.arange(1, int(p_SU["n_SC"]) + 1)
    p_LU = pd.DataFrame({c: {row: _nan(d["p_LU"][row][c]) for row in d["p_LU"]} for c in ["A", "S", "IG", "NC"]},
                        index=list(d["p_LU"].keys()), dtype=float)
    p_SC = pd.DataFrame({int(c): {row: _nan(v) for row, v in col.items()} for c, col in d["p_SC"].items()}, dtype=float)
    p_struc = pd.DataFrame({"Upstream_SCs": [np.nan] * len(d["p_struc"]), "In_final_flux?": [np.nan] * len(d["p_struc"])},
                           index=pd.Index([int(i) for i in d["p_struc"]], name="Reach"))
    p_struc["Upstream_SCs"] = p_struc["Upstream_SCs"].astype(object)
    dynamic_options = p_SU[["Dynamic_EPC0", "Dynamic_effluent_inputs", "Dynamic_terrestrialP_inputs",
                            "Dynamic_erodibility"]].copy()
    return p_SU, dynamic_options, p, p_LU, p_SC, p_struc


def load_met(st_dt="2004-01-01", end_dt="2004-12-31", D_snow_0=0.0, f_DDSM=2.74, inc_snowmelt=True,
             data_dir=DATA_DIR):
    z = np.load(os.path.join(data_dir, "tarland_met.npz"))
    idx = pd.date_range(str(z["day0"]), periods=int(z["n"]), freq="D", name="Date")
    met = pd.DataFrame({"T_air": z["T_air"], "PET": z["PET"], "Precipitation": z["Precipitation"]}, index=idx)
    met = met.truncate(before=st_dt, after=end_dt)
    if inc_snowmelt:
        met = snow_hydrol_inputs(D_snow_0, f_DDSM, met)
    else:
        met = met.rename(columns={"Precipitation": "P"})
    return met


def load_obs(st_dt="2004-01-01", end_dt="2004-12-31", data_dir=DATA_DIR):
    z = np.load(os.path.join(data_dir, "tarland_obs.npz"))
    epoch = pd.Timestamp("1970-01-01")
    q = pd.DataFrame({"Q": z["Q"]}, index=pd.DatetimeIndex(epoch + pd.to_timedelta(z["q_days"], unit="D"), name="Date"))
    chem_cols = [c for c in ("SRP", "SS", "TDP", "TP", "PP") if c in z.files]
    chem = pd.DataFrame({c: z[c] for c in chem_cols},
                        index=pd.DatetimeIndex(epoch + pd.to_timedelta(z["chem_days"], unit="D"), name="Date"))
    q = q.truncate(before=st_dt, after=end_dt)
    chem = chem.truncate(before=st_dt, after=end_dt)
    # the shipped workbook lists the chemistry file first (its two obs paths are swapped)
    return {1: pd.concat([chem, q], axis=1, sort=True)}


def load(st_dt="2004-01-01", end_dt="2004-12-31", dynamic="n"):
    """Same 8-tuple as ``read_input_data``: (p_SU, dynamic_options, p, p_LU, p_SC, p_struc, met_df, obs_dict)."""
    p_SU, dyn, p, p_LU, p_SC, p_struc = load_parameters()
    p_SU["st_dt"], p_SU["end_dt"] = st_dt, end_dt
    dyn["Dynamic_EPC0"] = dynamic
    dyn["Dynamic_erodibility"] = dynamic
    met = load_met(st_dt, end_dt, p["D_snow_0"], p["f_DDSM"], p_SU["inc_snowmelt"] == "y")
    obs = load_obs(st_dt, end_dt)
    return p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs
