"""Input side of the SimplyP hot path: workbook/met/obs reader, snow module, PET.

Entry points keep the reference's names, argument order and return shapes
(reference ``simplyP/inputs.py``: ``read_input_data`` :19-155,
``snow_hydrol_inputs`` :159-210, ``daily_PET`` :232-312 and the Thornthwaite
helpers :315-508), so notebooks written against the reference keep working.
The implementation is new: the workbook is parsed with the stdlib reader in
``_xlsx.py`` (no openpyxl/xlrd), the snow pack is a vectorised pre-pass plus a
scalar scan over plain numpy arrays, and PET uses current pandas APIs.

These functions are host-side pre-processing of the *shared* forcing (one met
series for the whole catchment, reference ``inputs.py:90``); their output
columns ``P``, ``PET`` and ``D_snow_end`` are what the CUDA integrator consumes.
"""
from __future__ import annotations

import calendar
import math
import os

import numpy as np
import pandas as pd

from ._xlsx import Workbook

__all__ = ["read_input_data", "snow_hydrol_inputs", "daily_PET"]


# --------------------------------------------------------------------------- paths
def _resolve_path(path, workbook_path):
    """Resolve a data path written in the workbook.

    The shipped workbook holds Windows-style paths relative to the notebook's
    working directory (e.g. ``..\\..\\Example_Data\\...``).  We accept those on any
    OS and additionally look relative to the workbook's own directory.
    """
    if not isinstance(path, str):
        return path
    candidates = [path, path.replace("\\", os.sep)]
    base = os.path.dirname(os.path.abspath(workbook_path))
    candidates += [os.path.join(base, c) for c in list(candidates)]
    for c in candidates:
        if os.path.exists(c):
            return os.path.normpath(c)
    return path.replace("\\", os.sep)


def _read_obs_workbook(path, st_dt, end_dt):
    """One workbook with a sheet per reach id -> ``{reach: DataFrame}`` truncated to the run period."""
    wb = Workbook(path)
    out = {}
    for name in wb.sheet_names:
        reach = int(name)
        df = wb.read_table(name, index_col=0)
        idx = df.index
        if not isinstance(idx, pd.DatetimeIndex):
            if pd.api.types.is_numeric_dtype(idx):
                idx = pd.to_datetime(np.asarray(idx, dtype="float64"), unit="D", origin="1899-12-30")
            else:
                idx = pd.to_datetime(idx)
            df.index = idx
        df.index.name = "Date"
        df = df.sort_index().truncate(before=st_dt, after=end_dt)
        out[reach] = df
    return out


def read_input_data(params_fpath):
    """Read a SimplyP Excel parameter workbook plus the met and observation files it names.

    Returns the same 8-tuple as the reference (``inputs.py:155``):
    ``(p_SU, dynamic_options, p, p_LU, p_SC, p_struc, met_df, obs_dict)``.
    """
    wb = Workbook(params_fpath)

    # Setup sheet: parameter name in column A, value in column C (ref inputs.py:44-45)
    p_SU = wb.read_table("Setup", usecols="A,C")["Value"]
    dynamic_options = p_SU[["Dynamic_EPC0", "Dynamic_effluent_inputs",
                            "Dynamic_terrestrialP_inputs", "Dynamic_erodibility"]]

    # Constant sheet: name in B, value in E (ref inputs.py:57-58)
    p = wb.read_table("Constant", usecols="B,E")["Value"]
    p = p.astype(object)

    # Land-use sheet: name in B, classes A,S,IG,NC in E:H (ref inputs.py:61)
    p_LU = wb.read_table("LU", usecols="B,E,F,G,H")

    # Sub-catchment sheet: one column per SC starting at E (ref inputs.py:65-71)
    n_SC = int(p_SU["n_SC"])
    p["SC_list"] = np.arange(1, n_SC + 1)
    first = ord("E") - ord("A")
    sc_cols = ",".join(_excel_col(first + i) for i in range(n_SC))
    p_SC = wb.read_table("SC_reach", usecols="B," + sc_cols)
    p_SC.columns = [int(c) if _is_intlike(c) else c for c in p_SC.columns]

    # Reach structure (ref inputs.py:76-77)
    p_struc = wb.read_table("Reach_structure", usecols="A,B,C")
    p_struc.columns = ["Upstream_SCs", "In_final_flux?"]

    if n_SC != len(p_struc["Upstream_SCs"]):
        raise ValueError("The number of sub-catchments specified in your 'Setup' parameter sheet doesn't \n"
                         "match the number of rows in your 'Reach_structure' sheet")
    if n_SC != len(p_SC.columns):
        raise ValueError("The number of columns in your 'SC_reach' sheet should match the number of "
                         "sub-catchments specified in your 'Setup' parameter sheet")
    print("Parameter values successfully read in")

    # ---- met data (ref inputs.py:91-105)
    met_path = _resolve_path(p_SU["metdata_fpath"], params_fpath)
    met_df = pd.read_csv(met_path, parse_dates=True, dayfirst=True, index_col=0)
    met_df = met_df.truncate(before=p_SU["st_dt"], after=p_SU["end_dt"])
    print("Input meteorological data read in")

    if p_SU["inc_snowmelt"] == "y":
        met_df = snow_hydrol_inputs(p["D_snow_0"], p["f_DDSM"], met_df)
        print("Snow accumulation and melt module run to estimate snowmelt inputs to the soil")
    else:
        met_df = met_df.rename(columns={"Precipitation": "P"})

    if "PET" not in met_df.columns:
        met_df = daily_PET(latitude=p["latitude"], met_df=met_df)
        print("PET estimated using the Thornthwaite method")

    # ---- observations (ref inputs.py:118-152)
    q_obs, chem_obs = {}, {}
    if isinstance(p_SU.get("Qobsdata_fpath"), str):
        q_obs = _read_obs_workbook(_resolve_path(p_SU["Qobsdata_fpath"], params_fpath),
                                   p_SU["st_dt"], p_SU["end_dt"])
        print("Observed discharge data read in")
    if isinstance(p_SU.get("chemObsData_fpath"), str):
        chem_obs = _read_obs_workbook(_resolve_path(p_SU["chemObsData_fpath"], params_fpath),
                                      p_SU["st_dt"], p_SU["end_dt"])
        print("Observed water chemistry data read in")

    obs_dict = {}
    for SC in p["SC_list"]:
        frames = [d[SC] for d in (q_obs, chem_obs) if SC in d]
        if frames:
            obs_dict[int(SC)] = pd.concat(frames, axis=1, sort=True)

    return (p_SU, dynamic_options, p, p_LU, p_SC, p_struc, met_df, obs_dict)


def _excel_col(idx):
    s = ""
    idx += 1
    while idx:
        idx, rem = divmod(idx - 1, 26)
        s = chr(ord("A") + rem) + s
    return s


def _is_intlike(x):
    try:
        return float(x) == int(float(x))
    except (TypeError, ValueError):
        return False


# --------------------------------------------------------------------------- snow
def snow_hydrol_inputs(D_snow_0, f_DDSM, met_df):
    """Degree-day snow accumulation and melt (reference ``inputs.py:159-210``).

    Adds columns ``P_snow, P_rain, P_melt, D_snow_start, D_snow_end, P`` where
    ``P`` = rain + melt is the hydrological input to the soil box (mm/day).
    Precipitation counts as snow when ``T_air < 0``; potential melt is
    ``f_DDSM * T_air`` clipped at 0 and limited by the pack depth at the start of
    the day.  The pack recursion is inherently serial, so it runs as one scalar
    scan over numpy arrays (10,957 days take ~2 ms).
    """
    met_df = met_df.copy()
    precip = met_df["Precipitation"].to_numpy(dtype="float64")
    t_air = met_df["T_air"].to_numpy(dtype="float64")

    p_snow = np.where(t_air < 0, precip, 0.0)
    p_snow = np.where(np.isnan(p_snow), 0.0, p_snow)
    p_rain = precip - p_snow
    melt_pot = f_DDSM * (t_air - 0)
    melt_pot = np.where(melt_pot < 0, 0.0, melt_pot)

    n = len(met_df)
    d_start = np.empty(n)
    d_end = np.empty(n)
    p_melt = np.empty(n)
    depth = float(D_snow_0)
    for i in range(n):
        d_start[i] = depth
        melt = min(melt_pot[i], depth)
        p_melt[i] = melt
        depth = depth + p_snow[i] - melt
        d_end[i] = depth

    met_df["P_snow"] = p_snow
    met_df["P_rain"] = p_rain
    met_df["P_melt"] = p_melt
    met_df["D_snow_start"] = d_start
    met_df["D_snow_end"] = d_end
    met_df["P"] = p_rain + p_melt
    return met_df


# --------------------------------------------------------------------------- PET (Thornthwaite 1948)
_MONTHDAYS = (31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31)
_LEAP_MONTHDAYS = (31, 29, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31)


def deg2rad(degrees):
    return degrees * (math.pi / 180.0)


def check_latitude_rad(latitude):
    lo, hi = deg2rad(-90.0), deg2rad(90.0)
    if not lo <= latitude <= hi:
        raise ValueError("latitude outside valid range {0!r} to {1!r} rad: {2!r}".format(lo, hi, latitude))


def check_doy(doy):
    if not 1 <= doy <= 366:
        raise ValueError("Day of the year (doy) must be in range 1-366: {0!r}".format(doy))


def check_sunset_hour_angle_rad(sha):
    lo, hi = 0.0, deg2rad(180)
    if not lo <= sha <= hi:
        raise ValueError("sunset hour angle outside valid range {0!r} to {1!r} rad: {2!r}".format(lo, hi, sha))


def check_sol_dec_rad(sd):
    lo, hi = deg2rad(-23.5), deg2rad(23.5)
    if not lo <= sd <= hi:
        raise ValueError("solar declination outside valid range {0!r} to {1!r} rad: {2!r}".format(lo, hi, sd))


def sol_dec(day_of_year):
    """Solar declination [rad], FAO-56 eq. 24 (reference ``inputs.py:370-379``)."""
    check_doy(day_of_year)
    return 0.409 * math.sin((2.0 * math.pi / 365.0) * day_of_year - 1.39)


def sunset_hour_angle(latitude, sol_dec):
    """Sunset hour angle [rad], FAO-56 eq. 25, arccos argument clipped to [-1, 1] (ref ``inputs.py:381-402``)."""
    check_latitude_rad(latitude)
    check_sol_dec_rad(sol_dec)
    cos_sha = -math.tan(latitude) * math.tan(sol_dec)
    return math.acos(min(max(cos_sha, -1.0), 1.0))


def daylight_hours(sha):
    """Daylight hours from the sunset hour angle, FAO-56 eq. 34 (ref ``inputs.py:404-414``)."""
    check_sunset_hour_angle_rad(sha)
    return (24.0 / math.pi) * sha


def monthly_mean_daylight_hours(latitude, year=None):
    """Mean daylight hours of each calendar month at ``latitude`` [rad] (ref ``inputs.py:416-445``)."""
    check_latitude_rad(latitude)
    month_days = _LEAP_MONTHDAYS if (year is not None and calendar.isleap(year)) else _MONTHDAYS
    out = []
    doy = 1
    for mdays in month_days:
        total = 0.0
        for _ in range(mdays):
            total += daylight_hours(sunset_hour_angle(latitude, sol_dec(doy)))
            doy += 1
        out.append(total / mdays)
    return out


def annual_thornthwaite(monthly_t, monthly_mean_dlh, year=None):
    """Monthly PET [mm/month] for one year by Thornthwaite (1948) (ref ``inputs.py:447-508``).

    ``PET = 1.6 (L/12)(N/30)(10 Ta / I)^a`` cm/month with heat index
    ``I = sum((Ta/5)^1.514)`` over months with ``Ta > 0`` and
    ``a = 6.75e-7 I^3 - 7.71e-5 I^2 + 1.792e-2 I + 0.49239``; negative monthly
    temperatures count as zero.
    """
    if len(monthly_t) != 12:
        raise ValueError("monthly_t should be length 12 but is length {0}.".format(len(monthly_t)))
    if len(monthly_mean_dlh) != 12:
        raise ValueError("monthly_mean_dlh should be length 12 but is length {0}.".format(len(monthly_mean_dlh)))
    month_days = _LEAP_MONTHDAYS if (year is not None and calendar.isleap(year)) else _MONTHDAYS

    adj = [t * (t >= 0) for t in monthly_t]
    heat = 0.0
    for ta in adj:
        if ta / 5.0 > 0.0:
            heat += (ta / 5.0) ** 1.514
    a = (6.75e-07 * heat ** 3) - (7.71e-05 * heat ** 2) + (1.792e-02 * heat) + 0.49239
    return [1.6 * (L / 12.0) * (N / 30.0) * ((10.0 * ta / heat) ** a) * 10.0
            for ta, L, N in zip(adj, monthly_mean_dlh, month_days)]


def _daylight_table_is_leap(years, strict_reference_quirks):
    """Per year: does Thornthwaite's day-length factor use the leap-year table?  The reference picks the table with
    ``monthly_mean_dlh = monthly_mean_dlh_leap`` in the leap branch and ``monthly_mean_dlh = monthly_mean_dlh`` in the
    other (``inputs.py:269-273``): the first leap year overwrites the variable the non-leap branch keeps, so every
    later year uses the leap-year table too (up to 0.5 % in March-December).  Reproduced under strict quirks."""
    out, seen = [], False
    for y in years:
        leap = calendar.isleap(int(y))
        seen = seen or leap
        out.append(seen if strict_reference_quirks else leap)
    return out


def daily_PET(latitude, met_df, strict_reference_quirks=True):
    """Daily PET [mm/day] from ``T_air`` by Thornthwaite, monthly values placed on the 16th of
    each month and linearly interpolated to days (reference ``inputs.py:232-312``; pinned to the unmodified
    reference wrapper by ``tests/golden/ref_pet_daily.npz``).  ``strict_reference_quirks=False`` picks the
    daylight-hours table by the calendar instead of the reference's sticky choice.
    """
    latitude = deg2rad(latitude)
    dlh_normal = monthly_mean_daylight_hours(latitude, year=1983)
    dlh_leap = monthly_mean_daylight_hours(latitude, year=1984)

    t_month = met_df["T_air"].groupby([met_df.index.year, met_df.index.month]).mean()
    pet_m = []
    years = sorted(set(met_df.index.year))
    for year, use_leap in zip(years, _daylight_table_is_leap(years, strict_reference_quirks)):
        vals = t_month.loc[year].to_numpy()
        if len(vals) < 12:
            raise ValueError("PET calc requires input met data for whole calendar years."
                             "Year {0!r} does not contain 12 months. Check input met data,"
                             "or change the start/end dates in the parameter file".format(year))
        dlh = dlh_leap if use_leap else dlh_normal
        pet_m.extend(annual_thornthwaite(vals, dlh, year=year))

    start = met_df.index[0].date()
    end = met_df.index[-1].date()
    idx_m = pd.date_range(start=start, end=end, freq="MS") + pd.DateOffset(days=15)
    pet_df = pd.DataFrame({"PET": pet_m}, index=idx_m)
    pet_df["PET"] = pet_df["PET"] / pet_df.index.daysinmonth

    if "PET" in met_df.columns:
        met_df = met_df.drop(columns=["PET"])
    met_df = met_df.join(pet_df)
    met_df["PET"] = met_df["PET"].interpolate(method="linear", limit=32, limit_direction="both")
    return met_df


def daily_PET_device(latitude, met_df, engine=None, strict_reference_quirks=True):
    """:func:`daily_PET` with the arithmetic on the GPU (``thornthwaite_kernel`` behind the C-ABI entry point
    ``simplyp_thornthwaite_pet_device``): same arguments, same returned frame, same ``ValueError`` for a record
    that is not made of whole calendar years (reference ``inputs.py:232-312``).  No CPU fallback."""
    from .engine import Engine

    if not -90.0 <= latitude <= 90.0:
        check_latitude_rad(deg2rad(latitude))
    idx = met_df.index
    years = sorted(set(idx.year))
    counts = met_df["T_air"].groupby([idx.year, idx.month]).size()
    for year in years:
        if len(counts.loc[year]) < 12:
            raise ValueError("PET calc requires input met data for whole calendar years."
                             "Year {0!r} does not contain 12 months. Check input met data,"
                             "or change the start/end dates in the parameter file".format(year))
    month_start = np.concatenate([[0], np.cumsum(counts.to_numpy())]).astype(np.int32)
    leap = np.array([(1 if calendar.isleap(y) else 0) | (2 if t else 0)
                     for y, t in zip(years, _daylight_table_is_leap(years, strict_reference_quirks))], dtype=np.int32)
    eng = engine or Engine()
    pet = eng.thornthwaite_pet(met_df["T_air"].to_numpy(dtype=np.float64, copy=True), month_start, leap, float(latitude))
    out = met_df.drop(columns=["PET"]) if "PET" in met_df.columns else met_df.copy()
    out["PET"] = pet.cpu().numpy()
    return out
