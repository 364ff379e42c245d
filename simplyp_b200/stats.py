"""Goodness-of-fit table, same layout as the reference's ``goodness_of_fit_stats``
(``simplyP/visualise_results.py:387-474``).

For a single run this is post-processing of DataFrames that already live on the host, done with
numpy.  For ensembles the same statistics (all but Spearman's r) are reduced on the device by the
calibration kernel; ``stats_frame`` turns one member's device statistics into the same table.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

from . import packing as pk

STATS_COLUMNS = ["N obs", "NSE", "log NSE", "Spearmans r", "r2", "Bias (%)", "nRMSD (%)"]
SIM_COLUMN = {"Q": "Q_cumecs", "SS": "SS_mgl", "PP": "PP_mgl", "TP": "TP_mgl", "TDP": "TDP_mgl", "SRP": "SRP_mgl"}


def _rank(a):
    """Average ranks (ties share the mean rank), as pandas' ``corr(method='spearman')`` uses."""
    order = np.argsort(a, kind="mergesort")
    ranks = np.empty(len(a), dtype=float)
    sa = a[order]
    i = 0
    n = len(a)
    while i < n:
        j = i
        while j + 1 < n and sa[j + 1] == sa[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    return ranks


def gof_one(obs, sim):
    """Statistics of one variable; ``obs``/``sim`` are pandas Series indexed by date."""
    n_obs = int(obs.notnull().sum())                                   # :429 (before alignment)
    if n_obs <= 10:                                                     # :431
        return None
    tdf = pd.concat([obs, sim], axis=1).dropna(how="any")               # :435-436
    o = tdf.iloc[:, 0].to_numpy(dtype=float)
    s = tdf.iloc[:, 1].to_numpy(dtype=float)
    lo, ls = np.log(o), np.log(s)
    NSE = 1 - (np.sum((o - s) ** 2) / np.sum((o - np.mean(o)) ** 2))    # :441
    log_NSE = 1 - (np.sum((lo - ls) ** 2) / np.sum((lo - np.mean(lo)) ** 2))   # :442-443
    spearman = float(np.corrcoef(_rank(o), _rank(s))[0, 1])             # :444-445
    r2 = float(np.corrcoef(o, s)[0, 1] ** 2)                            # :446-447
    pbias = 100 * np.sum(s - o) / np.sum(o)                             # :448
    nrmsd = 100 * np.mean(np.abs(s - o)) / np.std(o)                    # :449
    return [n_obs, NSE, log_NSE, spearman, r2, pbias, nrmsd]


def goodness_of_fit_stats(p_SU, df_R_dict, obs_dict):
    """Tabulate (and optionally save) goodness-of-fit statistics per reach and variable."""
    if p_SU["run_mode"] != "scenario" and len(obs_dict) > 0:
        frames = []
        for SC in df_R_dict.keys():
            if SC not in obs_dict:
                continue
            rows, names = [], []
            for var in pk.VAR_KINDS:
                if var not in obs_dict[SC].columns:
                    continue
                res = gof_one(obs_dict[SC][var], df_R_dict[SC][SIM_COLUMN[var]])
                if res is None:
                    continue
                rows.append(res)
                names.append(var)
            df = pd.DataFrame(rows, columns=STATS_COLUMNS, index=names)
            df["Reach"] = SC
            frames.append(df)
        stats_df_allSC = pd.concat(frames)
        if p_SU.get("save_stats_csv", "n") == "y":
            stats_df_allSC.to_csv(os.path.join(p_SU["output_fpath"], "GoF_stats.csv"))
        return stats_df_allSC
    print("No observations read in, therefore cannot calculate model performance statistics")
    return None


def gof_table_from_device(member_stats, labels, n_obs=None):
    """One member's device statistics [V][NSTAT] -> the reference's goodness-of-fit table (same columns and
    index as ``goodness_of_fit_stats``, ``visualise_results.py:451-470``).  Needs ``rank_stats`` for Spearman's r.
    ``n_obs[(reach, var)]`` overrides 'N obs' (the reference reports the count BEFORE aligning with the run, :429)."""
    st = np.asarray(member_stats, dtype=float)
    frames = []
    for reach in dict.fromkeys(r for r, _v in labels):
        rows, names = [], []
        for k, (r, var) in enumerate(labels):
            if r != reach:
                continue
            n = st[k, 0] if n_obs is None else n_obs[(r, var)]
            rows.append([int(n), st[k, 1], st[k, 2], st[k, 8], st[k, 4], st[k, 5], st[k, 6]])
            names.append(var)
        df = pd.DataFrame(rows, columns=STATS_COLUMNS, index=names)
        df["Reach"] = reach
        frames.append(df)
    return pd.concat(frames) if frames else None


def stats_frame(member_stats, labels):
    """One member's device statistics [V][8] -> DataFrame indexed by (reach, variable)."""
    idx = pd.MultiIndex.from_tuples(labels, names=["Reach", "Variable"])
    return pd.DataFrame(np.asarray(member_stats), index=idx, columns=pk.STAT_NAMES)
