"""Generates the committed fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What it writes (all small, all derived — no reference source is copied):
  tarland_inputs.json   parameter workbook contents (Setup / Constant / LU / SC_reach / Reach_structure)
  tarland_met.npz       met series 1981-2010 (T_air, PET, Precipitation)
  tarland_obs.npz       observed Q and chemistry of reach 1 (whole record)
  shipped_golden.npz    the reference's shipped example outputs (Example_Data/Example_Output/*.csv)
  ref_tarland2004.npz   live reference runs of Tarland 2004: Dynamic_* n/y x (reference tolerance, tight)
  ref_gof.json          the reference's goodness_of_fit_stats on those runs
  kat.json              known answers of f_x / ode_f / discretized_soilP from the live reference
  ref_network.npz       live reference run of a 5-reach branching network with mixed NC land (tight tol)
  ref_ensemble.npz      live reference runs of a 12-member Latin-hypercube sample (tight tol) + fit statistics
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DATA = os.path.join(ROOT, "simplyp_b200", "data", "tarland")      # example data of the package (tarland.py)
sys.path.insert(0, ROOT)

from oracle import reference_live as rl          # noqa: E402
import simplyp_b200 as sp                         # noqa: E402
from simplyp_b200 import ensemble as ens          # noqa: E402

REF = "/root/reference"
CR = os.path.join(REF, "Current_Release", "v0-2A")
TIGHT = (1e-10, 1e-13)


def read_tarland(st="2004-01-01", end="2004-12-31"):
    cwd = os.getcwd()
    os.chdir(CR)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            out = sp.read_input_data("Parameters_v0-2A_Tarland.xlsx")
    finally:
        os.chdir(cwd)
    return out


def frame_to_dict(df):
    return {"columns": [str(c) for c in df.columns], "values": df.to_numpy(dtype=float)}


def main():
    p_SU, dyn, p, p_LU, p_SC, p_struc, met_df, obs_dict = read_tarland()

    # ---------------------------------------------------------------- inputs
    inputs = {
        "p_SU": {k: (v if isinstance(v, str) else (None if pd.isnull(v) else float(v))) for k, v in p_SU.items()},
        "p": {k: float(v) for k, v in p.items() if k != "SC_list"},
        "p_LU": {row: {c: (None if pd.isnull(p_LU.loc[row, c]) else float(p_LU.loc[row, c])) for c in p_LU.columns}
                 for row in p_LU.index},
        "p_SC": {str(c): {row: (None if pd.isnull(p_SC.loc[row, c]) else float(p_SC.loc[row, c])) for row in p_SC.index}
                 for c in p_SC.columns},
        "p_struc": {str(i): {"Upstream_SCs": None, "In_final_flux?": None} for i in p_struc.index},
    }
    with open(os.path.join(DATA, "tarland_inputs.json"), "w") as f:
        json.dump(inputs, f, indent=1)

    met_all = pd.read_csv(os.path.join(REF, "Example_Data", "Tarland_Scotland", "Tarland_MetData_1981-2010.csv"),
                          parse_dates=True, dayfirst=True, index_col=0)
    np.savez_compressed(os.path.join(DATA, "tarland_met.npz"),
                        day0=str(met_all.index[0].date()), n=len(met_all),
                        T_air=met_all["T_air"].to_numpy(float), PET=met_all["PET"].to_numpy(float),
                        Precipitation=met_all["Precipitation"].to_numpy(float))

    from simplyp_b200.inputs import _read_obs_workbook
    obs_dir = os.path.join(REF, "Example_Data", "Tarland_Scotland", "Observations")
    q = _read_obs_workbook(os.path.join(obs_dir, "Coull_DailyMeanQ.xlsx"), None, None)[1]
    chem = _read_obs_workbook(os.path.join(obs_dir, "Coull_ChemObs.xlsx"), None, None)[1]
    epoch = pd.Timestamp("1970-01-01")
    np.savez_compressed(os.path.join(DATA, "tarland_obs.npz"),
                        q_days=((q.index - epoch).days).to_numpy(), Q=q["Q"].to_numpy(float),
                        chem_days=((chem.index - epoch).days).to_numpy(),
                        **{c: chem[c].to_numpy(float) for c in chem.columns})

    # ---------------------------------------------------------------- shipped golden outputs
    eo = os.path.join(REF, "Example_Data", "Example_Output")
    tc = pd.read_csv(os.path.join(eo, "Results_TC_SC1.csv"), index_col=0, parse_dates=True)
    rr = pd.read_csv(os.path.join(eo, "Instream_results_Reach1.csv"), index_col=0, parse_dates=True)
    gof = pd.read_csv(os.path.join(eo, "GoF_stats.csv"), index_col=0)
    np.savez_compressed(os.path.join(HERE, "shipped_golden.npz"),
                        tc_cols=np.array(tc.columns, dtype=str), tc=tc.to_numpy(float),
                        r_cols=np.array(rr.columns, dtype=str), r=rr.to_numpy(float),
                        gof_index=np.array(gof.index, dtype=str), gof_cols=np.array(gof.columns, dtype=str),
                        gof=gof.to_numpy(float), first_day=str(tc.index[0].date()), n_days=len(tc))

    # ---------------------------------------------------------------- live reference, Tarland 2004
    runs = {}
    gofs = {}
    for dy in ("n", "y"):
        d2 = dyn.copy()
        d2["Dynamic_EPC0"] = dy
        d2["Dynamic_erodibility"] = dy
        for label, (rt, at) in (("reftol", (None, None)), ("tight", TIGHT)):
            TC, R, Kf, info = rl.run_simply_p(met_df, p_struc, p_SU, p_LU, p_SC, p, d2, rtol=rt, atol=at)
            key = "dyn%s_%s" % (dy, label)
            runs[key + "_tc"] = TC[1].to_numpy(float)
            runs[key + "_r"] = R[1].to_numpy(float)
            runs[key + "_tc_cols"] = np.array(TC[1].columns, dtype=str)
            runs[key + "_r_cols"] = np.array(R[1].columns, dtype=str)
            runs[key + "_Kf"] = Kf
            g = rl.goodness_of_fit_stats(p_SU, R, obs_dict)
            gofs[key] = {"index": list(g.index), "columns": list(g.columns),
                         "values": g.to_numpy(float).tolist()}
            print(key, "Kf", Kf, rl.solver_counters(reset=True))
    np.savez_compressed(os.path.join(HERE, "ref_tarland2004.npz"), **runs)
    with open(os.path.join(HERE, "ref_gof.json"), "w") as f:
        json.dump(gofs, f, indent=1)

    # ---------------------------------------------------------------- known answers (Appendix F of SURVEY.md)
    spm = rl.load()
    kat = {"f_x": [], "ode_f": None, "discretized_soilP": None}
    for x in (289.9, 290.0, 291.0, 291.45, 292.9, 293.0):
        kat["f_x"].append([x, 290.0, 0.01, float(spm.f_x(x, 290.0, 0.01))])
    kat["f_x"].append([0.5, 0.4, 0.01, float(spm.f_x(0.5, 0.4, 0.01))])
    kat["f_x"].append([0.402, 0.4, 0.01, float(spm.f_x(0.402, 0.4, 0.01))])
    Esus = pd.Series({"A": 582.295081967213, "S": 252.00000000000003, "IG": 432.0})
    T_s = pd.Series({"A": 2.0, "S": 10.0})
    y0 = [290.0, 290.0, 76.03868471953578, 0.31179540372581704, 1.6711798839458414, 0, 0, 0, 0, 0, 0, 0]
    ode_params = [1.48, 0.22, 0.015879897193062383, 0.0296, 0.0, Esus, 0.0, 0.0, 0.0,
                  0.5, 0.2, 0.3, 0.5, 0.0, 0.0, 0.0, 0.0, "None",
                  0.02, 1.0, 0.7, T_s, 65.0, 290.0, 10000.0, 51.7, 0.5, 0.42, 1500.0, 2.0,
                  5.170000000000001, 0.0, 2873227.5, 0.0, 4911500000.0, 0.1, 0.02, 1.6, 4287739.5, "y", 0.4]
    kat["ode_f"] = {"y0": y0, "dy": [float(v) for v in spm.ode_f(np.array(y0, dtype=float), 0.0, ode_params)]}
    y1 = [291.17, 291.2, 74.9, 0.2589, 1.213, 0.3, 115.4, 20.0, 0.3122, 0.2, 0.2327, 0.1]
    kat["ode_f2"] = {"y0": y1, "dy": [float(v) for v in spm.ode_f(np.array(y1, dtype=float), 0.0, ode_params)]}
    args = (10., 51.7, 1, 0.00011315280464216634, 95e6 * 51.7, 5.17, 0.21, 0.0296, 291.17, 1499.3, 2873362.5)
    kat["discretized_soilP"] = {"args": list(args), "out": [float(v) for v in spm.discretized_soilP(*args)]}
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    # ---------------------------------------------------------------- branching network with mixed NC land
    net = make_network_inputs(p, p_LU, p_SC, p_struc)
    p_n, p_LU_n, p_SC_n, p_struc_n = net
    met_n = met_df.iloc[:150]
    d2 = dyn.copy()
    d2["Dynamic_EPC0"] = "y"
    d2["Dynamic_erodibility"] = "y"
    p_SU_n = p_SU.copy()
    TC, R, Kf, info = rl.run_simply_p(met_n, p_struc_n, p_SU_n, p_LU_n, p_SC_n, p_n, d2, rtol=TIGHT[0], atol=TIGHT[1])
    out = {"Kf": Kf, "n_days": len(met_n)}
    for SC in p_n["SC_list"]:
        out["tc_%d" % SC] = TC[SC].to_numpy(float)
        out["r_%d" % SC] = R[SC].to_numpy(float)
        out["tc_cols_%d" % SC] = np.array(TC[SC].columns, dtype=str)
        out["r_cols_%d" % SC] = np.array(R[SC].columns, dtype=str)
    np.savez_compressed(os.path.join(HERE, "ref_network.npz"), **out)
    print("network done", rl.solver_counters(reset=True))

    # ---------------------------------------------------------------- 12-member ensemble (tight tolerance)
    M = 12
    samples = ens.latin_hypercube(M, seed=20260101)
    d2 = dyn.copy()
    d2["Dynamic_EPC0"] = "y"
    d2["Dynamic_erodibility"] = "y"
    cols = ["Q_cumecs", "SS_mgl", "TDP_mgl", "PP_mgl", "TP_mgl", "SRP_mgl", "Msus_kg/day", "TDP_kg/day", "PP_kg/day"]
    series = np.zeros((M, len(met_df), len(cols)))
    gof_rows = []
    for i in range(M):
        pi, pLUi, pSCi = ens.apply_member_to_pandas(samples, i, p, p_LU, p_SC)
        TC, R, Kf, info = rl.run_simply_p(met_df, p_struc, p_SU, pLUi, pSCi, pi, d2, rtol=TIGHT[0], atol=TIGHT[1])
        series[i] = R[1][cols].to_numpy(float)
        g = rl.goodness_of_fit_stats(p_SU, R, obs_dict)
        gof_rows.append(g.loc[["Q", "SS", "TDP", "PP", "TP", "SRP"], ["N obs", "NSE", "log NSE", "Spearmans r", "r2",
                                                                   "Bias (%)", "nRMSD (%)"]].to_numpy(float))
        print("member", i, rl.solver_counters(reset=True))
    np.savez_compressed(os.path.join(HERE, "ref_ensemble.npz"), series=series, cols=np.array(cols, dtype=str),
                        gof=np.array(gof_rows), gof_vars=np.array(["Q", "SS", "TDP", "PP", "TP", "SRP"], dtype=str),
                        sample_names=np.array(list(samples.keys()), dtype=str),
                        sample_values=np.array([samples[k] for k in samples]))


def make_network_inputs(p, p_LU, p_SC, p_struc):
    from tests.golden.networks import network5_inputs
    return network5_inputs(p, p_LU, p_SC, p_struc)


if __name__ == "__main__":
    main()
