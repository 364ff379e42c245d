"""Round-2 fixtures from the UNMODIFIED reference (build container only; needs /root/reference):

    python tests/golden/make_golden_r2.py

  ref_modes.npz       Tarland 2004, both dynamic options on, odeint forced to rtol=1e-10/atol=1e-13:
                      (a) run_mode='val' — Kf read from p['Kf'] (model.py:449-453), here 1.9e-4, not the 'cal' value;
                      (b) step_len=0.5 — every forcing record integrated over half a day (model.py:193,640)
  ref_waterbody.npz   the reference's sum_to_waterbody (model.py:851-900) on its own run of the 5-reach network with
                      reaches 3, 4, 5 flagged In_final_flux? == 1: its inputs (the 4 summed columns of the 3 reaches)
                      and its output frame
  ref_csv.json        what the reference writes with save_output_csvs == 'y' (model.py:815-825): file names, header
                      lines, row counts, first index cell
  ref_pet_daily.npz   the reference's daily_PET wrapper (inputs.py:232-312) on Tarland air temperature 2001-2007
                      (leap year 2004 in the middle: the years after it keep the leap-year daylight table) and
                      1981-1983 (no leap year), latitude 57.1
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_live as rl          # noqa: E402
from simplyp_b200 import tarland                  # noqa: E402
from tests.golden.networks import network5_inputs  # noqa: E402

TIGHT = dict(rtol=1e-10, atol=1e-13)
VAL_KF = 1.9e-4


def frames(prefix, TC, R):
    return {prefix + "_tc": TC.to_numpy(float), prefix + "_tc_cols": np.array([str(c) for c in TC.columns]),
            prefix + "_r": R.to_numpy(float), prefix + "_r_cols": np.array([str(c) for c in R.columns])}


def main():
    out = {}
    # ---------------------------------------------------------------- run modes
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    p_SU_val = p_SU.copy()
    p_SU_val["run_mode"] = "val"
    p_val = p.copy()
    p_val["Kf"] = VAL_KF
    TC, R, Kf, _ = rl.run_simply_p(met, p_struc, p_SU_val, p_LU, p_SC, p_val, dyn, **TIGHT)
    assert Kf == VAL_KF
    out.update(frames("val", TC[1], R[1]))
    out["val_Kf"] = VAL_KF
    TC, R, Kf, _ = rl.run_simply_p(met, p_struc, p_SU, p_LU, p_SC, p, dyn, step_len=0.5, **TIGHT)
    out.update(frames("half", TC[1], R[1]))
    np.savez_compressed(os.path.join(HERE, "ref_modes.npz"), **out)
    print("ref_modes.npz written")

    # ---------------------------------------------------------------- sum_to_waterbody
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    p, p_LU, p_SC, p_struc = network5_inputs(p, p_LU, p_SC, p_struc)
    met = met.iloc[:150]
    TC, R, Kf, _ = rl.run_simply_p(met, p_struc, p_SU, p_LU, p_SC, p, dyn, **TIGHT)
    p_struc = p_struc.copy()
    p_struc["In_final_flux?"] = [np.nan, np.nan, 1.0, 1.0, 1.0]
    sp = rl.load()
    with contextlib.redirect_stdout(io.StringIO()):
        wb = sp.sum_to_waterbody(p_struc, 5, R, p["f_TDP"])
        none = sp.sum_to_waterbody(p_struc.assign(**{"In_final_flux?": [np.nan] * 4 + [1.0]}), 5, R, p["f_TDP"])
    assert none is None
    cols_in = ["Q_cumecs", "Msus_kg/day", "TDP_kg/day", "PP_kg/day"]
    np.savez_compressed(os.path.join(HERE, "ref_waterbody.npz"),
                        reaches=np.array([3, 4, 5]), cols_in=np.array(cols_in),
                        inputs=np.stack([R[r][cols_in].to_numpy(float) for r in (3, 4, 5)]),
                        wb=wb.to_numpy(float), wb_cols=np.array([str(c) for c in wb.columns]), f_TDP=float(p["f_TDP"]))
    print("ref_waterbody.npz written", list(wb.columns))

    # ---------------------------------------------------------------- CSV writer
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    p, p_LU, p_SC, p_struc = network5_inputs(p, p_LU, p_SC, p_struc)
    met = met.iloc[:20]
    rec = {}
    with tempfile.TemporaryDirectory() as tmp:
        p_SU = p_SU.copy()
        p_SU["save_output_csvs"] = "y"
        p_SU["output_fpath"] = tmp
        rl.run_simply_p(met, p_struc, p_SU, p_LU, p_SC, p, dyn)
        for name in sorted(os.listdir(tmp)):
            with open(os.path.join(tmp, name)) as f:
                lines = f.read().splitlines()
            rec[name] = {"header": lines[0], "n_rows": len(lines) - 1, "first_index": lines[1].split(",")[0],
                         "last_index": lines[-1].split(",")[0]}
    with open(os.path.join(HERE, "ref_csv.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print("ref_csv.json written", sorted(rec))

    # ---------------------------------------------------------------- daily_PET wrapper
    z = np.load(os.path.join(tarland.DATA_DIR, "tarland_met.npz"))
    idx = pd.date_range(str(z["day0"]), periods=int(z["n"]), freq="D")
    t_air = pd.DataFrame({"T_air": z["T_air"].astype(float)}, index=idx)
    pet = {}
    for a, b in (("2001", "2007"), ("1981", "1983")):
        got = rl.daily_PET(57.1, t_air[a:b])
        pet["pet_%s_%s" % (a, b)] = got["PET"].to_numpy(float)
        assert list(got.columns) == ["T_air", "PET"] and got.index.equals(t_air[a:b].index)
    # a frame that already carries a PET column: the reference replaces it (inputs.py:300-304)
    both = t_air["2003":"2004"].copy()
    both["PET"] = 1.0
    got = rl.daily_PET(57.1, both)
    pet["pet_2003_2004_replaced"] = got["PET"].to_numpy(float)
    pet["replaced_cols"] = np.array([str(c) for c in got.columns])
    np.savez_compressed(os.path.join(HERE, "ref_pet_daily.npz"), latitude=57.1, **pet)
    print("ref_pet_daily.npz written")


if __name__ == "__main__":
    main()
