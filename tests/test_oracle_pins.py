"""Pins oracle/simplyp_oracle.py (the CPU restatement) against the reference's own vectors:
shipped golden CSVs, known-answer values and whole runs of the unmodified reference (fixtures made by
tests/golden/make_golden.py).  CPU tier."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import simplyp_oracle as orc
from tests.util import max_rel


def _z(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_known_answers(golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "kat.json")))
    for x, thr, reld, want in kat["f_x"]:
        assert orc.f_x(x, thr, reld) == pytest.approx(want, rel=0, abs=1e-15)
    q = (1.48, 0.22, 0.015879897193062383, 0.0296, 0.0, 582.295081967213, 252.00000000000003, 432.0, 0.0, 0.0, 0.0,
         0.5, 0.2, 0.3, 0.5, 0.0, 0.0, 0.0, 0.0, False,
         0.02, 1.0, 0.7, 2.0, 10.0, 65.0, 290.0, 10000.0, 51.7, 0.5, 0.42, 2.0,
         5.170000000000001, 0.0, 2873227.5, 0.0, 4911500000.0, 0.1, 0.02, 1.6, 4287739.5, 0.4)
    for key in ("ode_f", "ode_f2"):
        got = orc.ode_f(np.array(kat[key]["y0"], dtype=float), 0.0, q)
        assert np.array_equal(np.array(got, dtype=float), np.array(kat[key]["dy"]))
    got = orc.discretized_soilP(*kat["discretized_soilP"]["args"])
    assert tuple(float(g) for g in got) == tuple(kat["discretized_soilP"]["out"])


@pytest.mark.parametrize("dy", ["n", "y"])
def test_oracle_equals_reference_at_reference_tolerance(golden_dir, dy):
    """Same LSODA, same arithmetic order -> bit-identical to the unmodified reference at rtol=0.01."""
    from simplyp_b200 import tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic=dy)
    TC, R, Kf, info = orc.run_simply_p(met, p_struc, p_SU, p_LU, p_SC, p, dyn, rtol=0.01)
    z = _z(golden_dir, "ref_tarland2004.npz")
    key = "dyn%s_reftol" % dy
    assert list(TC[1].columns) == list(z[key + "_tc_cols"])
    assert list(R[1].columns) == list(z[key + "_r_cols"])
    assert np.array_equal(TC[1].to_numpy(float), z[key + "_tc"])
    assert np.array_equal(R[1].to_numpy(float), z[key + "_r"])
    assert Kf == float(z[key + "_Kf"]) == 0.00011315280464216634


def test_oracle_vs_shipped_golden_csvs(golden_dir):
    """The reference's shipped example output (made with Dynamic_*='y' and an older SciPy): agreement to the
    reference's own solver tolerance in-stream, much tighter for the slow soil-P and exact-algebra columns."""
    from simplyp_b200 import tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    TC, R, Kf, info = orc.run_simply_p(met, p_struc, p_SU, p_LU, p_SC, p, dyn, rtol=0.01)
    z = _z(golden_dir, "shipped_golden.npz")
    r = pd.DataFrame(z["r"], columns=list(z["r_cols"]))
    tc = pd.DataFrame(z["tc"], columns=list(z["tc_cols"]))
    for c in r.columns:
        assert max_rel(R[1][c].to_numpy(), r[c].to_numpy()) < 5e-3, c
    for c, tol in (("D_snow", 1e-12), ("C_cover_A", 1e-12), ("Qq", 1e-12), ("P_labile_A_kg", 1e-6),
                   ("EPC0_A_mgl", 1e-6), ("TDPs_A_mgl", 1e-6), ("Vg", 1e-3), ("VsA", 1e-3), ("VsS", 1e-3)):
        assert max_rel(TC[1][c].to_numpy(), tc[c].to_numpy()) < tol, c
    assert R[1]["Q_cumecs"].iloc[0] == pytest.approx(0.750563676990669, rel=2e-3)


@pytest.mark.parametrize("dy", ["n", "y"])
def test_oracle_equals_reference_at_tight_tolerance(golden_dir, dy):
    from simplyp_b200 import tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic=dy)
    met = met.iloc[:120]
    TC, R, Kf, info = orc.run_simply_p(met, p_struc, p_SU, p_LU, p_SC, p, dyn, rtol=1e-10, atol=1e-13, mxstep=50000)
    z = _z(golden_dir, "ref_tarland2004.npz")
    key = "dyn%s_tight" % dy
    assert max_rel(R[1].to_numpy(float), z[key + "_r"][:120]) < 1e-12
    want_tc = pd.DataFrame(z[key + "_tc"][:120], columns=list(z[key + "_tc_cols"]))
    for c in TC[1].columns:
        assert max_rel(TC[1][c].to_numpy(float), want_tc[c].to_numpy()) < 1e-12, c


def test_oracle_network_equals_reference(golden_dir):
    """5-reach branching network with mixed newly-converted land (exercises routing, the area scaling of
    upstream flow, TDPeff blank -> 0 and the leaked NC_type quirk)."""
    from tests.golden.networks import network5_inputs
    from simplyp_b200 import tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    p, p_LU, p_SC, p_struc = network5_inputs(p, p_LU, p_SC, p_struc)
    z = _z(golden_dir, "ref_network.npz")
    n = 60
    TC, R, Kf, info = orc.run_simply_p(met.iloc[:n], p_struc, p_SU, p_LU, p_SC, p, dyn, rtol=1e-10, atol=1e-13,
                                       mxstep=50000)
    for SC in (1, 2, 3, 4, 5):
        assert list(TC[SC].columns) == list(z["tc_cols_%d" % SC]), SC
        assert max_rel(R[SC].to_numpy(float), z["r_%d" % SC][:n]) < 1e-11, SC
        assert max_rel(TC[SC].to_numpy(float), z["tc_%d" % SC][:n]) < 1e-11, SC
    assert Kf == float(z["Kf"])


def test_oracle_gof_equals_reference(golden_dir):
    from simplyp_b200 import tarland
    gof = json.load(open(os.path.join(golden_dir, "ref_gof.json")))["dyny_tight"]
    z = _z(golden_dir, "ref_tarland2004.npz")
    R = pd.DataFrame(z["dyny_tight_r"], columns=list(z["dyny_tight_r_cols"]))
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    R.index = met.index
    simcol = {"Q": "Q_cumecs", "SS": "SS_mgl", "TDP": "TDP_mgl", "PP": "PP_mgl", "TP": "TP_mgl", "SRP": "SRP_mgl"}
    for var, row in zip(gof["index"], gof["values"]):
        o = obs[1][var].reindex(met.index).to_numpy(float)
        st = orc.gof_stats(o, R[simcol[var]].to_numpy(float))
        got = [st["n"], st["NSE"], st["log_NSE"], st["spearman_r"], st["r2"], st["pbias"], st["nRMSD"]]
        assert np.allclose(got, row[:7], rtol=1e-10, atol=1e-12), var


def test_snow_restatement_matches_shipped_golden(golden_dir):
    from simplyp_b200 import tarland
    met = tarland.load_met(inc_snowmelt=False)
    P, D_end, melt = orc.snow_hydrol_inputs(0.0, 2.74, met["P"].to_numpy(), met["T_air"].to_numpy())
    z = _z(golden_dir, "shipped_golden.npz")
    tc = pd.DataFrame(z["tc"], columns=list(z["tc_cols"]))
    assert np.allclose(D_end, tc["D_snow"].to_numpy(), rtol=0, atol=1e-12)
    assert np.allclose(0.02 * P, tc["Qq"].to_numpy(), rtol=1e-13, atol=0)
    assert abs(P.sum() - 953.26) < 0.01 and abs(D_end.max() - 42.74) < 1e-9


def test_thornthwaite_helpers_equal_the_reference(golden_dir):
    """Host Thornthwaite helpers (simplyp_b200/inputs.py) against values produced by the UNMODIFIED reference helpers
    (inputs.py:315-508; fixture by tests/golden/make_pet_golden.py): bit-identical."""
    import calendar
    import json
    from simplyp_b200 import inputs
    with open(os.path.join(golden_dir, "ref_pet.json")) as f:
        ref = json.load(f)
    lat = inputs.deg2rad(ref["latitude_deg"])
    assert inputs.monthly_mean_daylight_hours(lat, 1983) == ref["dlh_normal"]
    assert inputs.monthly_mean_daylight_hours(lat, 1984) == ref["dlh_leap"]
    for year, rec in ref["years"].items():
        dlh = ref["dlh_leap"] if calendar.isleap(int(year)) else ref["dlh_normal"]
        assert inputs.annual_thornthwaite(rec["monthly_t"], dlh, year=int(year)) == rec["pet_mm_month"], year
