"""CPU tier, round 2: the oracle and the host-side functions pinned to outputs of the UNMODIFIED reference that round 1
left unpinned — run_mode 'val', step_len != 1, the Gaussian log-likelihood expression, sum_to_waterbody and the
daily_PET wrapper (fixtures by tests/golden/make_golden_r2.py) — plus the host build of the quad program on the same
fixtures."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import simplyp_oracle as orc
from tests import hostemu, parity
from tests.util import max_rel


def _modes(golden_dir, key):
    z = np.load(os.path.join(golden_dir, "ref_modes.npz"))
    tc = pd.DataFrame(z[key + "_tc"], columns=[str(c) for c in z[key + "_tc_cols"]])
    r = pd.DataFrame(z[key + "_r"], columns=[str(c) for c in z[key + "_r_cols"]])
    return z, tc, r


def _mode_inputs(golden_dir, key):
    from simplyp_b200 import tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    z, tc, r = _modes(golden_dir, key)
    step_len = 1.0
    if key == "val":
        p_SU["run_mode"] = "val"
        p["Kf"] = float(z["val_Kf"])
    else:
        step_len = 0.5
    return (p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs), step_len, tc, r


@pytest.mark.parametrize("key", ["val", "half"])
def test_oracle_run_modes_equal_the_reference(golden_dir, key):
    """run_mode='val' reads Kf from the parameter file (model.py:449-453); step_len=0.5 integrates every forcing record
    over half a day (model.py:193,640).  Oracle vs the unmodified reference, both LSODA at 1e-10: <= 1e-12."""
    (p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs), step_len, tc, r = _mode_inputs(golden_dir, key)
    n = 150
    TC, R, Kf, _ = orc.run_simply_p(met.iloc[:n], p_struc, p_SU, p_LU, p_SC, p, dyn, step_len=step_len, rtol=1e-10,
                                    atol=1e-13, mxstep=50000)
    assert list(R[1].columns) == list(r.columns) and list(TC[1].columns) == list(tc.columns)
    assert max_rel(R[1].to_numpy(float), r.to_numpy()[:n]) < 1e-12
    for c in tc.columns:
        assert max_rel(TC[1][c].to_numpy(float), tc[c].to_numpy()[:n]) < 1e-12, c
    if key == "val":
        assert Kf == p["Kf"] != 0.00011315280464216634
        # the fixture really differs from the 'cal' run (another Kf moves the soil-P columns)
        zc = np.load(os.path.join(golden_dir, "ref_tarland2004.npz"))
        cal = pd.DataFrame(zc["dyny_tight_tc"], columns=[str(c) for c in zc["dyny_tight_tc_cols"]])
        assert max_rel(cal["P_labile_A_kg"].to_numpy(), tc["P_labile_A_kg"].to_numpy()) > 1e-4


def check_mode(runner, golden_dir, key):
    from simplyp_b200 import model as spm, packing as pk
    inputs, step_len, tc, r = _mode_inputs(golden_dir, key)
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = inputs
    nc_types = pk.validate_land_use(p_SC, p["SC_list"])
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo, step_len)
    assert opt.run_mode_cal == (0 if key == "val" else 1) and opt.step_len == step_len
    out, diag = runner(pk.forcing_matrix(met), pk.member_vector(p, p_LU)[None], pk.sc_matrix(p_SC, topo.sc_ids)[None],
                       topo.parent_offsets, topo.parent_ids, opt)
    TC, R = parity.frames_from_raw(out[0], topo, p_SC, p, nc_types, met)
    tc.index = r.index = met.index
    parity.assert_frames_close(TC[1], R[1], tc, r, "mode %s" % key)
    assert int(diag[0, 0, 3]) == 0


@pytest.mark.parametrize("key", ["val", "half"])
def test_quad_program_run_modes(golden_dir, key):
    check_mode(hostemu.run_quad, golden_dir, key)


def test_log_likelihood_is_the_reference_expression():
    """The only definition of the Gaussian log-likelihood in the reference repository is
    ``Development/2016/MCMC.ipynb:233-242``: ``sigma_e = m*sim; ll = sum(norm(sim, sigma_e).logpdf(obs))`` with NaN
    -> -inf.  The oracle's closed form must BE that expression (scipy's norm), not merely resemble it."""
    from scipy.stats import norm
    rng = np.random.default_rng(2)
    for m in (0.05, 0.3, 1.0):
        sim = rng.lognormal(0.0, 1.0, 400)
        obs = sim * np.exp(rng.normal(0, 0.4, 400))
        obs[rng.random(400) < 0.3] = np.nan            # days without an observation are skipped, as in the notebook's join
        ok = ~np.isnan(obs)
        want = float(np.sum(norm(sim[ok], m * sim[ok]).logpdf(obs[ok])))
        got = orc.gaussian_log_likelihood(obs, sim, m)
        assert got == pytest.approx(want, rel=1e-13)
    sim[5] = -1.0                                       # a negative simulated value: scale < 0 -> NaN -> -inf
    obs[5] = 1.0
    assert orc.gaussian_log_likelihood(obs, sim, 0.3) == -np.inf


def test_sum_to_waterbody_equals_the_reference(golden_dir):
    """Host sum_to_waterbody against the reference's own (model.py:851-900) on the reference's own reach frames:
    same columns in the same order, identical values; one flagged reach -> None (model.py:898-900)."""
    import simplyp_b200 as sp
    z = np.load(os.path.join(golden_dir, "ref_waterbody.npz"))
    idx = pd.date_range("2004-01-01", periods=z["inputs"].shape[1])
    cols = [str(c) for c in z["cols_in"]]
    R = {int(r): pd.DataFrame(z["inputs"][k], columns=cols, index=idx) for k, r in enumerate(z["reaches"])}
    R[1] = R[2] = R[3] * 0.5                            # unflagged reaches must not count
    ps = pd.DataFrame({"Upstream_SCs": [np.nan] * 5, "In_final_flux?": [np.nan, np.nan, 1.0, 1.0, 1.0]},
                      index=[1, 2, 3, 4, 5])
    got = sp.sum_to_waterbody(ps, 5, R, float(z["f_TDP"]))
    assert [str(c) for c in got.columns] == [str(c) for c in z["wb_cols"]]
    assert np.array_equal(got.to_numpy(float), z["wb"])
    ps1 = ps.copy()
    ps1["In_final_flux?"] = [np.nan] * 4 + [1.0]
    assert sp.sum_to_waterbody(ps1, 5, R, 0.7) is None


def test_daily_pet_equals_the_reference_wrapper(golden_dir):
    """Host daily_PET against the reference's own wrapper (inputs.py:232-312: monthly resampling, the 16th-of-month
    placement, the two-sided interpolation) over 2001-2007.  2004 is a leap year: the reference keeps the leap-year
    daylight table for 2005-2007 as well (inputs.py:269-273 overwrites the table the non-leap branch keeps); with
    strict quirks the host function reproduces that, without them it differs there and only there."""
    import simplyp_b200 as sp
    from simplyp_b200 import tarland
    z = np.load(os.path.join(tarland.DATA_DIR, "tarland_met.npz"))
    idx = pd.date_range(str(z["day0"]), periods=int(z["n"]), freq="D")
    t_air = pd.DataFrame({"T_air": z["T_air"].astype(float)}, index=idx)
    ref = np.load(os.path.join(golden_dir, "ref_pet_daily.npz"))
    lat = float(ref["latitude"])
    for a, b in (("2001", "2007"), ("1981", "1983")):
        got = sp.daily_PET(lat, t_air[a:b])
        assert list(got.columns) == ["T_air", "PET"]
        assert max_rel(got["PET"].to_numpy(), ref["pet_%s_%s" % (a, b)]) < 1e-13, (a, b)
    cal = sp.daily_PET(lat, t_air["2001":"2007"], strict_reference_quirks=False)
    rel = np.abs(cal["PET"].to_numpy() / ref["pet_2001_2007"] - 1.0)
    years = cal.index.year.to_numpy()
    assert rel[years <= 2004][:-20].max() < 1e-13            # (the last days of 2004 interpolate towards January 2005)
    assert 1e-4 < rel[years >= 2005].max() < 1e-2
    both = t_air["2003":"2004"].copy()
    both["PET"] = 1.0
    got = sp.daily_PET(lat, both)
    assert [str(c) for c in got.columns] == [str(c) for c in ref["replaced_cols"]]
    assert max_rel(got["PET"].to_numpy(), ref["pet_2003_2004_replaced"]) < 1e-13
