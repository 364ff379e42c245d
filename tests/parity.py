"""Parity checks shared by the CPU tier (host build of the kernel arithmetic) and the GPU tier (C-ABI).

`runner(forcing, member_params, sc_params, parent_offsets, parent_ids, opt) -> (out, diag)` is either
tests.hostemu.run or simplyp_b200._cabi.run_host."""
import os

import numpy as np
import pandas as pd

from simplyp_b200 import ensemble as ens
from simplyp_b200 import model as spm
from simplyp_b200 import packing as pk
from tests.util import FLOW_CONC_COLS, max_mixed, max_rel

# Tolerance pair the product defaults to; the 1e-5 parity bound of north_star is asserted at it.
RTOL, ATOL = spm.DEFAULT_RTOL, spm.DEFAULT_ATOL
PARITY = 1e-5


def frames_from_raw(out_m, topo, p_SC, p, nc_types, met, with_snow=True):
    TC, R = {}, {}
    for i, SC in enumerate(topo.sc_ids):
        tc, r = spm.raw_to_frames(out_m[i], met.index, float(p_SC.loc["A_catch", SC]), p["Msoil_m2"], p["f_TDP"],
                                  nc_types[SC], met["D_snow_end"] if with_snow else None)
        TC[SC], R[SC] = tc, r
    return TC, R


def run_single(runner, inputs, rtol=RTOL, atol=ATOL, n_days=None):
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = inputs
    if n_days:
        met = met.iloc[:n_days]
    nc_types = pk.validate_land_use(p_SC, p["SC_list"])
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo, 1.0, rtol, atol)
    out, diag = runner(pk.forcing_matrix(met), pk.member_vector(p, p_LU)[None], pk.sc_matrix(p_SC, topo.sc_ids)[None],
                       topo.parent_offsets, topo.parent_ids, opt)
    TC, R = frames_from_raw(out[0], topo, p_SC, p, nc_types, met)
    return TC, R, diag, met


def assert_frames_close(TC, R, want_tc, want_r, label=""):
    """Flows and concentrations within PARITY relative; everything else within a mixed abs/rel bound."""
    assert list(R.columns) == list(want_r.columns), label
    assert list(TC.columns) == list(want_tc.columns), label
    for c in R.columns:
        if c in FLOW_CONC_COLS:
            assert max_rel(R[c].to_numpy(), want_r[c].to_numpy()) <= PARITY, (label, c)
        else:
            assert max_mixed(R[c].to_numpy(), want_r[c].to_numpy(), PARITY, 1e-3) <= 1.0, (label, c)
    for c in TC.columns:
        assert max_mixed(TC[c].to_numpy(), want_tc[c].to_numpy(), PARITY, 1e-3) <= 1.0, (label, c)


def check_tarland(runner, golden_dir, dy, n_days=None):
    from simplyp_b200 import tarland
    inputs = tarland.load(dynamic=dy)
    TC, R, diag, met = run_single(runner, inputs, n_days=n_days)
    z = np.load(os.path.join(golden_dir, "ref_tarland2004.npz"))
    key = "dyn%s_tight" % dy
    n = len(met)
    want_tc = pd.DataFrame(z[key + "_tc"][:n], columns=list(z[key + "_tc_cols"]), index=met.index)
    want_r = pd.DataFrame(z[key + "_r"][:n], columns=list(z[key + "_r_cols"]), index=met.index)
    assert_frames_close(TC[1], R[1], want_tc, want_r, "tarland dyn=%s" % dy)
    assert int(diag[0, 0, 3]) == 0
    return TC, R, diag


def check_network(runner, golden_dir, n_days=150):
    from simplyp_b200 import tarland
    from tests.golden.networks import network5_inputs
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    p, p_LU, p_SC, p_struc = network5_inputs(p, p_LU, p_SC, p_struc)
    TC, R, diag, met = run_single(runner, (p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs), n_days=n_days)
    z = np.load(os.path.join(golden_dir, "ref_network.npz"))
    for SC in (1, 2, 3, 4, 5):
        want_tc = pd.DataFrame(z["tc_%d" % SC][:n_days], columns=list(z["tc_cols_%d" % SC]), index=met.index)
        want_r = pd.DataFrame(z["r_%d" % SC][:n_days], columns=list(z["r_cols_%d" % SC]), index=met.index)
        assert_frames_close(TC[SC], R[SC], want_tc, want_r, "network SC %d" % SC)
    assert not np.any(diag[..., 3])


def ensemble_fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "ref_ensemble.npz"))
    samples = {str(k): z["sample_values"][i] for i, k in enumerate(z["sample_names"])}
    return z, samples


def check_ensemble_series(runner, golden_dir, members=None):
    """Full-output run of the Latin-hypercube fixture members vs the live-reference series."""
    from simplyp_b200 import tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    z, samples = ensemble_fixture(golden_dir)
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo, 1.0, RTOL, ATOL)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    idx = list(range(member.shape[0])) if members is None else list(members)
    out, diag = runner(pk.forcing_matrix(met), member[idx], sc[idx], topo.parent_offsets, topo.parent_ids, opt)
    cols = [str(c) for c in z["cols"]]
    worst = 0.0
    for j, i in enumerate(idx):
        _tc, r = spm.raw_to_frames(out[j, 0], met.index, float(sc[i, 0, pk.SC_INDEX["A_catch"]]), p["Msoil_m2"],
                                   p["f_TDP"], "None", None)
        for k, c in enumerate(cols):
            e = max_rel(r[c].to_numpy(), z["series"][i, :, k])
            worst = max(worst, e)
            assert e <= PARITY, (i, c, e)
    assert not np.any(diag[..., 3])
    return worst


def check_stiff_chain(runner, area, n_days=45, max_steps_per_day=None):
    """A stiff outlet reach (tests/golden/stiff_chain.py) against the oracle: LSODA at rtol=1e-10, which integrates
    it with BDF.  Returns the outlet reach's step attempts per day."""
    from oracle import simplyp_oracle as orc
    from simplyp_b200 import tarland
    from tests.golden.stiff_chain import stiff_chain_inputs
    p_SU, dyn, p, p_LU, p_SC0, p_struc0, met, obs = tarland.load(dynamic="y")
    p, p_SC, p_struc = stiff_chain_inputs(p, p_SC0[1], area)
    met = met.iloc[:n_days]
    TC, R, diag, met = run_single(runner, (p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs))
    TCo, Ro, _Kf, _ = orc.run_simply_p(met, p_struc, p_SU, p_LU.copy(), p_SC.copy(), p, dyn, rtol=1e-10, atol=1e-13,
                                       mxstep=500000)
    for SC in (1, 2, 3):
        assert_frames_close(TC[SC], R[SC], TCo[SC], Ro[SC], "stiff chain SC %d (area %g)" % (SC, area))
    assert not np.any(diag[..., 3])
    per_day = diag[0, 2, 0] / float(n_days)
    if max_steps_per_day is not None:
        assert per_day <= max_steps_per_day, per_day
    return per_day


def check_shipped_golden(runner, golden_dir):
    """The reference's own shipped example output (Example_Data/Example_Output/*.csv: 2004, both dynamic options on,
    made with LSODA rtol=0.01 and an older SciPy) — its only golden vectors.  They pin a result only to the
    reference's solver noise (1e-3..3.4e-3 in-stream between SciPy versions); the converged solution
    computed here must lie within that noise of them, and much closer on the slow / exact-algebra columns."""
    from simplyp_b200 import tarland
    inputs = tarland.load(dynamic="y")
    TC, R, diag, met = run_single(runner, inputs)
    z = np.load(os.path.join(golden_dir, "shipped_golden.npz"), allow_pickle=True)
    r = pd.DataFrame(z["r"], columns=[str(c) for c in z["r_cols"]])
    tc = pd.DataFrame(z["tc"], columns=[str(c) for c in z["tc_cols"]])
    for c in r.columns:
        assert max_rel(R[1][c].to_numpy(), r[c].to_numpy()) < 1e-2, c
    for c, tol in (("D_snow", 1e-12), ("C_cover_A", 1e-12), ("Qq", 1e-12), ("P_labile_A_kg", 1e-6),
                   ("EPC0_A_mgl", 1e-6), ("TDPs_A_mgl", 1e-6), ("Vg", 1e-3), ("VsA", 1e-3), ("VsS", 1e-3)):
        assert max_rel(TC[1][c].to_numpy(), tc[c].to_numpy()) < tol, c
