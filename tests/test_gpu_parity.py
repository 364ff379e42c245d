"""GPU tier: the CUDA path, called through the C-ABI (ctypes), against the oracle and the live-reference
fixtures.  Parity bound: relative error <= 1e-5 on daily flows and concentrations at the product's default
tolerances (rtol=1e-7, atol=1e-10) versus the reference's odeint path run at tight tolerance
(rtol=1e-10, atol=1e-13); see tests/parity.py."""
import json
import os
import sys

import numpy as np
import pandas as pd
import pytest

from tests import parity
from tests.util import max_rel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cabi():
    from simplyp_b200 import _cabi
    _cabi.require_device()          # fails loudly if the extension or the GPU is missing: no fallback
    return _cabi


@pytest.mark.parametrize("dy", ["n", "y"])
def test_tarland_2004_vs_reference(cabi, golden_dir, dy):
    parity.check_tarland(cabi.run_host, golden_dir, dy)


def test_branching_network_vs_reference(cabi, golden_dir):
    parity.check_network(cabi.run_host, golden_dir)


def test_ensemble_series_vs_reference(cabi, golden_dir):
    worst = parity.check_ensemble_series(cabi.run_host, golden_dir)
    assert worst <= parity.PARITY


def test_drop_in_run_simply_p(cabi, golden_dir):
    """The reference-facing entry point: same return tuple, column layout, mutations, Kf."""
    import simplyp_b200 as sp
    from simplyp_b200 import tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    TC, R, Kf, info = sp.run_simply_p(met, p_struc, p_SU, p_LU, p_SC, p, dyn, verbose=False)
    z = np.load(os.path.join(golden_dir, "ref_tarland2004.npz"))
    assert list(TC[1].columns) == list(z["dyny_tight_tc_cols"]) and len(TC[1].columns) == 20
    assert list(R[1].columns) == list(z["dyny_tight_r_cols"]) and len(R[1].columns) == 17
    assert Kf == 0.00011315280464216634
    assert {"EPC0_0", "Plab0", "TDPs0"} <= set(p_LU.index) and {"f_A", "f_NC_A", "NC_type"} <= set(p_SC.index)
    assert p_SC.loc["NC_type", 1] == "None"
    assert info["nfe"] > 0 and info["message"].startswith("Integration successful")
    want = pd.DataFrame(z["dyny_tight_r"], columns=list(z["dyny_tight_r_cols"]))
    for c in ("Q_cumecs", "SS_mgl", "TDP_mgl", "PP_mgl", "TP_mgl", "SRP_mgl"):
        assert max_rel(R[1][c].to_numpy(), want[c].to_numpy()) <= 1e-5, c
    # goodness-of-fit table of the reference on the reference run vs ours on our run
    gof = json.load(open(os.path.join(golden_dir, "ref_gof.json")))["dyny_tight"]
    got = sp.goodness_of_fit_stats(p_SU, R, obs)
    assert list(got.index) == gof["index"]
    assert np.allclose(got.to_numpy(float)[:, :7], np.array(gof["values"])[:, :7], rtol=2e-5, atol=2e-6)


def test_fused_statistics_vs_reference(cabi, golden_dir):
    """Calibration mode: on-device NSE / log-NSE / r2 / bias / nRMSD vs the reference's goodness_of_fit_stats
    on its own (tight-tolerance) runs of the same Latin-hypercube members; log-likelihood vs the oracle."""
    from oracle import simplyp_oracle as orc
    from simplyp_b200 import ensemble as ens, model as spm, packing as pk, tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    z, samples = parity.ensemble_fixture(golden_dir)
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    obs_m, desc, labels = pk.obs_arrays(obs, topo, met.index, None)
    assert [l[1] for l in labels] == ["Q", "SS", "TDP", "PP", "TP", "SRP"]
    stats, diag = cabi.calibrate_host(pk.forcing_matrix(met), member, sc, topo.parent_offsets, topo.parent_ids,
                                      obs_m, desc, opt)
    assert not np.any(diag[..., 3])
    gof = z["gof"]     # [M][6 vars][N obs, NSE, log NSE, Spearman, r2, bias, nRMSD]
    cols = [str(c) for c in z["cols"]]
    for i in range(member.shape[0]):
        for v, (_sc, var) in enumerate(labels):
            want = gof[i, v]
            got = stats[i, v]
            assert got[0] == want[0], (i, var)                                   # n pairs
            for k_got, k_want, name in ((1, 1, "NSE"), (2, 2, "log NSE"), (4, 4, "r2"), (5, 5, "bias"), (6, 6, "nRMSD")):
                tol = 2e-5 * max(1.0, abs(want[k_want]))
                assert abs(got[k_got] - want[k_want]) <= tol, (i, var, name, got[k_got], want[k_want])
        # Gaussian log-likelihood against the oracle's restatement on the reference series
        for v, var, col in ((0, "Q", "Q_cumecs"), (2, "TDP", "TDP_mgl")):
            o = obs[1][var].reindex(met.index).to_numpy(float)
            sim = z["series"][i, :, cols.index(col)]
            m = member[i, pk.MEMBER_INDEX["err_m:" + var]]
            want = orc.gaussian_log_likelihood(o, sim, m)
            # (the oracle's expression is pinned to scipy's norm(sim, m*sim).logpdf(obs) — MCMC.ipynb:233-236 — in
            # tests/test_reference_pins_r2.py; sim here is the reference's own series, the device integrates its own)
            assert abs(stats[i, v, 3] - want) <= 1e-5 * max(1.0, abs(want)), (i, var, stats[i, v, 3], want)


def test_device_pointer_api_matches_host_api(cabi, golden_dir):
    import torch
    from simplyp_b200 import ensemble as ens, model as spm, packing as pk, tarland
    from simplyp_b200.engine import Engine
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    met = met.iloc[:90]
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    samples = ens.latin_hypercube(257, seed=3)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    forcing = pk.forcing_matrix(met)
    obs_m, desc, labels = pk.obs_arrays(obs, topo, met.index, ("Q", "TDP"))
    out_h, diag_h = cabi.run_host(forcing, member, sc, topo.parent_offsets, topo.parent_ids, opt)
    st_h, _ = cabi.calibrate_host(forcing, member, sc, topo.parent_offsets, topo.parent_ids, obs_m, desc, opt)
    eng = Engine(0)
    out_d, diag_d = eng.run(eng.to_device(forcing), eng.to_device(member), eng.to_device(sc), topo.parent_offsets,
                            topo.parent_ids, opt)
    st_d, _ = eng.calibrate(eng.to_device(forcing), eng.to_device(member), eng.to_device(sc), topo.parent_offsets,
                            topo.parent_ids, eng.to_device(obs_m), eng.to_device(desc), opt)
    torch.cuda.synchronize()
    assert np.array_equal(out_h, out_d.cpu().numpy())          # bitwise: same kernel, same inputs
    assert np.array_equal(diag_h, diag_d.cpu().numpy())
    assert np.array_equal(st_h, st_d.cpu().numpy(), equal_nan=True)   # Spearman column is NaN without rank_stats
    # calibration statistics are consistent with the full-output series of the same run
    q = out_h[:, 0, :, 5] * 51.7 * 1000 / 86400
    o = obs_m[0]
    ok = ~np.isnan(o)
    nse = 1 - ((o[ok] - q[:, ok]) ** 2).sum(axis=1) / ((o[ok] - o[ok].mean()) ** 2).sum()
    assert np.allclose(nse, st_h[:, 0, 1], rtol=1e-10, atol=1e-10)


def test_full_size_ensemble_properties(cabi):
    """BASELINE config 2 at full size (10^4 members, 2004): properties that need no oracle.
    (a) shard invariance: any contiguous block of members integrated alone is bitwise the same;
    (b) the reach volume/flow invariant Vr = L/(a_Q 86400) Qr^(1-b_Q) implied by ode_f (:127-131) holds;
    (c) statistics converge under tolerance refinement (default vs 10x tighter)."""
    from simplyp_b200 import ensemble as ens, model as spm, packing as pk, tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    samples = ens.latin_hypercube(10000, seed=20260101)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    forcing = pk.forcing_matrix(met)
    obs_m, desc, labels = pk.obs_arrays(obs, topo, met.index, ("Q", "TDP"))
    args = (topo.parent_offsets, topo.parent_ids, obs_m, desc)
    st, dg = cabi.calibrate_host(forcing, member, sc, *args, opt)
    assert np.all(np.isfinite(st[..., :3])) and not np.any(dg[..., 3])
    lo, hi = 3331, 5017
    st2, dg2 = cabi.calibrate_host(forcing, member[lo:hi], sc[lo:hi], *args, opt)
    assert np.array_equal(st2, st[lo:hi], equal_nan=True) and np.array_equal(dg2, dg[lo:hi])
    opt_t = spm.make_options(p_SU, p, dyn, topo, rtol=opt.rtol / 10, atol=opt.atol / 10)
    st3, _ = cabi.calibrate_host(forcing, member[:2048], sc[:2048], *args, opt_t)
    sse, sse3 = st[:2048, :, 7], st3[:, :, 7]
    assert np.max(np.abs(sse - sse3) / np.abs(sse3)) < 2e-5
    out, dgo = cabi.run_host(forcing[:120], member[:512], sc[:512], topo.parent_offsets, topo.parent_ids, opt)
    Qr, Vr = out[:, 0, :, 4], out[:, 0, :, 3]
    aQ, bQ = member[:512, pk.MEMBER_INDEX["a_Q"]][:, None], member[:512, pk.MEMBER_INDEX["b_Q"]][:, None]
    L = sc[:512, 0, pk.SC_INDEX["L_reach"]][:, None]
    assert np.max(np.abs(Vr - L / (aQ * 86400.0) * Qr ** (1 - bQ)) / Vr) < 1e-5


def test_synthetic_network_vs_oracle(cabi):
    """64-reach branching network (config-3 generator, small) against the oracle on a 40-day window."""
    from oracle import simplyp_oracle as orc
    from simplyp_b200 import model as spm, packing as pk, synthetic, tarland
    p_SU, dyn, p, p_LU, p_SC0, p_struc0, met, obs = tarland.load(dynamic="y")
    p, p_SC, p_struc = synthetic.random_network(p, p_SC0[1], n_sc=64, seed=3)
    met = met.iloc[:40]
    TC, R, diag, met = parity.run_single(cabi.run_host, (p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs))
    TCo, Ro, Kf, _ = orc.run_simply_p(met, p_struc, p_SU, p_LU, p_SC, p, dyn, rtol=1e-10, atol=1e-13, mxstep=50000)
    for SC in p["SC_list"]:
        parity.assert_frames_close(TC[SC], R[SC], TCo[SC], Ro[SC], "synthetic SC %d" % SC)


def test_network_ensemble_lanes_share_a_warp(cabi):
    """5 members x 5 reaches = 25 threads in ONE warp, parents and children side by side: the routing
    wavefront must neither deadlock nor depend on who shares a warp (each member alone gives the same bits)."""
    from simplyp_b200 import ensemble as ens, model as spm, packing as pk, tarland
    from tests.golden.networks import network5_inputs
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    p, p_LU, p_SC, p_struc = network5_inputs(p, p_LU, p_SC, p_struc)
    pk.validate_land_use(p_SC, p["SC_list"])
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    samples = ens.latin_hypercube(5, seed=11)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    forcing = pk.forcing_matrix(met.iloc[:200])
    out, dg = cabi.run_host(forcing, member, sc, topo.parent_offsets, topo.parent_ids, opt)
    assert np.all(np.isfinite(out)) and not np.any(dg[..., 3])
    for i in range(5):
        out1, _ = cabi.run_host(forcing, member[i:i + 1], sc[i:i + 1], topo.parent_offsets, topo.parent_ids, opt)
        assert np.array_equal(out1[0], out[i]), i
    # calibration mode of the same network: statistics at the outlet equal those of the full-output series
    obs5 = {5: obs[1]}
    obs_m, desc, labels = pk.obs_arrays(obs5, topo, met.iloc[:200].index, ("Q", "TDP"))
    st, _ = cabi.calibrate_host(forcing, member, sc, topo.parent_offsets, topo.parent_ids, obs_m, desc, opt)
    A5 = sc[:, 4, pk.SC_INDEX["A_catch"]][:, None]
    q = out[:, 4, :, 5] * A5 * 1000 / 86400
    o = obs_m[0]
    ok = ~np.isnan(o)
    nse = 1 - ((o[ok] - q[:, ok]) ** 2).sum(axis=1) / ((o[ok] - o[ok].mean()) ** 2).sum()
    assert np.allclose(nse, st[:, 0, 1], rtol=1e-10, atol=1e-10)


def test_epoch_sweep_is_invisible_in_the_results(cabi, monkeypatch):
    """Networks are swept in epochs (units of >= 256 days x a block of items, epoch-major, the midnight state of every item
    handed over through the workspace) so that the main-stem chains of one epoch run beside the headwaters of the next.
    The hand-over is exact: 64 reaches x 3 members x 700 days in epochs of 128 days (6 epochs), of 256 days (3) and in
    one piece (also the default of a launch this small: all its blocks are resident at once) give the same bits — full output, diagnostics, and the fused statistics of a calibration run."""
    from simplyp_b200 import ensemble as ens, model as spm, packing as pk, synthetic, tarland
    p_SU, dyn, p, p_LU, p_SC0, p_struc0, met, obs = tarland.load("2003-06-01", "2005-04-30", dynamic="y")
    assert len(met) == 700
    p, p_SC, p_struc = synthetic.random_network(p, p_SC0[1], n_sc=64, seed=3)
    pk.validate_land_use(p_SC, p["SC_list"])
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    samples = ens.latin_hypercube(3, seed=5)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    forcing = pk.forcing_matrix(met)
    outlet = topo.sc_ids[-1]
    obs_m, desc, labels = pk.obs_arrays({outlet: obs[1]}, topo, met.index, ("Q", "TDP"))
    res = {}
    for days in ("0", "128", "256", None):
        if days is None:
            monkeypatch.delenv("SIMPLYP_EPOCH_DAYS", raising=False)
        else:
            monkeypatch.setenv("SIMPLYP_EPOCH_DAYS", days)
        out, dg = cabi.run_host(forcing, member, sc, topo.parent_offsets, topo.parent_ids, opt)
        st, dgc = cabi.calibrate_host(forcing, member, sc, topo.parent_offsets, topo.parent_ids, obs_m, desc, opt)
        assert not np.any(dg[..., 3]) and not np.any(dgc[..., 3]) and np.all(np.isfinite(out))
        res[days] = (out, dg, st, dgc)
    for days in ("128", "256", None):
        for a, b in zip(res["0"], res[days]):
            assert np.array_equal(a, b, equal_nan=True), days
    assert res["0"][1][..., 0].min() > 700 * 3


def test_fused_gather_entry_point_on_one_rank(cabi):
    """simplyp_calibrate_gather_device (the calibration kernel that stores every member's statistics into all ranks'
    gather buffers over peer memory) with ONE rank: a peer-visible allocation, its IPC handle, three passes through the
    two alternating buffer sets, and statistics bit-identical to the plain entry point.  With N > 1 ranks the path is
    checked against NCCL under torchrun (scripts/check_peer_gather.py: bitwise on 2 and 8 GPUs)."""
    import torch
    import torch.distributed as dist
    from simplyp_b200 import ensemble as ens, model as spm, packing as pk, tarland
    from simplyp_b200.engine import Engine
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1)
    try:
        p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load("2004-01-01", "2004-06-30", dynamic="y")
        topo = pk.build_topology(p_struc, p["SC_list"])
        opt = spm.make_options(p_SU, p, dyn, topo)
        samples = ens.latin_hypercube(700, seed=13)
        member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
        obs_m, desc, labels = pk.obs_arrays(obs, topo, met.index, ("Q", "TDP"))
        eng = Engine(0)
        args = (eng.to_device(pk.forcing_matrix(met)), eng.to_device(member), eng.to_device(sc), topo.parent_offsets,
                topo.parent_ids, eng.to_device(obs_m), eng.to_device(desc), opt)
        want, dg0 = eng.calibrate(*args)
        pg = ens.PeerGather(700, (obs_m.shape[0], pk.NSTAT), eng.device)
        assert len(cabi.ipc_export(pg.base)) == 64
        try:
            for _ in range(3):
                got, dg = eng.calibrate(*args, peer_gather=pg)
                torch.cuda.synchronize()
                assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(want, nan=-7.0))
                assert torch.equal(dg, dg0) and int(dg[..., 3].max().item()) == 0
            with pytest.raises(cabi.SimplypError):
                opt.rank_stats = 1
                eng.calibrate(*args, peer_gather=pg)
        finally:
            opt.rank_stats = 0
            pg.close()
    finally:
        if created:
            dist.destroy_process_group()


def test_long_record_wraps_the_forcing_ring(cabi):
    """1,300 days = 11 forcing tiles through the 4-slot TMA ring (slots are re-armed and refilled), with
    members of very different speed in one block; member 0 is checked against the oracle over the whole window."""
    from oracle import simplyp_oracle as orc
    from simplyp_b200 import ensemble as ens, model as spm, packing as pk, tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load("1995-01-01", "1998-07-23", dynamic="y")
    assert len(met) == 1300
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    samples = ens.latin_hypercube(160, seed=21)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    forcing = pk.forcing_matrix(met)
    out, dg = cabi.run_host(forcing, member, sc, topo.parent_offsets, topo.parent_ids, opt)
    assert np.all(np.isfinite(out)) and not np.any(dg[..., 3])
    spd = dg[:, 0, 0] / 1300.0
    assert spd.max() / spd.min() > 2.0            # fast and slow members really share blocks
    out1, _ = cabi.run_host(forcing, member[7:8], sc[7:8], topo.parent_offsets, topo.parent_ids, opt)
    assert np.array_equal(out1[0], out[7])
    p0, pLU0, pSC0 = ens.apply_member_to_pandas(samples, 0, p, p_LU, p_SC)
    _TC, Ro, _Kf, _ = orc.run_simply_p(met, p_struc, p_SU, pLU0, pSC0, p0, dyn, rtol=1e-10, atol=1e-13, mxstep=50000)
    _tc, r = spm.raw_to_frames(out[0, 0], met.index, float(sc[0, 0, 0]), p["Msoil_m2"], p["f_TDP"], "None", None)
    for c in ("Q_cumecs", "SS_mgl", "TDP_mgl", "PP_mgl", "TP_mgl", "SRP_mgl"):
        assert max_rel(r[c].to_numpy(), Ro[1][c].to_numpy()) <= 1e-5, c


def test_edge_cases(cabi):
    from simplyp_b200 import model as spm, packing as pk, tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="n")
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    member, sc, forcing = pk.member_vector(p, p_LU)[None], pk.sc_matrix(p_SC, topo.sc_ids)[None], pk.forcing_matrix(met)
    out0, _ = cabi.run_host(forcing[:0], member, sc, topo.parent_offsets, topo.parent_ids, opt)    # empty period
    assert out0.shape == (1, 1, 0, 25)
    out1, _ = cabi.run_host(forcing[:1], member, sc, topo.parent_offsets, topo.parent_ids, opt)    # one day
    out3, _ = cabi.run_host(forcing[:3], member, sc, topo.parent_offsets, topo.parent_ids, opt)
    assert np.array_equal(out1[0, 0, 0], out3[0, 0, 0])
    # a dry spell: zero rain and zero PET for 200 days keeps every state finite and flow positive
    dry = forcing.copy()
    dry[:, 0] = 0.0
    dry[:, 1] = 0.0
    outd, dg = cabi.run_host(dry[:200], member, sc, topo.parent_offsets, topo.parent_ids, opt)
    assert np.all(np.isfinite(outd)) and np.all(outd[0, 0, :, 5] > 0) and not np.any(dg[..., 3])
    # bad arguments come back as errors, not crashes
    with pytest.raises(cabi.SimplypError):
        cabi.run_host(forcing, member, sc, np.array([0, 1], dtype=np.int32), np.array([0], dtype=np.int32), opt)


def _device_run(w, n_days=None, keep=None):
    """Full-output run through the device-pointer C-ABI (torch holds the buffers); returns (out, diag) on the GPU."""
    import torch
    from simplyp_b200 import packing as pk
    from simplyp_b200.engine import Engine
    eng = Engine(0)
    forcing = w["forcing"] if n_days is None else w["forcing"][:n_days]
    topo, sc = w["topo"], w["sc"]
    if keep is not None:                      # prune to a set of reaches that is closed under "upstream of"
        from simplyp_b200 import packing as pk2
        p_struc = w["p_struc"].loc[[topo.sc_ids[i] for i in keep]]
        topo = pk2.build_topology(p_struc, [topo.sc_ids[i] for i in keep])
        sc = sc[:, keep]
    opt = w["opt"]
    out, diag = eng.run(eng.to_device(forcing), eng.to_device(w["member"]), eng.to_device(sc), topo.parent_offsets,
                        topo.parent_ids, opt)
    torch.cuda.synchronize()
    return out, diag


def _upstream_closure(topo, i):
    keep, stack = set(), [i]
    while stack:
        j = stack.pop()
        if j in keep:
            continue
        keep.add(j)
        stack.extend(int(p) for p in topo.parent_ids[topo.parent_offsets[j]:topo.parent_offsets[j + 1]])
    return sorted(keep)


@pytest.mark.parametrize("cfg", [3, 5])
def test_scale_configs_full_size_properties(cabi, cfg):
    """BASELINE configs 3 (256 reaches, 30 years) and 5 (4096 reaches x 3 land uses, 50 years, 15 GB of daily
    output kept in HBM) at FULL size, through properties that need no oracle:
    (a) every value finite, no integrator status bit;
    (b) causality in time: the first 300 days of the long run equal a 300-day run, bitwise;
    (c) locality in the reach DAG: a reach's series depends only on its upstream sub-network — the network pruned
        to the upstream closure of one reach gives the same bits for those reaches;
    (d) the reach volume stays on ode_f's invariant curve Vr = L/(a_Q 86400) Qr^(1-b_Q) (:127-131, :457-459)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import exp_configs
    from simplyp_b200 import packing as pk
    w = exp_configs.build(cfg)
    # the reference's leaked NC_type (model.py:442,676) couples every reach to the LAST one of the run order, so
    # locality only holds with that quirk off
    w["opt"].strict_quirks = 0
    topo = w["topo"]
    out, diag = _device_run(w)
    assert bool(torch.isfinite(out).all().item()) and int(diag[..., 3].max().item()) == 0
    out_p, _ = _device_run(w, n_days=300)
    assert bool(torch.equal(out[:, :, :300], out_p))
    target = topo.n_sc // 3
    keep = sorted(set(_upstream_closure(topo, target)) | {0})     # reach 0 is p['SC_Qr0'] (initial flow, :386)
    assert 1 <= len(keep) < topo.n_sc
    out_k, _ = _device_run(w, n_days=300, keep=keep)
    assert bool(torch.equal(out_k, out[:, keep][:, :, :300]))
    Qr, Vr = out[0, :, :, 4], out[0, :, :, 3]
    aQ, bQ = w["member"][0, pk.MEMBER_INDEX["a_Q"]], w["member"][0, pk.MEMBER_INDEX["b_Q"]]
    L = torch.from_numpy(w["sc"][0, :, pk.SC_INDEX["L_reach"]]).to(out.device)[:, None]
    rel = ((Vr - L / (aQ * 86400.0) * Qr ** (1 - bQ)).abs() / Vr).max().item()
    assert rel < 1e-5, rel


def test_device_goodness_of_fit_table_with_spearman(cabi, golden_dir):
    """the reference's whole goodness-of-fit table (visualise_results.py:441-449, incl.
    Spearman's r on average ranks) reduced on the device for every member of an ensemble, against
    (a) the reference's own table for the shipped run (fixture ref_gof.json, member = shipped parameters) and
    (b) the host table (simplyp_b200.stats.gof_one, itself pinned to the reference) built from the same members'
        full-output series."""
    from simplyp_b200 import ensemble as ens, model as spm, packing as pk, stats as sps, tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    opt.rank_stats = 1
    samples = ens.latin_hypercube(40, seed=5)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    member[0] = pk.member_vector(p, p_LU)            # member 0 = the shipped parameter set
    sc[0] = pk.sc_matrix(p_SC, topo.sc_ids)
    forcing = pk.forcing_matrix(met)
    variables = ("Q", "SS", "TDP", "PP", "TP", "SRP")
    obs_m, desc, labels = pk.obs_arrays(obs, topo, met.index, variables)
    st, dg = cabi.calibrate_host(forcing, member, sc, topo.parent_offsets, topo.parent_ids, obs_m, desc, opt)
    assert st.shape == (40, len(labels), pk.NSTAT) and not np.any(dg[..., 3])
    out, _ = cabi.run_host(forcing, member, sc, topo.parent_offsets, topo.parent_ids, opt)
    for i in (0, 7, 23, 39):
        _tc, r = spm.raw_to_frames(out[i, 0], met.index, float(sc[i, 0, pk.SC_INDEX["A_catch"]]), p["Msoil_m2"],
                                   float(member[i, pk.MEMBER_INDEX["f_TDP"]]), "None", None)
        table = sps.gof_table_from_device(st[i], labels)
        for k, (reach, var) in enumerate(labels):
            want = sps.gof_one(obs[reach][var], r[sps.SIM_COLUMN[var]])
            got = table.loc[var, sps.STATS_COLUMNS].to_numpy(dtype=float)
            assert np.allclose(got, np.asarray(want, dtype=float), rtol=2e-6, atol=2e-6), (i, var, got, want)
    # (a) the reference's own table (dynamic options on, shipped parameters), at the 1e-5 parity bound
    ref = json.load(open(os.path.join(golden_dir, "ref_gof.json")))
    tab0 = sps.gof_table_from_device(st[0], labels)
    tight = ref["dyny_tight"]                      # unmodified reference, odeint at rtol=1e-10
    for var, row in zip(tight["index"], tight["values"]):
        for col, val in zip(tight["columns"], row):
            if col in sps.STATS_COLUMNS:
                assert abs(tab0.loc[var, col] - val) <= 1e-5 * max(1.0, abs(val)), (var, col, tab0.loc[var, col], val)


def test_sum_to_waterbody_on_device(cabi):
    """sum_to_waterbody (model.py:851-900) as a device reduction over the flagged reaches for
    every member and day, against the host implementation (pinned to the reference) on the 5-reach network."""
    import torch
    from simplyp_b200 import ensemble as ens, model as spm, packing as pk, tarland
    from simplyp_b200.engine import Engine
    from tests.golden.networks import network5_inputs
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    p, p_LU, p_SC, p_struc = network5_inputs(p, p_LU, p_SC, p_struc)
    p_struc = p_struc.copy()
    p_struc["In_final_flux?"] = [np.nan, np.nan, 1.0, 1.0, 1.0]
    nc_types = pk.validate_land_use(p_SC, p["SC_list"])
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    samples = ens.latin_hypercube(6, seed=3)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    met = met.iloc[:120]
    eng = Engine(0)
    d_mem, d_sc = eng.to_device(member), eng.to_device(sc)
    out, _ = eng.run(eng.to_device(pk.forcing_matrix(met)), d_mem, d_sc, topo.parent_offsets, topo.parent_ids, opt)
    reaches = [i for i, s in enumerate(topo.sc_ids) if p_struc.loc[s, "In_final_flux?"] == 1]
    wb = eng.sum_to_waterbody(out, d_mem, d_sc, reaches).cpu().numpy()
    out_h = out.cpu().numpy()
    for i in range(6):
        R = {}
        for k, SC in enumerate(topo.sc_ids):
            _tc, R[SC] = spm.raw_to_frames(out_h[i, k], met.index, float(sc[i, k, pk.SC_INDEX["A_catch"]]),
                                           p["Msoil_m2"], float(member[i, pk.MEMBER_INDEX["f_TDP"]]), nc_types[SC], None)
        want = spm.sum_to_waterbody(p_struc, len(topo.sc_ids), R, float(member[i, pk.MEMBER_INDEX["f_TDP"]]))
        for j, col in enumerate(Engine.WATERBODY_COLUMNS):
            assert np.allclose(wb[i, :, j], want[col].to_numpy(), rtol=1e-12, atol=0), (i, col)


def test_snow_parameters_in_the_ensemble(cabi):
    """D_snow_0 and f_DDSM sampled per member, snow_hydrol_inputs (inputs.py:159-210) fused into
    the kernel.  Every member must equal, bitwise, the run of that member alone on forcing pre-processed on the host
    with its own snow parameters (the reference's order of operations)."""
    from simplyp_b200 import ensemble as ens, inputs as spi, model as spm, packing as pk, tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load("2003-10-01", "2004-09-30", dynamic="y")
    topo = pk.build_topology(p_struc, p["SC_list"])
    ranges = dict(ens.TARLAND_RANGES)
    ranges["D_snow_0"] = (0.0, 40.0)
    ranges["f_DDSM"] = (1.0, 4.0)
    samples = ens.latin_hypercube(48, ranges=ranges, seed=9)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    opt = spm.make_options(p_SU, p, dyn, topo)
    opt.snow_on_device = 1
    raw = met[["T_air", "PET", "Precipitation"]]
    out, dg = cabi.run_host(pk.forcing_matrix(raw, raw_snow=True), member, sc, topo.parent_offsets, topo.parent_ids, opt)
    assert np.all(np.isfinite(out)) and not np.any(dg[..., 3])
    opt0 = spm.make_options(p_SU, p, dyn, topo)
    melted = 0.0
    for i in (0, 13, 47):
        met_i = spi.snow_hydrol_inputs(samples["D_snow_0"][i], samples["f_DDSM"][i], raw)
        melted += float(met_i["P_melt"].sum())
        out_i, _ = cabi.run_host(pk.forcing_matrix(met_i), member[i:i + 1], sc[i:i + 1], topo.parent_offsets,
                                 topo.parent_ids, opt0)
        assert np.array_equal(out_i[0], out[i]), i
    assert melted > 0


@pytest.mark.parametrize("area", [0.05, 5.0])
def test_stiff_reach_vs_oracle(cabi, area):
    """The Rosenbrock path of the quad kernel (stiff main-stem-like reach) against LSODA/BDF at tight tolerance."""
    per_day = parity.check_stiff_chain(cabi.run_host, area, max_steps_per_day=60)
    assert per_day > 20


def test_retired_scalar_kernel_is_refused(cabi):
    """lanes_per_item = 1 (the round-1 one-thread-per-item kernel) is no longer on the product ABI."""
    from simplyp_b200 import model as spm, packing as pk, tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    opt.lanes_per_item = 1
    with pytest.raises(cabi.SimplypError):
        cabi.run_host(pk.forcing_matrix(met.iloc[:5]), pk.member_vector(p, p_LU)[None],
                      pk.sc_matrix(p_SC, topo.sc_ids)[None], topo.parent_offsets, topo.parent_ids, opt)


def test_against_the_reference_shipped_csvs(cabi, golden_dir):
    """The CUDA path against the reference's own shipped golden CSVs (to the reference's solver noise)."""
    parity.check_shipped_golden(cabi.run_host, golden_dir)


@pytest.mark.gpu
def test_thornthwaite_pet_on_device(cabi, golden_dir):
    """SURVEY §8f rank 3 (PET half): daily_PET (inputs.py:232-312) as a device kernel.  Against the host function
    (whose helpers are pinned bit-identical to the reference's, test_oracle_pins) over the 30-year Tarland record,
    and against the reference helpers' monthly values on the 16th of each month; tolerance 1e-12 relative (CUDA's
    pow/sin/tan/acos differ from libm's in the last bits).  Error behaviour equals the host function's."""
    import calendar
    import json
    import pandas as pd
    import simplyp_b200 as sp
    from simplyp_b200 import inputs, tarland
    z = np.load(os.path.join(tarland.DATA_DIR, "tarland_met.npz"), allow_pickle=True)
    idx = pd.date_range("1981-01-01", periods=len(z["T_air"]), freq="D")
    met = pd.DataFrame({"T_air": z["T_air"].astype(float), "PET": z["PET"].astype(float)}, index=idx)
    want = sp.daily_PET(57.1, met)
    got = inputs.daily_PET_device(57.1, met)
    assert list(got.columns) == list(want.columns) and got.index.equals(want.index)
    assert np.allclose(got["PET"].to_numpy(), want["PET"].to_numpy(), rtol=1e-12, atol=0)
    assert "PET" in met.columns and np.array_equal(met["PET"].to_numpy(), z["PET"].astype(float))   # input untouched
    with open(os.path.join(golden_dir, "ref_pet.json")) as f:
        ref = json.load(f)
    for year, rec in ref["years"].items():
        # ref_pet.json holds the reference HELPERS' values with the calendar-correct daylight table; the reference's
        # wrapper keeps the leap-year table after its first leap year (1984 in this record, inputs.py:269-273), so the
        # non-leap years after it are pinned by ref_pet_daily.npz instead (test_device_pet_equals_the_reference_wrapper)
        if int(year) > 1984 and not calendar.isleap(int(year)):
            continue
        for mon in range(12):
            n = calendar.monthrange(int(year), mon + 1)[1]
            assert got.loc["%s-%02d-16" % (year, mon + 1), "PET"] == pytest.approx(rec["pet_mm_month"][mon] / n, rel=1e-12)
    # a southern latitude and a record with months below zero (counted as zero, inputs.py:489)
    cold = pd.DataFrame({"T_air": z["T_air"].astype(float) - 4.0}, index=idx)
    a, b = inputs.daily_PET_device(-33.9, cold["1981":"1989"]), sp.daily_PET(-33.9, cold["1981":"1989"])
    assert np.allclose(a["PET"].to_numpy(), b["PET"].to_numpy(), rtol=1e-12, atol=0)
    with pytest.raises(ValueError):
        inputs.daily_PET_device(57.1, met.iloc[:400])
    with pytest.raises(ValueError):
        inputs.daily_PET_device(95.0, met)


def test_ensemble_container_equals_one_shot_run(cabi, tmp_path):
    """the full-output ensemble container (chunked runs, pinned asynchronous D2H, one .npy per
    raw variable) holds bit-for-bit what a single full-output launch of all members returns, for a chunk size that
    does not divide the ensemble (ragged last chunk) and for a column subset."""
    from simplyp_b200 import ensemble as ens, model as spm, packing as pk, tarland
    from simplyp_b200.engine import Engine
    from tests.golden.networks import network5_inputs
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    p, p_LU, p_SC, p_struc = network5_inputs(p, p_LU, p_SC, p_struc)
    met = met.iloc[:90]
    samples = ens.latin_hypercube(7, seed=11)
    man = ens.run_ensemble_to_dir(met, p_struc, p_SU, p_LU, p_SC, p, dyn, samples, str(tmp_path / "all"),
                                  members_per_chunk=3)
    topo = pk.build_topology(p_struc, p["SC_list"])
    p_SC2 = p_SC.copy(deep=True)
    pk.validate_land_use(p_SC2, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC2, topo.sc_ids), samples)
    eng = Engine(0)
    out, diag = eng.run(eng.to_device(pk.forcing_matrix(met)), eng.to_device(member), eng.to_device(sc),
                        topo.parent_offsets, topo.parent_ids, opt)
    out, diag = out.cpu().numpy(), diag.cpu().numpy()
    assert man["shape"] == [7, 5, 90] and len(man["variables"]) == pk.NOUT
    for j, col in enumerate(pk.RAW_COLS):
        got = np.load(os.path.join(str(tmp_path / "all"), man["variables"][col]))
        assert got.shape == (7, 5, 90) and np.array_equal(got, out[:, :, :, j]), col
    d = np.load(os.path.join(str(tmp_path / "all"), "diag.npy"))
    assert np.array_equal(d[:, :, 3], diag[:, :, 3]) and (d[:, :, 0] > 0).all()
    with open(os.path.join(str(tmp_path / "all"), "manifest.json")) as f:
        assert json.load(f)["sub_catchments"] == [int(s) for s in topo.sc_ids]
    man2 = ens.run_ensemble_to_dir(met, p_struc, p_SU, p_LU, p_SC, p, dyn, samples, str(tmp_path / "two"),
                                   columns=["Qr", "TDP_kg/day"], members_per_chunk=100)
    assert sorted(man2["variables"]) == ["Qr", "TDP_kg/day"]
    got = np.load(os.path.join(str(tmp_path / "two"), "TDP_kg_per_day.npy"))
    assert np.array_equal(got, out[:, :, :, pk.RAW_COLS.index("TDP_kg/day")])


@pytest.mark.parametrize("M", [4800, 6001, 9500, 10000, 13999])
def test_planned_placement_is_invisible_in_the_results(cabi, M, monkeypatch):
    """Latency-bound ensembles (more blocks than SMs, fewer than 3 per SM) are placed on the SMs by plan (claim by
    %smid, reversed partner blocks, warps led by one heavy member; beyond 2 blocks per SM the 168-register build with
    three light blocks resident on some SMs and 3 x 148 blocks launched, of which the spare ones leave): the
    statistics and diagnostics of every member must be bit-identical to the plain launch.  Sizes: a few partner
    blocks only (4800), a ragged last block (6001), one, seventeen and 142 SMs with three blocks (9500, 10000, 13999:
    the last one ragged as well)."""
    import torch
    import bench
    from simplyp_b200 import model as spm, packing as pk
    from simplyp_b200.engine import Engine
    eng = Engine(0)
    w = bench.build_workload("2004", M)
    opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"], 1.0, None, None)
    forcing = w["forcing"][:100]
    obs_m = np.ascontiguousarray(w["obs_m"][:, :100])
    args = (eng.to_device(forcing), eng.to_device(w["member"]), eng.to_device(w["sc"]), w["topo"].parent_offsets,
            w["topo"].parent_ids, eng.to_device(obs_m), eng.to_device(w["desc"]), opt)
    res = {}
    for plan in ("0", "1", "no pilot"):
        monkeypatch.setenv("SIMPLYP_SM_PLAN", "0" if plan == "0" else "1")
        opt.pilot_days = -1 if plan == "no pilot" else 0
        stats, diag = eng.calibrate(*args)
        torch.cuda.synchronize()
        res[plan] = (stats.cpu().numpy(), diag.cpu().numpy())
    opt.pilot_days = 0
    assert np.array_equal(res["0"][0], res["1"][0], equal_nan=True)
    assert np.array_equal(res["0"][1], res["1"][1])
    # the pilot is the first 8 days of the run itself (the main launch continues from its midnight state in another
    # member order): a run without pilot, integrated in one piece, must give the same bits
    assert np.array_equal(res["no pilot"][0], res["1"][0], equal_nan=True)
    assert np.array_equal(res["no pilot"][1], res["1"][1])
    assert (res["1"][1][:, 0, 0] > 0).all() and (res["1"][1][:, 0, 3] == 0).all()      # every member ran, status clean
    if M == 6001:                                  # the full-output mode takes the same placement
        runs = {}
        for plan in ("0", "1", "no pilot"):
            monkeypatch.setenv("SIMPLYP_SM_PLAN", "0" if plan == "0" else "1")
            opt.pilot_days = -1 if plan == "no pilot" else 0
            out, diag = eng.run(args[0][:70], args[1], args[2], args[3], args[4], opt)
            torch.cuda.synchronize()
            runs[plan] = (out.cpu().numpy(), diag.cpu().numpy())
            del out
        opt.pilot_days = 0
        assert np.array_equal(runs["0"][0], runs["1"][0]) and np.array_equal(runs["0"][1], runs["1"][1])
        assert np.array_equal(runs["no pilot"][0], runs["1"][0]) and np.array_equal(runs["no pilot"][1], runs["1"][1])
        assert np.isfinite(runs["1"][0]).all()


def test_device_entry_points_capture_into_a_cuda_graph(cabi):
    """include/simplyp_b200.h: the *_device entry points enqueue on the caller's stream WITHOUT synchronising — also
    for networks, whose stiff/non-stiff grouping of the reaches is now made on the device.  Proof: a full-output run
    and a calibration run of a 5-reach network are captured into CUDA graphs (a synchronising call would abort the
    capture) and the replays give the bits of the direct calls."""
    import torch
    from simplyp_b200 import ensemble as ens, model as spm, packing as pk, tarland
    from simplyp_b200.engine import Engine
    from tests.golden.networks import network5_inputs
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    p, p_LU, p_SC, p_struc = network5_inputs(p, p_LU, p_SC, p_struc)
    pk.validate_land_use(p_SC, p["SC_list"])
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    samples = ens.latin_hypercube(37, seed=4)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    met = met.iloc[:150]
    obs_m, desc, labels = pk.obs_arrays({5: obs[1]}, topo, met.index, ("Q", "TDP"))
    eng = Engine(0)
    d_f, d_m, d_s = eng.to_device(pk.forcing_matrix(met)), eng.to_device(member), eng.to_device(sc)
    d_o, d_d = eng.to_device(obs_m), eng.to_device(desc)
    po, pid = topo.parent_offsets, topo.parent_ids
    out_ref, diag_ref = eng.run(d_f, d_m, d_s, po, pid, opt)                 # direct calls (also the warm-up: the
    st_ref, _ = eng.calibrate(d_f, d_m, d_s, po, pid, d_o, d_d, opt)         # topology image and workspace exist now)
    torch.cuda.synchronize()
    out_ref, diag_ref, st_ref = out_ref.clone(), diag_ref.clone(), st_ref.clone()
    out_g, diag_g = torch.zeros_like(out_ref), torch.zeros_like(diag_ref)
    st_g = torch.zeros_like(st_ref)
    dg2 = torch.zeros_like(diag_ref)
    g_run, g_cal = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_run):
        eng.run(d_f, d_m, d_s, po, pid, opt, out=out_g, diag=diag_g)
    with torch.cuda.graph(g_cal):
        eng.calibrate(d_f, d_m, d_s, po, pid, d_o, d_d, opt, stats=st_g, diag=dg2)
    for _ in range(2):
        out_g.zero_(); diag_g.zero_(); st_g.zero_()
        g_run.replay()
        g_cal.replay()
        torch.cuda.synchronize()
        assert torch.equal(out_g, out_ref) and torch.equal(diag_g, diag_ref)
        assert torch.equal(st_g.nan_to_num(nan=-7.0), st_ref.nan_to_num(nan=-7.0))
    assert int(diag_ref[..., 3].max().item()) == 0


def test_host_entry_points_from_concurrent_threads(cabi):
    """The *_host entry points keep one buffer cache per device behind a per-device lock: host threads that drive
    different devices (or, on a one-GPU box, the same device) get the results of the single-threaded calls while
    their ensembles — of different sizes, so the caches are re-allocated under each other's feet — interleave."""
    import threading
    from simplyp_b200 import ensemble as ens, model as spm, packing as pk, tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    met = met.iloc[:60]
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    forcing = pk.forcing_matrix(met)
    obs_m, desc, labels = pk.obs_arrays(obs, topo, met.index, ("Q", "TDP"))
    n_dev = cabi.load().simplyp_device_count()
    jobs = []
    for t, M in enumerate((300, 1100, 64, 700)):
        samples = ens.latin_hypercube(M, seed=40 + t)
        member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
        jobs.append((member, sc, t % n_dev))
    want = [(cabi.run_host(forcing, m, s, topo.parent_offsets, topo.parent_ids, opt, device=d)[0],
             cabi.calibrate_host(forcing, m, s, topo.parent_offsets, topo.parent_ids, obs_m, desc, opt, device=d)[0])
            for m, s, d in jobs]
    got, errors = [None] * len(jobs), []

    def work(i):
        try:
            m, s, d = jobs[i]
            for _ in range(3):
                out, _dg = cabi.run_host(forcing, m, s, topo.parent_offsets, topo.parent_ids, opt, device=d)
                st, _dg = cabi.calibrate_host(forcing, m, s, topo.parent_offsets, topo.parent_ids, obs_m, desc, opt, device=d)
            got[i] = (out, st)
        except Exception as e:       # noqa: BLE001 - surfaced below
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for i in range(len(jobs)):
        assert np.array_equal(got[i][0], want[i][0]), i
        assert np.array_equal(got[i][1], want[i][1], equal_nan=True), i


def test_spearman_of_a_record_longer_than_shared_memory(cabi):
    """Spearman's r for series of more than 12,800 observed days (round 1 returned NaN there): 14,000 days of
    synthetic daily observations; the device value against the host table (pandas ranks, pinned to the reference)."""
    from simplyp_b200 import ensemble as ens, inputs as spi, model as spm, packing as pk, stats as sps, synthetic, tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met0, obs0 = tarland.load(dynamic="y")
    met = spi.snow_hydrol_inputs(p["D_snow_0"], p["f_DDSM"], synthetic.synthetic_met(14000, seed=5))
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    samples = ens.latin_hypercube(3, seed=8)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    forcing = pk.forcing_matrix(met)
    out, dg = cabi.run_host(forcing, member, sc, topo.parent_offsets, topo.parent_ids, opt)
    assert not np.any(dg[..., 3])
    rng = np.random.default_rng(0)
    q0 = out[0, 0, :, 5] * float(sc[0, 0, pk.SC_INDEX["A_catch"]]) * 1000 / 86400
    q_obs = np.round(q0 * np.exp(rng.normal(0, 0.3, len(q0))), 3)          # rounded: ties in the observations
    q_obs[rng.random(len(q0)) < 0.02] = np.nan
    obs = {1: pd.DataFrame({"Q": q_obs}, index=met.index)}
    obs_m, desc, labels = pk.obs_arrays(obs, topo, met.index, ("Q",))
    assert np.isfinite(obs_m[0]).sum() > 12800
    opt.rank_stats = 1
    st, _ = cabi.calibrate_host(forcing, member, sc, topo.parent_offsets, topo.parent_ids, obs_m, desc, opt)
    for i in range(3):
        _tc, r = spm.raw_to_frames(out[i, 0], met.index, float(sc[i, 0, pk.SC_INDEX["A_catch"]]), p["Msoil_m2"],
                                   float(member[i, pk.MEMBER_INDEX["f_TDP"]]), "None", None)
        want = dict(zip(sps.STATS_COLUMNS, sps.gof_one(obs[1]["Q"], r["Q_cumecs"])))
        # 2e-6 like the short-record test: the device forms Q_cumecs as Qr*A*(1000/86400), pandas as Qr*A*1000/86400;
        # a last-bit difference can swap two neighbouring ranks out of 13,700
        assert abs(st[i, 0, 8] - want["Spearmans r"]) <= 2e-6, (i, st[i, 0], want)     # SIMPLYP_ST_SPEARMAN = 8


@pytest.mark.parametrize("key", ["val", "half"])
def test_run_modes_vs_reference(cabi, golden_dir, key):
    """run_mode='val' (Kf from p['Kf'], model.py:449-453) and step_len=0.5 (model.py:193,640) through the C-ABI against
    the unmodified reference's runs (fixture ref_modes.npz, odeint at rtol=1e-10)."""
    from tests.test_reference_pins_r2 import check_mode
    check_mode(cabi.run_host, golden_dir, key)


def test_csv_writer_matches_the_reference(cabi, golden_dir, tmp_path):
    """p_SU.save_output_csvs == 'y' (model.py:815-825): same files, same header lines (column order after
    sort_index, the six dropped reach columns), same number of rows and index cells as the reference writes."""
    import simplyp_b200 as sp
    from simplyp_b200 import tarland
    from tests.golden.networks import network5_inputs
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    p, p_LU, p_SC, p_struc = network5_inputs(p, p_LU, p_SC, p_struc)
    p_SU["save_output_csvs"] = "y"
    p_SU["output_fpath"] = str(tmp_path)
    TC, R, Kf, info = sp.run_simply_p(met.iloc[:20], p_struc, p_SU, p_LU, p_SC, p, dyn, verbose=False)
    ref = json.load(open(os.path.join(golden_dir, "ref_csv.json")))
    assert sorted(os.listdir(tmp_path)) == sorted(ref)
    for name, rec in ref.items():
        lines = open(os.path.join(str(tmp_path), name)).read().splitlines()
        assert lines[0] == rec["header"], name
        assert len(lines) - 1 == rec["n_rows"], name
        assert lines[1].split(",")[0] == rec["first_index"] and lines[-1].split(",")[0] == rec["last_index"], name
    got = pd.read_csv(os.path.join(str(tmp_path), "Instream_results_Reach5.csv"), index_col=0)
    assert np.allclose(got["TDP_mgl"].to_numpy(), R[5]["TDP_mgl"].to_numpy(), rtol=1e-14)


@pytest.mark.parametrize("cfg", [3, 5])
def test_scale_configs_full_topology_vs_oracle(cabi, golden_dir, cfg):
    """BASELINE configs 3 (256 reaches) and 5 (4096 reaches, 487 levels) at their FULL topology against the oracle
    (LSODA at rtol=1e-10 — BDF on the main stem) over 730 / 120 days: fixture tests/golden/ref_config{3,5}.npz
    (make_scale_golden.py) holds the outlet, the stiffest and the deepest reaches and a spread of others.  This is the
    check of the Rosenbrock path (run at 30x the tolerance, simplyp_quad.cuh) through every level of the network:
    daily flows and concentrations <= 1e-5 relative, the other columns within the mixed bound."""
    from simplyp_b200 import model as spm, packing as pk, synthetic
    z = np.load(os.path.join(golden_dir, "ref_config%d.npz" % cfg))
    w = synthetic.scale_config(cfg, n_days=int(z["n_days"]))
    assert w["topo"].n_sc == int(z["n_sc"]) and np.array_equal(w["forcing"][:5], z["forcing_head"])
    out, diag = _device_run(w)
    assert int(diag[..., 3].max().item()) == 0
    sel = [int(s) for s in z["reaches"]]
    got = out[0, sel].cpu().numpy()
    met, p, p_SC, topo = w["met"], w["p"], w["p_SC"], w["topo"]
    nc_types = pk.validate_land_use(p_SC.copy(), p["SC_list"])
    worst = 0.0
    for k, s in enumerate(sel):
        SC = topo.sc_ids[s]
        A = float(p_SC.loc["A_catch", SC])
        tc, r = spm.raw_to_frames(got[k], met.index, A, p["Msoil_m2"], p["f_TDP"], nc_types[SC], met["D_snow_end"])
        tco, ro = spm.raw_to_frames(z["raw"][k], met.index, A, p["Msoil_m2"], p["f_TDP"], nc_types[SC], met["D_snow_end"])
        parity.assert_frames_close(tc, r, tco, ro, "config %d reach %d (level %d)" % (cfg, s, int(z["levels"][k])))
        worst = max(worst, max(max_rel(r[c].to_numpy(), ro[c].to_numpy()) for c in ("Q_cumecs", "SS_mgl", "TDP_mgl", "PP_mgl")))
    # every reach of the network on the last day of the window (flow), as a checksum over the reaches not sampled
    qr_last = out[0, :, -1, 5].cpu().numpy()
    assert max_rel(qr_last, z["qr_last_day"]) <= 1e-5
    # the stiff path was really taken: the stiffest sampled reach needs fewer than 100 attempts per day
    spd = diag[0, sel, 0].cpu().numpy() / float(z["n_days"])
    assert spd.max() < 100, spd
    print("config %d: worst flow/concentration error over %d sampled reaches %.2e" % (cfg, len(sel), worst))


def _oracle_member(args):
    i, n_total, n_days = args
    import bench
    from oracle import simplyp_oracle as orc
    from simplyp_b200 import ensemble as ens
    w = bench.build_workload("2004", n_total)
    pi, pLUi, pSCi = ens.apply_member_to_pandas(w["samples"], i, w["p"], w["p_LU"], w["p_SC"])
    _TC, R, _Kf, _ = orc.run_simply_p(w["met"].iloc[:n_days], w["p_struc"], w["p_SU"], pLUi, pSCi, pi, w["dyn"],
                                      rtol=1e-10, atol=1e-13, mxstep=50000)
    return i, R[1][["Q_cumecs", "SS_mgl", "TDP_mgl", "PP_mgl", "Qr", "Msus_kg/day", "TDP_kg/day", "PP_kg/day"]].to_numpy(float)


def test_bench_ensemble_members_vs_oracle(cabi):
    """The ensemble bench.py times (10^4 Latin-hypercube members, seed 20260101), integrated as ONE launch of all 10^4
    (cost pilot, planned placement, lock-step warps of unequal members), against the oracle port on 256 of its members
    spread over the whole cost order (every 39th): daily flows and concentrations <= 1e-5 relative.  The oracle runs
    on the GPU box's host cores (LSODA at rtol=1e-10), one member per task."""
    import multiprocessing as mp
    import bench
    from simplyp_b200 import model as spm, packing as pk
    M = 10000
    w = bench.build_workload("2004", M)
    opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"])
    cores = os.cpu_count() or 1
    picked = list(range(0, M, 39))[:256] if cores >= 8 else list(range(0, M, 39))[:32 * cores]
    out, dg = cabi.run_host(w["forcing"], w["member"], w["sc"], w["topo"].parent_offsets, w["topo"].parent_ids, opt)
    assert not np.any(dg[..., 3])
    with mp.get_context("spawn").Pool(min(cores, 32)) as pool:
        res = pool.map(_oracle_member, [(i, M, 366) for i in picked], chunksize=1)
    worst = []
    for i, want in res:
        A = float(w["sc"][i, 0, pk.SC_INDEX["A_catch"]])
        _tc, r = spm.raw_to_frames(out[i, 0], w["met"].index, A, w["p"]["Msoil_m2"], w["p"]["f_TDP"], "None", None)
        got = r[["Q_cumecs", "SS_mgl", "TDP_mgl", "PP_mgl", "Qr", "Msus_kg/day", "TDP_kg/day", "PP_kg/day"]].to_numpy(float)
        e = max(max_rel(got[:, k], want[:, k]) for k in range(got.shape[1]))
        assert e <= 1e-5, (i, e)
        worst.append(e)
    print("bench ensemble: %d members vs oracle, worst %.2e, median %.2e" % (len(worst), max(worst), float(np.median(worst))))


def test_device_pet_equals_the_reference_wrapper(cabi, golden_dir):
    """thornthwaite_kernel against the reference's own daily_PET wrapper (fixture ref_pet_daily.npz by
    make_golden_r2.py), including the leap-year daylight table the reference keeps after its first leap year."""
    from simplyp_b200 import inputs, tarland
    z = np.load(os.path.join(tarland.DATA_DIR, "tarland_met.npz"))
    idx = pd.date_range(str(z["day0"]), periods=int(z["n"]), freq="D")
    t_air = pd.DataFrame({"T_air": z["T_air"].astype(float)}, index=idx)
    ref = np.load(os.path.join(golden_dir, "ref_pet_daily.npz"))
    for a, b in (("2001", "2007"), ("1981", "1983")):
        got = inputs.daily_PET_device(float(ref["latitude"]), t_air[a:b])
        assert max_rel(got["PET"].to_numpy(), ref["pet_%s_%s" % (a, b)]) < 1e-12, (a, b)
