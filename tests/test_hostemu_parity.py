"""CPU tier: the kernels' per-item arithmetic and control flow (compiled for the host by tests/hostemu,
a test harness — the product has no CPU path) against the live-reference fixtures.  Both device programs are
covered: the quad program (4 lanes per item, simplyp_quad.cuh — the default kernel) and the one-thread-per-item
program (simplyp_thread.cuh)."""
import numpy as np
import pytest

from tests import hostemu, parity

RUNNERS = {"quad": hostemu.run_quad, "scalar": hostemu.run}


@pytest.mark.parametrize("prog", ["quad", "scalar"])
@pytest.mark.parametrize("dy", ["n", "y"])
def test_tarland_2004(golden_dir, dy, prog):
    parity.check_tarland(RUNNERS[prog], golden_dir, dy)


@pytest.mark.parametrize("prog", ["quad", "scalar"])
def test_branching_network(golden_dir, prog):
    parity.check_network(RUNNERS[prog], golden_dir)


@pytest.mark.parametrize("prog", ["quad", "scalar"])
def test_ensemble_members(golden_dir, prog):
    parity.check_ensemble_series(RUNNERS[prog], golden_dir, members=[0, 5, 11])


def test_quad_rhs_equals_scalar_rhs(golden_dir):
    """ode_f in quad form (u = ln Qr, Vr integrated, lane coefficients) against the scalar rhs() at random states."""
    from simplyp_b200 import ensemble as ens, packing as pk, tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    topo = pk.build_topology(p_struc, p["SC_list"])
    samples = ens.latin_hypercube(16, seed=11)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    rng = np.random.default_rng(5)
    worst = 0.0
    for i in range(16):
        fc = member[i, pk.MEMBER_INDEX["fc"]]
        y7 = [fc * rng.uniform(0.97, 1.03), fc * rng.uniform(0.97, 1.03), rng.uniform(5, 150), rng.uniform(0.05, 8),
              rng.uniform(0, 5e3), rng.uniform(0, 3), rng.uniform(0, 3)]
        us = [rng.uniform(0, 2), rng.uniform(0, 100), rng.uniform(0, 1), rng.uniform(0, 1)]
        worst = max(worst, hostemu.quad_rhs_check(member[i], sc[i, 0], rng.uniform(0, 30), rng.uniform(0, 4),
                                                  float(rng.integers(1, 366)), us, y7))
    assert worst < 1e-12, worst


def test_table_exp_accuracy():
    """sp_exp_tab (64-entry table, one-FMA reduction, degree-4 polynomial) against libm: relative accuracy 4e-14 over
    the arguments the RHS visits (|x| < 40), 1e-13 out to |x| = 700 — the integration tolerance it serves is 1e-7."""
    x = np.linspace(-40, 30, 20001)
    assert np.max(np.abs(hostemu.exp_tab(x) - np.exp(x)) / np.exp(x)) < 5e-14
    x = np.random.default_rng(1).uniform(-700, 700, 20000)
    assert np.max(np.abs(hostemu.exp_tab(x) - np.exp(x)) / np.exp(x)) < 1e-13


def test_snow_on_device_equals_host_preprocessing(golden_dir):
    """the degree-day snow recursion (inputs.py:159-210) fused into the day-start code with
    per-member D_snow_0 / f_DDSM gives the same bits as running snow_hydrol_inputs on the host first."""
    from simplyp_b200 import inputs as spi, model as spm, packing as pk, tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
    topo = pk.build_topology(p_struc, p["SC_list"])
    for D0, fd in ((0.0, 2.74), (35.0, 1.3)):
        p2 = p.copy(deep=True)
        p2["D_snow_0"], p2["f_DDSM"] = D0, fd
        met2 = spi.snow_hydrol_inputs(D0, fd, met[["T_air", "PET", "Precipitation"]])
        member = pk.member_vector(p2, p_LU)[None]
        sc = pk.sc_matrix(p_SC, topo.sc_ids)[None]
        opt = spm.make_options(p_SU, p2, dyn, topo)
        want, _ = hostemu.run_quad(pk.forcing_matrix(met2), member, sc, topo.parent_offsets, topo.parent_ids, opt)
        opt.snow_on_device = 1
        got, _ = hostemu.run_quad(pk.forcing_matrix(met2, raw_snow=True), member, sc, topo.parent_offsets,
                                  topo.parent_ids, opt)
        assert np.array_equal(got, want)
        assert met2["P_melt"].sum() > 0          # the period really has snow


@pytest.mark.parametrize("area", [0.05, 1.0, 5.0])
def test_stiff_reach_takes_the_rosenbrock_path(area):
    """Main-stem-like reach (flow 300 ... 32,000 mm/d over its own area): the quad program switches that reach to
    Kaps-Rentrop 4(3) with the exact Jacobian; parity bound as everywhere, and far fewer attempts than the explicit
    pair needs (one-thread-per-item program: 112 ... 690 per day)."""
    per_day = parity.check_stiff_chain(hostemu.run_quad, area, max_steps_per_day=60)
    assert per_day > 20


@pytest.mark.parametrize("prog", ["quad", "scalar"])
def test_against_the_reference_shipped_csvs(golden_dir, prog):
    parity.check_shipped_golden(RUNNERS[prog], golden_dir)


def test_record_continued_from_a_stored_midnight_state_is_bitwise_the_same():
    """The cost pilot integrates the first days of the record and hands every member's midnight state to the main
    launch (QuadCarry): days [0, k) + [k, D) must give exactly the bits of [0, D) — outputs and counters."""
    from simplyp_b200 import ensemble as ens, model as spm, packing as pk, tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load("2004-01-01", "2004-03-10", dynamic="y")
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    samples = ens.latin_hypercube(6, seed=5)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    forcing = pk.forcing_matrix(met)
    want, dg = hostemu.run_quad(forcing, member, sc, topo.parent_offsets, topo.parent_ids, opt)
    for split in (1, 8, 37):
        got, dg2 = hostemu.run_quad_split(forcing, member, sc, topo.parent_offsets, topo.parent_ids, opt, split)
        assert np.array_equal(got, want) and np.array_equal(dg2, dg), split


@pytest.mark.parametrize("resident", [2, 3])
@pytest.mark.parametrize("n_sm", [148, 132])
def test_placement_plan_covers_every_member_and_block_once(n_sm, resident):
    """Arithmetic of the planned placement (simplyp_plan.cuh): for every ensemble size in the planned range the member
    layout is a permutation and every virtual block sits in exactly one list.  2 resident blocks per SM: only the Q
    overflow lists hold two blocks and the launch has one block per list.  3 resident blocks per SM: no list holds two
    blocks, the launch has 3 n_sm blocks, and the Q SMs of the light blocks have a first-, second- and third-list.  The
    heaviest block's partner is the lightest of the partner region."""
    for M in (32 * n_sm + 1, 4800, 6001, 32 * 2 * n_sm - 7, 32 * 2 * n_sm, 32 * 2 * n_sm + 1, 9500, 10000,
              32 * (2 * n_sm + n_sm // 4)):
        B = (M + 31) // 32
        for solo in (0, 24):
            res = hostemu.plan(M, n_sm, solo, resident)
            if B <= n_sm:
                assert res is None
                continue
            idx, lst, pos, (nY, nP, Q, n_lists, n_launch, res_used) = res
            assert np.array_equal(np.sort(idx), np.arange(M)), (M, solo)              # a permutation of the members
            assert Q == max(0, B - 2 * n_sm) and nY == n_sm - Q and nY + nP + 3 * Q == B
            assert res_used == (3 if (resident == 3 and Q > 0) else 2)
            assert (lst >= 0).all() and len(np.unique(lst)) == n_lists
            counts = np.bincount(lst, minlength=3 * n_sm)
            if res_used == 2:
                assert n_lists == n_launch == n_sm + nP + Q
                assert (counts[:nY] == 1).all() and (counts[nY:n_sm] == 2).all() and counts.max() <= 2
                assert sorted(pos[lst == nY].tolist()) == ([0, 1] if Q else sorted(pos[lst == nY].tolist()))
            else:
                assert n_lists == B and n_launch == 3 * n_sm and counts.max() == 1 and (pos == 0).all()
                assert (counts[:n_sm] == 1).all() and (counts[n_sm + nY:2 * n_sm] == 1).all()
                assert (counts[2 * n_sm:2 * n_sm + nY] == 0).all() and (counts[2 * n_sm + nY:] == 1).all()
                # the three blocks of a light SM are the ones the chained plan puts into one slot and beside it
                chain = hostemu.plan(M, n_sm, solo, 2)[1]
                for x in range(Q):
                    trio = sorted(int(np.where(lst == c * n_sm + nY + x)[0][0]) for c in range(3))
                    assert trio == sorted(np.where((chain == nY + x) | (chain == n_sm + nY + x))[0].tolist())
            # rank 0 leads virtual block 0; its partner block (list n_sm + 0) holds the lightest ranks of the partner region
            assert idx[0] == 0
            partner_block = int(np.where(lst == n_sm)[0][0])
            ranks_in_partner = np.where(idx // 32 == partner_block)[0]
            other = np.where(lst == n_sm + nP - 1)[0]
            if nP > 1 and solo == 0 and not (Q == 0 and M % 32):
                assert ranks_in_partner.min() > np.where(idx // 32 == int(other[0]))[0].max()
