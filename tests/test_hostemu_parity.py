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
    """sp_exp_tab (table + degree-6 polynomial) against libm over the range the RHS visits."""
    x = np.concatenate([np.linspace(-40, 20, 20001), np.random.default_rng(1).uniform(-700, 700, 20000)])
    got = hostemu.exp_tab(x)
    want = np.exp(x)
    assert np.max(np.abs(got - want) / want) < 5e-16      # 2 ulp
