"""CPU tier: the kernels' per-thread arithmetic and control flow (compiled for the host by tests/hostemu,
a test harness — the product has no CPU path) against the live-reference fixtures."""
import pytest

from tests import hostemu, parity


@pytest.mark.parametrize("dy", ["n", "y"])
def test_tarland_2004(golden_dir, dy):
    parity.check_tarland(hostemu.run, golden_dir, dy)


def test_branching_network(golden_dir):
    parity.check_network(hostemu.run, golden_dir)


def test_ensemble_members(golden_dir):
    parity.check_ensemble_series(hostemu.run, golden_dir, members=[0, 5, 11])
