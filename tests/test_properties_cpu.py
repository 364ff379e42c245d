"""CPU tier, property tests (hypothesis) of the host logic around the C-ABI: topology packing, level computation,
member sharding, Latin-hypercube sampling, observation alignment."""
import numpy as np
import pandas as pd
from hypothesis import given, settings, strategies as st

from simplyp_b200 import _cabi, ensemble as ens, packing as pk


@st.composite
def trees(draw):
    """Random upstream-first networks: reach i drains into one j in (i, min(i+5, n)]."""
    n = draw(st.integers(min_value=1, max_value=40))
    down = {i: draw(st.integers(min_value=i + 1, max_value=min(i + 5, n))) for i in range(1, n)}
    ups = {i: [] for i in range(1, n + 1)}
    for i, j in down.items():
        ups[j].append(i)
    cells = []
    for i in range(1, n + 1):
        u = sorted(ups[i])
        cells.append(np.nan if not u else (float(u[0]) if len(u) == 1 else ", ".join(str(x) for x in u)))
    p_struc = pd.DataFrame({"Upstream_SCs": pd.Series(cells, index=range(1, n + 1), dtype=object),
                            "In_final_flux?": np.nan}, index=pd.Index(range(1, n + 1), name="Reach"))
    return n, ups, p_struc


@settings(max_examples=60, deadline=None)
@given(trees())
def test_topology_csr_and_levels(tree):
    n, ups, p_struc = tree
    topo = pk.build_topology(p_struc, np.arange(1, n + 1))
    assert topo.n_sc == n and topo.parent_offsets[0] == 0 and topo.parent_offsets[-1] == sum(len(u) for u in ups.values())
    for i in range(n):
        got = sorted(int(p) + 1 for p in topo.parent_ids[topo.parent_offsets[i]:topo.parent_offsets[i + 1]])
        assert got == sorted(ups[i + 1])
        assert all(p < i for p in topo.parent_ids[topo.parent_offsets[i]:topo.parent_offsets[i + 1]])
    n_levels, lv = _cabi.topology_levels(topo.parent_offsets, topo.parent_ids)
    want = np.zeros(n, dtype=int)
    for i in range(n):
        ps = topo.parent_ids[topo.parent_offsets[i]:topo.parent_offsets[i + 1]]
        want[i] = 0 if len(ps) == 0 else 1 + max(want[p] for p in ps)
    assert np.array_equal(lv, want) and n_levels == want.max() + 1


@settings(max_examples=100, deadline=None)
@given(st.integers(min_value=0, max_value=10 ** 6), st.integers(min_value=1, max_value=16))
def test_shards_partition_the_members(n, world):
    bounds = [ens.shard_bounds(n, world, r) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n
    assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
    sizes = [hi - lo for lo, hi in bounds]
    assert max(sizes) - min(sizes) <= 1


@settings(max_examples=25, deadline=None)
@given(st.integers(min_value=1, max_value=300), st.integers(min_value=0, max_value=2 ** 31 - 1))
def test_latin_hypercube_is_stratified(n, seed):
    s = ens.latin_hypercube(n, seed=seed)
    for name, (lo, hi) in ens.TARLAND_RANGES.items():
        x = (s[name] - lo) / (hi - lo)
        assert np.all((x >= 0) & (x <= 1))
        assert sorted(np.floor(x * n).clip(max=n - 1).astype(int)) == list(range(n))     # one sample per stratum


@settings(max_examples=40, deadline=None)
@given(st.integers(min_value=11, max_value=200), st.integers(min_value=0, max_value=10 ** 6))
def test_observations_align_with_the_run_period(n_obs, seed):
    rng = np.random.default_rng(seed)
    days = pd.date_range("2004-01-01", periods=366, freq="D")
    pick = np.sort(rng.choice(366, size=min(n_obs, 366), replace=False))
    vals = rng.uniform(0.1, 5, size=len(pick))
    obs = {1: pd.DataFrame({"Q": vals}, index=days[pick])}
    p_struc = pd.DataFrame({"Upstream_SCs": pd.Series([np.nan], index=[1], dtype=object), "In_final_flux?": [np.nan]},
                           index=pd.Index([1], name="Reach"))
    topo = pk.build_topology(p_struc, [1])
    window = days[40:300]
    m, desc, labels = pk.obs_arrays(obs, topo, window, ("Q", "TDP"))
    assert labels == [(1, "Q")] and desc.tolist() == [[0, pk.VAR_INDEX["Q"]]]
    inside = (pick >= 40) & (pick < 300)
    assert np.array_equal(np.where(~np.isnan(m[0]))[0], pick[inside] - 40)
    assert np.array_equal(m[0][~np.isnan(m[0])], vals[inside])
