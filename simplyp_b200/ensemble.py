"""Parameter ensembles: sampling, packing, sharding over GPUs, fused calibration statistics.

The reference has no ensemble driver in its release; the only construct of this kind is the
development notebook's emcee sampler over an IPython.parallel pool
(``Development/2016/MCMC.ipynb:29-31,380``), i.e. independent parameter sets evaluated in parallel.
Here the members of an ensemble are the threads of one kernel launch, and the ranks of a
``torch.distributed`` job each take a contiguous block of members; the only
collective is one all-gather of the per-member fit statistics.
"""
from __future__ import annotations

import numpy as np

from . import packing as pk

#: Latin-hypercube ranges bracketing the Tarland values.
#: key: member field (``packing.MEMBER_FIELDS``), or ``sc:<row>`` for a p_SC row applied to every SC.
TARLAND_RANGES = {
    "f_quick": (0.0, 0.1), "alpha": (0.5, 1.2), "fc": (100.0, 400.0), "beta": (0.3, 0.9), "T_g": (20.0, 150.0),
    "T_s:A": (1.0, 8.0), "T_s:S": (3.0, 20.0), "a_Q": (0.2, 0.9), "b_Q": (0.3, 0.6), "Qg_min": (0.05, 0.6),
    "E_M": (300.0, 3000.0), "k_M": (1.2, 2.5), "E_PP": (1.0, 3.0), "TDPg": (0.0, 0.05),
    "EPC0_init_mgl:A": (0.03, 0.3), "sc:TDPeff": (0.0, 0.5),
    "err_m:Q": (0.05, 1.0), "err_m:TDP": (0.05, 1.0),
}


def latin_hypercube(n, ranges=None, seed=20260101):
    """``n`` stratified samples per parameter; returns dict name -> array [n]."""
    ranges = TARLAND_RANGES if ranges is None else ranges
    rng = np.random.default_rng(seed)
    out = {}
    for name, (lo, hi) in ranges.items():
        strata = (rng.permutation(n) + rng.random(n)) / n
        out[name] = lo + (hi - lo) * strata
    return out


def pack_members(base_member, base_sc, samples):
    """Broadcast the base parameter vectors over the ensemble and apply the sampled columns.

    base_member [NP_MEMBER], base_sc [S][NP_SC], samples: dict name -> [M].
    Returns (member_params [M][NP_MEMBER], sc_params [1 or M][S][NP_SC]).
    """
    M = len(next(iter(samples.values()))) if samples else 1
    member = np.repeat(np.asarray(base_member, dtype=np.float64)[None, :], M, axis=0)
    sc = np.asarray(base_sc, dtype=np.float64)[None, :, :]
    per_member_sc = any(k.startswith("sc:") for k in samples)
    if per_member_sc:
        sc = np.repeat(sc, M, axis=0)
    for name, vals in samples.items():
        vals = np.asarray(vals, dtype=np.float64)
        if name.startswith("sc:"):
            row = name[3:]
            if "@" in row:                       # sc:<row>@<position in run order>
                row, at = row.split("@")
                sc[:, int(at), pk.SC_INDEX[row]] = vals
            else:
                sc[:, :, pk.SC_INDEX[row]] = vals[:, None]
        else:
            member[:, pk.MEMBER_INDEX[name]] = vals
    return np.ascontiguousarray(member), np.ascontiguousarray(sc)


def apply_member_to_pandas(samples, i, p, p_LU, p_SC):
    """Copies of the reference's pandas objects with member ``i``'s sampled values written in — this is how
    the oracle / the reference is run on the same member."""
    p, p_LU, p_SC = p.copy(deep=True), p_LU.copy(deep=True), p_SC.copy(deep=True)
    for name, vals in samples.items():
        v = float(vals[i])
        if name.startswith("err_m:"):
            continue
        if name.startswith("sc:"):
            row = name[3:]
            if "@" in row:
                row, at = row.split("@")
                p_SC.loc[row, p_SC.columns[int(at)]] = v
            else:
                p_SC.loc[row, :] = v
        elif ":" in name:
            row, col = name.split(":")
            p_LU.loc[row, col] = v
        else:
            p[name] = v
    return p, p_LU, p_SC


def shard_bounds(n_members, world_size, rank):
    """Contiguous block [lo, hi) of members for ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(n_members, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_stats(local_stats, n_members, group=None):
    """All-gather the per-member statistics [M_local][V][8] of every rank into [M][V][8] (rank order).

    Uses ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).  Blocks are padded to the largest
    shard so one ``all_gather_into_tensor`` suffices.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local_stats
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_members, world, r) for r in range(world)]
    mmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mmax,) + tuple(local_stats.shape[1:]), dtype=local_stats.dtype, device=local_stats.device)
    pad[: local_stats.shape[0]] = local_stats
    gathered = torch.empty((world * mmax,) + tuple(local_stats.shape[1:]), dtype=local_stats.dtype,
                           device=local_stats.device)
    dist.all_gather_into_tensor(gathered, pad, group=group)
    parts = [gathered[r * mmax: r * mmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


def calibrate_ensemble(met_df, p_struc, p_SU, p_LU, p_SC, p, dynamic_options, obs_dict, samples,
                       variables=("Q", "TDP"), step_len=1.0, rtol=None, atol=None, engine=None, gather=True,
                       snow_on_device=False, rank_stats=False):
    """Fused-statistics run of an ensemble; this rank integrates its shard and (optionally) all-gathers.

    ``snow_on_device``: ``met_df`` carries the raw ``Precipitation`` and ``T_air``; the degree-day snow module
    (reference ``inputs.py:159-210``) runs per member on the device, so ``D_snow_0`` and ``f_DDSM`` may be sampled.
    ``rank_stats``: also reduce Spearman's r (the whole table of ``goodness_of_fit_stats``).

    Returns ``(stats [M][V][8] torch tensor on the device, labels [(reach, variable)], diag)``.
    """
    import torch
    import torch.distributed as dist

    from .engine import Engine
    from .model import make_options

    p_LU, p_SC = p_LU.copy(deep=True), p_SC.copy(deep=True)
    pk.validate_land_use(p_SC, p["SC_list"])
    pk.check_erosion_windows(p)
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = make_options(p_SU, p, dynamic_options, topo, step_len, rtol, atol)
    opt.snow_on_device = 1 if snow_on_device else 0
    opt.rank_stats = 1 if rank_stats else 0
    member, sc = pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    M = member.shape[0]
    rank, world = (dist.get_rank(), dist.get_world_size()) if (dist.is_available() and dist.is_initialized()) else (0, 1)
    lo, hi = shard_bounds(M, world, rank)
    obs, desc, labels = pk.obs_arrays(obs_dict, topo, met_df.index, variables)
    eng = engine or Engine()
    d_forc = eng.to_device(pk.forcing_matrix(met_df, raw_snow=bool(snow_on_device)))
    d_mem = eng.to_device(member[lo:hi])
    d_sc = eng.to_device(sc if sc.shape[0] == 1 else sc[lo:hi])
    d_obs = eng.to_device(obs)
    d_desc = eng.to_device(desc)
    stats, diag = eng.calibrate(d_forc, d_mem, d_sc, topo.parent_offsets, topo.parent_ids, d_obs, d_desc, opt)
    if gather and world > 1:
        stats = all_gather_stats(stats, M)
    torch.cuda.synchronize(eng.device)
    return stats, labels, diag
