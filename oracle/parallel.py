"""TEST INFRASTRUCTURE — the oracle's ``run_network`` spread over host processes, level by level.

The reference integrates a network sub-catchment by sub-catchment, upstream first (``model.py:365``), and a reach
reads only the finished daily series of its DIRECT parents (``:508-544``).  Reaches of one topological level are
therefore independent given the levels above them; this driver hands each of them to a worker process
(``run_network(only=[sc], preset={parents})``) and assembles the same arrays the serial call returns — bit-identical,
because each reach is integrated by the same code on the same inputs.  Used by the fixture generator
(``tests/golden/make_scale_golden.py``) and by ``bench.py``'s CPU legs for the network configurations.
"""
from __future__ import annotations

import multiprocessing as mp

import numpy as np

from . import simplyp_oracle as orc

_G = {}


def _init(args):
    _G["args"] = args


def _one(task):
    SC, preset = task
    a = _G["args"]
    raw = orc.run_network(a["P"], a["PET"], a["doy"], a["p"], a["p_LU"], a["p_SC"], a["upstream"], only=[SC], preset=preset,
                          **a["kw"])
    i = raw["sc_ids"].index(SC)
    return SC, raw["ode"][i], raw["nonode"][i], raw["Kf"][SC], raw["nfe"], raw["nst"], raw["nc_type"][SC]


def levels_of(sc_ids, upstream):
    lvl = {}
    for SC in sc_ids:                                   # run order is upstream-first
        lvl[SC] = 1 + max((lvl[u] for u in upstream.get(SC, [])), default=-1)
    return lvl


def run_network_parallel(forcing_P, forcing_PET, doy, p, p_LU, p_SC, upstream, processes=None, **kw):
    """Same arguments and result as ``simplyp_oracle.run_network`` (no ``only``/``preset``)."""
    sc_ids = list(p_SC.keys())
    lvl = levels_of(sc_ids, upstream)
    D = len(forcing_P) if kw.get("n_days") is None else int(kw["n_days"])
    ode = np.zeros((len(sc_ids), D, 12))
    non = np.zeros((len(sc_ids), D, 13))
    pos = {SC: i for i, SC in enumerate(sc_ids)}
    Kf, nc_type, nfe, nst = {}, {}, 0, 0
    args = dict(P=np.asarray(forcing_P, float), PET=np.asarray(forcing_PET, float), doy=np.asarray(doy), p=p, p_LU=p_LU,
                p_SC=p_SC, upstream=upstream, kw=kw)
    with mp.get_context("fork").Pool(processes or mp.cpu_count(), initializer=_init, initargs=(args,)) as pool:
        for L in range(max(lvl.values()) + 1):
            tasks = [(SC, {u: (ode[pos[u]], non[pos[u]]) for u in upstream.get(SC, [])}) for SC in sc_ids if lvl[SC] == L]
            for SC, o, n, kf, f, s, nc in pool.imap_unordered(_one, tasks):
                ode[pos[SC]], non[pos[SC]] = o, n
                Kf[SC], nc_type[SC] = kf, nc
                nfe += f
                nst += s
    return {"ode": ode, "nonode": non, "Kf": Kf, "nfe": nfe, "nst": nst, "sc_ids": sc_ids, "nc_type": nc_type}
